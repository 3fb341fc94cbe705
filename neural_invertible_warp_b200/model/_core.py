"""Shared implementation behind the per-model ``Graph`` / ``NeRF`` classes.

The reference states the same render arithmetic four times (model/nerf.py:243-483,
model/nerf_inn_llff.py:485-818, model/nerf_inn_dtu.py:363-680, model/nerf_gaussian.py); here it
exists once.  ``NeRFCore`` owns the reference's parameter objects (``mlp_feat.{i}``, ``mlp_rgb.{i}``
``nn.Linear`` -- same names, shapes and initialisation, so ``state_dict`` round-trips) and maps
``forward_samples`` / ``composite`` onto the CUDA operators; ``RenderCore`` is the ray pipeline
(ray generation -> stratified depths -> fused encoding+MLP -> compositing -> optional inverse-CDF
resampling and second pass) that ``Graph.render`` / ``render_local`` / ``render_by_slices`` share.

Tensors cross this layer in the reference's shapes ([B,P,3], [B,P,N,1] ...); the kernels see them
flattened over rays.
"""
import contextlib
import math

import torch
from torch import nn

from .. import camera
from .. import functional as F
from ..config import AttrDict

edict = AttrDict   # the reference returns EasyDict; AttrDict has the same access semantics


def get_layer_dims(layers):
    """reference util.py (get_layer_dims): [(k_in, k_out)] of consecutive entries."""
    return list(zip(layers[:-1], layers[1:]))


def _arch_supported(opt):
    a = opt.arch
    feat = list(a.layers_feat)
    rgb = list(a.layers_rgb)
    ok = (len(feat) == 9 and all(w == 256 for w in feat[1:]) and list(a.skip) == [4]
          and len(rgb) == 3 and rgb[1] == 128 and rgb[2] == 3 and a.posenc and a.posenc.L_3D == 10
          and a.posenc.L_view == 4 and bool(opt.nerf.view_dep) and a.density_activ == "softplus")
    return ok


class NeRFCore(nn.Module):
    """Parameter-compatible with reference ``model.nerf.NeRF`` (model/nerf.py:367-483)."""

    has_progress = False   # BARF variants own a ``progress`` Parameter (model/barf.py:254)
    # opt-in of the engine (engine.use_flat_gradients): the MLP kernels accumulate parameter gradients straight into the
    # flat bucket the parameters' ``.grad`` are slices of, instead of returning them through autograd
    accumulate_grads_in_place = False

    def __init__(self, opt):
        super().__init__()
        self.define_network(opt)
        if self.has_progress:
            self.progress = nn.Parameter(torch.tensor(0.))

    # -- construction ---------------------------------------------------------------------
    def define_network(self, opt):
        if not _arch_supported(opt):
            raise RuntimeError(
                "niw_b200: the CUDA path implements the reference architecture only (8x256 feature MLP, skip at "
                "layer 4, 128-wide view-dependent RGB head, posenc L_3D=10 / L_view=4, softplus density); got "
                "arch=%r -- there is no fallback" % (dict(opt.arch),))
        d3 = 3 + 6 * opt.arch.posenc.L_3D
        dv = 3 + 6 * opt.arch.posenc.L_view
        self.mlp_feat = nn.ModuleList()
        self.total_param = 0
        dims = get_layer_dims(opt.arch.layers_feat)
        for li, (k_in, k_out) in enumerate(dims):
            if li == 0:
                k_in = d3
            if li in opt.arch.skip:
                k_in += d3
            if li == len(dims) - 1:
                k_out += 1
            lin = nn.Linear(k_in, k_out)
            if opt.arch.tf_init:
                self.tensorflow_init_weights(opt, lin, out="first" if li == len(dims) - 1 else None)
            self.mlp_feat.append(lin)
            self.total_param += lin.weight.numel()
        self.mlp_rgb = nn.ModuleList()
        dims = get_layer_dims(opt.arch.layers_rgb)
        for li, (k_in, k_out) in enumerate(dims):
            if li == 0:
                k_in = opt.arch.layers_feat[-1] + dv
            lin = nn.Linear(k_in, k_out)
            if opt.arch.tf_init:
                self.tensorflow_init_weights(opt, lin, out="all" if li == len(dims) - 1 else None)
            self.mlp_rgb.append(lin)
            self.total_param += lin.weight.numel()

    def tensorflow_init_weights(self, opt, linear, out=None):
        """Xavier-uniform with ReLU gain on hidden rows (model/nerf.py:404-414)."""
        gain = math.sqrt(2.0)
        with torch.no_grad():
            if out == "all":
                nn.init.xavier_uniform_(linear.weight)
            elif out == "first":
                nn.init.xavier_uniform_(linear.weight[:1])
                nn.init.xavier_uniform_(linear.weight[1:], gain=gain)
            else:
                nn.init.xavier_uniform_(linear.weight, gain=gain)
            nn.init.zeros_(linear.bias)

    # -- helpers ----------------------------------------------------------------------------
    def mlp_parameters(self):
        """The 20 MLP parameter tensors in state_dict order (weights then bias per layer)."""
        ps = []
        for lin in list(self.mlp_feat) + list(self.mlp_rgb):
            ps += [lin.weight, lin.bias]
        return ps

    def _is_flat(self, tensors, base):
        """True if ``tensors`` are consecutive contiguous slices starting at data pointer ``base``."""
        off = 0
        for t in tensors:
            if t is None or not t.is_contiguous() or t.data_ptr() != base + 4 * off:
                return False
            off += t.numel()
        return True

    def flat_parameters(self):
        """The 530 052 MLP parameters as ONE fp32 vector in state_dict order, without a copy: the
        ``nn.Linear`` parameters are kept as views into a flat buffer owned by this module (re-made
        whenever they stop being views, e.g. after ``.to(device)``; ``load_state_dict`` and the
        optimisers update in place and keep them).  state_dict keys and shapes are untouched."""
        ps = self.mlp_parameters()
        flat = getattr(self, "_flat_values", None)
        if (flat is None or flat.data_ptr() != ps[0].data_ptr()) and self._is_flat([p.data for p in ps], ps[0].data_ptr()):
            # already consecutive inside someone else's buffer (engine.FlatAdam): view it, do not move it
            flat = ps[0].data.as_strided((sum(p.numel() for p in ps),), (1,))
            self._flat_values = flat
        if flat is None or flat.device != ps[0].device or not self._is_flat([p.data for p in ps], flat.data_ptr()):
            with torch.no_grad():
                flat = torch.cat([p.data.reshape(-1) for p in ps])
                off = 0
                for p in ps:
                    p.data = flat[off:off + p.numel()].view_as(p)
                    off += p.numel()
            self._flat_values = flat
        return flat

    def flat_grad_pointer(self):
        """Device pointer of a flat fp32 gradient buffer whose consecutive slices are the ``.grad`` of the
        20 MLP parameters, or None when their gradients are laid out otherwise (then autograd accumulates
        them).  If no parameter has a gradient yet, this module's own zeroed flat buffer is attached."""
        ps = self.mlp_parameters()
        if any(not p.requires_grad for p in ps):
            return None
        grads = [p.grad for p in ps]
        if all(g is None for g in grads):
            fg = getattr(self, "_flat_grads", None)
            if fg is None or fg.device != ps[0].device:
                fg = torch.zeros(sum(p.numel() for p in ps), device=ps[0].device)
                self._flat_grads = fg
            else:
                fg.zero_()
            off = 0
            for p in ps:
                p.grad = fg[off:off + p.numel()].view_as(p)
                off += p.numel()
            return fg.data_ptr()
        if grads[0] is not None and self._is_flat(grads, grads[0].data_ptr()):
            return grads[0].data_ptr()
        return None

    def c2f_schedule(self, opt):
        """(progress device scalar, (start, end)) of the BARF coarse-to-fine encoding (model/barf.py:256-268),
        or (None, None) for plain NeRF / unset ``barf_c2f``.  The band weights are evaluated on the
        device from the ``progress`` Parameter: no host read per step."""
        if self.has_progress and opt.get("barf_c2f") is not None:
            return self.progress.data, tuple(opt.barf_c2f)
        return None, None

    _warned_default = False

    @staticmethod
    def precision(opt, mode=None):
        """MLP operand precision of a call.  ``arch.mlp_precision`` (fp32 | bf16 | bf16x3), when the YAML sets it, rules
        every mode (``arch.mlp_precision_eval`` overrides it for val / eval renders).  An unmodified reference YAML sets
        neither; then the defaults follow the contract of BASELINE.json: rendering whose outputs are the product
        (val / eval / direct calls) runs the split-precision tensor-core path (bf16x3: rgb / depth / opacity within 1e-3
        of the reference's fp32), optimisation steps (train / test-optim) run BF16 operands (gradients within 1e-2
        relative) -- announced once, because training numerics then differ from the reference's fp32 GEMMs."""
        a = opt.arch
        get = a.get if hasattr(a, "get") else (lambda k, d=None: getattr(a, k, d))
        p = get("mlp_precision", None)
        if mode in ("val", "eval") and get("mlp_precision_eval", None) is not None:
            return get("mlp_precision_eval")
        if p is not None:
            return p
        if mode in ("train", "test-optim"):
            if not NeRFCore._warned_default:
                NeRFCore._warned_default = True
                import warnings
                warnings.warn("niw_b200: arch.mlp_precision is not set: optimisation steps run the MLP with BF16 tensor-core "
                              "operands (FP32 accumulate; parameter gradients within 1e-2 relative of the reference's fp32 "
                              "path), val / eval renders with split BF16 operands (bf16x3, outputs within 1e-3).  Set "
                              "arch.mlp_precision=bf16x3 or fp32 for reference-tight training numerics.", stacklevel=3)
            return "bf16"
        return "bf16x3"

    def _check_mode(self, opt, mode):
        if opt.nerf.density_noise_reg and mode == "train":
            raise RuntimeError("niw_b200: nerf.density_noise_reg is not implemented in the CUDA path")

    # -- reference API ----------------------------------------------------------------------
    def forward(self, opt, points_3D, ray_unit=None, mode=None):
        """model/nerf.py:416-447 on explicit points: each point is treated as a one-sample ray
        (x = p + 0 * v), so the same fused kernel serves this entry point."""
        self._check_mode(opt, mode)
        if ray_unit is None:
            raise AssertionError("view-dependent network: ray_unit is required")
        shape = points_3D.shape[:-1]
        pts = points_3D.reshape(-1, 3)
        view = ray_unit.expand_as(points_3D).reshape(-1, 3)
        depth = torch.zeros(pts.shape[0], 1, device=pts.device)
        progress, c2f = self.c2f_schedule(opt)
        rgb, sigma = F.nerf_forward_samples(self.flat_parameters(), pts, view, depth, progress, c2f,
                                            self.precision(opt, mode), module=self)
        return rgb.view(*shape, 3), sigma.view(*shape)

    def forward_samples(self, opt, center, ray, depth_samples, mode=None, prepacked=None):
        """model/nerf.py:449-456: center, ray [B,P,3], depth_samples [B,P,N,1] ->
        rgb_samples [B,P,N,3], density_samples [B,P,N]."""
        self._check_mode(opt, mode)
        B, P, N = depth_samples.shape[:3]
        progress, c2f = self.c2f_schedule(opt)
        rgb, sigma = F.nerf_forward_samples(self.flat_parameters(), center.reshape(B * P, 3), ray.reshape(B * P, 3),
                                            depth_samples.reshape(B * P, N), progress, c2f, self.precision(opt, mode),
                                            module=self, prepacked=prepacked)
        return rgb.view(B, P, N, 3), sigma.view(B, P, N)

    def prepack(self, opt, n_rays, n_samples, mode="train"):
        """Workspace of the coming ``forward_samples`` call with the weight streams already packed (current stream);
        None when the precision has nothing to hoist.  Training-ness is decided as that call will decide it."""
        training = torch.is_grad_enabled() and any(p.requires_grad for p in self.mlp_parameters())
        progress, c2f = self.c2f_schedule(opt)
        return F.nerf_prepack(self.flat_parameters(), n_rays, n_samples, progress, c2f, self.precision(opt, mode), training), training

    def composite(self, opt, ray, rgb_samples, density_samples, depth_samples, want_prob=True):
        """model/nerf.py:458-474 -> rgb [B,P,3], depth [B,P,1], opacity [B,P,1], prob [B,P,N,1]
        (``want_prob=False``, used by the render pipeline when nothing samples from the weights: prob is None)."""
        B, P, N = density_samples.shape
        bg = opt.data.bgcolor if opt.nerf.setbg_opaque else None
        target = F.mse_target
        if target is not None and F.composite_mse_supported(N) and target.image.shape[0] == B and \
                (target.ray_idx is None or len(target.ray_idx) == P):
            # train-mode render of a Graph that will compare these colours with the target's pixels (model/nerf.py:276-288):
            # the loss comes out of the compositor's epilogue; ``compute_loss`` finds it by the identity of ``rgb``
            rgb, depth, opacity, prob, loss = F.composite_mse(
                ray.reshape(B * P, 3), rgb_samples.reshape(B * P, N, 3), density_samples.reshape(B * P, N),
                depth_samples.reshape(B * P, N), target, B, P, bg, want_prob=want_prob)
            rgb = rgb.view(B, P, 3)
            target.put(rgb, loss)
        else:
            rgb, depth, opacity, prob = F.composite(ray.reshape(B * P, 3), rgb_samples.reshape(B * P, N, 3),
                                                    density_samples.reshape(B * P, N), depth_samples.reshape(B * P, N), bg,
                                                    want_prob=want_prob)
            rgb = rgb.view(B, P, 3)
        return rgb, depth.view(B, P, 1), opacity.view(B, P, 1), (prob.view(B, P, N, 1) if want_prob else None)

    def positional_encoding(self, opt, input, L):
        """model/nerf.py:476-483 (+ the BARF weighting of model/barf.py:256-268 in subclasses with
        ``progress``).  API utility in PyTorch: the render path computes the encoding inside the
        MLP kernel and never materialises it."""
        freq = 2 ** torch.arange(L, dtype=torch.float32, device=input.device) * math.pi
        spectrum = input[..., None] * freq
        enc = torch.stack([spectrum.sin(), spectrum.cos()], dim=-2).reshape(*input.shape[:-1], -1)
        if self.has_progress and opt.get("barf_c2f") is not None:
            start, end = opt.barf_c2f
            alpha = (self.progress.data - start) / (end - start) * L
            k = torch.arange(L, dtype=torch.float32, device=input.device)
            w = (1 - ((alpha - k).clamp(min=0, max=1) * math.pi).cos()) / 2
            enc = (enc.reshape(-1, L) * w).reshape(enc.shape)
        return enc


class RenderCore(nn.Module):
    """Ray pipeline shared by every Graph.  Subclasses provide ``self.nerf`` (and ``nerf_fine``)."""

    # -- depth sampling ---------------------------------------------------------------------
    def sample_depth(self, opt, batch_size, num_rays=None, depth_range=None):
        """model/nerf.py:334-344 (DTU: nerf_inn_dtu.py:524-546 takes ``depth_range``).  The uniform
        draws come from ``torch.rand`` exactly as in the reference, so a shared seed gives shared
        jitter; the arithmetic runs in one kernel with the reference's rounding sequence."""
        rng = opt.nerf.depth.range if depth_range is None else depth_range
        if not (torch.is_tensor(rng) and rng.is_cuda):           # a device tensor [min, max] is read by the kernel itself
            rng = [float(rng[0]), float(rng[1])]
        num_rays = num_rays or opt.H * opt.W
        N = opt.nerf.sample_intvs
        u = torch.rand(batch_size, num_rays, N, 1, device=opt.device) if opt.nerf.sample_stratified else None
        d = F.sample_stratified(None if u is None else u.view(-1), batch_size * num_rays, N, rng,
                                opt.nerf.depth.param, device=opt.device)
        return d.view(batch_size, num_rays, N, 1)

    def sample_depth_from_pdf(self, opt, pdf):
        """model/nerf.py:346-365: pdf [B,P,N] -> fine depths [B,P,Nf,1] (bit-exact bins)."""
        B, P, N = pdf.shape
        fine, _, _ = F.sample_pdf_merge(pdf.reshape(B * P, N), None, opt.nerf.sample_intvs_fine, opt.nerf.depth.range,
                                        want_merged=False)
        return fine.view(B, P, -1, 1)

    # -- rays -> pixels ---------------------------------------------------------------------
    def prefetch_render(self, opt, B, P, depth_range=None):
        """The part of the render pass that does not depend on the rays -- the stratified depth samples and the coarse
        network's packed BF16 weight streams -- launched on a side stream so that it overlaps the pose / warp kernels
        that produce the rays (~19 us of small kernels at C2).  Returns a token for ``_render_rays(prefetched=...)``."""
        cur = torch.cuda.current_stream()
        side = getattr(self, "_side_stream", None)
        if side is None:
            side = self._side_stream = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            depth_samples = self.sample_depth(opt, B, num_rays=P, depth_range=depth_range)
            ws, training = self.nerf.prepack(opt, B * P, depth_samples.shape[2])
        return depth_samples, ws, training, side

    def _render_rays(self, opt, center, ray, intr, mode, depth_range=None, prefetched=None):
        """Everything after ray generation in model/nerf.py:301-319."""
        B, P = ray.shape[:2]
        if opt.camera.ndc:
            center, ray = camera.convert_NDC(opt, center, ray, intr=intr)
        ws = None
        if prefetched is not None:
            depth_samples, ws, training, side = prefetched
            cur = torch.cuda.current_stream()
            cur.wait_stream(side)
            if not torch.cuda.is_current_stream_capturing():
                for t in (depth_samples, ws):       # allocated on the side stream, consumed (and freed) on this one
                    if t is not None:
                        t.record_stream(cur)
            if ws is not None and training != (torch.is_grad_enabled() and (center.requires_grad or ray.requires_grad or
                                                                             any(p.requires_grad for p in self.nerf.mlp_parameters()))):
                ws = None                            # the call will decide otherwise: let it pack for itself
        else:
            depth_samples = self.sample_depth(opt, B, num_rays=P, depth_range=depth_range)
        rgb_s, sigma_s = self.nerf.forward_samples(opt, center, ray, depth_samples, mode=mode, prepacked=ws)
        rgb, depth, opacity, prob = self.nerf.composite(opt, ray, rgb_s, sigma_s, depth_samples,
                                                        want_prob=bool(opt.nerf.fine_sampling))
        ret = edict(rgb=rgb, depth=depth, opacity=opacity)
        if opt.nerf.fine_sampling:
            N = depth_samples.shape[2]
            with torch.no_grad():
                # inverse-CDF resampling fused with the cat + sort of model/nerf.py:313-315
                _, _, merged = F.sample_pdf_merge(prob.reshape(B * P, N), depth_samples.reshape(B * P, N),
                                                  opt.nerf.sample_intvs_fine, opt.nerf.depth.range, want_fine=False)
                depth_samples = merged.view(B, P, -1, 1)
            rgb_s, sigma_s = self.nerf_fine.forward_samples(opt, center, ray, depth_samples, mode=mode)
            rgb_f, depth_f, opacity_f, _ = self.nerf_fine.composite(opt, ray, rgb_s, sigma_s, depth_samples, want_prob=False)
            ret.update(rgb_fine=rgb_f, depth_fine=depth_f, opacity_fine=opacity_f)
        return ret

    def _render_pose(self, opt, pose, intr=None, ray_idx=None, mode=None, depth_range=None, idx_start=0, num=None):
        """model/nerf.py:293-319.  Only the requested pixels are generated (no full-frame grid, no
        NaN retry loop / host sync: the kernel cannot produce the NaN the reference guards against)."""
        center, ray = camera.get_center_and_ray(opt, pose, intr=intr, ray_idx=ray_idx, idx_start=idx_start, num=num)
        return self._render_rays(opt, center, ray, intr, mode, depth_range=depth_range)

    def _render_local(self, opt, ray, center, intr=None, ray_idx=None, mode=None, depth_range=None, prefetched=None):
        """model/nerf_inn_llff.py:581-612 / nerf_inn_dtu.py:420-456: render given world-frame rays."""
        if ray_idx is not None:
            center, ray = center[:, ray_idx], ray[:, ray_idx]
        return self._render_rays(opt, center, ray, intr, mode, depth_range=depth_range, prefetched=prefetched)

    def _slices(self, opt, render_slice):
        """model/nerf.py:321-332: ``rand_rays`` pixels at a time, concatenated along the ray axis."""
        keys = ["rgb", "depth", "opacity"]
        if opt.nerf.fine_sampling:
            keys += ["rgb_fine", "depth_fine", "opacity_fine"]
        parts = {k: [] for k in keys}
        HW = opt.H * opt.W
        for c in range(0, HW, opt.nerf.rand_rays):
            ret = render_slice(c, min(opt.nerf.rand_rays, HW - c))
            for k in keys:
                parts[k].append(ret[k])
        return edict({k: torch.cat(v, dim=1) for k, v in parts.items()})

    # -- losses -----------------------------------------------------------------------------
    def L1_loss(self, pred, label=0):
        return (pred.contiguous() - label).abs().mean()

    def MSE_loss(self, pred, label=0):
        return ((pred.contiguous() - label) ** 2).mean()

    def _image_losses(self, opt, var, mode):
        """model/nerf.py:276-288: the pixel gather + squared error run in one kernel that never
        materialises ``image[:, ray_idx]``."""
        loss = edict()
        B = len(var.idx)
        ray_idx = var.ray_idx if (opt.nerf.rand_rays and mode in ["train", "test-optim"]) else None
        image = var.image.view(B, 3, opt.H, opt.W)
        # a train-mode render leaves the losses of its composite calls with the target it was given (``_loss_target``)
        target = getattr(var, "_mse_target", None)
        fused = (lambda rgb: target.take(rgb)) if target is not None and target.ray_idx is ray_idx else (lambda rgb: None)
        if opt.loss_weight.render is not None:
            loss.render = fused(var.rgb)
            if loss.render is None:
                loss.render = F.mse_gather(var.rgb, image, ray_idx)
        if opt.loss_weight.render_fine is not None:
            if not opt.nerf.fine_sampling:
                raise AssertionError("loss_weight.render_fine needs nerf.fine_sampling")
            loss.render_fine = fused(var.rgb_fine)
            if loss.render_fine is None:
                loss.render_fine = F.mse_gather(var.rgb_fine, image, ray_idx)
        return loss

    @contextlib.contextmanager
    def _loss_target(self, opt, var, mode):
        """Around the train-mode render of ``forward``: the compositor may compute the image loss of ``compute_loss``
        (model/nerf.py:276-288) in its epilogue (functional.MseTarget).  Only when gradients are on and a render loss is
        configured; ``var._mse_target`` carries the results to ``_image_losses``."""
        var._mse_target = None
        on = (F.fused_loss and mode in ["train", "test-optim"] and opt.nerf.rand_rays and torch.is_grad_enabled()
              and (opt.loss_weight.render is not None or opt.loss_weight.render_fine is not None)
              and var.get("image") is not None and var.get("ray_idx") is not None)
        if not on:
            yield
            return
        target = F.MseTarget(var.image.view(len(var.idx), 3, opt.H, opt.W), var.ray_idx)
        saved, F.mse_target = F.mse_target, target
        try:
            yield
        finally:
            F.mse_target = saved
        var._mse_target = target
