"""Operator layer: torch.autograd.Functions over the C ABI (include/niw_b200.h).

Each function mirrors one reference operation (file:line in the docstrings) on CUDA tensors.
PyTorch is used for device memory, streams and autograd bookkeeping only; all arithmetic on the
ray/sample axes happens in the library.  Non-CUDA inputs raise: there is no CPU fallback.
"""
import ctypes
import os
import math

import torch

from . import _lib
from ._lib import NIW_PREC_BF16, NIW_PREC_BF16X3, NIW_PREC_FP32, NIW_NERF_PARAMS, NIW_NVP_BLOCK_FLOATS  # noqa: F401

_c = ctypes


def _p(t):
    return None if t is None else _c.c_void_p(t.data_ptr())


def _stream():
    return _c.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32(t, name):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("niw_b200: %s must be a CUDA tensor (no CPU fallback)" % name)
    if t.dtype != torch.float32:
        raise RuntimeError("niw_b200: %s must be float32, got %s" % (name, t.dtype))
    return t.contiguous()


def _idx(ray_idx, device):
    if ray_idx is None:
        return None
    if not ray_idx.is_cuda:
        ray_idx = ray_idx.to(device)
    return ray_idx.to(torch.int64).contiguous()


PRECISIONS = {"fp32": NIW_PREC_FP32, "bf16": NIW_PREC_BF16, "bf16x3": NIW_PREC_BF16X3}


class KernelTimer:
    """CUDA-event timing of named library calls on the launching stream (bench.py's roofline leg).
    Disabled by default (zero overhead); ``with KernelTimer() as t`` arms it; ``t.totals()`` syncs
    and returns {name: (calls, total_ms)}."""
    active = None

    def __init__(self):
        self.events = []

    def __enter__(self):
        KernelTimer.active = self
        return self

    def __exit__(self, *exc):
        KernelTimer.active = None

    def totals(self):
        torch.cuda.synchronize()
        out = {}
        for name, a, b in self.events:
            n, ms = out.get(name, (0, 0.0))
            out[name] = (n + 1, ms + a.elapsed_time(b))
        return out


class _timed:
    def __init__(self, name):
        self.name, self.t = name, KernelTimer.active

    def __enter__(self):
        if self.t is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.a.record()

    def __exit__(self, *exc):
        if self.t is not None:
            b = torch.cuda.Event(enable_timing=True)
            b.record()
            self.t.events.append((self.name, self.a, b))


def precision_code(p):
    if isinstance(p, int):
        return p
    try:
        return PRECISIONS[str(p).lower()]
    except KeyError:
        raise RuntimeError("niw_b200: unknown arch.mlp_precision %r (fp32 | bf16 | bf16x3)" % (p,))


# --------------------------------------------------------------------------------------------
# ray generation
# --------------------------------------------------------------------------------------------

class _RaygenPose(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pose, intr, ray_idx, idx_start, P, H, W):
        lib = _lib.load()
        pose, intr = _f32(pose, "pose"), _f32(intr, "intr")
        B = pose.shape[0]
        center = torch.empty(B, P, 3, device=pose.device, dtype=torch.float32)
        ray = torch.empty_like(center)
        _lib.check(lib.niw_raygen_pose_fwd(_p(pose), _p(intr), _p(ray_idx), idx_start, B, P, H, W, _p(center),
                                           _p(ray), _stream()))
        ctx.save_for_backward(pose, intr, ray_idx)
        ctx.dims = (idx_start, B, P, H, W)
        return center, ray

    @staticmethod
    def backward(ctx, d_center, d_ray):
        pose, intr, ray_idx = ctx.saved_tensors
        idx_start, B, P, H, W = ctx.dims
        d_pose = torch.empty_like(pose)
        dc = None if d_center is None else d_center.contiguous()
        dr = None if d_ray is None else d_ray.contiguous()
        _lib.check(_lib.load().niw_raygen_pose_bwd(_p(pose), _p(intr), _p(ray_idx), idx_start, B, P, H, W, _p(dc),
                                                   _p(dr), _p(d_pose), _stream()))
        return d_pose, None, None, None, None, None, None


def raygen_pose(pose, intr, H, W, ray_idx=None, idx_start=0, num=None):
    """camera.get_center_and_ray followed by ``[:, ray_idx]`` (camera.py:419-443,
    model/nerf.py:298-300).  pose [B,3,4], intr [B,3,3] -> center, ray [B,P,3].  With
    ``ray_idx=None`` pixels ``idx_start .. idx_start+num-1`` (default: the whole frame)."""
    ray_idx = _idx(ray_idx, pose.device)
    P = int(ray_idx.numel()) if ray_idx is not None else int(num if num is not None else H * W - idx_start)
    return _RaygenPose.apply(pose, intr, ray_idx, int(idx_start), P, int(H), int(W))


def raygen_unwarped(intr, H, W, ray_idx=None, pose_init=None, idx_start=0, num=None, shared_center=False):
    """camera.get_unwarped_center_and_ray (camera.py:359-390) -> pts [B,2P,3] = [grid ; centre], or with
    ``shared_center`` [B,P+1,3] = [grid ; the one centre row all P would repeat]."""
    lib = _lib.load()
    intr = _f32(intr, "intr")
    pose_init = _f32(pose_init, "pose_init")
    ray_idx = _idx(ray_idx, intr.device)
    P = int(ray_idx.numel()) if ray_idx is not None else int(num if num is not None else H * W - idx_start)
    B = intr.shape[0]
    nc = 1 if shared_center else P
    pts = torch.empty(B, P + nc, 3, device=intr.device, dtype=torch.float32)
    _lib.check(lib.niw_raygen_unwarped(_p(intr), _p(pose_init), _p(ray_idx), int(idx_start), B, P, nc, int(H), int(W),
                                       _p(pts), _stream()))
    return pts


# --------------------------------------------------------------------------------------------
# NVP warp
# --------------------------------------------------------------------------------------------

def index_map_args(index_map, Pt):
    """(offset, split, jump) of the C ABI's point-index map; None = identity."""
    if index_map is None:
        return 0, int(Pt), 0
    o, sp, j = (int(v) for v in index_map)
    return o, sp, j


# ray shard of the current step, set by engine._ShardedRandperm: (index of this rank's first ray in the global per-image
# ray list, global rays per image).  The warp's annealing quirk is keyed on the position in the GLOBAL point list.
ray_shard = None


class _NvpWarp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, wpack, code_bias, pts, alpha_ratio, index_map):
        lib = _lib.load()
        wpack, code_bias, pts = _f32(wpack, "wpack"), _f32(code_bias, "code_bias"), _f32(pts, "pts")
        B, Pt = pts.shape[0], pts.shape[1]
        if wpack.numel() != 3 * NIW_NVP_BLOCK_FLOATS or code_bias.numel() != 3 * 2 * B * 128:
            raise RuntimeError("niw_b200: NVP warp supports 3 blocks x hidden 128 x 6 bands only")
        out = torch.empty_like(pts)
        im = index_map_args(index_map, Pt)
        _lib.check(lib.niw_nvp_warp_fwd(_p(wpack), _p(code_bias), _p(pts), float(alpha_ratio), B, Pt, *im, _p(out), _stream()))
        ctx.save_for_backward(wpack, code_bias, pts)
        ctx.alpha, ctx.im = float(alpha_ratio), im
        return out

    @staticmethod
    def backward(ctx, d_out):
        wpack, code_bias, pts = ctx.saved_tensors
        B, Pt = pts.shape[0], pts.shape[1]
        d_w = torch.empty_like(wpack)
        d_cb = torch.empty_like(code_bias)
        ov = backward_overlap                   # engine: the MLP weight-gradient pass is running on another stream
        _lib.check(_lib.load().niw_nvp_warp_bwd(_p(wpack), _p(code_bias), _p(pts), ctx.alpha, B, Pt, *ctx.im,
                                                _p(d_out.contiguous()), _p(d_w), _p(d_cb),
                                                int(ov.side_ctas) if ov is not None and ov.used else 0, _stream()))
        return d_w, d_cb, None, None, None


def nvp_warp(wpack, code_bias, pts, alpha_ratio, index_map=None):
    """DeformNetwork.forward on packed effective weights (model/nvp/nvp_ndr.py:365-468).
    pts [B,Pt,3] (no gradient: the reference detaches them, barf_inn_llff.py:328-330)."""
    return _NvpWarp.apply(wpack, code_bias, pts, alpha_ratio, index_map)


# --------------------------------------------------------------------------------------------
# depth sampling
# --------------------------------------------------------------------------------------------

class DeviceUniform:
    """Stand-in for ``torch.rand(B, P, N, 1, device=cuda)`` inside ``engine.device_ray_draws``: no tensor is drawn; the
    stratified-sampling kernel that consumes it draws its uniforms itself (``niw_sample_stratified_rng``, Philox4x32-10)
    from ``seed`` and the caller-owned device counter ``rng`` (int64 [2], advanced by the kernel).  Anything but
    ``view`` / ``numel`` / ``shape`` is unsupported on purpose: another consumer of the draw would need real numbers."""

    def __init__(self, shape, rng, seed):
        self.shape, self.rng, self.seed = tuple(int(x) for x in shape), rng, int(seed) & (2 ** 64 - 1)
        self.device = rng.device

    def numel(self):
        n = 1
        for x in self.shape:
            n *= x
        return n

    def view(self, *shape):
        return self


def sample_stratified(u, n_rays, N, depth_range, param, device=None):
    """Graph.sample_depth (model/nerf.py:334-344).  u: [n_rays*N] uniforms, None (0.5) or a ``DeviceUniform`` (drawn inside
    the kernel).  Returns depth [n_rays, N].  Bit-exact with the reference's fp32 CPU evaluation of the same uniforms."""
    lib = _lib.load()
    if isinstance(u, DeviceUniform):
        if u.numel() != n_rays * N:
            raise RuntimeError("niw_b200: uniforms have %d elements, expected %d" % (u.numel(), n_rays * N))
        if param not in ("metric", "inverse"):
            raise KeyError(param)
        depth = torch.empty(n_rays, N, device=u.device, dtype=torch.float32)
        if torch.is_tensor(depth_range) and depth_range.is_cuda:
            rng_dev, scale, dmin = _f32(depth_range.reshape(-1)[:2], "depth_range"), 0.0, 0.0
        else:
            rng_dev, dmin = None, float(depth_range[0])
            scale = float(depth_range[1]) - dmin
        _lib.check(lib.niw_sample_stratified_rng(n_rays, N, scale, dmin, _p(rng_dev), int(param == "inverse"), u.seed, _p(u.rng),
                                                 _p(depth), _stream()))
        return depth
    u = _f32(u, "u")
    device = u.device if u is not None else device
    if u is not None and u.numel() != n_rays * N:
        raise RuntimeError("niw_b200: uniforms have %d elements, expected %d" % (u.numel(), n_rays * N))
    if param not in ("metric", "inverse"):
        raise KeyError(param)
    depth = torch.empty(n_rays, N, device=device, dtype=torch.float32)
    if torch.is_tensor(depth_range) and depth_range.is_cuda:
        # [min, max] stays on the device: no host read, the step remains capturable
        rng = _f32(depth_range.reshape(-1)[:2], "depth_range")
        _lib.check(lib.niw_sample_stratified_dev(_p(u), n_rays, N, _p(rng), int(param == "inverse"), _p(depth), _stream()))
        return depth
    dmin, dmax = float(depth_range[0]), float(depth_range[1])
    _lib.check(lib.niw_sample_stratified(_p(u), n_rays, N, dmax - dmin, dmin, int(param == "inverse"), _p(depth),
                                         _stream()))
    return depth


_TABLES = {}


def _pdf_tables(N, Nf, depth_range, device):
    """unif = 0.5*(grid[:-1]+grid[1:]) and bins = linspace(dmin,dmax,N+1) evaluated by torch on the
    CPU in fp32, exactly as the reference does (model/nerf.py:352-356), cached per configuration."""
    key = (N, Nf, float(depth_range[0]), float(depth_range[1]), str(device))
    if key not in _TABLES:
        grid = torch.linspace(0, 1, Nf + 1)
        unif = 0.5 * (grid[:-1] + grid[1:])
        bins = torch.linspace(float(depth_range[0]), float(depth_range[1]), N + 1)
        _TABLES[key] = (unif.to(device), bins.to(device))
    return _TABLES[key]


def sample_pdf_merge(pdf, depth_coarse, Nf, depth_range, want_idx=False, want_fine=True, want_merged=True):
    """Graph.sample_depth_from_pdf (model/nerf.py:346-365) fused with the cat+sort of
    model/nerf.py:313-315.  pdf, depth_coarse [R,N] -> (fine [R,Nf], idx [R,Nf] int64, merged [R,N+Nf])."""
    lib = _lib.load()
    pdf = _f32(pdf, "pdf")
    R, N = pdf.shape
    depth_coarse = _f32(depth_coarse, "depth_coarse")
    unif, bins = _pdf_tables(N, Nf, depth_range, pdf.device)
    fine = torch.empty(R, Nf, device=pdf.device, dtype=torch.float32) if want_fine else None
    idx = torch.empty(R, Nf, device=pdf.device, dtype=torch.int64) if want_idx else None
    merged = torch.empty(R, N + Nf, device=pdf.device, dtype=torch.float32) if want_merged else None
    _lib.check(lib.niw_sample_pdf_merge(_p(pdf), _p(depth_coarse), _p(unif), _p(bins), R, N, Nf, _p(fine), _p(idx),
                                        _p(merged), _stream()))
    return fine, idx, merged


# --------------------------------------------------------------------------------------------
# compositing
# --------------------------------------------------------------------------------------------

class _Composite(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ray, rgb_s, sigma, depth_s, bg, want_prob):
        lib = _lib.load()
        ray, rgb_s, sigma, depth_s = _f32(ray, "ray"), _f32(rgb_s, "rgb_samples"), _f32(sigma, "density_samples"), \
            _f32(depth_s, "depth_samples")
        R, N = sigma.shape
        dev = ray.device
        rgb = torch.empty(R, 3, device=dev)
        depth = torch.empty(R, device=dev)
        opacity = torch.empty(R, device=dev)
        ctx.set_materialize_grads(False)     # unused outputs (depth / opacity under an rgb-only loss) arrive as None, not zeros
        # one of the two per-sample outputs is normally enough (24 B/sample): the weights when the caller samples
        # from them (fine pass), the transmittance when a backward pass will follow
        need_grad = any(ctx.needs_input_grad[:3])
        prob = torch.empty(R, N, device=dev) if want_prob else None
        trans = torch.empty(R, N, device=dev) if need_grad else None
        with _timed("composite_fwd"):
            _lib.check(lib.niw_composite_fwd(_p(ray), _p(rgb_s), _p(sigma), _p(depth_s), R, N, float(bg), _p(rgb),
                                             _p(depth), _p(opacity), _p(prob), _p(trans), _stream()))
        if need_grad:
            ctx.save_for_backward(ray, rgb_s, sigma, depth_s, trans, *([prob] if want_prob else []))
        ctx.bg = float(bg)
        if prob is None:
            prob = torch.empty(0, device=dev)
        ctx.mark_non_differentiable(prob)
        return rgb, depth, opacity, prob

    @staticmethod
    def backward(ctx, d_rgb, d_depth, d_opacity, _d_prob):
        ray, rgb_s, sigma, depth_s, trans, *rest = ctx.saved_tensors
        prob = rest[0] if rest else None
        if d_rgb is None and d_depth is None and d_opacity is None:
            return None, None, None, None, None, None
        R, N = sigma.shape
        d_rgb_s = torch.empty_like(rgb_s)
        d_sigma = torch.empty_like(sigma)
        d_ray = torch.empty_like(ray)
        c = lambda t: None if t is None else t.contiguous()
        d_rgb, d_depth, d_opacity = c(d_rgb), c(d_depth), c(d_opacity)
        with _timed("composite_bwd"):
            _lib.check(_lib.load().niw_composite_bwd(_p(ray), _p(rgb_s), _p(sigma), _p(depth_s), _p(prob), _p(trans), R,
                                                     N, ctx.bg, _p(d_rgb), _p(d_depth), _p(d_opacity), _p(d_rgb_s),
                                                     _p(d_sigma), _p(d_ray), _stream()))
        return d_ray, d_rgb_s, d_sigma, None, None, None


def composite(ray, rgb_samples, density_samples, depth_samples, bgcolor=None, want_prob=True):
    """NeRF.composite (model/nerf.py:458-474) on flattened rays: ray [R,3], rgb_samples [R,N,3],
    density_samples [R,N], depth_samples [R,N] -> rgb [R,3], depth [R], opacity [R], prob [R,N]
    (``want_prob=False``: an empty tensor -- the render pipeline asks for the weights only when it resamples from them).
    No gradient flows to depth_samples (they come from torch.rand / no_grad in the reference)."""
    bg = -1.0 if bgcolor is None else float(bgcolor)
    return _Composite.apply(ray, rgb_samples, density_samples, depth_samples, bg, bool(want_prob))


# --------------------------------------------------------------------------------------------
# compositor with the loss head in its epilogue
# --------------------------------------------------------------------------------------------

class MseTarget:
    """What the rendered colours of the coming ``composite`` calls will be compared with (model/nerf.py:276-288):
    ``image`` [B,3,H,W] and the pixel subset ``ray_idx`` (or None: pixels idx_start ..).  A Graph's ``forward`` sets it
    (``functional.mse_target``) for a train-mode render; ``composite_mse`` then leaves each call's loss here, and the
    Graph's ``compute_loss`` picks it up by the identity of the colour tensor it is about to compare."""

    def __init__(self, image, ray_idx=None, idx_start=0):
        self.image, self.ray_idx, self.idx_start = _f32(image, "image"), ray_idx, int(idx_start)
        self.losses = []

    def put(self, rgb, loss):
        self.losses.append((rgb, loss))

    def take(self, rgb):
        for t, loss in self.losses:
            if t is rgb:
                return loss
        return None


mse_target = None
# NIW_FUSED_LOSS=0: the loss head stays a kernel of its own (niw_mse_gather) -- the parity partner of the fused form
fused_loss = os.environ.get("NIW_FUSED_LOSS", "1") != "0"
_mse_scratch = {}


def _mse_scratch_for(dev):
    """Ticket + per-block partial sums of niw_composite_fwd_mse: zeroed once, then kept (the ticket resets itself)."""
    key = (dev.type, dev.index)
    buf = _mse_scratch.get(key)
    if buf is None:
        buf = torch.zeros(_lib.load().niw_composite_mse_scratch_floats(), device=dev)
        if not torch.cuda.is_current_stream_capturing():     # (memory of a graph's private pool must not outlive the graph)
            _mse_scratch[key] = buf
    return buf


def composite_mse_supported(N):
    return fused_loss and int(N) in (64, 128, 192, 256)


class _CompositeMse(torch.autograd.Function):
    """_Composite + the MSE against the gathered ground-truth pixels in the same two launches: the forward kernel also
    writes the loss and d_unit = d loss / d rgb, the backward kernel adds d_loss * d_unit to the incoming d_rgb."""

    @staticmethod
    def forward(ctx, ray, rgb_s, sigma, depth_s, bg, want_prob, image, ray_idx, idx_start, B, P):
        lib = _lib.load()
        ray, rgb_s, sigma, depth_s = _f32(ray, "ray"), _f32(rgb_s, "rgb_samples"), _f32(sigma, "density_samples"), \
            _f32(depth_s, "depth_samples")
        R, N = sigma.shape
        dev = ray.device
        H, W = image.shape[-2:]
        rgb = torch.empty(R, 3, device=dev)
        depth = torch.empty(R, device=dev)
        opacity = torch.empty(R, device=dev)
        loss = torch.empty((), device=dev)
        d_unit = torch.empty(R, 3, device=dev)
        ctx.set_materialize_grads(False)
        need_grad = any(ctx.needs_input_grad[:3])
        prob = torch.empty(R, N, device=dev) if want_prob else None
        trans = torch.empty(R, N, device=dev) if need_grad else None
        with _timed("composite_fwd"):
            _lib.check(lib.niw_composite_fwd_mse(_p(ray), _p(rgb_s), _p(sigma), _p(depth_s), R, N, float(bg), _p(rgb),
                                                 _p(depth), _p(opacity), _p(prob), _p(trans), _p(image), _p(ray_idx),
                                                 int(idx_start), int(B), int(P), int(H), int(W), _p(d_unit),
                                                 _p(_mse_scratch_for(dev)), _p(loss), _stream()))
        if need_grad:
            ctx.save_for_backward(ray, rgb_s, sigma, depth_s, trans, d_unit, *([prob] if want_prob else []))
        ctx.bg = float(bg)
        if prob is None:
            prob = torch.empty(0, device=dev)
        ctx.mark_non_differentiable(prob)
        return rgb, depth, opacity, prob, loss

    @staticmethod
    def backward(ctx, d_rgb, d_depth, d_opacity, _d_prob, d_loss):
        ray, rgb_s, sigma, depth_s, trans, d_unit, *rest = ctx.saved_tensors
        prob = rest[0] if rest else None
        none = (None,) * 11
        if d_rgb is None and d_depth is None and d_opacity is None and d_loss is None:
            return none
        R, N = sigma.shape
        d_rgb_s = torch.empty_like(rgb_s)
        d_sigma = torch.empty_like(sigma)
        d_ray = torch.empty_like(ray)
        c = lambda t: None if t is None else t.contiguous()
        d_rgb, d_depth, d_opacity = c(d_rgb), c(d_depth), c(d_opacity)
        lib = _lib.load()
        with _timed("composite_bwd"):
            if d_loss is None:
                _lib.check(lib.niw_composite_bwd(_p(ray), _p(rgb_s), _p(sigma), _p(depth_s), _p(prob), _p(trans), R, N, ctx.bg,
                                                 _p(d_rgb), _p(d_depth), _p(d_opacity), _p(d_rgb_s), _p(d_sigma), _p(d_ray),
                                                 _stream()))
            else:
                d_loss = _f32(d_loss, "d_loss")
                _lib.check(lib.niw_composite_bwd_mse(_p(ray), _p(rgb_s), _p(sigma), _p(depth_s), _p(prob), _p(trans), R, N,
                                                     ctx.bg, _p(d_rgb), _p(d_depth), _p(d_opacity), _p(d_unit), _p(d_loss),
                                                     _p(d_rgb_s), _p(d_sigma), _p(d_ray), _stream()))
        return (d_ray, d_rgb_s, d_sigma) + (None,) * 8


def composite_mse(ray, rgb_samples, density_samples, depth_samples, target, B, P, bgcolor=None, want_prob=True):
    """``composite`` plus mean((rgb - image[:, :, ray_idx])**2) of ``target`` (an ``MseTarget``) from the same kernel:
    -> rgb [R,3], depth [R], opacity [R], prob [R,N], loss [] (R = B P rays, image by image)."""
    bg = -1.0 if bgcolor is None else float(bgcolor)
    return _CompositeMse.apply(ray, rgb_samples, density_samples, depth_samples, bg, bool(want_prob), target.image,
                               _idx(target.ray_idx, ray.device), target.idx_start, int(B), int(P))


# --------------------------------------------------------------------------------------------
# positional encoding + MLP
# --------------------------------------------------------------------------------------------

def band_weights(progress, c2f, L):
    """BARF coarse-to-fine weights (model/barf.py:260-264) as a host list of L floats (fp32 math
    mirrors the reference: alpha, clamp, *pi, cos in fp32)."""
    if c2f is None:
        return [1.0] * L
    start, end = c2f
    alpha = (torch.tensor(float(progress), dtype=torch.float32) - start) / (end - start) * L
    k = torch.arange(L, dtype=torch.float32)
    w = (1 - ((alpha - k).clamp(min=0, max=1) * math.pi).cos()) / 2
    return [float(x) for x in w]


class BackwardOverlap:
    """Step-level scheduling of the training backward (engine.train_step): the MLP weight-gradient pass (niw_nerf_bwd_dw:
    streams the tile records, HBM-bound, nothing downstream of it but the optimiser) is enqueued on a side stream with
    ``dw_ctas`` CTAs, while the pose / warp backward that consumes d_center / d_ray continues on the launching stream on
    the remaining SMs (``side_ctas`` caps the warp backward's grid to them).  Only for a module the engine has opted in
    (``module.overlap_weight_gradients``: the LAST NeRF network of the backward pass; an earlier one would push the next
    network's persistent dX kernel onto the few free SMs).  ``join()`` makes the current stream wait for the side work."""

    def __init__(self, dw_ctas=None, side_ctas=None):
        import os
        self.stream = torch.cuda.Stream()
        n = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
        self.side_ctas = int(os.environ.get("NIW_OVERLAP_SIDE_CTAS", side_ctas if side_ctas is not None else 20))
        self.dw_ctas = int(os.environ.get("NIW_OVERLAP_DW_CTAS", dw_ctas if dw_ctas is not None else n - self.side_ctas))
        self.used = False

    def __enter__(self):
        global backward_overlap
        self.used = False
        backward_overlap = self
        return self

    def __exit__(self, *exc):
        global backward_overlap
        backward_overlap = None
        self.join()

    def join(self):
        if self.used:
            torch.cuda.current_stream().wait_stream(self.stream)
            self.used = False


backward_overlap = None


class _NerfSamples(torch.autograd.Function):
    """``flat`` is the 530 052-float parameter vector the kernels read.  Two ways to receive its
    gradient: (a) ``flat`` itself requires grad (functional use, tests) -> returned through autograd;
    (b) ``module`` is a NeRFCore whose ``nn.Linear`` parameters are passed as ``*params`` so autograd knows
    the output depends on them.  Their gradients are returned through autograd as well (so
    ``torch.autograd.grad``, ``backward(inputs=...)`` and gradient hooks behave) UNLESS the engine has opted the
    module in to in-kernel accumulation (``module.accumulate_grads_in_place``, set by ``engine.train_step`` /
    ``engine.use_flat_gradients`` when the parameters' ``.grad`` are slices of one flat bucket): then the kernels
    accumulate straight into that bucket and None is returned (no per-parameter accumulate launches).  Every
    training-mode forward call of such a module is counted in ``module._pending_backward``; the module is reported
    to ``_grads_ready`` (engine.overlap_allreduce) only when the last of those calls has been back-propagated."""

    @staticmethod
    def forward(ctx, flat, center, ray, depth, progress, c2f, precision, training, module, prepacked, *params):
        lib = _lib.load()
        flat, center, ray, depth = _f32(flat, "params"), _f32(center, "center"), _f32(ray, "ray"), _f32(depth, "depth")
        if flat.numel() != NIW_NERF_PARAMS:
            raise RuntimeError("niw_b200: the MLP kernels implement the 8x256/skip-4/128-RGB architecture "
                               "(%d parameters); got %d" % (NIW_NERF_PARAMS, flat.numel()))
        R, N = depth.shape
        nbytes = lib.niw_nerf_workspace_bytes(R, N, precision, int(training))
        if nbytes == 0:
            raise RuntimeError("niw_b200: MLP precision %d is not available in this build" % precision)
        flags = int(training)
        if prepacked is not None:
            # weight streams already packed into this workspace (nerf_prepack, possibly on another stream)
            if prepacked.numel() != nbytes or precision not in (NIW_PREC_BF16, NIW_PREC_BF16X3):
                raise RuntimeError("niw_b200: prepacked workspace does not match this call")
            ws, flags = prepacked, flags | _lib.NIW_NERF_PREPACKED
        else:
            ws = torch.empty(nbytes, dtype=torch.uint8, device=depth.device)
        rgb = torch.empty(R, N, 3, device=depth.device)
        sigma = torch.empty(R, N, device=depth.device)
        c0, c1 = (float(c2f[0]), float(c2f[1])) if progress is not None else (0.0, 1.0)
        with _timed("nerf_fwd"):
            _lib.check(lib.niw_nerf_fwd(_p(flat), _p(center), _p(ray), _p(depth), R, N, _p(progress), c0, c1, precision,
                                        flags, _p(ws), nbytes, _p(rgb), _p(sigma), _stream()))
        if training:
            ctx.save_for_backward(flat, center, ray, depth, ws)
            ctx.cfg = (precision, nbytes, module, len(params))
            # weight gradients wanted at all?  (test-time pose refinement differentiates through a frozen network)
            ctx.need_dw = bool(flat.requires_grad or any(p.requires_grad for p in params))
            if module is not None and params:
                module._pending_backward = getattr(module, "_pending_backward", 0) + 1
        return rgb, sigma

    @staticmethod
    def backward(ctx, d_rgb, d_sigma):
        flat, center, ray, depth, ws = ctx.saved_tensors
        precision, nbytes, module, n_params = ctx.cfg
        R, N = depth.shape
        in_place = module is not None and n_params and getattr(module, "accumulate_grads_in_place", False)
        target = module.flat_grad_pointer() if in_place else None
        if target is None:
            d_params = torch.zeros_like(flat)
            dp = _p(d_params)
        else:
            d_params = None
            dp = _c.c_void_p(target)
        both = torch.empty((2,) + tuple(center.shape), device=center.device, dtype=center.dtype)     # adjacent: one memset
        d_center, d_ray = both[0], both[1]
        if d_rgb is None:
            d_rgb = torch.zeros(R, N, 3, device=depth.device)
        if d_sigma is None:
            d_sigma = torch.zeros(R, N, device=depth.device)
        d_rgb, d_sigma = d_rgb.contiguous(), d_sigma.contiguous()
        lib = _lib.load()
        ov = backward_overlap
        split = (ov is not None and target is not None and precision in (NIW_PREC_BF16, NIW_PREC_BF16X3)
                 and getattr(module, "overlap_weight_gradients", False) and KernelTimer.active is None)
        if not ctx.need_dw and precision in (NIW_PREC_BF16, NIW_PREC_BF16X3):
            # frozen network: only the activation-gradient chain (d_center / d_ray); the weight-gradient GEMMs are skipped
            scratch = torch.zeros(NIW_NERF_PARAMS, device=depth.device)     # the chain adds two head-bias sums here
            with _timed("nerf_bwd"):
                _lib.check(lib.niw_nerf_bwd_dx(_p(flat), _p(center), _p(ray), _p(depth), R, N, precision, _p(ws), nbytes,
                                               _p(d_rgb), _p(d_sigma), _p(scratch), _p(d_center), _p(d_ray), _stream()))
            return (None, d_center, d_ray, None, None, None, None, None, None, None) + (None,) * n_params
        if split:
            # dX on this stream; dW on the side stream, ordered after it, while this stream goes on to the pose / warp backward
            _lib.check(lib.niw_nerf_bwd_dx(_p(flat), _p(center), _p(ray), _p(depth), R, N, precision, _p(ws), nbytes,
                                           _p(d_rgb), _p(d_sigma), dp, _p(d_center), _p(d_ray), _stream()))
            cur = torch.cuda.current_stream()
            ov.stream.wait_stream(cur)
            ov.used = True
            with torch.cuda.stream(ov.stream):
                _lib.check(lib.niw_nerf_bwd_dw(R, N, precision, _p(ws), nbytes, dp, ov.dw_ctas, _stream()))
                if not torch.cuda.is_current_stream_capturing():
                    ws.record_stream(ov.stream)
        else:
            with _timed("nerf_bwd"):
                _lib.check(lib.niw_nerf_bwd(_p(flat), _p(center), _p(ray), _p(depth), R, N, precision,
                                            _p(ws), nbytes, _p(d_rgb), _p(d_sigma), dp, _p(d_center),
                                            _p(d_ray), _stream()))
        if module is not None and n_params:
            module._pending_backward = max(getattr(module, "_pending_backward", 1) - 1, 0)
            ready = getattr(module, "_grads_ready", None)       # engine.overlap_allreduce: this module's gradients are final
            if target is not None and ready is not None and module._pending_backward == 0:
                if split:
                    with torch.cuda.stream(ov.stream):          # "final" once the side stream's dW has run: order the reduce after it
                        ready(module)
                else:
                    ready(module)
        if n_params and d_params is not None:
            # slow path: the module's gradients are not one flat buffer -> hand slices back to autograd
            pg = tuple(g.view_as(p) for g, p in zip(torch.split(d_params, [p.numel() for p in module.mlp_parameters()]),
                                                     module.mlp_parameters()))
            return (None, d_center, d_ray, None, None, None, None, None, None, None) + pg
        return (d_params if not n_params else None, d_center, d_ray, None, None, None, None, None, None, None) + (None,) * n_params


def _progress_tensor(progress, c2f, device):
    if c2f is None:
        return None
    if progress is None:
        raise RuntimeError("niw_b200: barf_c2f needs the progress scalar")
    if not torch.is_tensor(progress):
        return torch.tensor(float(progress), dtype=torch.float32, device=device)
    return _f32(progress.detach(), "progress")


def nerf_forward_samples(params, center, ray, depth, progress=None, c2f=None, precision=NIW_PREC_FP32, training=None,
                         module=None, prepacked=None):
    """NeRF.forward_samples (model/nerf.py:449-456 -> :416-447, with camera.py:517-521 and the BARF
    encoding model/barf.py:256-268).  params: flat [530052] fp32; center/ray [R,3]; depth [R,N]
    -> rgb [R,N,3], sigma [R,N].  ``progress`` (device scalar tensor, or a float) and ``c2f`` = (start, end)
    select the coarse-to-fine band weights, evaluated on the device; None = no annealing.  With
    ``module`` (a NeRFCore) the parameter gradients are accumulated into the module's flat gradient
    buffer by the kernels (see _NerfSamples).  ``prepacked``: a workspace from ``nerf_prepack`` for this very call."""
    progress = _progress_tensor(progress, c2f, depth.device)
    mparams = tuple(module.mlp_parameters()) if module is not None else ()
    if training is None:
        training = torch.is_grad_enabled() and (params.requires_grad or center.requires_grad or ray.requires_grad
                                                or any(p.requires_grad for p in mparams))
    return _NerfSamples.apply(params, center, ray, depth, progress, c2f, precision_code(precision), bool(training),
                              module, prepacked, *mparams)


def nerf_prepack(params, R, N, progress=None, c2f=None, precision=NIW_PREC_BF16, training=True):
    """The ray-independent part of ``nerf_forward_samples`` (BF16 path): allocate the call's workspace and pack the
    weight streams into it on the CURRENT stream -- which may be a side stream running while the rays are still being
    produced.  Returns the workspace to pass as ``prepacked`` (``training`` must be what that call will use), or None
    when there is nothing to hoist (FP32 path)."""
    precision = precision_code(precision)
    if precision not in (NIW_PREC_BF16, NIW_PREC_BF16X3):
        return None
    lib = _lib.load()
    params = _f32(params, "params")
    progress = _progress_tensor(progress, c2f, params.device)
    nbytes = lib.niw_nerf_workspace_bytes(R, N, precision, int(bool(training)))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=params.device)
    c0, c1 = (float(c2f[0]), float(c2f[1])) if progress is not None else (0.0, 1.0)
    _lib.check(lib.niw_nerf_pack(_p(params), _p(progress), c0, c1, precision, int(bool(training)), R, N, _p(ws), nbytes,
                                 _stream()))
    return ws


# --------------------------------------------------------------------------------------------
# loss head
# --------------------------------------------------------------------------------------------

class _MseGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rgb, image, ray_idx, idx_start):
        lib = _lib.load()
        rgb, image = _f32(rgb, "rgb"), _f32(image, "image")
        B, P = rgb.shape[0], rgb.shape[1]
        H, W = image.shape[-2:]
        loss = (torch.zeros if lib.niw_mse_gather_needs_zero(B, P) else torch.empty)((), device=rgb.device)
        d_rgb = torch.empty_like(rgb)
        scale = 1.0 / (B * P * 3)
        _lib.check(lib.niw_mse_gather(_p(image), _p(rgb), _p(ray_idx), idx_start, B, P, H, W, scale, _p(loss), _p(d_rgb),
                                      _stream()))
        ctx.save_for_backward(d_rgb)
        return loss

    @staticmethod
    def backward(ctx, g):
        (d_rgb,) = ctx.saved_tensors
        return d_rgb * g, None, None, None


def mse_gather(rgb, image, ray_idx=None, idx_start=0):
    """MSE between rendered rgb [B,P,3] and image[:, :, ray_idx] (model/nerf.py:276-283,
    model/base.py:209-211), without materialising the gathered pixels."""
    return _MseGather.apply(rgb, image, _idx(ray_idx, rgb.device), int(idx_start))


def image_metrics(rgb, image, H, W):
    """PSNR and SSIM of rendered views (reference model/nerf.py:176-183, pytorch_ssim.ssim): rgb [B,H*W,3] as the
    renderer returns it, image [B,3,H,W] -> (psnr [B], ssim [B]) device tensors (no host read)."""
    lib = _lib.load()
    rgb, image = _f32(rgb, "rgb"), _f32(image, "image")
    B = image.shape[0]
    if rgb.numel() != B * H * W * 3 or image.numel() != B * 3 * H * W:
        raise RuntimeError("niw_b200: image_metrics expects rgb [B,%d,3] and image [B,3,%d,%d]" % (H * W, H, W))
    out = torch.empty(B, 2, device=image.device, dtype=torch.float32)
    _lib.check(lib.niw_image_metrics(_p(rgb), _p(image), B, int(H), int(W), _p(out), _stream()))
    n = 3.0 * H * W
    return -10.0 * torch.log10(out[:, 0] / n), out[:, 1] / n


def depth_metrics(pred, gt, valid=None, scale=1.0):
    """core/metrics.py:64-111: masked mean absolute depth error and RMSE of a rendered depth map; with ``scale != 1``
    the better of the scaled / unscaled prediction, as the reference.  Returns (abs_err, rmse) device scalars."""
    lib = _lib.load()
    pred, gt = _f32(pred, "pred").reshape(-1), _f32(gt, "gt").reshape(-1)
    if valid is not None:
        valid = valid.reshape(-1).to(torch.uint8).contiguous()
    out = torch.empty(5, device=pred.device, dtype=torch.float32)
    _lib.check(lib.niw_depth_metrics(_p(pred), _p(gt), _p(valid), pred.numel(), float(scale), _p(out), _stream()))
    n = out[0]
    abs_e, rmse = out[1] / (n + 1e-6), torch.sqrt(out[2] / n)
    if scale != 1.0:
        abs_e, rmse = torch.minimum(abs_e, out[3] / (n + 1e-6)), torch.minimum(rmse, torch.sqrt(out[4] / n))
    return abs_e, rmse


def kabsch(x, y):
    """roma.rigid_points_registration(x, y) as the reference uses it (model/nerf_inn_llff.py:569,
    model/pose_models/inn.py:100): the least-squares R [B,3,3], t [B,3] with y ~ R x + t, for x, y [B,M,3].  No gradient
    (the reference detaches the result), no host synchronisation."""
    lib = _lib.load()
    x, y = _f32(x.detach(), "x"), _f32(y.detach(), "y")
    B, M = x.shape[0], x.shape[1]
    if x.shape != y.shape or x.shape[-1] != 3:
        raise RuntimeError("niw_b200: kabsch expects two [B,M,3] point sets")
    R = torch.empty(B, 3, 3, device=x.device, dtype=torch.float32)
    t = torch.empty(B, 3, device=x.device, dtype=torch.float32)
    _lib.check(lib.niw_kabsch(_p(x), _p(y), B, M, _p(R), _p(t), _stream()))
    return R, t


# process group over which per-image point lists are sharded (data-parallel ray shards); set by the engine for the
# duration of a sharded forward + loss.  While set, ``kabsch_sharded`` sums the fit's sufficient statistics over it.
data_parallel_group = None


def kabsch_stats(x, y):
    """Per-image sufficient statistics of the rigid fit over THIS rank's rows: [B,16] float64 =
    (rows, sum x, sum y, sum x_a y_c).  Summed over ranks they determine the fit of the whole list."""
    lib = _lib.load()
    x, y = _f32(x.detach(), "x"), _f32(y.detach(), "y")
    B, M = x.shape[0], x.shape[1]
    if x.shape != y.shape or x.shape[-1] != 3:
        raise RuntimeError("niw_b200: kabsch expects two [B,M,3] point sets")
    stats = torch.empty(B, 16, device=x.device, dtype=torch.float64)
    _lib.check(lib.niw_kabsch_stats(_p(x), _p(y), B, M, _p(stats), _stream()))
    return stats


def kabsch_solve(stats):
    """R [B,3,3], t [B,3] from (rank-summed) ``kabsch_stats``."""
    lib = _lib.load()
    if not stats.is_cuda or stats.dtype != torch.float64 or stats.shape[-1] != 16:
        raise RuntimeError("niw_b200: kabsch_solve expects a CUDA float64 [B,16] statistics tensor")
    stats = stats.contiguous()
    B = stats.shape[0]
    R = torch.empty(B, 3, 3, device=stats.device, dtype=torch.float32)
    t = torch.empty(B, 3, device=stats.device, dtype=torch.float32)
    _lib.check(lib.niw_kabsch_solve(_p(stats), B, _p(R), _p(t), _stream()))
    return R, t


def kabsch_sharded(x, y, group=None):
    """``kabsch`` of point lists whose rows are spread over the ranks of ``group``: statistics, ONE all-reduce of 16
    doubles per image, solve -- every rank obtains the fit of the whole list, as the single-GPU step computes it."""
    import torch.distributed as dist
    stats = kabsch_stats(x, y)
    dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return kabsch_solve(stats)


def sample_pixels(n, k, counter, seed=0):
    """First k entries of a random permutation of range(n) (the reference's ``torch.randperm(n)[:k]``,
    model/nerf.py:268) in O(k): int64 [k] on ``counter``'s device.  ``counter`` is a zero-initialised int64 [1]
    device tensor owned by the caller; the kernel advances it (fresh draw per call / per graph replay)."""
    if not counter.is_cuda:
        raise RuntimeError("niw_b200: counter must be a CUDA tensor (no CPU fallback)")
    if counter.dtype != torch.int64 or counter.numel() != 1:
        raise ValueError("sample_pixels: counter must be an int64 tensor with one element")
    out = torch.empty(int(k), dtype=torch.int64, device=counter.device)
    _lib.check(_lib.load().niw_sample_pixels(int(n), int(k), int(seed) & (2 ** 64 - 1), _p(counter), _p(out), _stream()))
    return out


class _RaysFromWarp(torch.autograd.Function):
    """warped [B,P+nc,3] = [grid rows ; centre rows] -> (ray = grid - centre, centre), both contiguous [B,P,3]
    (model/barf_inn_llff.py:352-356); nc = P (the reference's list) or 1 (the centre evaluated once per image: backward
    sums the P per-ray centre gradients into that row).  One launch each way (the eager slice / sub / expand / sum / cat
    chain is ~8 tiny ones)."""

    @staticmethod
    def forward(ctx, warped, P, nc):
        warped = _f32(warped, "warped")
        B = warped.shape[0]
        if warped.shape[1] != P + nc:
            raise RuntimeError("niw_b200: rays_from_warp expects [B,%d,3], got %s" % (P + nc, tuple(warped.shape)))
        ctx.dims = (B, P, nc)
        ctx.set_materialize_grads(False)
        ray = torch.empty(B, P, 3, device=warped.device, dtype=torch.float32)
        center = torch.empty_like(ray)
        _lib.check(_lib.load().niw_rays_from_warp_fwd(_p(warped), B, P, nc, _p(ray), _p(center), _stream()))
        return ray, center

    @staticmethod
    def backward(ctx, d_ray, d_center):
        if d_ray is None and d_center is None:
            return None, None, None
        B, P, nc = ctx.dims
        like = d_ray if d_ray is not None else d_center
        d_warped = torch.empty(B, P + nc, 3, device=like.device, dtype=torch.float32)
        c = lambda t: None if t is None else t.contiguous()
        _lib.check(_lib.load().niw_rays_from_warp_bwd(_p(c(d_ray)), _p(c(d_center)), B, P, nc, _p(d_warped), _stream()))
        return d_warped, None, None


def rays_from_warp(warped, P):
    return _RaysFromWarp.apply(warped, int(P), int(P))


def rays_from_warp_shared(warped, P):
    return _RaysFromWarp.apply(warped, int(P), 1)


# rows of the per-image point list that the embedder's annealing quirk touches (embedder.py:46-49: [d, d (2 NF + 1)),
# d = 2 or 1, NF = 6): a centre row may be shared only if every centre row of the full list lies beyond them
NVP_QUIRK_ROWS = 2 * (2 * 6 + 1)


# train-mode warped ray generation as ONE launch (nvp.DeformNetwork.warped_rays, csrc/nvp.cu niw_nvp_rays_fwd) instead of
# raygen_unwarped -> warp -> rays_from_warp.  Off by default: inside the captured C2 step it measured 3.6 us SLOWER
# (0.6725 vs 0.6688 ms, three alternating runs of 100 steps) -- the un-warped grid kernel already runs
# beside the weight pack of the side stream, and the centre row costs every CTA a ninth warp.  NIW_FUSED_RAYS=1 turns it on
# (two launches fewer per step where the host's launch rate is the limit: eager calls from the reference's engine).
fused_warped_rays = os.environ.get("NIW_FUSED_RAYS", "0") == "1"


def shared_center_ok(P, shard=None):
    """May the P identical centre rows of an image be evaluated once?  Only if none of the centre rows of the
    (global) per-image list is one the annealing quirk touches."""
    P_global = shard[1] if shard is not None else P
    return P_global >= NVP_QUIRK_ROWS


def warp_point_list(pts, P, shard=None):
    """The point list handed to the warp and its index map, for pts = [grid rows (P) ; centre rows (P or 1)].
    Returns (pts', index_map, shared).  ``shared``: the centre is evaluated once per image ([grid ; centre]), exact
    because all centre rows are equal by construction and none of them is an annealed row; a ray shard maps its rows
    to their positions in the global per-image list."""
    offset, P_global = shard if shard is not None else (0, P)
    if shared_center_ok(P, shard):
        # local row n < P -> offset + n; the shared centre row (local index P) -> a centre row of the global list
        wpts = pts if pts.shape[1] == P + 1 else torch.cat([pts[:, :P], pts[:, P:P + 1]], dim=1)
        return wpts, (offset, P, P_global - P), True
    if pts.shape[1] != 2 * P:
        raise RuntimeError("niw_b200: a shared centre row needs at least %d rays per image" % NVP_QUIRK_ROWS)
    if shard is None:
        return pts, None, False
    return pts, (offset, P, P_global - P), False


def tc_selftest(A, Bm, variant=0):
    """D = A . Bm^T through the tcgen05 staging used by the MLP kernel (A [128,K], Bm [N,K]; variant 4 is the
    CTA-pair form, cta_group::2, with A [256,K])."""
    lib = _lib.load()
    A, Bm = _f32(A, "A"), _f32(Bm, "B")
    N, K = Bm.shape
    if A.shape[0] != (256 if int(variant) == 4 else 128):
        raise ValueError("tc_selftest: A must have %d rows for variant %d" % (256 if int(variant) == 4 else 128, variant))
    D = torch.zeros(A.shape[0], N, device=A.device)
    _lib.check(lib.niw_tc_selftest(_p(A), _p(Bm), N, K, int(variant), _p(D), _stream()))
    return D
