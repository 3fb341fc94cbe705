"""Host-side mirror of the reference's NVP warp network (model/nvp/nvp_ndr.py:229-468).

``DeformNetwork`` keeps the reference's parameter names (``lin{b}_{a,b}_0.{weight_g,weight_v,bias}``,
``lin{b}_{a,b}_1.{weight,bias}``, ``lin{b}_c.{weight,bias}``), constructor signature and
initialisation, so reference checkpoints load unchanged.  ``forward`` resolves the weight-norm
re-parametrisation and folds the per-image latent code into per-image first-layer biases with a
few B-sized PyTorch ops (autograd handles their backward), then evaluates all points in one CUDA
kernel (csrc/nvp.cu).
"""
import ctypes
import math
import warnings

import torch
from torch import nn

from . import _lib
from . import functional as F

N_FREQ = 6
HID = 128


def pack_effective(p, code, n_blocks=3, n_freq=N_FREQ, prefix=""):
    """(state-dict-style params, code [B,D]) -> (wpack [3*5636], code_bias [3,2,B,128]).

    w = g * v / ||v||_row is torch's legacy ``nn.utils.weight_norm`` (dim=0) used at
    model/nvp/nvp_ndr.py:291-292,335-336; code_b = lin_c(code) + code is :382.
    """
    chunks, biases = [], []
    for b in range(n_blocks):
        cb = torch.addmm(p[f"{prefix}lin{b}_c.bias"], code, p[f"{prefix}lin{b}_c.weight"].t()) + code
        for part, emb in (("a", 2 * (1 + 2 * n_freq)), ("b", 1 + 2 * n_freq)):
            name = f"{prefix}lin{b}_{part}_0"
            if name + ".weight_g" in p:
                v, g = p[name + ".weight_v"], p[name + ".weight_g"]
                w0 = v * (g / v.norm(dim=1, keepdim=True))
            else:
                w0 = p[name + ".weight"]
            w1 = w0[:, :emb]
            if part == "a":      # the packed image keeps W1a with an odd row stride (27), include/niw_b200.h
                w1 = torch.nn.functional.pad(w1, (0, 1))
            chunks += [w1.reshape(-1), p[f"{prefix}lin{b}_{part}_1.weight"].reshape(-1),
                       p[f"{prefix}lin{b}_{part}_1.bias"].reshape(-1)]
            biases.append(torch.addmm(p[name + ".bias"], cb, w0[:, emb:].t()))
        used = sum(c.numel() for c in chunks) - b * F.NIW_NVP_BLOCK_FLOATS
        chunks.append(code.new_zeros(F.NIW_NVP_BLOCK_FLOATS - used))     # pad the block to a multiple of 4 floats
    B = code.shape[0]
    return torch.cat(chunks), torch.stack(biases).view(n_blocks, 2, B, -1)


def _pack_forward(lib, params, code, B, dev):
    wpack = torch.empty(3 * F.NIW_NVP_BLOCK_FLOATS, device=dev)
    code_bias = torch.empty(3, 2, B, HID, device=dev)
    cb = torch.empty(3, B, HID, device=dev)
    ptrs = (ctypes.c_void_p * len(params))(*[p.data_ptr() for p in params])
    _lib.check(lib.niw_nvp_pack_fwd(ptrs, F._p(code), B, F._p(wpack), F._p(code_bias), F._p(cb), F._stream()))
    return wpack, code_bias, cb


class _NvpNetwork(torch.autograd.Function):
    """DeformNetwork.forward as four kernels: pack (weight-norm + code projection + per-image biases),
    warp; warp backward, pack backward.  With ``module.accumulate_grads_in_place`` (engine.use_flat_gradients) the
    parameter gradients are accumulated by the pack-backward kernel straight into the parameters' ``.grad``
    buffers (allocated zero-filled when absent) instead of being returned through autograd: that saves ~30 tiny
    accumulate launches per step.  Otherwise they are returned through autograd.  The latent code's gradient is
    always returned normally."""

    @staticmethod
    def forward(ctx, code, pts, alpha_ratio, module, index_map, *params):
        lib = _lib.load()
        code = F._f32(code, "deformation_code")
        pts = F._f32(pts, "input_pts")
        B, Pt = pts.shape[0], pts.shape[1]
        if code.shape != (B, HID):
            raise RuntimeError("niw_b200 DeformNetwork: latent code must be [B=%d, %d], got %s" % (B, HID, tuple(code.shape)))
        dev = pts.device
        pre = module._take_prepacked(code) if module is not None else None
        if pre is not None:
            # the pack ran earlier on a side stream (DeformNetwork.prepack): wait for it here, nothing to launch
            wpack, code_bias, cb, done = pre
            torch.cuda.current_stream().wait_event(done)
            if not torch.cuda.is_current_stream_capturing():
                for t in (wpack, code_bias, cb):
                    t.record_stream(torch.cuda.current_stream())
        else:
            wpack, code_bias, cb = _pack_forward(lib, params, code, B, dev)
        out = torch.empty_like(pts)
        im = F.index_map_args(index_map, Pt)
        _lib.check(lib.niw_nvp_warp_fwd(F._p(wpack), F._p(code_bias), F._p(pts), float(alpha_ratio), B, Pt, *im, F._p(out),
                                        F._stream()))
        ctx.save_for_backward(code, pts, wpack, code_bias, cb)
        ctx.module, ctx.alpha, ctx.im = module, float(alpha_ratio), im
        return out

    @staticmethod
    def backward(ctx, d_out):
        code, pts, wpack, code_bias, cb = ctx.saved_tensors
        d_code, grads = _warp_backward(ctx.module, ctx.alpha, ctx.im, code, pts, wpack, code_bias, cb, d_out)
        return (d_code, None, None, None, None) + grads


def _warp_backward(module, alpha, im, code, pts, wpack, code_bias, cb, d_out):
    """Warp backward + pack backward for the gradient ``d_out`` of the warped point list; returns d_code and the tuple of
    parameter gradients (Nones when the engine opted into accumulation in place into ``.grad``)."""
    lib = _lib.load()
    B, Pt = pts.shape[0], pts.shape[1]
    d_w = torch.empty_like(wpack)
    d_cb = torch.empty_like(code_bias)
    d_out = d_out.contiguous()
    ov = F.backward_overlap                 # engine: the MLP weight-gradient pass is running on another stream
    _lib.check(lib.niw_nvp_warp_bwd(F._p(wpack), F._p(code_bias), F._p(pts), alpha, B, Pt, *im, F._p(d_out),
                                    F._p(d_w), F._p(d_cb), int(ov.side_ctas) if ov is not None and ov.used else 0,
                                    F._stream()))
    params = module.ordered_parameters()
    in_place = getattr(module, "accumulate_grads_in_place", False)
    if in_place:
        for p in params:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        grads = [p.grad for p in params]
    else:
        # gradients go back through autograd (torch.autograd.grad, backward(inputs=...), hooks all behave)
        grads = [torch.zeros_like(p) for p in params]
    ptrs = (ctypes.c_void_p * len(params))(*[p.data_ptr() for p in params])
    gptrs = (ctypes.c_void_p * len(params))(*[g.data_ptr() for g in grads])
    d_code = torch.empty_like(code)
    _lib.check(lib.niw_nvp_pack_bwd(ptrs, gptrs, F._p(code), F._p(cb), F._p(d_w), F._p(d_cb), B, F._p(d_code), F._stream()))
    return d_code, ((None,) * len(params) if in_place else tuple(grads))


class _NvpRays(torch.autograd.Function):
    """Warped ray generation in ONE launch (csrc/nvp.cu ``niw_nvp_rays_fwd``; model/barf_inn_llff.py:325-364): pixel indices
    -> (ray, warped centre, warped grid points, un-warped [grid ; centre] list).  Backward: the three kernels of the
    separate path (rays-from-warp backward, warp backward, pack backward)."""

    @staticmethod
    def forward(ctx, code, intr, pose_init, ray_idx, alpha_ratio, module, index_map, H, W, *params):
        lib = _lib.load()
        code = F._f32(code, "deformation_code")
        intr, pose_init = F._f32(intr, "intr"), F._f32(pose_init, "pose_init")
        ray_idx = F._idx(ray_idx, intr.device)
        B, P = intr.shape[0], int(ray_idx.numel())
        dev = intr.device
        pre = module._take_prepacked(code)
        if pre is not None:
            wpack, code_bias, cb, done = pre
            torch.cuda.current_stream().wait_event(done)
            if not torch.cuda.is_current_stream_capturing():
                for t in (wpack, code_bias, cb):
                    t.record_stream(torch.cuda.current_stream())
        else:
            wpack, code_bias, cb = _pack_forward(lib, params, code, B, dev)
        pts = torch.empty(B, P + 1, 3, device=dev)
        warped = torch.empty(B, P + 1, 3, device=dev)
        ray = torch.empty(B, P, 3, device=dev)
        center = torch.empty(B, P, 3, device=dev)
        im = F.index_map_args(index_map, P + 1)
        _lib.check(lib.niw_nvp_rays_fwd(F._p(wpack), F._p(code_bias), F._p(intr), F._p(pose_init), F._p(ray_idx), 0,
                                        float(alpha_ratio), B, P, int(H), int(W), *im, F._p(pts), F._p(warped), F._p(ray),
                                        F._p(center), F._stream()))
        ctx.save_for_backward(code, pts, wpack, code_bias, cb)
        ctx.module, ctx.alpha, ctx.im, ctx.P = module, float(alpha_ratio), im, P
        ctx.mark_non_differentiable(pts)
        return ray, center, warped[:, :P], pts

    @staticmethod
    def backward(ctx, d_ray, d_center, d_grid, _d_pts):
        code, pts, wpack, code_bias, cb = ctx.saved_tensors
        B, P = pts.shape[0], ctx.P
        d_warped = torch.empty(B, P + 1, 3, device=pts.device)
        c = lambda t: None if t is None else t.contiguous()
        if d_ray is None and d_center is None:
            d_warped.zero_()
        else:
            _lib.check(_lib.load().niw_rays_from_warp_bwd(F._p(c(d_ray)), F._p(c(d_center)), B, P, 1, F._p(d_warped), F._stream()))
        if d_grid is not None:
            d_warped[:, :P] += d_grid
        d_code, grads = _warp_backward(ctx.module, ctx.alpha, ctx.im, code, pts, wpack, code_bias, cb, d_warped)
        return (d_code, None, None, None, None, None, None, None, None) + grads


class DeformNetwork(nn.Module):
    """Drop-in for ``model.nvp.nvp_ndr.DeformNetwork`` restricted to what the target models
    instantiate (barf_inn_llff.py:54-55, pose_models/inn.py:23-27): d_in=3, n_blocks=3,
    n_layers=1, skip_in=[], multires=6, weight_norm=True, softplus.  Anything else raises."""

    accumulate_grads_in_place = False      # engine.use_flat_gradients

    def __init__(self, d_feature, d_in, d_out_1, d_out_2, n_blocks, d_hidden, n_layers, skip_in=(4,),
                 multires=0, weight_norm=True, actfn="softplus"):
        super().__init__()
        ok = (d_in == 3 and d_out_1 == 1 and d_out_2 == 3 and n_blocks == 3 and d_hidden == HID and n_layers == 1
              and len(tuple(skip_in)) == 0 and multires == N_FREQ and weight_norm and actfn == "softplus"
              and d_feature == HID)
        if not ok:
            raise RuntimeError("niw_b200 DeformNetwork: only the configuration used by barf_inn_llff / barf_inn_dtu "
                               "is implemented in CUDA (3 blocks, hidden 128, latent 128, 6 bands, weight-norm, softplus)")
        self.n_blocks, self.d_feature = n_blocks, d_feature
        for b in range(n_blocks):
            for part, ori in (("a", 2), ("b", 1)):
                emb = ori * (1 + 2 * multires)
                n_out = d_out_1 if part == "a" else d_out_2
                lin0 = nn.Linear(emb + d_feature, d_hidden)
                nn.init.constant_(lin0.bias, 0.0)
                nn.init.normal_(lin0.weight[:, :ori], 0.0, math.sqrt(2) / math.sqrt(d_hidden))
                nn.init.constant_(lin0.weight[:, ori:], 0.0)
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")      # legacy weight_norm: the reference's parametrisation / key names
                    lin0 = nn.utils.weight_norm(lin0)
                lin1 = nn.Linear(d_hidden, n_out)
                nn.init.constant_(lin1.bias, 0.0)
                nn.init.constant_(lin1.weight, 0.0)
                setattr(self, f"lin{b}_{part}_0", lin0)
                setattr(self, f"lin{b}_{part}_1", lin1)
        for b in range(n_blocks):
            lin = nn.Linear(d_feature, d_feature)
            nn.init.constant_(lin.bias, 0.0)
            nn.init.constant_(lin.weight, 0.0)
            setattr(self, f"lin{b}_c", lin)
        # weight_norm recomputes ``weight`` from (g, v) in a forward pre-hook; the CUDA path reads g and
        # v directly, so the hooks (4 launches each) are dropped -- the parameters stay weight_g / weight_v
        for m in self.modules():
            for k, hook in list(m._forward_pre_hooks.items()):
                if type(hook).__name__ == "WeightNorm":
                    del m._forward_pre_hooks[k]

    def ordered_parameters(self):
        """The 36 tensors in the order of the C ABI's pointer table (include/niw_b200.h)."""
        out = []
        for b in range(self.n_blocks):
            for part in ("a", "b"):
                l0, l1 = getattr(self, f"lin{b}_{part}_0"), getattr(self, f"lin{b}_{part}_1")
                out += [l0.weight_v, l0.weight_g, l0.bias, l1.weight, l1.bias]
            lc = getattr(self, f"lin{b}_c")
            out += [lc.weight, lc.bias]
        return out

    def _params(self):
        p = {}
        for b in range(self.n_blocks):
            for part in ("a", "b"):
                l0, l1 = getattr(self, f"lin{b}_{part}_0"), getattr(self, f"lin{b}_{part}_1")
                p[f"lin{b}_{part}_0.weight_g"], p[f"lin{b}_{part}_0.weight_v"] = l0.weight_g, l0.weight_v
                p[f"lin{b}_{part}_0.bias"] = l0.bias
                p[f"lin{b}_{part}_1.weight"], p[f"lin{b}_{part}_1.bias"] = l1.weight, l1.bias
            lc = getattr(self, f"lin{b}_c")
            p[f"lin{b}_c.weight"], p[f"lin{b}_c.bias"] = lc.weight, lc.bias
        return p

    def prepack(self, deformation_code, stream):
        """The part of ``forward`` that depends on the parameters and the latent codes only (weight-norm resolution, code
        projection, per-image first-layer biases: ``niw_nvp_pack_fwd``, ~15 us) launched on ``stream`` NOW, so that it
        overlaps whatever produces the points (pixel draw, un-warped grid).  The next ``forward`` with the same code
        tensor picks the result up and waits for it on its own stream; its backward is unchanged."""
        code = F._f32(deformation_code, "deformation_code")
        params = self.ordered_parameters()
        with torch.cuda.stream(stream):
            packed = _pack_forward(_lib.load(), params, code.detach(), code.shape[0], code.device)
            done = torch.cuda.Event()
            done.record(stream)
        self._prepacked = (code.data_ptr(), code._version, tuple(code.shape)) + (packed + (done,),)

    def _take_prepacked(self, code):
        pre, self._prepacked = getattr(self, "_prepacked", None), None
        if pre is not None and pre[:3] == (code.data_ptr(), code._version, tuple(code.shape)):
            return pre[3]
        return None

    def warped_rays(self, deformation_code, intr, pose_init, ray_idx, H, W, alpha_ratio=0, index_map=None):
        """The train-mode ray generation of model/barf_inn_llff.py:325-364 in one launch: un-warped grid points of the pixels
        ``ray_idx`` (camera.py:359-390, initial pose ``pose_init`` [B,3,4] or None), this network's warp, ray = warped grid -
        warped camera centre.  The centre is warped once per image, so ``functional.shared_center_ok`` must hold.  Returns
        (ray [B,P,3], center [B,P,3], warped grid [B,P,3], un-warped list [B,P+1,3] = [grid ; centre], no gradient)."""
        params = self.ordered_parameters()
        for p in params:
            if not p.is_contiguous():
                raise RuntimeError("niw_b200 DeformNetwork: parameters must be contiguous")
        return _NvpRays.apply(deformation_code, intr, pose_init, ray_idx, alpha_ratio, self, index_map, H, W, *params)

    def forward(self, deformation_code, input_pts, alpha_ratio=0, index_map=None):
        """deformation_code [B,D], input_pts [B,P,1,3] -> [B,P,1,3]  (nvp_ndr.py:365).  ``index_map`` (offset, split,
        jump): where the given points sit in the point list the reference would have built (include/niw_b200.h) --
        the embedder's annealing quirk is keyed on that position; None = the list is the reference's."""
        squeeze = input_pts.dim() == 4
        pts = input_pts[:, :, 0] if squeeze else input_pts
        params = self.ordered_parameters()
        for p in params:
            if not p.is_contiguous():
                raise RuntimeError("niw_b200 DeformNetwork: parameters must be contiguous")
        out = _NvpNetwork.apply(deformation_code, pts.detach(), alpha_ratio, self, index_map, *params)
        return out[:, :, None] if squeeze else out
