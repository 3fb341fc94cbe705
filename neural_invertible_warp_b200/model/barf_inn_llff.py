"""Drop-in for reference ``model/barf_inn_llff.py`` (Graph :273-424, NeRF :427-442): the per-image
pose is an invertible NVP warp of the camera-frame pixel grid and camera centre."""
import math

import torch

from .. import camera
from .. import functional as F
from . import barf, nerf_inn_llff
from ._core import NeRFCore


class NeRF(NeRFCore):
    has_progress = True


class Graph(nerf_inn_llff.Graph):

    def __init__(self, opt):
        super().__init__(opt)
        self.nerf = NeRF(opt)
        if opt.nerf.fine_sampling:
            self.nerf_fine = NeRF(opt)
        self.pose_eye = torch.eye(3, 4).to(opt.device)

    def _initial_pose(self, opt, var):
        """Initial world->camera poses and whether the un-warped points live in that world frame
        (blender) or in the camera frame (LLFF: pose_init None, barf_inn_llff.py:309-323)."""
        if opt.data.dataset == "blender":
            if opt.camera.noise_type == "barf":
                var.pose_noise = self.pose_noise[var.idx]
                pose = camera.pose.compose([var.pose_noise, var.pose])
            elif opt.camera.noise_type == "l2g":
                var.pose_noise = self.pose_noise[var.idx]
                pose = camera.pose.compose([var.pose, var.pose_noise])
            else:
                pose = var.pose
            return pose, pose
        return self.pose_eye[None].repeat(len(var.idx), 1, 1), None

    def get_pose_init(self, opt, var, mode=None, ind=None, iter=None):
        """barf_inn_llff.py:282-302."""
        if mode == "train":
            return self._initial_pose(opt, var)[0]

    def _warp_network(self):
        """``self.warp_mlp`` is attached by the engine (model/barf_inn_llff.py:54-55).  Behind the reference's engine it must be
        this package's fused ``DeformNetwork`` -- ``dropin.install_dropin`` swaps ``model.nvp.nvp_ndr.DeformNetwork`` -- not the
        reference's eager module: there is no non-CUDA-kernel path."""
        from ..nvp import DeformNetwork
        if not isinstance(self.warp_mlp, DeformNetwork):
            raise RuntimeError("niw_b200: graph.warp_mlp is %s.%s, not neural_invertible_warp_b200.nvp.DeformNetwork -- call "
                               "neural_invertible_warp_b200.dropin.install_dropin(reference_root) before the engine builds its "
                               "networks (it swaps model.nvp.nvp_ndr.DeformNetwork)"
                               % (type(self.warp_mlp).__module__, type(self.warp_mlp).__name__))
        return self.warp_mlp

    def _prefetch_pose(self, opt, var):
        """The warp network's weight pack depends on its parameters and the latent codes only: it runs on the render side
        stream while the pixel draw and the un-warped grid are produced (``DeformNetwork.prepack``)."""
        from ..nvp import DeformNetwork
        if not isinstance(getattr(self, "warp_mlp", None), DeformNetwork) or not torch.is_grad_enabled() \
                or opt.warp_latent.enc_type != "l2fbarf":      # (other encodings build a new code tensor per call)
            return
        cur = torch.cuda.current_stream()
        # (a stream of its own: on the render side stream the depth samples and the MLP weight pack of ``prefetch_render``
        # would queue behind this pack instead of running beside it)
        side = getattr(self, "_pose_stream", None)
        if side is None:
            side = self._pose_stream = torch.cuda.Stream()
        side.wait_stream(cur)
        self.warp_mlp.prepack(self._latent(opt), side)

    def _latent(self, opt):
        if opt.warp_latent.enc_type == "l2fbarf":
            return self.warp_latent.weight
        if opt.warp_latent.enc_type == "posenc":
            return self.positional_encoding(opt, self.frame_id, opt.warp_latent.posenc.freq_len)
        raise NotImplementedError("warp_latent.enc_type=%r" % (opt.warp_latent.enc_type,))

    def get_pose(self, opt, var, mode=None, ind=None, iter=None):
        """barf_inn_llff.py:305-399.  train: (ray, center_3D, grid_3D, alpha_ratio), each [B,P,3]."""
        if mode == "train":
            _, pose_init = self._initial_pose(opt, var)
            P = len(var.ray_idx)
            if opt.inn.real_nvp.c2f == True:   # noqa: E712  (the reference compares with == True)
                alpha_ratio = max(min(iter / opt.inn.real_nvp.max_pe_iter, 1), 0)
            else:
                alpha_ratio = 1
            shared = F.shared_center_ok(P, F.ray_shard)
            if shared and F.fused_warped_rays:
                # grid points, warp and ray construction in one launch (csrc/nvp.cu niw_nvp_rays_fwd)
                offset, P_global = F.ray_shard if F.ray_shard is not None else (0, P)
                ray, center_3D, grid_3D, pts = self._warp_network().warped_rays(
                    self._latent(opt), var.intr, pose_init, var.ray_idx, opt.H, opt.W, alpha_ratio=alpha_ratio,
                    index_map=(offset, P, P_global - P))
                var.grid_cam, var.center_cam = pts[:, :P], pts[:, P:].expand(-1, P, -1)
                return ray, center_3D, grid_3D, alpha_ratio
            # [grid ; centre] rows for the sampled pixels only, no gradient (:325-330, :348)
            with torch.no_grad():
                pts = camera.unwarped_points(opt, var.intr, ray_idx=var.ray_idx, pose_init=pose_init, shared_center=shared)
            var.grid_cam, var.center_cam = pts[:, :P], (pts[:, P:].expand(-1, P, -1) if shared else pts[:, P:])
            # the warp sees [grid rows ; centre] with the centre evaluated once per image when that is exact, and a ray
            # shard's rows at their positions in the global list (functional.warp_point_list)
            wpts, index_map, shared = F.warp_point_list(pts, P, F.ray_shard)
            warped = self._warp_network().forward(self._latent(opt), wpts.unsqueeze(2), alpha_ratio=alpha_ratio,
                                                  index_map=index_map)[:, :, 0]
            ray, center_3D = F.rays_from_warp_shared(warped, P) if shared else F.rays_from_warp(warped, P)
            return ray, center_3D, warped[:, :P], alpha_ratio
        if mode == "render_train":
            # the reference's branch calls warp_mlp.forward with a wrong arity (:378, SURVEY.md A.6 iii)
            # and has no live caller; implemented with the evident intent (image ``ind``, alpha 1).
            with torch.no_grad():
                pts = camera.unwarped_points(opt, var.intr[ind][None])
            P = pts.shape[1] // 2
            warped = self._warp_network().forward(self._latent(opt)[ind][None], pts.unsqueeze(2), alpha_ratio=1)[:, :, 0]
            return warped[:, :P] - warped[:, P:], warped[:, P:]
        if mode in ["val", "eval", "test-optim"]:
            return barf.Graph._aligned_test_pose(self, opt, var, mode)
        return var.pose

    def positional_encoding(self, opt, input, L):
        """barf_inn_llff.py:416-423 (un-weighted encoding of small per-image inputs)."""
        return _plain_encoding(input, L)


def _plain_encoding(x, L):
    freq = 2 ** torch.arange(L, dtype=torch.float32, device=x.device) * math.pi
    spectrum = x[..., None] * freq
    return torch.stack([spectrum.sin(), spectrum.cos()], dim=-2).reshape(*x.shape[:-1], -1)
