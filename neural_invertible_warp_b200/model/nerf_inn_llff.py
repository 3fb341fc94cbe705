"""Drop-in for reference ``model/nerf_inn_llff.py`` (Graph :485-703): the training path renders
from rays produced by the invertible warp (``get_pose`` of the BARF-INN subclass), the evaluation
path from aligned poses."""
import numpy as np
import torch

from .. import camera
from . import base
from ._core import NeRFCore, edict


class NeRF(NeRFCore):
    pass


class Graph(base.Graph):

    def __init__(self, opt):
        super().__init__(opt)
        self.nerf = NeRF(opt)
        if opt.nerf.fine_sampling:
            self.nerf_fine = NeRF(opt)

    def _rescale_depth_range_l2g(self, opt):
        """model/nerf_inn_llff.py:495-503 (blender + l2g noise only)."""
        depth_min, depth_max = opt.nerf.depth.range
        position = camera.pose.invert(self.global_rigid.weight.data.detach().clone().view(-1, 3, 4))[..., -1]
        diameter = (position[self.idx_grid[..., 0]] - position[self.idx_grid[..., 1]]).norm(dim=-1).max()
        opt.nerf.depth.range = [(depth_min / (depth_max + depth_min)) * diameter,
                                (depth_max / (depth_max + depth_min)) * diameter]

    def _prefetch_pose(self, opt, var):
        """Hook: work of ``get_pose`` that depends neither on the pixel draw nor on the rays (none here)."""

    def forward(self, opt, var, mode=None, iter=None):
        """model/nerf_inn_llff.py:493-546."""
        if opt.data.dataset == "blender" and opt.camera.noise_type == "l2g":
            self._rescale_depth_range_l2g(opt)
        batch_size = len(var.idx)
        if opt.nerf.rand_rays and mode == "train":
            self._prefetch_pose(opt, var)           # (barf_inn_llff: the warp network's weight pack, on the side stream)
        if opt.nerf.rand_rays and mode in ["train", "test-optim"]:
            var.ray_idx = torch.randperm(opt.H * opt.W, device=opt.device)[:opt.nerf.rand_rays // batch_size]
            if mode == "train":
                # depth samples and packed MLP weights do not depend on the rays: a side stream prepares them while the
                # warp kernels of get_pose run (same draws, same values; only the launch order differs)
                pre = self.prefetch_render(opt, batch_size, len(var.ray_idx)) if not opt.camera.ndc else None
                ray, center, grid_3D, alpha_ratio = self.get_pose(opt, var, mode=mode, iter=iter)
                with self._loss_target(opt, var, mode):     # the image loss rides in the compositor's epilogue
                    ret = self._render_local(opt, ray, center, intr=var.intr, mode=mode, prefetched=pre)
                # the un-warped points were generated inside get_pose (one kernel instead of the
                # reference's two full-frame grids, nerf_inn_llff.py:519 and barf_inn_llff.py:325)
                ret.update(grid_3D=grid_3D, center=center, grid_cam=var.grid_cam, center_cam=var.center_cam,
                           inn_posenc_alpha=alpha_ratio)
            else:
                pose = self.get_pose(opt, var, mode=mode)
                with self._loss_target(opt, var, mode):
                    ret = self.render(opt, pose, intr=var.intr, ray_idx=var.ray_idx, mode=mode)
        elif mode == "render_train":
            ind = np.random.choice(len(var.idx))
            ray, center = self.get_pose(opt, var, mode=mode, ind=ind)
            ret = self.render_by_slices_local(opt, ray, center, intr=var.intr[ind][None], mode=mode) \
                if opt.nerf.rand_rays else self.render_local(opt, ray, center, intr=var.intr, mode=mode)
            var.render_train_idx = ind
        else:
            pose = self.get_pose(opt, var, mode=mode)
            ret = self.render_by_slices(opt, pose, intr=var.intr, mode=mode) if opt.nerf.rand_rays else \
                self.render(opt, pose, intr=var.intr, mode=mode)
        var.update(ret)
        return var

    def compute_loss(self, opt, var, mode=None):
        """model/nerf_inn_llff.py:548-573."""
        loss = self._image_losses(opt, var, mode)
        if opt.loss_weight.global_alignment is not None and mode == "train":
            source = torch.cat([var.grid_cam, var.center_cam], dim=1)
            target = torch.cat([var.grid_3D, var.center], dim=1)
            with torch.no_grad():
                R_global, t_global = camera.rigid_points_registration(target, source)
                svd_poses = torch.cat((R_global, t_global[..., None]), -1)
            self.global_rigid.weight.data = svd_poses.detach().clone().view(-1, 12)
            loss.global_alignment = self.MSE_loss(target, camera.cam2world(source, svd_poses))
        return loss

    def get_pose(self, opt, var, mode=None):
        return var.pose

    def render_local(self, opt, ray, center, intr=None, ray_idx=None, mode=None):
        """model/nerf_inn_llff.py:581-612."""
        return self._render_local(opt, ray, center, intr=intr, ray_idx=ray_idx, mode=mode)

    def render_by_slices_local(self, opt, ray, center, intr=None, mode=None):
        """model/nerf_inn_llff.py:614-625."""
        def one(c, n):
            idx = torch.arange(c, c + n, device=opt.device)
            return self._render_local(opt, ray, center, intr=intr, ray_idx=idx, mode=mode)
        return self._slices(opt, one)

    def render(self, opt, pose, intr=None, ray_idx=None, mode=None):
        """model/nerf_inn_llff.py:627-656.  The reference unpacks three values from
        ``forward_samples`` here for BARF models and raises ValueError; this
        implements the evident two-value intent."""
        return self._render_pose(opt, pose, intr=intr, ray_idx=ray_idx, mode=mode)

    def render_by_slices(self, opt, pose, intr=None, mode=None):
        """model/nerf_inn_llff.py:658-669."""
        return self._slices(opt, lambda c, n: self._render_pose(opt, pose, intr=intr, mode=mode, idx_start=c, num=n))
