"""Drop-in for reference ``model/nerf.py`` (Graph :243-365, NeRF :367-483): NeRF with given poses."""
import numpy as np
import torch

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

    def forward(self, opt, var, mode=None):
        """model/nerf.py:251-274."""
        batch_size = len(var.idx)
        if opt.nerf.rand_rays and mode in ["train", "test-optim"]:
            pose = self.get_pose(opt, var, mode=mode)
            var.ray_idx = torch.randperm(opt.H * opt.W, device=opt.device)[:opt.nerf.rand_rays // batch_size]
            with self._loss_target(opt, var, mode):     # the image loss of compute_loss rides in the compositor's epilogue
                ret = self.render(opt, pose, intr=var.intr, ray_idx=var.ray_idx, mode=mode)
        elif mode == "render_train":
            ind = np.random.choice(len(var.idx))
            pose = self.get_pose(opt, var, mode=mode, ind=ind)
            ret = self.render_by_slices(opt, pose[ind][None], intr=var.intr[ind][None], mode=mode)
            var.render_train_idx = ind
        else:
            pose = self.get_pose(opt, var, mode=mode)
            ret = self.render_by_slices(opt, pose, intr=var.intr, mode=mode) if opt.nerf.rand_rays else \
                self.render(opt, pose, intr=var.intr, mode=mode)
        var.update(ret)
        return var

    def compute_loss(self, opt, var, mode=None):
        """model/nerf.py:276-288."""
        return self._image_losses(opt, var, mode)

    def get_pose(self, opt, var, mode=None, ind=None):
        return var.pose

    def render(self, opt, pose, intr=None, ray_idx=None, mode=None):
        """model/nerf.py:293-319 -> edict(rgb, depth, opacity[, *_fine]) of [B,P,K]."""
        return self._render_pose(opt, pose, intr=intr, ray_idx=ray_idx, mode=mode)

    def render_by_slices(self, opt, pose, intr=None, mode=None):
        """model/nerf.py:321-332.  Each slice is a contiguous pixel range, generated directly."""
        return self._slices(opt, lambda c, n: self._render_pose(opt, pose, intr=intr, mode=mode, idx_start=c, num=n))
