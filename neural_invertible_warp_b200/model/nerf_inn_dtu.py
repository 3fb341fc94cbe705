"""Drop-in for reference ``model/nerf_inn_dtu.py`` (Graph :363-567): as ``nerf_inn_llff`` with a
per-scene metric ``depth_range`` carried by ``var`` and the warp owned by ``pose_net``."""
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

    def forward(self, opt, var, mode=None, iter=None):
        """model/nerf_inn_dtu.py:371-396."""
        batch_size = len(var.idx)
        depth_range = opt.nerf.depth.range if opt.nerf.depth.param == "inverse" else var.depth_range[0]
        if torch.is_tensor(depth_range) and not depth_range.is_cuda:
            depth_range = [float(v) for v in depth_range.tolist()]
        # (a CUDA tensor [min, max] goes to the sampler kernel as it is -- the reference unpacks it element-wise on the
        # host, nerf_inn_dtu.py:536; here there is no host read, so the step can be captured in a CUDA graph)
        if opt.nerf.rand_rays and mode in ["train", "test-optim"]:
            var.ray_idx = torch.randperm(opt.H * opt.W, device=opt.device)[:opt.nerf.rand_rays // batch_size]
            if mode == "test-optim":
                # the reference unpacks (ray, center, grid) from get_pose here as well, but in this mode get_pose returns
                # the aligned test POSE (barf_inn_dtu.py:546-564) and the unpack raises; the evident intent -- the BARF
                # test-time refinement (barf_inn_dtu.py:468-483), rendering the drawn pixels from that pose -- is implemented
                pose_w2c = self.get_pose(opt, var, mode=mode)
                with self._loss_target(opt, var, mode):
                    ret = self.render(opt, pose_w2c, intr=var.intr, ray_idx=var.ray_idx, mode=mode, depth_range=depth_range)
                var.update(ret)
                return var
            ray, center, grid_3d = self.get_pose(opt, var, mode=mode, iter=iter)
            with self._loss_target(opt, var, mode):     # the image losses ride in the compositors' epilogues
                ret = self.render_local(opt, ray, center, intr=var.intr, mode=mode, depth_range=depth_range)
            ret.update(grid_local=grid_3d, center_local=center, grid_init=self.pose_net.grid_init,
                       center_init=self.pose_net.center_init)
        else:
            pose_w2c = self.get_pose(opt, var, mode=mode)
            ret = self.render_by_slices(opt, pose_w2c, intr=var.intr, mode=mode, depth_range=depth_range) \
                if opt.nerf.rand_rays else self.render(opt, pose_w2c, intr=var.intr, mode=mode, depth_range=depth_range)
        var.update(ret)
        return var

    def compute_loss(self, opt, var, mode=None):
        """model/nerf_inn_dtu.py:398-415."""
        loss = self._image_losses(opt, var, mode)
        if mode == "train" and opt.loss_weight.global_alignment is not None:
            target = torch.cat([var.grid_local, var.center_local], dim=1)
            source = torch.cat([var.grid_init, var.center_init], dim=1)
            loss.global_alignment = self.MSE_loss(target, camera.cam2world(source, self.pose_net.get_w2c_poses()))
        return loss

    def get_pose(self, opt, var, mode=None):
        return var.pose

    def render_local(self, opt, ray, center, intr=None, ray_idx=None, mode=None, depth_range=None):
        """model/nerf_inn_dtu.py:420-456."""
        return self._render_local(opt, ray, center, intr=intr, ray_idx=ray_idx, mode=mode, depth_range=depth_range)

    def render(self, opt, pose, intr=None, ray_idx=None, mode=None, depth_range=None):
        """model/nerf_inn_dtu.py:472-509."""
        return self._render_pose(opt, pose, intr=intr, ray_idx=ray_idx, mode=mode, depth_range=depth_range)

    def render_by_slices(self, opt, pose, intr=None, mode=None, depth_range=None):
        """model/nerf_inn_dtu.py:511-522 (``render_by_slices_local`` :458-470 is the same loop)."""
        return self._slices(opt, lambda c, n: self._render_pose(opt, pose, intr=intr, mode=mode, idx_start=c, num=n,
                                                                depth_range=depth_range))

    render_by_slices_local = render_by_slices
