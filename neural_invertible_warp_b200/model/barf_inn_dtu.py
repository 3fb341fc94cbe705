"""Drop-in for reference ``model/barf_inn_dtu.py`` (Graph :524-591, NeRF :593-622)."""
import torch

from .. import camera
from . import nerf_inn_dtu
from ._core import NeRFCore


class NeRF(NeRFCore):
    has_progress = True


def backtrack_from_aligning_the_trajectory(pose_GT_w2c, sim3):
    """utils/geometry/align_trajectories.py:94-101: express GT test poses in the optimised frame."""
    c2w = camera.pose.invert(pose_GT_w2c)
    Rt = sim3.R.transpose(-2, -1)
    R_al = Rt @ c2w[:, :3, :3]
    t_al = Rt / sim3.s @ (c2w[:, :3, 3:4] - sim3.t)
    return camera.pose.invert(camera.pose(R=R_al, t=t_al.reshape(-1, 3)))


class Graph(nerf_inn_dtu.Graph):

    def __init__(self, opt, pose_net):
        super().__init__(opt)
        self.pose_net = pose_net
        self.nerf = NeRF(opt)
        if opt.nerf.fine_sampling:
            self.nerf_fine = NeRF(opt)
        self.pose_eye = torch.eye(3, 4).to(opt.device)

    def get_pose(self, opt, var, mode=None, iter=None):
        return self.get_w2c_pose(opt, var, mode, iter)

    def get_w2c_pose(self, opt, var, mode=None, iter=None):
        """model/barf_inn_dtu.py:538-564."""
        if mode == "train":
            if iter is None:
                raise AssertionError("ERROR: Iteration is needed for the c2f embedding in INN")
            return self.pose_net.get_warped_rays_in_world(var, mode, iter)
        if mode in ["val", "eval", "test-optim", "test"]:
            sim3 = self.pose_net.sim3_est_to_gt_c2w
            if sim3.type != "traj_align":
                raise NotImplementedError if sim3.type == "align_to_first" else ValueError(sim3.type)
            pose = backtrack_from_aligning_the_trajectory(var.pose, sim3)
            if opt.optim.test_photo and mode != "val":
                pose = camera.pose.compose([var.pose_refine_test, pose])
            return pose
        raise ValueError(mode)

    def get_c2w_pose(self, opt, var, mode=None):
        return camera.pose.invert(self.get_w2c_pose(opt, var, mode))
