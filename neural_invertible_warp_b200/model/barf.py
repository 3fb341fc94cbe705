"""Drop-in for reference ``model/barf.py`` (Graph :208-248, NeRF :250-268): bundle-adjusting NeRF
with a learnable se(3) correction per image and coarse-to-fine positional encoding."""
import torch

from .. import camera
from . import nerf


class NeRF(nerf.NeRF):
    has_progress = True    # ``progress`` Parameter: checkpointed, drives the c2f band weights


class Graph(nerf.Graph):

    def __init__(self, opt):
        super().__init__(opt)
        self.nerf = NeRF(opt)
        if opt.nerf.fine_sampling:
            self.nerf_fine = NeRF(opt)
        self.pose_eye = torch.eye(3, 4).to(opt.device)

    def _aligned_test_pose(self, opt, var, mode):
        """model/barf.py:235-246: map a GT test pose into the optimised frame through ``self.sim3``
        (set by the engine's pose pre-alignment), optionally composed with the test-time refinement."""
        sim3 = self.sim3
        center = torch.zeros(1, 1, 3, device=opt.device)
        center = camera.cam2world(center, var.pose)[:, 0]
        center_aligned = (center - sim3.t0) / sim3.s0 @ sim3.R * sim3.s1 + sim3.t1
        R_aligned = var.pose[..., :3] @ sim3.R
        t_aligned = (-R_aligned @ center_aligned[..., None])[..., 0]
        pose = camera.pose(R=R_aligned, t=t_aligned)
        if opt.optim.test_photo and mode != "val":
            pose = camera.pose.compose([var.pose_refine_test, pose])
        return pose

    def get_pose(self, opt, var, mode=None, ind=None):
        """model/barf.py:217-248.  B-sized Lie algebra in PyTorch; its gradient arrives from the
        ray-generation kernel's d_pose."""
        if mode == "train":
            if opt.data.dataset == "blender":
                if opt.camera.noise:
                    var.pose_noise = self.pose_noise[var.idx]
                    pose = camera.pose.compose([var.pose_noise, var.pose])
                else:
                    pose = var.pose
            else:
                pose = self.pose_eye
            var.se3_refine = self.se3_refine.weight[var.idx]
            pose = camera.pose.compose([camera.lie.se3_to_SE3(var.se3_refine), pose])
        elif mode == "render_train":
            var.se3_refine = self.se3_refine.weight[ind][None]
            pose = camera.pose.compose([camera.lie.se3_to_SE3(var.se3_refine), self.pose_eye])
        elif mode in ["val", "eval", "test-optim"]:
            pose = self._aligned_test_pose(opt, var, mode)
        else:
            pose = var.pose
        return pose
