"""Drop-in for reference ``model/pose_models/inn.py`` (INNPoseParams :9-102): owns the DTU warp
(latent codes, NVP network, the Kabsch-fitted global poses)."""
import torch
from torch import nn

from ... import camera
from ...nvp import DeformNetwork


class INNPoseParams(nn.Module):

    def __init__(self, opt, num_poses, initial_poses_w2c, device="cuda"):
        super().__init__()
        self.opt = opt
        self.num_poses = num_poses
        self.device = opt.device
        self.initial_poses_w2c = initial_poses_w2c
        self.init_poses_embed()

    def init_poses_embed(self):
        rn = self.opt.inn.real_nvp
        self.pose_latent = nn.Embedding(self.num_poses, rn.latent_dim).to(self.opt.device)
        self.pose_embedding = DeformNetwork(d_feature=rn.latent_dim, d_in=3, d_out_1=1, d_out_2=3, n_blocks=3,
                                            d_hidden=rn.d_hidden, n_layers=1, skip_in=[], multires=rn.multires,
                                            weight_norm=True, actfn=self.opt.inn.actfn).to(self.opt.device)
        self.pose_global = nn.Embedding(self.num_poses, 12).to(self.opt.device)

    def get_w2c_poses(self):
        return self.pose_global.weight.data.detach().clone().view(-1, 3, 4)

    def get_warped_rays_in_world(self, var, mode=None, iter=None):
        """inn.py:63-77: un-warped grid/centre in the initial-pose world frame -> NVP warp ->
        (ray, center, grid) [B,P,3]; then the rigid fit of :96-102."""
        if mode != "train":
            raise AssertionError("get_warped_rays_in_world is a training-path function")
        P = len(var.ray_idx)
        from ... import functional as F
        shared = F.shared_center_ok(P, F.ray_shard)
        with torch.no_grad():
            pts = camera.unwarped_points(self.opt, var.intr, ray_idx=var.ray_idx, pose_init=self.initial_poses_w2c,
                                         shared_center=shared)
        self.grid_init, self.center_init = pts[:, :P], (pts[:, P:].expand(-1, P, -1) if shared else pts[:, P:])
        wpts, index_map, shared = F.warp_point_list(pts, P, F.ray_shard)
        out = self.forward_inn(self.center_init, self.grid_init, iter, _pts=wpts, _index_map=index_map)[:, :, 0]
        grid_pred = out[:, :P]
        ray, center_pred = F.rays_from_warp_shared(out, P) if shared else F.rays_from_warp(out, P)
        self.solve_for_global_transformation(grid_pred, center_pred)
        return ray, center_pred, grid_pred

    def forward_inn(self, centers, grids, iter, _pts=None, _index_map=None):
        """inn.py:81-93 -> warped [B,2P,1,3] ([grid rows ; centre rows])."""
        rn = self.opt.inn.real_nvp
        alpha_ratio = max(min(iter / rn.max_pe_iter, 1), 0) if rn.c2f == True else 1   # noqa: E712
        pts = _pts if _pts is not None else torch.cat([grids, centers], dim=1)
        return self.pose_embedding.forward(self.pose_latent.weight, pts.unsqueeze(2), alpha_ratio=alpha_ratio,
                                           index_map=_index_map)

    def solve_for_global_transformation(self, grid_pred, center_pred):
        """inn.py:96-102 (roma.rigid_points_registration -> batched Kabsch in ``camera``)."""
        with torch.no_grad():
            source = torch.cat([self.grid_init, self.center_init], dim=1)
            target = torch.cat([grid_pred, center_pred], dim=1)
            R, t = camera.rigid_points_registration(target, source)
            self.pose_global.weight.data = torch.cat((R, t[..., None]), -1).view(-1, 12)
