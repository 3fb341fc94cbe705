"""Camera algebra for the drop-in path (mirror of the reference's top-level ``camera.py``).

Split of work, following SURVEY.md section 8:

* everything that is *per ray* runs in CUDA (csrc/raygen.cu) and never materialises the full
  ``B x HW`` pixel grid the reference builds (camera.py:430-443): ``get_center_and_ray`` and
  ``get_unwarped_center_and_ray`` generate only the requested pixels;
* everything that is *per image* (``B`` poses: Lie algebra, pose composition) stays in PyTorch on the
  device -- a handful of 3x3 operations whose autograd PyTorch handles; the Kabsch fit of the
  global-alignment loss (no gradient) is one kernel (csrc/kabsch.cu).

Names, argument meaning and return conventions follow the reference (``camera.pose``,
``camera.lie``, ``camera.cam2world`` ...), so model code written against it reads the same.
"""
import math

import torch

from . import functional as F


# --------------------------------------------------------------------------------------------
# poses [..., 3, 4] = [R | t], world -> camera      (reference camera.py:64-112)
# --------------------------------------------------------------------------------------------

class Pose:
    def __call__(self, R=None, t=None):
        if R is None and t is None:
            raise AssertionError("Pose(): need R and/or t")
        if R is not None and not torch.is_tensor(R):
            R = torch.tensor(R)
        if t is not None and not torch.is_tensor(t):
            t = torch.tensor(t)
        if R is None:
            R = torch.eye(3, device=t.device).expand(*t.shape[:-1], 3, 3)
        if t is None:
            t = torch.zeros(R.shape[:-1], device=R.device)
        if R.shape[:-1] != t.shape or tuple(R.shape[-2:]) != (3, 3):
            raise AssertionError("Pose(): R [...,3,3] and t [...,3] expected")
        return torch.cat([R.float(), t.float()[..., None]], dim=-1)

    def invert(self, pose, use_inverse=False):
        R, t = pose[..., :3], pose[..., 3:]
        Ri = torch.linalg.inv(R) if use_inverse else R.transpose(-1, -2)
        return self(R=Ri, t=(-Ri @ t)[..., 0])

    def compose_pair(self, pose_a, pose_b):
        """x -> pose_b(pose_a(x))."""
        Ra, ta = pose_a[..., :3], pose_a[..., 3:]
        Rb, tb = pose_b[..., :3], pose_b[..., 3:]
        return self(R=Rb @ Ra, t=(Rb @ ta + tb)[..., 0])

    def compose(self, pose_list):
        out = pose_list[0]
        for p in pose_list[1:]:
            out = self.compose_pair(out, p)
        return out


class Lie:
    """so(3)/se(3) <-> SO(3)/SE(3) with the reference's 10-term series (camera.py:193-272)."""

    @staticmethod
    def _series(x, first_factorial_arg, nth=10):
        # sum_i (-1)^i x^(2i) / (2i + first_factorial_arg)!   evaluated term by term in fp32
        out = torch.zeros_like(x)
        x2 = x * x
        term = torch.ones_like(x) / math.factorial(first_factorial_arg)
        for i in range(nth + 1):
            if i > 0:
                a = 2 * i + first_factorial_arg
                term = term * x2 / float((a - 1) * a)
            out = out + (term if i % 2 == 0 else -term)
        return out

    def taylor_A(self, x, nth=10):   # sin(x)/x
        return self._series(x, 1, nth)

    def taylor_B(self, x, nth=10):   # (1-cos(x))/x^2
        return self._series(x, 2, nth)

    def taylor_C(self, x, nth=10):   # (x-sin(x))/x^3
        return self._series(x, 3, nth)

    def skew_symmetric(self, w):
        w0, w1, w2 = w.unbind(dim=-1)
        z = torch.zeros_like(w0)
        return torch.stack([torch.stack([z, -w2, w1], dim=-1),
                            torch.stack([w2, z, -w0], dim=-1),
                            torch.stack([-w1, w0, z], dim=-1)], dim=-2)

    def so3_to_SO3(self, w):
        wx = self.skew_symmetric(w)
        th = w.norm(dim=-1)[..., None, None]
        eye = torch.eye(3, device=w.device, dtype=torch.float32)
        return eye + self.taylor_A(th) * wx + self.taylor_B(th) * (wx @ wx)

    def SO3_to_so3(self, R, eps=1e-7):
        tr = R[..., 0, 0] + R[..., 1, 1] + R[..., 2, 2]
        th = ((tr - 1) / 2).clamp(-1 + eps, 1 - eps).acos()[..., None, None] % math.pi
        lnR = 1 / (2 * self.taylor_A(th) + 1e-8) * (R - R.transpose(-2, -1))
        return torch.stack([lnR[..., 2, 1], lnR[..., 0, 2], lnR[..., 1, 0]], dim=-1)

    def se3_to_SE3(self, wu):
        w, u = wu.split([3, 3], dim=-1)
        wx = self.skew_symmetric(w)
        th = w.norm(dim=-1)[..., None, None]
        eye = torch.eye(3, device=w.device, dtype=torch.float32)
        A, B, C = self.taylor_A(th), self.taylor_B(th), self.taylor_C(th)
        wx2 = wx @ wx
        R = eye + A * wx + B * wx2
        V = eye + B * wx + C * wx2
        return torch.cat([R, V @ u[..., None]], dim=-1)

    def SE3_to_se3(self, Rt, eps=1e-8):
        R, t = Rt.split([3, 1], dim=-1)
        w = self.SO3_to_so3(R)
        wx = self.skew_symmetric(w)
        th = w.norm(dim=-1)[..., None, None]
        eye = torch.eye(3, device=w.device, dtype=torch.float32)
        A, B = self.taylor_A(th), self.taylor_B(th)
        invV = eye - 0.5 * wx + (1 - A / (2 * B)) / (th ** 2 + eps) * (wx @ wx)
        return torch.cat([w, (invV @ t)[..., 0]], dim=-1)


pose = Pose()
lie = Lie()
_POSE = pose      # functions below take a ``pose`` argument (the reference's name), which shadows the module-level object


def to_hom(X):
    return torch.cat([X, torch.ones_like(X[..., :1])], dim=-1)


def world2cam(X, pose):
    return to_hom(X) @ pose.transpose(-1, -2)


def cam2img(X, cam_intr):
    return X @ cam_intr.transpose(-1, -2)


def img2cam(X, cam_intr):
    return X @ torch.linalg.inv(cam_intr).transpose(-1, -2)


def cam2world(X, pose):
    return to_hom(X) @ _POSE.invert(pose).transpose(-1, -2)


# --------------------------------------------------------------------------------------------
# ray generation (CUDA)
# --------------------------------------------------------------------------------------------

def _check_perspective(opt):
    if opt.camera.model != "perspective":   # the reference asserts the same (camera.py:427)
        raise AssertionError("only the perspective camera model is supported")


def _bcast_intr(intr, B):
    return intr if intr.shape[0] == B else intr.expand(B, 3, 3)


def get_center_and_ray(opt, pose, intr=None, ray_idx=None, idx_start=0, num=None):
    """reference camera.py:419-443 followed by the ``[:, ray_idx]`` of model/nerf.py:298-300.

    pose [B,3,4] (or [3,4], broadcast over the images of ``intr``), intr [B,3,3] ->
    center, ray [B,P,3].  ``ray_idx=None`` renders pixels ``idx_start .. idx_start+num-1``
    (whole frame by default).  Differentiable with respect to ``pose``.
    """
    _check_perspective(opt)
    B = max(pose.shape[0] if pose.dim() == 3 else 1, intr.shape[0])
    if pose.dim() == 2 or pose.shape[0] != B:
        pose = pose.expand(B, 3, 4)
    return F.raygen_pose(pose, _bcast_intr(intr, B), opt.H, opt.W, ray_idx=ray_idx, idx_start=idx_start, num=num)


def get_unwarped_center_and_ray(opt, intr=None, ray_idx=None, pose_init=None, idx_start=0, num=None):
    """reference camera.py:359-390: camera-frame (or ``pose_init`` world-frame) pixel grid and
    centre, no gradient.  Returns (center_3D, grid_3D), each [B,P,3]."""
    _check_perspective(opt)
    B = intr.shape[0]
    if pose_init is not None and (pose_init.dim() == 2 or pose_init.shape[0] != B):
        pose_init = pose_init.expand(B, 3, 4)
    with torch.no_grad():
        pts = unwarped_points(opt, intr, ray_idx=ray_idx, pose_init=pose_init, idx_start=idx_start, num=num)
    P = pts.shape[1] // 2
    return pts[:, P:], pts[:, :P]


def unwarped_points(opt, intr, ray_idx=None, pose_init=None, idx_start=0, num=None, shared_center=False):
    """[grid rows ; centre rows] [B,2P,3] -- the concatenation barf_inn_llff.py:348 feeds the warp; with
    ``shared_center`` [B,P+1,3]: the centre row (identical for every ray of an image) once."""
    return F.raygen_unwarped(intr, opt.H, opt.W, ray_idx=ray_idx, pose_init=pose_init, idx_start=idx_start, num=num,
                             shared_center=shared_center)


def get_3D_points_from_depth(opt, center, ray, depth, multi_samples=False):
    """x = c + d v (reference camera.py:517-521).  Utility only: the render path evaluates this
    inside the fused encoding+MLP kernel and never stores the points."""
    if multi_samples:
        center, ray = center[:, :, None], ray[:, :, None]
    return center + ray * depth


def convert_NDC(opt, center, ray, intr, near=1):
    """reference camera.py:523-540 (per-ray, PyTorch; ``camera.ndc`` is false in every target YAML)."""
    center = center + (near - center[..., 2:]) / ray[..., 2:] * ray
    cx, cy, cz = center.unbind(dim=-1)
    rx, ry, rz = ray.unbind(dim=-1)
    sx = (intr[:, 0, 0] / intr[:, 0, 2])[:, None]
    sy = (intr[:, 1, 1] / intr[:, 1, 2])[:, None]
    center_ndc = torch.stack([sx * (cx / cz), sy * (cy / cz), 1 - 2 * near / cz], dim=-1)
    ray_ndc = torch.stack([sx * (rx / rz - cx / cz), sy * (ry / rz - cy / cz), 2 * near / cz], dim=-1)
    return center_ndc, ray_ndc


# --------------------------------------------------------------------------------------------
# rigid registration (replaces roma.rigid_points_registration, roma==1.4.1:
# model/nerf_inn_llff.py:569, model/pose_models/inn.py:100)
# --------------------------------------------------------------------------------------------

def rigid_points_registration(x, y):
    """roma's convention: the least-squares R, t with y ~ R x + t (batched Kabsch with the det
    fix).  The reference calls it as (target, source), i.e. it fits the world->camera map
    source ~ R target + t, which ``cam2world`` then inverts.  x, y [B,M,3] -> R [B,3,3], t [B,3]."""
    if F.data_parallel_group is not None:
        # the rows of every image's list are sharded over data-parallel ranks: fit the WHOLE list (15 sums per image
        # all-reduced), not this rank's shard
        group = None if F.data_parallel_group is True else F.data_parallel_group
        return F.kabsch_sharded(x, y, group)
    return F.kabsch(x, y)              # one kernel, no SVD library call, capturable (csrc/kabsch.cu); CPU tensors raise
