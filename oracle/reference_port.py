"""CPU oracle for the NeRF ray-render hot path (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

A functional, state-free restatement in plain PyTorch (CPU, fp32 by default, fp64 on request)
of the arithmetic of the reference ``sfchng/neural_invertible_warp`` hot path.  Every function
names the reference lines it follows.  Parity status: the reference ships *no* tests or golden
vectors for this path, so this port is pinned against outputs of the real
reference executed in the build container (``oracle/make_golden.py`` -> ``tests/golden/*.pt``;
``tests/test_oracle_golden.py`` re-checks every fixture, and re-runs the live reference when
/root/reference is present).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import this module.  The product package never does: it fails loudly when
its CUDA library is missing instead of falling back to this code.

Parameters are passed as flat ``dict[str, Tensor]`` using the reference's state_dict key names
(``mlp_feat.{i}.weight`` ..., ``lin{b}_a_0.weight_g`` ...), so a reference checkpoint can be fed
directly.
"""
import math

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# camera algebra  (camera.py)
# --------------------------------------------------------------------------------------


def pixel_centers(H, W, dtype=torch.float32):
    """Pixel-centre grid [(x+.5, y+.5)] in row-major order.  camera.py:430-433."""
    ys = torch.arange(H, dtype=dtype) + 0.5
    xs = torch.arange(W, dtype=dtype) + 0.5
    yy, xx = torch.meshgrid(ys, xs, indexing="ij")
    return torch.stack([xx, yy], dim=-1).reshape(-1, 2)


def homogeneous(X):
    """camera.py:330-333."""
    return torch.cat([X, torch.ones_like(X[..., :1])], dim=-1)


def invert_pose(pose):
    """[R|t] -> [R^T | -R^T t].  camera.py:89-95 (use_inverse=False)."""
    R, t = pose[..., :3], pose[..., 3:]
    Rt = R.transpose(-1, -2)
    return torch.cat([Rt, -Rt @ t], dim=-1)


def cam_to_world(X, pose):
    """camera.py:343-346."""
    return homogeneous(X) @ invert_pose(pose).transpose(-1, -2)


def image_to_cam(X, intr):
    """camera.py:341-342."""
    return X @ intr.inverse().transpose(-1, -2)


def compose_pair(pose_a, pose_b):
    """pose_b o pose_a.  camera.py:105-112."""
    Ra, ta = pose_a[..., :3], pose_a[..., 3:]
    Rb, tb = pose_b[..., :3], pose_b[..., 3:]
    return torch.cat([Rb @ Ra, Rb @ ta + tb], dim=-1)


def _taylor(x, kind, nth=10):
    """Series for sin(x)/x ('A'), (1-cos x)/x^2 ('B'), (x-sin x)/x^3 ('C').  camera.py:249-272."""
    out = torch.zeros_like(x)
    denom = 1.0
    for i in range(nth + 1):
        if kind == "A":
            if i > 0:
                denom *= (2 * i) * (2 * i + 1)
        elif kind == "B":
            denom *= (2 * i + 1) * (2 * i + 2)
        else:
            denom *= (2 * i + 2) * (2 * i + 3)
        out = out + (-1) ** i * x ** (2 * i) / denom
    return out


def skew(w):
    """camera.py:241-247."""
    w0, w1, w2 = w.unbind(dim=-1)
    O = torch.zeros_like(w0)
    return torch.stack([torch.stack([O, -w2, w1], dim=-1),
                        torch.stack([w2, O, -w0], dim=-1),
                        torch.stack([-w1, w0, O], dim=-1)], dim=-2)


def se3_to_SE3(wu):
    """camera.py:215-226."""
    w, u = wu.split([3, 3], dim=-1)
    wx = skew(w)
    theta = w.norm(dim=-1)[..., None, None]
    eye = torch.eye(3, dtype=wu.dtype)
    A, B, C = _taylor(theta, "A"), _taylor(theta, "B"), _taylor(theta, "C")
    R = eye + A * wx + B * wx @ wx
    V = eye + B * wx + C * wx @ wx
    return torch.cat([R, V @ u[..., None]], dim=-1)


def center_and_ray(H, W, pose, intr, ray_idx=None):
    """World-frame camera centre and (un-normalised) ray per pixel.

    camera.py:419-443, followed by the sub-selection of model/nerf.py:298-300.
    """
    B = pose.shape[0]
    grid = pixel_centers(H, W, pose.dtype).repeat(B, 1, 1)
    grid_cam = image_to_cam(homogeneous(grid), intr)
    center_cam = torch.zeros_like(grid_cam)
    grid_w = cam_to_world(grid_cam, pose)
    center_w = cam_to_world(center_cam, pose)
    ray = grid_w - center_w
    if ray_idx is not None:
        center_w, ray = center_w[:, ray_idx], ray[:, ray_idx]
    return center_w, ray


def unwarped_center_and_grid(H, W, intr, ray_idx=None, pose_init=None):
    """Camera-frame (or initial-pose world-frame) centre and pixel grid.  camera.py:359-390."""
    B = intr.shape[0]
    grid = pixel_centers(H, W, intr.dtype).repeat(B, 1, 1)
    grid_3D = image_to_cam(homogeneous(grid), intr)
    center_3D = torch.zeros_like(grid_3D)
    if pose_init is not None:
        grid_3D = cam_to_world(grid_3D, pose_init)
        center_3D = cam_to_world(center_3D, pose_init)
    if ray_idx is not None:
        center_3D, grid_3D = center_3D[:, ray_idx], grid_3D[:, ray_idx]
    return center_3D, grid_3D


def to_ndc(center, ray, intr, near=1.0):
    """camera.py:523-540."""
    center = center + (near - center[..., 2:]) / ray[..., 2:] * ray
    cx, cy, cz = center.unbind(dim=-1)
    rx, ry, rz = ray.unbind(dim=-1)
    sx = (intr[:, 0, 0] / intr[:, 0, 2])[:, None]
    sy = (intr[:, 1, 1] / intr[:, 1, 2])[:, None]
    c = torch.stack([sx * (cx / cz), sy * (cy / cz), 1 - 2 * near / cz], dim=-1)
    r = torch.stack([sx * (rx / rz - cx / cz), sy * (ry / rz - cy / cz), 2 * near / cz], dim=-1)
    return c, r


# --------------------------------------------------------------------------------------
# depth sampling  (model/nerf.py:334-365)
# --------------------------------------------------------------------------------------


def stratified_depth(u, n_intvs, depth_range, param):
    """u: uniforms [B,R,N,1] (or the scalar 0.5 when not stratified).  model/nerf.py:334-344."""
    dmin, dmax = depth_range
    k = torch.arange(n_intvs, dtype=torch.float32)[None, None, :, None]
    s = (u + k) if torch.is_tensor(u) else (k + u)
    d = s / n_intvs * (dmax - dmin) + dmin
    if param == "inverse":
        d = 1 / (d + 1e-8)
    elif param != "metric":
        raise KeyError(param)
    return d


def pdf_depth(pdf, n_intvs, n_fine, depth_range, return_idx=False):
    """Inverse-CDF resampling.  pdf [B,R,N] -> depth [B,R,Nf,1].  model/nerf.py:346-365.

    Note torch.cumsum on CPU fp32 accumulates in higher precision and rounds each output
    (SURVEY.md fact 10); this function inherits that behaviour from torch itself.
    """
    dmin, dmax = depth_range
    cdf = torch.cat([torch.zeros_like(pdf[..., :1]), pdf.cumsum(dim=-1)], dim=-1)
    edges = torch.linspace(0, 1, n_fine + 1)
    mid = 0.5 * (edges[:-1] + edges[1:])
    u = mid.repeat(*cdf.shape[:-1], 1)
    idx = torch.searchsorted(cdf, u, right=True)
    bins = torch.linspace(dmin, dmax, n_intvs + 1).repeat(*cdf.shape[:-1], 1)
    lo_i, hi_i = (idx - 1).clamp(min=0), idx.clamp(max=n_intvs)
    d_lo, d_hi = bins.gather(2, lo_i), bins.gather(2, hi_i)
    c_lo, c_hi = cdf.gather(2, lo_i), cdf.gather(2, hi_i)
    t = (u - c_lo) / (c_hi - c_lo + 1e-8)
    d = (d_lo + t * (d_hi - d_lo))[..., None]
    return (d, idx) if return_idx else d


def merge_depth(coarse, fine):
    """cat + ascending sort.  model/nerf.py:313-315."""
    return torch.cat([coarse, fine], dim=2).sort(dim=2).values


# --------------------------------------------------------------------------------------
# positional encoding + MLP  (model/nerf.py:416-483, model/barf.py:256-268)
# --------------------------------------------------------------------------------------


def fourier_features(x, L):
    """[..., C] -> [..., 2*C*L] laid out per coordinate as [sin k=0..L-1, cos k=0..L-1].
    model/nerf.py:476-483."""
    freq = 2 ** torch.arange(L, dtype=torch.float32) * math.pi
    spec = x[..., None] * freq.to(x.dtype)
    enc = torch.stack([spec.sin(), spec.cos()], dim=-2)
    return enc.reshape(*x.shape[:-1], -1)


def c2f_band_weights(progress, c2f, L):
    """BARF coarse-to-fine band weights.  model/barf.py:260-264.  ``None`` c2f -> all ones."""
    if c2f is None:
        return torch.ones(L, dtype=torch.float32)
    start, end = c2f
    alpha = (torch.as_tensor(progress, dtype=torch.float32) - start) / (end - start) * L
    k = torch.arange(L, dtype=torch.float32)
    return (1 - ((alpha - k).clamp(min=0, max=1) * math.pi).cos()) / 2


def barf_encoding(x, L, progress, c2f):
    """Fourier features with the band weights applied (model/barf.py:256-268) and the raw
    coordinates prepended un-weighted (model/nerf.py:419)."""
    enc = fourier_features(x, L)
    if c2f is not None:
        w = c2f_band_weights(progress, c2f, L).to(x.dtype)
        enc = (enc.reshape(-1, L) * w).reshape(enc.shape)
    return torch.cat([x, enc], dim=-1)


def nerf_mlp(p, points, ray_unit, *, L_3D=10, L_view=4, skip=(4,), progress=0.0, c2f=None,
             density_activ="softplus", prefix=""):
    """8x256 feature MLP with skip, density head, view-dependent RGB head.
    model/nerf.py:416-447 (ReLU after *every* mlp_feat layer, softplus on the pre-ReLU column 0
    of the last one, sigmoid on RGB)."""
    n_feat = len([k for k in p if k.startswith(prefix + "mlp_feat.") and k.endswith(".weight")])
    n_rgb = len([k for k in p if k.startswith(prefix + "mlp_rgb.") and k.endswith(".weight")])
    enc = barf_encoding(points, L_3D, progress, c2f)
    h = enc
    density = None
    for li in range(n_feat):
        if li in skip:
            h = torch.cat([h, enc], dim=-1)
        h = F.linear(h, p[f"{prefix}mlp_feat.{li}.weight"], p[f"{prefix}mlp_feat.{li}.bias"])
        if li == n_feat - 1:
            density = getattr(F, density_activ)(h[..., 0])
            h = h[..., 1:]
        h = F.relu(h)
    if ray_unit is not None:
        h = torch.cat([h, barf_encoding(ray_unit, L_view, progress, c2f)], dim=-1)
    for li in range(n_rgb):
        h = F.linear(h, p[f"{prefix}mlp_rgb.{li}.weight"], p[f"{prefix}mlp_rgb.{li}.bias"])
        if li != n_rgb - 1:
            h = F.relu(h)
    return h.sigmoid(), density


def sample_points(center, ray, depth):
    """x = c + d v and the unit view direction.  camera.py:517-521, model/nerf.py:449-456."""
    pts = center[:, :, None] + ray[:, :, None] * depth
    unit = F.normalize(ray, dim=-1)[:, :, None, :].expand_as(pts)
    return pts, unit


# --------------------------------------------------------------------------------------
# volume compositing  (model/nerf.py:458-474)
# --------------------------------------------------------------------------------------


def composite(ray, rgb_s, sigma_s, depth_s, bgcolor=None):
    """ray [B,R,3], rgb_s [B,R,N,3], sigma_s [B,R,N], depth_s [B,R,N,1]
    -> rgb [B,R,3], depth [B,R,1], opacity [B,R,1], prob [B,R,N,1]."""
    length = ray.norm(dim=-1, keepdim=True)
    d = depth_s[..., 0]
    intv = torch.cat([d[..., 1:] - d[..., :-1], torch.full_like(d[..., :1], 1e10)], dim=2)
    sd = sigma_s * (intv * length)
    alpha = 1 - (-sd).exp()
    shifted = torch.cat([torch.zeros_like(sd[..., :1]), sd[..., :-1]], dim=2)
    T = (-shifted.cumsum(dim=2)).exp()
    prob = (T * alpha)[..., None]
    depth = (depth_s * prob).sum(dim=2)
    rgb = (rgb_s * prob).sum(dim=2)
    opacity = prob.sum(dim=2)
    if bgcolor is not None:
        rgb = rgb + bgcolor * (1 - opacity)
    return rgb, depth, opacity, prob


# --------------------------------------------------------------------------------------
# NVP invertible warp  (model/nvp/nvp_ndr.py:229-468, model/nvp/embedder.py)
# --------------------------------------------------------------------------------------

# focus axis / the two other axes for coupling block b (form 0, i.e. blocks 0..2): z, y, x.
# model/nvp/nvp_ndr.py:389-399
NVP_AXES = {0: (2, (0, 1)), 1: (1, (0, 2)), 2: (0, (1, 2))}


def nvp_embed(x, alpha_ratio, n_freq):
    """Annealed Fourier embedding *including the reference's dim-1 quirk*.

    x is [B, P, 1, d].  Output order [x, sin(f0 x), cos(f0 x), sin(f1 x), ...], f_k = 2^k pi
    (model/nvp/embedder.py:19-34).  The annealing loop slices ``output[:, a:b]`` of a 4-D
    tensor, i.e. along the *point* axis: for band i the points with index in
    [(2i+1)d, (2i+3)d) get every channel multiplied by w_i (embedder.py:41-50).
    """
    d = x.shape[-1]
    parts = [x]
    for k in range(n_freq):
        f = (2.0 ** k) * math.pi
        parts += [torch.sin(x * f), torch.cos(x * f)]
    out = torch.cat(parts, dim=-1)
    scale = torch.ones(x.shape[1], dtype=x.dtype)
    for i in range(n_freq):
        w = (1.0 - math.cos(math.pi * max(min(alpha_ratio * n_freq - i, 1.0), 0.0))) * 0.5
        scale[(2 * i + 1) * d:(2 * i + 3) * d] *= w
    return out * scale[None, :, None, None]


def _wn_linear(p, name, x):
    """weight_norm'd Linear: w = g * v / ||v||_row.  model/nvp/nvp_ndr.py:291-292 (torch's
    legacy nn.utils.weight_norm, dim=0)."""
    if name + ".weight_g" in p:
        v, g = p[name + ".weight_v"], p[name + ".weight_g"]
        w = v * (g / v.norm(dim=1, keepdim=True))
    else:
        w = p[name + ".weight"]
    return F.linear(x, w, p[name + ".bias"])


def nvp_warp(p, code, pts, alpha_ratio, *, n_blocks=3, n_freq=6, beta=100.0, prefix=""):
    """DeformNetwork.forward for the configuration the target models instantiate
    (n_layers=1, skip_in=[], softplus(beta=100)).  code [B,D], pts [B,P,1,3] -> [B,P,1,3].
    model/nvp/nvp_ndr.py:365-468."""
    q = {k[len(prefix):]: v for k, v in p.items() if k.startswith(prefix)}
    x = pts
    B, P = pts.shape[:2]
    for b in range(n_blocks):
        foc, oth = NVP_AXES[b % 3]
        cb = F.linear(code, q[f"lin{b}_c.weight"], q[f"lin{b}_c.bias"]) + code
        cb = cb[:, None, None, :].expand(B, P, 1, cb.shape[-1])
        x_f, x_o = x[..., [foc]], x[..., list(oth)]
        # part a: the two "other" axes predict a shift of the focus axis
        h = torch.cat([nvp_embed(x_o, alpha_ratio, n_freq), cb], dim=-1)
        h = F.softplus(_wn_linear(q, f"lin{b}_a_0", h), beta=beta)
        x_f = x_f - _wn_linear(q, f"lin{b}_a_1", h)
        # part b: the shifted focus axis predicts a 2-D rotation+translation of the others
        h = torch.cat([nvp_embed(x_f, alpha_ratio, n_freq), cb], dim=-1)
        h = F.softplus(_wn_linear(q, f"lin{b}_b_0", h), beta=beta)
        out = _wn_linear(q, f"lin{b}_b_1", h)
        th, tr = out[..., 0], out[..., 1:]
        c, s = th.cos(), th.sin()
        y = x_o - tr
        # euler2rot_2dinv as assembled at nvp_ndr.py:166-174 is [[c, s], [-s, c]]
        y0 = c * y[..., 0] + s * y[..., 1]
        y1 = -s * y[..., 0] + c * y[..., 1]
        cols = [None, None, None]
        cols[foc] = x_f[..., 0]
        cols[oth[0]], cols[oth[1]] = y0, y1
        x = torch.stack(cols, dim=-1)
    return x


# --------------------------------------------------------------------------------------
# whole-path drivers (model/nerf.py:293-319, model/nerf_inn_llff.py:493-612)
# --------------------------------------------------------------------------------------


def render_rays(nerf_p, center, ray, u, cfg, *, progress, nerf_fine_p=None, depth_range=None,
                progress_fine=None):
    """center/ray [B,R,3], u [B,R,N,1] -> dict(rgb, depth, opacity[, *_fine], prob, depth_samples).
    cfg keys: N, Nf (or None), range, param, L_3D, L_view, skip, c2f, density_activ."""
    rng = depth_range if depth_range is not None else cfg["range"]
    d = stratified_depth(u, cfg["N"], rng, cfg["param"])
    kw = dict(L_3D=cfg["L_3D"], L_view=cfg["L_view"], skip=cfg["skip"], c2f=cfg["c2f"],
              density_activ=cfg.get("density_activ", "softplus"))
    pts, unit = sample_points(center, ray, d)
    rgb_s, sig_s = nerf_mlp(nerf_p, pts, unit, progress=progress, **kw)
    rgb, depth, opacity, prob = composite(ray, rgb_s, sig_s, d)
    out = dict(rgb=rgb, depth=depth, opacity=opacity, prob=prob, depth_samples=d)
    if cfg.get("Nf"):
        with torch.no_grad():
            # NB fine bins use cfg["range"] even when coarse used depth_range (nerf_inn_dtu.py:549)
            fine, idx = pdf_depth(prob[..., 0], cfg["N"], cfg["Nf"], cfg["range"], return_idx=True)
            d2 = merge_depth(d, fine)
        pts, unit = sample_points(center, ray, d2)
        pf = progress if progress_fine is None else progress_fine
        rgb_s, sig_s = nerf_mlp(nerf_fine_p, pts, unit, progress=pf, **kw)
        rgb_f, depth_f, op_f, _ = composite(ray, rgb_s, sig_s, d2)
        out.update(rgb_fine=rgb_f, depth_fine=depth_f, opacity_fine=op_f, fine_idx=idx,
                   depth_samples_fine=d2)
    return out


def warped_rays(nvp_p, code, H, W, intr, ray_idx, alpha_ratio, pose_init=None, **kw):
    """Training-time ray generation of the INN models.
    model/barf_inn_llff.py:325-364 / model/pose_models/inn.py:63-93."""
    center_cam, grid_cam = unwarped_center_and_grid(H, W, intr, ray_idx, pose_init)
    pts = torch.cat([grid_cam, center_cam], dim=1).detach()[:, :, None]
    out = nvp_warp(nvp_p, code, pts, alpha_ratio, **kw)[:, :, 0]
    P = grid_cam.shape[1]
    grid_3D, center_3D = out[:, :P], out[:, P:]
    return grid_3D - center_3D, center_3D, grid_3D, grid_cam, center_cam


def kabsch(x, y):
    """roma.rigid_points_registration(x, y) (roma==1.4.1, requirements.txt:1; absent from /root/reference): the
    least-squares rigid motion y ~ R x + t by the published Kabsch/Umeyama construction -- centroids,
    M = sum (y - ym)(x - xm)^T, SVD, determinant fix.  The reference calls it as (target, source):
    model/nerf_inn_llff.py:569, model/pose_models/inn.py:100.  x, y [B,M,3] -> R [B,3,3], t [B,3]."""
    xm, ym = x.mean(dim=-2, keepdim=True), y.mean(dim=-2, keepdim=True)
    M = (y - ym).transpose(-1, -2) @ (x - xm)
    U, _, Vh = torch.linalg.svd(M)
    d = torch.det(U @ Vh)
    D = torch.diag_embed(torch.stack([torch.ones_like(d), torch.ones_like(d), d], dim=-1))
    R = U @ D @ Vh
    return R, ym.squeeze(-2) - (R @ xm.transpose(-1, -2)).squeeze(-1)


def kabsch_stats(x, y):
    """Sufficient statistics of ``kabsch`` over the given rows, [B,16] float64 = (rows, sum x, sum y, sum x_a y_c): additive
    over any partition of the rows (the data-parallel form of the fit)."""
    xd, yd = x.double(), y.double()
    n = torch.full((x.shape[0], 1), float(x.shape[1]), dtype=torch.float64)
    return torch.cat([n, xd.sum(1), yd.sum(1), (xd.transpose(1, 2) @ yd).reshape(-1, 9)], dim=1)


def kabsch_from_stats(stats):
    """``kabsch`` from (summed) ``kabsch_stats``: M = sum y x^T - n ym xm^T."""
    n = stats[:, 0]
    xm, ym = stats[:, 1:4] / n[:, None], stats[:, 4:7] / n[:, None]
    Sxy = stats[:, 7:16].reshape(-1, 3, 3)                               # [a][c] = sum x_a y_c
    M = Sxy.transpose(1, 2) - n[:, None, None] * ym[:, :, None] * xm[:, None, :]
    U, _, Vh = torch.linalg.svd(M)
    d = torch.det(U @ Vh)
    D = torch.diag_embed(torch.stack([torch.ones_like(d), torch.ones_like(d), d], dim=-1))
    R = U @ D @ Vh
    return R.float(), (ym - (R @ xm[..., None])[..., 0]).float()


def mse(pred, label):
    """model/base.py:209-211."""
    return ((pred.contiguous() - label) ** 2).mean()


def gather_pixels(image, ray_idx):
    """image [B,3,H,W] -> [B,R,3] at flat pixel indices.  model/nerf.py:278-281."""
    B = image.shape[0]
    flat = image.reshape(B, 3, -1).permute(0, 2, 1)
    return flat if ray_idx is None else flat[:, ray_idx]


# --------------------------------------------------------------------------------------
# evaluation metrics  (model/nerf.py:176-183, external/pohsun_ssim/pytorch_ssim/__init__.py:7-37, core/metrics.py:64-111)
# --------------------------------------------------------------------------------------

def psnr(rgb_map, image):
    """-10 log10 MSE.  model/nerf.py:179 (MSE_loss = mean squared error, model/base.py:209-211)."""
    return -10 * ((rgb_map - image) ** 2).mean().log10()


def ssim(img1, img2, window_size=11, sigma=1.5):
    """pytorch_ssim.ssim with size_average=True: img [B,C,H,W]."""
    import math
    g = torch.tensor([math.exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)])
    g = (g / g.sum())[:, None]
    C = img1.shape[1]
    window = (g @ g.t()).float()[None, None].expand(C, 1, window_size, window_size).contiguous()
    conv = lambda t: torch.nn.functional.conv2d(t, window, padding=window_size // 2, groups=C)
    mu1, mu2 = conv(img1), conv(img2)
    mu1_sq, mu2_sq, mu12 = mu1 ** 2, mu2 ** 2, mu1 * mu2
    s1, s2, s12 = conv(img1 * img1) - mu1_sq, conv(img2 * img2) - mu2_sq, conv(img1 * img2) - mu12
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    return (((2 * mu12 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))).mean()


def depth_error(pred, gt, valid, scale=1.0):
    """core/metrics.py:64-111 (full image, B = 1): (abs_e, rmse), the better of scaled / unscaled when scale != 1."""
    gt, pred = gt.reshape(-1)[valid.reshape(-1)], pred.reshape(-1)[valid.reshape(-1)]

    def metric(d):
        a = (gt - d).abs()
        return (a.sum() / (a.nelement() + 1e-6)).item(), torch.sqrt(((d - gt) ** 2).mean()).item()
    a, r = metric(pred)
    if scale != 1.0:
        a2, r2 = metric(pred * scale)
        a, r = min(a, a2), min(r, r2)
    return a, r
