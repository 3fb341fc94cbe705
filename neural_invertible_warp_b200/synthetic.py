"""Deterministic synthetic inputs shaped like the reference's LLFF / DTU workloads.

There is no dataset on a GPU box, so benchmarks, smoke tests and parity tests use synthetic
cameras, images and random-init weights.  All factories draw from a
private CPU ``torch.Generator`` so that the same seed yields the same tensors everywhere; the
golden fixtures under ``tests/golden`` store only the seeds of their inputs.
"""
import math

import torch


def _gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def _uniform(g, shape, bound):
    return (torch.rand(shape, generator=g) * 2 - 1) * bound


def nerf_layer_shapes(L_3D=10, L_view=4, width=256, depth=8, skip=(4,), rgb_width=128, view_dep=True):
    """(name, out, in) for every Linear of the reference NeRF (model/nerf.py:373-402)."""
    d3 = 3 + 6 * L_3D
    dv = 3 + 6 * L_view
    shapes = []
    for li in range(depth):
        k_in = d3 if li == 0 else width
        if li in skip:
            k_in += d3
        k_out = width + (1 if li == depth - 1 else 0)
        shapes.append((f"mlp_feat.{li}", k_out, k_in))
    shapes.append(("mlp_rgb.0", rgb_width, width + (dv if view_dep else 0)))
    shapes.append(("mlp_rgb.1", 3, rgb_width))
    return shapes


def nerf_params(seed, bias_bound=0.05, **arch):
    """Xavier-uniform weights in the style of the reference's ``tf_init`` (model/nerf.py:404-414:
    ReLU gain on hidden layers, gain 1 on the density row and on the RGB output layer) plus small
    non-zero biases so that the bias paths are exercised."""
    g = _gen(seed)
    relu_gain = math.sqrt(2.0)
    p = {}
    shapes = nerf_layer_shapes(**arch)
    for name, k_out, k_in in shapes:
        bound = relu_gain * math.sqrt(6.0 / (k_in + k_out))
        w = _uniform(g, (k_out, k_in), bound)
        if name == shapes[-3][0]:  # last feature layer: density row has gain 1
            w[0] = w[0] / relu_gain
        if name == "mlp_rgb.1":
            w = w / relu_gain
        p[name + ".weight"] = w
        p[name + ".bias"] = _uniform(g, (k_out,), bias_bound)
    return p


def nvp_params(seed, d_feature=128, d_hidden=128, n_freq=6, n_blocks=3, out_scale=1e-2):
    """Parameters of ``DeformNetwork`` as instantiated by the target models (n_layers=1,
    skip_in=[], weight_norm on the first Linear of each part; model/nvp/nvp_ndr.py:230-340).
    The reference zero-initialises the output layers and the code projector (the warp starts as
    the identity); here they are perturbed (N(0, out_scale^2)) so gradients are exercised."""
    g = _gen(seed)
    p = {}
    emb_a, emb_b = 2 * (1 + 2 * n_freq), 1 * (1 + 2 * n_freq)
    for b in range(n_blocks):
        for part, emb, n_out, ori in (("a", emb_a, 1, 2), ("b", emb_b, 3, 1)):
            v = torch.zeros(d_hidden, emb + d_feature)
            v[:, :ori] = torch.randn(d_hidden, ori, generator=g) * math.sqrt(2) / math.sqrt(d_hidden)
            v[:, ori:] = torch.randn(d_hidden, emb + d_feature - ori, generator=g) * 0.05
            p[f"lin{b}_{part}_0.weight_v"] = v
            p[f"lin{b}_{part}_0.weight_g"] = v.norm(dim=1, keepdim=True) * (
                1 + 0.1 * torch.randn(d_hidden, 1, generator=g))
            p[f"lin{b}_{part}_0.bias"] = torch.randn(d_hidden, generator=g) * 0.01
            p[f"lin{b}_{part}_1.weight"] = torch.randn(n_out, d_hidden, generator=g) * out_scale
            p[f"lin{b}_{part}_1.bias"] = torch.randn(n_out, generator=g) * out_scale
        p[f"lin{b}_c.weight"] = torch.randn(d_feature, d_feature, generator=g) * out_scale
        p[f"lin{b}_c.bias"] = torch.randn(d_feature, generator=g) * out_scale
    return p


def intrinsics(B, H, W, focal_over_W):
    """[[f,0,W/2],[0,f,H/2],[0,0,1]] (LLFF-shaped: f=0.81 W; DTU-shaped: f=1.8 W)."""
    f = focal_over_W * W
    K = torch.tensor([[f, 0., W / 2.], [0., f, H / 2.], [0., 0., 1.]])
    return K[None].repeat(B, 1, 1)


def _rodrigues(w):
    th = w.norm(dim=-1, keepdim=True).clamp_min(1e-12)[..., None]
    k = w / th[..., 0]
    K = torch.zeros(*w.shape[:-1], 3, 3)
    K[..., 0, 1], K[..., 0, 2] = -k[..., 2], k[..., 1]
    K[..., 1, 0], K[..., 1, 2] = k[..., 2], -k[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -k[..., 1], k[..., 0]
    return torch.eye(3) + th.sin() * K + (1 - th.cos()) * (K @ K)


def llff_poses(seed, B, noise=0.05):
    """Identity world->camera poses perturbed by N(0, noise^2) rotation vectors / translations."""
    g = _gen(seed)
    wu = torch.randn(B, 6, generator=g) * noise
    return torch.cat([_rodrigues(wu[:, :3]), wu[:, 3:, None]], dim=-1)


def dtu_poses(seed, B, radius=3.0, noise=0.15):
    """Cameras on a ring looking at the origin from distance ``radius`` (DTU-like, metric depth
    range 1.2..5.2), world->camera."""
    g = _gen(seed)
    ang = torch.linspace(0, 2 * math.pi * (1 - 1 / B), B)
    w = torch.stack([torch.zeros(B), ang, torch.zeros(B)], dim=-1) + torch.randn(B, 3, generator=g) * noise
    R = _rodrigues(w)
    t = torch.tensor([0., 0., radius]).expand(B, 3) + torch.randn(B, 3, generator=g) * noise * 0.2
    return torch.cat([R, t[..., None]], dim=-1)


def images(seed, B, H, W):
    return torch.rand(B, 3, H, W, generator=_gen(seed))


def latent_codes(seed, B, dim=128, scale=0.1):
    return torch.randn(B, dim, generator=_gen(seed)) * scale


def uniforms(seed, *shape):
    return torch.rand(*shape, generator=_gen(seed))


def ray_indices(seed, H, W, n):
    return torch.randperm(H * W, generator=_gen(seed))[:n]
