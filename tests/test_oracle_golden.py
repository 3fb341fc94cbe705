"""The CPU oracle (oracle/reference_port.py) must reproduce every golden vector minted from the
real reference (oracle/make_golden.py).  This is what pins the oracle."""
import math

import pytest
import torch

from neural_invertible_warp_b200 import synthetic as syn
from neural_invertible_warp_b200 import config as cfgmod
from oracle import reference_port as ora

TOL = dict(rtol=2e-5, atol=2e-6)


def close(a, b, **kw):
    kw = {**TOL, **kw}
    torch.testing.assert_close(a, b, **kw)


def rel_l2(a, b, tol):
    """Per-tensor relative L2 error: the gradient metric of this repo."""
    err = (a.double() - b.double()).norm().item() / max(b.double().norm().item(), 1e-30)
    assert err <= tol, "rel-L2 %.3e > %.1e" % (err, tol)


def digest_close(named_grads, digest, rtol=2e-3):
    for k, d in digest.items():
        g = named_grads[k].detach().double().flatten()
        assert g.numel() == d["numel"], k
        scale = max(d["l2"], 1e-12)
        assert abs(g.norm().item() - d["l2"]) <= rtol * scale, (k, g.norm().item(), d["l2"])
        assert abs(g.sum().item() - d["sum"]) <= rtol * max(d["abssum"], 1e-12), k
        torch.testing.assert_close(g[:8].float(), d["head"], rtol=2.5 * rtol,
                                   atol=2.5 * rtol * max(d["head"].abs().max().item(),
                                                         3 * d["l2"] / math.sqrt(d["numel"])) + 1e-12)


def test_camera(golden):
    g = golden("camera")
    H, W = g["H"], g["W"]
    close(syn.llff_poses(g["pose_seed"], g["B"], noise=0.2)[..., :3], g["pose"][..., :3])
    c, r = ora.center_and_ray(H, W, g["pose"], g["intr"])
    close(c, g["center"]); close(r, g["ray"])
    cc, gc = ora.unwarped_center_and_grid(H, W, g["intr"], g["ray_idx"])
    close(cc, g["center_cam"]); close(gc, g["grid_cam"])
    cw, gw = ora.unwarped_center_and_grid(H, W, g["intr"], g["ray_idx"], g["pose"])
    close(cw, g["center_w"]); close(gw, g["grid_w"])
    SE3 = ora.se3_to_SE3(g["wu"])
    close(SE3, g["SE3"])
    close(ora.compose_pair(SE3[:3], g["pose"]), g["composed"])
    cn, rn = ora.to_ndc(c[:, g["ray_idx"]], r[:, g["ray_idx"]], g["intr"])
    close(cn, g["center_ndc"]); close(rn, g["ray_ndc"])
    pose = g["pose"].clone().requires_grad_(True)
    c2, r2 = ora.center_and_ray(H, W, pose, g["intr"])
    wc = syn.uniforms(g["wc_seed"], *c2.shape) - 0.5
    wr = syn.uniforms(g["wr_seed"], *r2.shape) - 0.5
    ((c2 * wc).sum() + (r2 * wr).sum()).backward()
    close(pose.grad, g["pose_grad"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("tag", ["llff", "dtu"])
def test_sampler(golden, tag):
    g = golden("sampler")[tag]
    d = ora.stratified_depth(g["u"], g["N"], g["range"], g["param"])
    assert torch.equal(d, g["depth"])
    fine, idx = ora.pdf_depth(g["pdf"], g["N"], g["Nf"], g["range"], return_idx=True)
    assert torch.equal(idx, g["idx"])
    assert torch.equal(fine, g["fine"])
    assert torch.equal(ora.merge_depth(d, fine), g["merged"])


def test_nerf_mlp(golden):
    g = golden("nerf_mlp")
    p = syn.nerf_params(g["param_seed"])
    for prog, case in g["cases"].items():
        c2f = None if prog == "no_c2f" else g["c2f"]
        pr = 1.0 if prog == "no_c2f" else prog
        q = {k: v.clone().requires_grad_(True) for k, v in p.items()}
        pts = g["points"].clone().requires_grad_(True)
        unit = g["ray_unit"].clone().requires_grad_(True)
        rgb, dens = ora.nerf_mlp(q, pts, unit, progress=pr, c2f=c2f)
        close(rgb, case["rgb"]); close(dens, case["density"], rtol=1e-4, atol=1e-5)
        if "grads" in case:
            wr = syn.uniforms(g["wr_seed"], *rgb.shape) - 0.5
            wd = syn.uniforms(g["wd_seed"], *dens.shape) - 0.5
            ((rgb * wr).sum() + (dens * wd).sum()).backward()
            close(pts.grad, case["d_points"], rtol=2e-3, atol=1e-4)
            close(unit.grad, case["d_unit"], rtol=2e-3, atol=1e-5)
            digest_close({k: v.grad for k, v in q.items()}, case["grads"])
            enc = ora.barf_encoding(g["points"], 10, pr, c2f)[..., 3:]
            close(enc[0, 0, :2], case["enc_head"])


def test_composite(golden):
    g = golden("composite")
    ins = [g[k].clone().requires_grad_(True) for k in ("ray", "rgb_samples", "sigma")]
    rgb, d, op, prob = ora.composite(ins[0], ins[1], ins[2], g["depth_samples"])
    close(rgb, g["rgb"]); close(d, g["depth"]); close(op, g["opacity"]); close(prob, g["prob"])
    w = [syn.uniforms(s, *t.shape) - 0.5 for s, t in zip(g["w_seeds"], (rgb, d, op))]
    ((rgb * w[0]).sum() + (d * w[1]).sum() + (op * w[2]).sum()).backward()
    close(ins[0].grad, g["d_ray"], rtol=1e-4, atol=1e-5)
    close(ins[1].grad, g["d_rgb_samples"])
    close(ins[2].grad, g["d_sigma"], rtol=1e-4, atol=1e-5)


def test_nvp(golden):
    g = golden("nvp")
    p = syn.nvp_params(g["param_seed"])
    code = syn.latent_codes(g["code_seed"], 2)
    for alpha, case in g["cases"].items():
        q = {k: v.clone().requires_grad_(True) for k, v in p.items()}
        cg = code.clone().requires_grad_(True)
        out = ora.nvp_warp(q, cg, g["pts"], alpha)
        close(out, case["out"], rtol=1e-5, atol=2e-6)
        w = syn.uniforms(g["w_seed"], *out.shape) - 0.5
        (out * w).sum().backward()
        rel_l2(cg.grad, case["d_code"], 2e-3)
        close(q["lin0_a_1.weight"].grad, case["d_a1w"], rtol=1e-3, atol=1e-6)
        close(q["lin2_b_1.weight"].grad, case["d_b1w"], rtol=1e-3, atol=1e-6)
        digest_close({k: v.grad for k, v in q.items()}, case["grads"])


def _cfg(N, Nf, rng, param, c2f=(0.1, 0.5)):
    return dict(N=N, Nf=Nf, range=rng, param=param, L_3D=10, L_view=4, skip=(4,), c2f=list(c2f))


def test_graph_barf(golden):
    """barf.Graph.forward(mode='train') + MSE + backward (model/nerf.py:251-288, barf.py:217-229)."""
    g = golden("graph_barf")
    B, H, W = g["B"], g["H"], g["W"]
    p = {k: v.requires_grad_(True) for k, v in syn.nerf_params(g["param_seed"]).items()}
    se3 = g["se3"].clone().requires_grad_(True)
    intr = syn.intrinsics(B, H, W, 0.81)
    image = syn.images(g["var_seed"], B, H, W)
    eye = torch.eye(3, 4)
    pose = ora.compose_pair(ora.se3_to_SE3(se3), eye)
    center, ray = ora.center_and_ray(H, W, pose, intr, g["ray_idx"])
    out = ora.render_rays(p, center, ray, g["u"], _cfg(g["N"], None, [1, 0], "inverse"), progress=g["progress"])
    close(out["rgb"], g["rgb"]); close(out["depth"], g["depth"], rtol=1e-4, atol=1e-4); close(out["opacity"], g["opacity"])
    loss = ora.mse(out["rgb"], ora.gather_pixels(image, g["ray_idx"]))
    close(loss, g["loss"])
    loss.backward()
    rel_l2(se3.grad, g["d_se3"], 5e-3)
    digest_close({k: v.grad for k, v in p.items()}, g["grads"])


@pytest.mark.parametrize("tag", ["p16", "p40"])
def test_graph_inn_llff(golden, tag):
    """barf_inn_llff train step: NVP-warped rays -> render_local -> MSE (+ global alignment)."""
    g = golden("graph_inn_llff")[tag]
    B, H, W = g["B"], g["H"], g["W"]
    p = {k: v.requires_grad_(True) for k, v in syn.nerf_params(g["nerf_seed"]).items()}
    q = {k: v.requires_grad_(True) for k, v in syn.nvp_params(g["nvp_seed"]).items()}
    code = syn.latent_codes(g["code_seed"], B).requires_grad_(True)
    intr = syn.intrinsics(B, H, W, 0.81)
    image = syn.images(g["var_seed"], B, H, W)
    alpha = max(min(g["iter"] / 100000, 1), 0)
    assert alpha == g["alpha_ratio"]
    ray, center, grid_3D, grid_cam, center_cam = ora.warped_rays(q, code, H, W, intr, g["ray_idx"], alpha)
    close(grid_3D, g["grid_3D"]); close(center, g["center"])
    out = ora.render_rays(p, center, ray, g["u"], _cfg(g["N"], None, [1, 0], "inverse"), progress=g["progress"])
    close(out["rgb"], g["rgb"]); close(out["opacity"], g["opacity"]); close(out["depth"], g["depth"], rtol=1e-4, atol=1e-4)
    l_render = ora.mse(out["rgb"], ora.gather_pixels(image, g["ray_idx"]))
    close(l_render, g["loss_render"])
    # global alignment (model/nerf_inn_llff.py:563-572): Kabsch fit of the warped points
    from oracle.ref_shim import _kabsch
    source = torch.cat([grid_cam, center_cam], dim=1)
    target = torch.cat([grid_3D, center], dim=1)
    R, t = _kabsch(target, source)
    svd = torch.cat([R, t[..., None]], dim=-1).detach()
    close(svd.reshape(B, 12), g["global_rigid"], rtol=1e-4, atol=1e-5)
    l_ga = ora.mse(target, ora.cam_to_world(source, svd))
    close(l_ga, g["loss_global_alignment"], rtol=1e-4, atol=1e-7)
    loss = l_render + 10 ** 2 * l_ga
    close(loss, g["loss"], rtol=1e-4, atol=1e-6)
    loss.backward()
    rel_l2(code.grad, g["d_code"], 5e-3)
    digest_close({k: v.grad for k, v in q.items()}, g["nvp_grads"], rtol=5e-2)  # fp32 noise floor of the ill-conditioned input path (fp64 check: 1-2%)
    digest_close({k: v.grad for k, v in p.items()}, g["grads"])


def test_graph_inn_dtu(golden):
    """barf_inn_dtu train step with hierarchical sampling (64... here 16 coarse + 32 fine)."""
    g = golden("graph_inn_dtu")
    B, H, W = g["B"], g["H"], g["W"]
    p = {k: v.requires_grad_(True) for k, v in syn.nerf_params(g["nerf_seed"]).items()}
    pf = {k: v.requires_grad_(True) for k, v in syn.nerf_params(g["nerf_fine_seed"]).items()}
    q = {k: v.requires_grad_(True) for k, v in syn.nvp_params(g["nvp_seed"]).items()}
    code = syn.latent_codes(g["code_seed"], B).requires_grad_(True)
    intr = syn.intrinsics(B, H, W, 1.8)
    image = syn.images(g["var_seed"], B, H, W)
    pose0 = syn.dtu_poses(g["var_seed"] + 1, B)
    alpha = max(min(g["iter"] / 100000, 1), 0)
    ray, center, grid_3D, grid_init, center_init = ora.warped_rays(q, code, H, W, intr, g["ray_idx"], alpha, pose_init=pose0)
    out = ora.render_rays(p, center, ray, g["u"], _cfg(g["N"], g["Nf"], [1.2, 5.2], "metric"), progress=g["progress"],
                          nerf_fine_p=pf, depth_range=[1.2, 5.2])
    close(out["rgb"], g["rgb"]); close(out["opacity"], g["opacity"]); close(out["depth"], g["depth"], rtol=1e-4, atol=1e-5)
    close(out["rgb_fine"], g["rgb_fine"]); close(out["opacity_fine"], g["opacity_fine"])
    close(out["depth_fine"], g["depth_fine"], rtol=1e-4, atol=1e-5)
    tgt = ora.gather_pixels(image, g["ray_idx"])
    loss = ora.mse(out["rgb"], tgt) + ora.mse(out["rgb_fine"], tgt)
    close(loss, g["loss"])
    loss.backward()
    rel_l2(code.grad, g["d_code"], 5e-3)
    digest_close({k: v.grad for k, v in q.items()}, g["nvp_grads"], rtol=5e-2)  # fp32 noise floor of the ill-conditioned input path (fp64 check: 1-2%)
    digest_close({k: v.grad for k, v in p.items()}, g["grads"])
    digest_close({k: v.grad for k, v in pf.items()}, g["grads_fine"])


def test_builtin_options_match_reference_yaml(golden):
    g = golden("options")
    for name, ref in g.items():
        mine = cfgmod.builtin_options(name)
        assert list(mine.data.image_size) == ref["image_size"]
        for sect in ("arch", "nerf", "camera", "loss_weight"):
            for k, v in ref[sect].items():
                if sect == "camera" and k not in mine[sect]:
                    continue
                assert mine[sect][k] == v, (name, sect, k, mine[sect].get(k), v)
        assert mine.max_iter == ref["max_iter"]
        if "inn" in ref:
            for k, v in ref["inn"]["real_nvp"].items():
                assert mine.inn.real_nvp[k] == v, (name, k)
            assert mine.inn.actfn == ref["inn"]["actfn"]
        if "warp_latent" in ref:
            assert mine.warp_latent.embed_dim == ref["warp_latent"]["embed_dim"]
            assert mine.warp_latent.enc_type == ref["warp_latent"]["enc_type"]


def test_metrics_oracle_matches_reference(golden):
    """the oracle's PSNR / SSIM / depth error against the reference's own pytorch_ssim and core/metrics.py."""
    g = golden("metrics")
    H, W, B = g["H"], g["W"], g["B"]
    image = syn.images(g["image_seed"], B, H, W)
    for b in range(B):
        rgb_map = g["rgb"][b:b + 1].view(-1, H, W, 3).permute(0, 3, 1, 2)
        assert abs(ora.psnr(rgb_map, image[b:b + 1]).item() - g["psnr"][b]) < 1e-5
        assert abs(ora.ssim(rgb_map, image[b:b + 1]).item() - g["ssim"][b]) < 1e-6
    a, r = ora.depth_error(g["depth"], g["depth_gt"], g["valid"], 1.0)
    assert abs(a - g["depth_err"][0]) < 1e-6 and abs(r - g["depth_err"][1]) < 1e-6
    a, r = ora.depth_error(g["depth"], g["depth_gt"], g["valid"], g["depth_scale"])
    assert abs(a - g["depth_err_scaled"][0]) < 1e-6 and abs(r - g["depth_err_scaled"][1]) < 1e-6
