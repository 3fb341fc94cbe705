"""GPU parity: every C-ABI operator against the CPU oracle (oracle/reference_port.py) and against
the golden vectors minted from the executed reference (tests/golden)."""
import math

import pytest
import torch

from neural_invertible_warp_b200 import synthetic as syn
from oracle import reference_port as ora

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def F():
    from neural_invertible_warp_b200 import functional
    return functional


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def flat_params(p):
    return torch.cat([p[k].reshape(-1) for k in nerf_keys()])


def nerf_keys():
    keys = []
    for i in range(8):
        keys += [f"mlp_feat.{i}.weight", f"mlp_feat.{i}.bias"]
    for i in range(2):
        keys += [f"mlp_rgb.{i}.weight", f"mlp_rgb.{i}.bias"]
    return keys


def unflatten(flat, like):
    out, off = {}, 0
    for k in nerf_keys():
        n = like[k].numel()
        out[k] = flat[off:off + n].view_as(like[k])
        off += n
    return out


# ------------------------------------------------------------------------------------------

def test_raygen_pose_golden(F, golden):
    g = golden("camera")
    H, W = g["H"], g["W"]
    pose = g["pose"].to(DEV).requires_grad_(True)
    intr = g["intr"].to(DEV)
    c, r = F.raygen_pose(pose, intr, H, W)
    torch.testing.assert_close(c.cpu(), g["center"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(r.cpu(), g["ray"], rtol=1e-5, atol=1e-5)
    wc = (syn.uniforms(g["wc_seed"], *c.shape) - 0.5).to(DEV)
    wr = (syn.uniforms(g["wr_seed"], *r.shape) - 0.5).to(DEV)
    ((c * wc).sum() + (r * wr).sum()).backward()
    assert rel_l2(pose.grad, g["pose_grad"]) < 1e-4
    c2, r2 = F.raygen_pose(pose.detach(), intr, H, W, ray_idx=g["ray_idx"].to(DEV))
    torch.testing.assert_close(c2.cpu(), g["center"][:, g["ray_idx"]], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(r2.cpu(), g["ray"][:, g["ray_idx"]], rtol=1e-5, atol=1e-5)
    c3, r3 = F.raygen_pose(pose.detach(), intr, H, W, idx_start=7, num=13)
    torch.testing.assert_close(r3.cpu(), g["ray"][:, 7:20], rtol=1e-5, atol=1e-5)


def test_raygen_unwarped_golden(F, golden):
    g = golden("camera")
    H, W = g["H"], g["W"]
    idx = g["ray_idx"].to(DEV)
    P = idx.numel()
    pts = F.raygen_unwarped(g["intr"].to(DEV), H, W, ray_idx=idx).cpu()
    torch.testing.assert_close(pts[:, :P], g["grid_cam"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(pts[:, P:], g["center_cam"], rtol=0, atol=0)
    pts = F.raygen_unwarped(g["intr"].to(DEV), H, W, ray_idx=idx, pose_init=g["pose"].to(DEV)).cpu()
    torch.testing.assert_close(pts[:, :P], g["grid_w"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(pts[:, P:], g["center_w"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("tag", ["llff", "dtu"])
def test_sampler_bit_exact_golden(F, golden, tag):
    g = golden("sampler")[tag]
    B, R, N = g["u"].shape[:3]
    d = F.sample_stratified(g["u"].to(DEV).reshape(-1), B * R, N, g["range"], g["param"])
    assert torch.equal(d.cpu().view(B, R, N, 1), g["depth"]), "stratified depths must be bit-exact"
    fine, idx, merged = F.sample_pdf_merge(g["pdf"].to(DEV).view(B * R, N), d, g["Nf"], g["range"], want_idx=True)
    assert torch.equal(idx.cpu().view(B, R, -1), g["idx"]), "importance-sampling bins must be bit-exact"
    assert torch.equal(fine.cpu().view(B, R, -1, 1), g["fine"])
    assert torch.equal(merged.cpu().view(B, R, -1, 1), g["merged"])


def test_sampler_unstratified_and_ragged(F):
    d = F.sample_stratified(None, 5, 7, [2.0, 6.0], "metric", device=DEV)
    ref = ora.stratified_depth(0.5, 7, [2.0, 6.0], "metric").expand(1, 5, 7, 1)
    assert torch.equal(d.cpu().view(1, 5, 7, 1), ref.contiguous())


@pytest.mark.parametrize("R,N,Nf", [(1000, 64, 128), (257, 128, 64), (33, 48, 80)])
def test_pdf_sampler_vs_oracle_random(F, R, N, Nf):
    gen = torch.Generator().manual_seed(R + N)
    pdf = torch.rand(1, R, N, generator=gen) ** 4
    pdf = pdf / pdf.sum(-1, keepdim=True) * torch.rand(1, R, 1, generator=gen)
    pdf[0, ::7, : N // 3] = 0
    u = torch.rand(1, R, N, 1, generator=gen)
    rng = [1.2, 5.2]
    coarse = ora.stratified_depth(u, N, rng, "metric")
    fine_ref, idx_ref = ora.pdf_depth(pdf, N, Nf, rng, return_idx=True)
    merged_ref = ora.merge_depth(coarse, fine_ref)
    fine, idx, merged = F.sample_pdf_merge(pdf[0].to(DEV), coarse[0, ..., 0].to(DEV), Nf, rng, want_idx=True)
    assert torch.equal(idx.cpu(), idx_ref[0])
    assert torch.equal(fine.cpu(), fine_ref[0, ..., 0])
    assert torch.equal(merged.cpu(), merged_ref[0, ..., 0])
    assert bool((merged[:, 1:] >= merged[:, :-1]).all())


def test_composite_golden(F, golden):
    g = golden("composite")
    B, R, N = g["sigma"].shape
    ins = [g[k].to(DEV).reshape(B * R, *g[k].shape[2:]).requires_grad_(True) for k in ("ray", "rgb_samples", "sigma")]
    depth_s = g["depth_samples"].to(DEV).reshape(B * R, N)
    rgb, d, op, prob = F.composite(ins[0], ins[1], ins[2], depth_s)
    torch.testing.assert_close(rgb.cpu().view(B, R, 3), g["rgb"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(d.cpu().view(B, R, 1), g["depth"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(op.cpu().view(B, R, 1), g["opacity"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(prob.cpu().view(B, R, N, 1), g["prob"], rtol=1e-5, atol=1e-7)
    w = [(syn.uniforms(s, *t.shape) - 0.5).to(DEV) for s, t in zip(g["w_seeds"], (g["rgb"], g["depth"], g["opacity"]))]
    ((rgb.view(B, R, 3) * w[0]).sum() + (d.view(B, R, 1) * w[1]).sum() + (op.view(B, R, 1) * w[2]).sum()).backward()
    assert rel_l2(ins[0].grad.view(B, R, 3), g["d_ray"]) < 1e-4
    assert rel_l2(ins[1].grad.view(B, R, N, 3), g["d_rgb_samples"]) < 1e-5
    assert rel_l2(ins[2].grad.view(B, R, N), g["d_sigma"]) < 1e-4


@pytest.mark.parametrize("R,N", [(300, 128), (77, 50), (5, 1), (64, 192), (129, 64), (41, 256), (9, 96)])
def test_composite_vs_oracle(F, R, N):
    gen = torch.Generator().manual_seed(R * N)
    ray = torch.randn(1, R, 3, generator=gen)
    rgb_s = torch.rand(1, R, N, 3, generator=gen)
    sig = torch.rand(1, R, N, generator=gen) * 4
    depth = (torch.rand(1, R, N, 1, generator=gen) + torch.arange(N)[None, None, :, None]) / N * 4 + 1
    ins_ref = [t.clone().requires_grad_(True) for t in (ray, rgb_s, sig)]
    ref = ora.composite(ins_ref[0], ins_ref[1], ins_ref[2], depth, bgcolor=1.0)
    ins = [t[0].to(DEV).requires_grad_(True) for t in (ray, rgb_s, sig)]
    out = F.composite(ins[0], ins[1], ins[2], depth[0, ..., 0].to(DEV), bgcolor=1.0)
    for a, b in zip(out, ref):
        torch.testing.assert_close(a.cpu().reshape(-1), b.detach().reshape(-1), rtol=2e-5, atol=2e-6)
    gw = [torch.rand(t.shape, generator=gen) - 0.5 for t in ref[:3]]
    sum((a * w).sum() for a, w in zip(ref[:3], gw)).backward()
    sum((a.reshape(w[0].shape) * w[0].to(DEV)).sum() for a, w in zip(out[:3], gw)).backward()
    for a, b in zip(ins, ins_ref):
        assert rel_l2(a.grad, b.grad[0]) < 2e-4


@pytest.mark.parametrize("N", [64, 128, 192, 100])
def test_composite_without_weights_matches(F, N):
    """want_prob=False (training without a fine pass: only the transmittance is kept, the backward recomputes the
    weights) must give the same outputs and gradients as the weight-keeping call, and an eval call (no grad) the
    same outputs with nothing saved."""
    gen = torch.Generator().manual_seed(N)
    R = 203
    ray = torch.randn(R, 3, generator=gen).to(DEV)
    rgb_s = torch.rand(R, N, 3, generator=gen).to(DEV)
    sig = (torch.rand(R, N, generator=gen) * 4).to(DEV)
    depth = ((torch.rand(R, N, generator=gen) + torch.arange(N)) / N * 4 + 1).to(DEV)
    gw = [torch.rand(R, 3, generator=gen).to(DEV) - 0.5, torch.rand(R, generator=gen).to(DEV) - 0.5, torch.rand(R, generator=gen).to(DEV) - 0.5]
    grads = []
    outs = []
    for want in (True, False):
        ins = [t.clone().requires_grad_(True) for t in (ray, rgb_s, sig)]
        out = F.composite(ins[0], ins[1], ins[2], depth, want_prob=want)
        assert out[3].numel() == (R * N if want else 0)
        sum((a * w).sum() for a, w in zip(out[:3], gw)).backward()
        grads.append([t.grad.clone() for t in ins])
        outs.append([t.detach().clone() for t in out[:3]])
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)
    for a, b in zip(grads[0], grads[1]):
        assert torch.equal(a, b), "recomputed weights must equal the stored ones bit for bit"
    with torch.no_grad():
        out = F.composite(ray, rgb_s, sig, depth, want_prob=True)
    for a, b in zip(out[:3], outs[0]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("N,Nf", [(64, 128), (128, 128), (64, 64), (128, 64)])
def test_pdf_sampler_fast_path_hard_cases(F, N, Nf):
    """Register-resident warp kernel (histogram search + merge path) on the cases that stress its shortcuts:
    peaked / all-zero / tiny-valued pdfs (searchsorted ties and empty bins), opaque and transparent rays, and a
    descending coarse list (must fall back to the sort and still be exact)."""
    gen = torch.Generator().manual_seed(N * 1000 + Nf)
    R = 600
    pdf = torch.rand(1, R, N, generator=gen) ** 6
    pdf = pdf / pdf.sum(-1, keepdim=True) * torch.rand(1, R, 1, generator=gen)
    pdf[0, 0::9] = 0                                               # transparent rays
    pdf[0, 1::9, :] = 0; pdf[0, 1::9, N // 2] = 1.0                # one opaque bin
    pdf[0, 2::9, : N - 3] = 0                                      # mass at the far end
    pdf[0, 3::9] *= 1e-30                                          # denormal-range weights
    pdf[0, 4::9] = torch.round(pdf[0, 4::9] * 256) / 256           # CDF values that tie with the query grid
    pdf[0, 5::9] = 1.0 / N                                         # uniform: every query lands on a bin edge
    u = torch.rand(1, R, N, 1, generator=gen)
    rng = [1.2, 5.2]
    coarse = ora.stratified_depth(u, N, rng, "metric")
    coarse[0, 7::9] = coarse[0, 7::9].flip(1)                      # descending -> not mergeable
    fine_ref, idx_ref = ora.pdf_depth(pdf, N, Nf, rng, return_idx=True)
    merged_ref = ora.merge_depth(coarse, fine_ref)
    fine, idx, merged = F.sample_pdf_merge(pdf[0].to(DEV), coarse[0, ..., 0].to(DEV), Nf, rng, want_idx=True)
    assert torch.equal(idx.cpu(), idx_ref[0])
    assert torch.equal(fine.cpu(), fine_ref[0, ..., 0])
    assert torch.equal(merged.cpu(), merged_ref[0, ..., 0])
    # descending bins (inverse-depth style range): fine samples come out descending -> sort fallback
    rng2 = [5.2, 1.2]
    fine_ref, idx_ref = ora.pdf_depth(pdf, N, Nf, rng2, return_idx=True)
    merged_ref = ora.merge_depth(coarse, fine_ref)
    fine, idx, merged = F.sample_pdf_merge(pdf[0].to(DEV), coarse[0, ..., 0].to(DEV), Nf, rng2, want_idx=True)
    assert torch.equal(idx.cpu(), idx_ref[0])
    assert torch.equal(fine.cpu(), fine_ref[0, ..., 0])
    assert torch.equal(merged.cpu(), merged_ref[0, ..., 0])


def test_pdf_sampler_large_batch_properties(F):
    """C4-sized batch (one 480x640 frame of rays is 307 200; here 200 000): bins bit-exact against the oracle on a
    strided subset, merged rows sorted and a permutation-invariant checksum (sum of coarse + fine) preserved."""
    gen = torch.Generator().manual_seed(5)
    R, N, Nf = 200000, 64, 128
    pdf = torch.rand(1, R, N, generator=gen) ** 8
    pdf = pdf / pdf.sum(-1, keepdim=True) * torch.rand(1, R, 1, generator=gen)
    u = torch.rand(1, R, N, 1, generator=gen)
    rng = [1.2, 5.2]
    coarse = ora.stratified_depth(u, N, rng, "metric")
    fine, idx, merged = F.sample_pdf_merge(pdf[0].to(DEV), coarse[0, ..., 0].to(DEV), Nf, rng, want_idx=True)
    sub = slice(0, R, 97)
    fine_ref, idx_ref = ora.pdf_depth(pdf[:, sub], N, Nf, rng, return_idx=True)
    assert torch.equal(idx.cpu()[sub], idx_ref[0])
    assert torch.equal(fine.cpu()[sub], fine_ref[0, ..., 0])
    assert bool((merged[:, 1:] >= merged[:, :-1]).all())
    both = torch.cat([coarse[0, ..., 0].to(DEV), fine], 1).sort(1).values
    assert torch.equal(both, merged)


@pytest.mark.parametrize("B,P", [(3, 4096), (2, 1027), (5, 6)])
def test_raygen_pose_shapes(F, B, P):
    """vectorised (P % 4 == 0) and scalar store paths against the oracle camera model"""
    H, W = 48, 64
    gen = torch.Generator().manual_seed(B * P)
    pose = syn.llff_poses(11, B)
    intr = syn.intrinsics(B, H, W, 0.81)
    idx = torch.randint(0, H * W, (P,), generator=gen)
    c_ref, r_ref = ora.center_and_ray(H, W, pose, intr)
    c, r = F.raygen_pose(pose.to(DEV), intr.to(DEV), H, W, ray_idx=idx.to(DEV))
    torch.testing.assert_close(c.cpu(), c_ref[:, idx], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(r.cpu(), r_ref[:, idx], rtol=1e-5, atol=1e-5)


def test_eval_metrics_golden_and_full_frame(F, golden):
    """PSNR / SSIM / depth error kernels against the reference's numbers (golden) and, on a full
    480x640 frame, against the oracle port."""
    g = golden("metrics")
    H, W, B = g["H"], g["W"], g["B"]
    image = syn.images(g["image_seed"], B, H, W)
    psnr, ssim = F.image_metrics(g["rgb"].to(DEV), image.to(DEV), H, W)
    torch.testing.assert_close(psnr.cpu(), torch.tensor(g["psnr"]), rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(ssim.cpu(), torch.tensor(g["ssim"]), rtol=1e-5, atol=2e-6)
    for scale, key in ((1.0, "depth_err"), (g["depth_scale"], "depth_err_scaled")):
        a, r = F.depth_metrics(g["depth"].to(DEV), g["depth_gt"].to(DEV), g["valid"].to(DEV), scale)
        assert abs(a.item() - g[key][0]) < 1e-5 and abs(r.item() - g[key][1]) < 1e-5
    H, W = 480, 640
    gen = torch.Generator().manual_seed(3)
    image = syn.images(7, 1, H, W)
    rgb = (image.permute(0, 2, 3, 1).reshape(1, H * W, 3) + 0.05 * torch.randn(1, H * W, 3, generator=gen)).clamp(0, 1)
    psnr, ssim = F.image_metrics(rgb.to(DEV), image.to(DEV), H, W)
    rgb_map = rgb.view(-1, H, W, 3).permute(0, 3, 1, 2)
    assert abs(psnr.item() - ora.psnr(rgb_map, image).item()) < 1e-3
    assert abs(ssim.item() - ora.ssim(rgb_map, image).item()) < 1e-5


def _nvp_pack(p, code):
    from neural_invertible_warp_b200.nvp import pack_effective
    return pack_effective(p, code)


@pytest.mark.parametrize("alpha", [0.05, 0.4, 1.0])
def test_nvp_golden(F, golden, alpha):
    g = golden("nvp")
    case = g["cases"][alpha]
    p = {k: v.to(DEV).requires_grad_(True) for k, v in syn.nvp_params(g["param_seed"]).items()}
    code = syn.latent_codes(g["code_seed"], 2).to(DEV).requires_grad_(True)
    wpack, code_bias = _nvp_pack(p, code)
    pts = g["pts"].to(DEV)[:, :, 0]
    out = F.nvp_warp(wpack, code_bias, pts, alpha)
    torch.testing.assert_close(out.cpu(), case["out"][:, :, 0], rtol=1e-5, atol=5e-6)
    w = (syn.uniforms(g["w_seed"], *case["out"].shape) - 0.5)[:, :, 0].to(DEV)
    (out * w).sum().backward()
    assert rel_l2(code.grad, case["d_code"]) < 2e-3
    assert rel_l2(p["lin0_a_1.weight"].grad, case["d_a1w"]) < 2e-3
    assert rel_l2(p["lin2_b_1.weight"].grad, case["d_b1w"]) < 2e-3
    for k, d in case["grads"].items():
        gk = p[k].grad.double().cpu().flatten()
        assert abs(gk.norm().item() - d["l2"]) <= 5e-3 * max(d["l2"], 1e-12), k
        assert abs(gk.sum().item() - d["sum"]) <= 5e-3 * max(d["abssum"], 1e-12), k


def test_nvp_vs_oracle_large(F):
    B, Pt = 3, 700            # crosses CTA and image boundaries (128 points per CTA)
    p_cpu = syn.nvp_params(5)
    code_cpu = syn.latent_codes(6, B)
    pts_cpu = torch.randn(B, Pt, 1, 3, generator=torch.Generator().manual_seed(7)) * 0.7
    q = {k: v.clone().requires_grad_(True) for k, v in p_cpu.items()}
    cg = code_cpu.clone().requires_grad_(True)
    ref = ora.nvp_warp(q, cg, pts_cpu, 0.3)
    w = torch.rand(ref.shape, generator=torch.Generator().manual_seed(8)) - 0.5
    (ref * w).sum().backward()
    p = {k: v.to(DEV).requires_grad_(True) for k, v in p_cpu.items()}
    code = code_cpu.to(DEV).requires_grad_(True)
    wpack, code_bias = _nvp_pack(p, code)
    out = F.nvp_warp(wpack, code_bias, pts_cpu[:, :, 0].to(DEV), 0.3)
    torch.testing.assert_close(out.cpu(), ref.detach()[:, :, 0], rtol=1e-5, atol=5e-6)
    (out * w[:, :, 0].to(DEV)).sum().backward()
    assert rel_l2(code.grad, cg.grad) < 2e-3
    for k in q:
        assert rel_l2(p[k].grad, q[k].grad) < 3e-3, k


@pytest.mark.parametrize("max_ctas,Pt", [(1, 700), (3, 700), (2, 37), (7, 333)])
def test_nvp_backward_on_a_capped_grid_runs_rounds(F, max_ctas, Pt):
    """csrc/nvp.cu nvp_bwd_kernel with the CTA budget the engine's backward overlap leaves it (``max_ctas``): every warp
    takes one point per round, the records of a round are summed into per-thread register accumulators after one barrier
    (two record buffers alternate), the per-image bias sums are flushed when a CTA's point range crosses an image
    boundary.  Odd and even round counts, the 32-round cap (more CTAs than the budget), partial last rounds: gradients
    equal the oracle's."""
    import types
    B = 3
    p_cpu = syn.nvp_params(5)
    code_cpu = syn.latent_codes(6, B)
    pts_cpu = torch.randn(B, Pt, 1, 3, generator=torch.Generator().manual_seed(7)) * 0.7
    q = {k: v.clone().requires_grad_(True) for k, v in p_cpu.items()}
    cg = code_cpu.clone().requires_grad_(True)
    ref = ora.nvp_warp(q, cg, pts_cpu, 0.3)
    w = torch.rand(ref.shape, generator=torch.Generator().manual_seed(8)) - 0.5
    (ref * w).sum().backward()
    p = {k: v.to(DEV).requires_grad_(True) for k, v in p_cpu.items()}
    code = code_cpu.to(DEV).requires_grad_(True)
    wpack, code_bias = _nvp_pack(p, code)
    out = F.nvp_warp(wpack, code_bias, pts_cpu[:, :, 0].to(DEV), 0.3)
    saved = F.backward_overlap
    F.backward_overlap = types.SimpleNamespace(side_ctas=max_ctas, used=True)
    try:
        (out * w[:, :, 0].to(DEV)).sum().backward()
    finally:
        F.backward_overlap = saved
    assert rel_l2(code.grad, cg.grad) < 2e-3
    for k in q:
        if q[k].numel() == 1:
            # a scalar bias gradient is a sum of O(1) terms over all points that may cancel to ~1e-4: absolute bound
            assert abs(float(p[k].grad) - float(q[k].grad)) <= 5e-5 + 3e-3 * abs(float(q[k].grad)), k
        else:
            assert rel_l2(p[k].grad, q[k].grad) < 3e-3, k


@pytest.mark.parametrize("alpha", [0.05, 0.4])
def test_nvp_ray_shards_and_shared_centre_equal_the_full_list(F, alpha):
    """The warp of a per-image list [grid rows (P) ; centre rows (P)] must not change when (a) the rays are split
    into contiguous shards that pass their position in the global list (the annealing quirk of embedder.py:46-49 is
    keyed on it) and (b) the P identical centre rows are evaluated once; outputs equal the full-list oracle and the
    summed parameter gradients equal the full-list gradients."""
    B, P = 2, 48
    gen = torch.Generator().manual_seed(11)
    grid = torch.randn(B, P, 3, generator=gen) * 0.5
    centre = (torch.randn(B, 1, 3, generator=gen) * 0.1).expand(-1, P, -1)
    pts_cpu = torch.cat([grid, centre], 1).contiguous()
    p_cpu, code_cpu = syn.nvp_params(5), syn.latent_codes(6, B)
    q = {k: v.clone().requires_grad_(True) for k, v in p_cpu.items()}
    cg = code_cpu.clone().requires_grad_(True)
    ref = ora.nvp_warp(q, cg, pts_cpu[:, :, None], alpha)[:, :, 0]
    w = torch.rand(ref.shape, generator=gen) - 0.5
    (ref * w).sum().backward()

    def run(parts):
        """parts: list of (point list, index_map, weights) evaluated separately with shared parameters"""
        p = {k: v.to(DEV).requires_grad_(True) for k, v in p_cpu.items()}
        code = code_cpu.to(DEV).requires_grad_(True)
        outs = []
        for pl, im, wl in parts:
            wpack, code_bias = _nvp_pack(p, code)
            o = F.nvp_warp(wpack, code_bias, pl.to(DEV), alpha, index_map=im)
            (o * wl.to(DEV)).sum().backward()
            outs.append(o.detach().cpu())
        return outs, p, code

    # (a) two ray shards of 24 rays, each [grid shard ; centre shard]
    half = P // 2
    parts = []
    for r in range(2):
        sl = slice(r * half, (r + 1) * half)
        pl = torch.cat([grid[:, sl], centre[:, sl]], 1).contiguous()
        wl = torch.cat([w[:, sl], w[:, P + r * half:P + (r + 1) * half]], 1).contiguous()
        wpts, im, shared = F.warp_point_list(pl, half, shard=(r * half, P))
        assert shared
        # shared centre row: its weight is the sum of the shard's centre-row weights
        wl = torch.cat([wl[:, :half], wl[:, half:].sum(1, keepdim=True)], 1)
        parts.append((wpts, im, wl))
    outs, p, code = run(parts)
    for r in range(2):
        sl = slice(r * half, (r + 1) * half)
        torch.testing.assert_close(outs[r][:, :half], ref.detach()[:, sl], rtol=1e-5, atol=5e-6)
        torch.testing.assert_close(outs[r][:, half:].expand(-1, half, -1), ref.detach()[:, P + r * half:P + (r + 1) * half],
                                   rtol=1e-5, atol=5e-6)
    assert rel_l2(code.grad, cg.grad) < 2e-3
    for k in q:
        assert rel_l2(p[k].grad, q[k].grad) < 3e-3, k
    # (b) without the index map the shards anneal the wrong rows: the outputs must differ (the map is not a no-op)
    o_wrong = F.nvp_warp(*_nvp_pack({k: v.to(DEV) for k, v in p_cpu.items()}, code_cpu.to(DEV)), parts[1][0].to(DEV), alpha)
    assert (o_wrong.cpu()[:, :half] - ref.detach()[:, half:P]).abs().max() > 1e-4


@pytest.mark.parametrize("progress", [0.2, 0.3, 1.0])
def test_nerf_mlp_fp32_golden(F, golden, progress):
    """NeRF.forward on the golden points: centre = point, zero depth, ray = view direction."""
    g = golden("nerf_mlp")
    case = g["cases"][progress]
    p_cpu = syn.nerf_params(g["param_seed"])
    flat = flat_params(p_cpu).to(DEV).requires_grad_(True)
    pts = g["points"].reshape(-1, 3).to(DEV).requires_grad_(True)
    unit = g["ray_unit"].reshape(-1, 3).to(DEV).requires_grad_(True)
    depth = torch.zeros(pts.shape[0], 1, device=DEV)
    prog, c2f = progress, g["c2f"]
    rgb, sigma = F.nerf_forward_samples(flat, pts, unit, depth, prog, c2f, "fp32", training=True)
    shp = g["points"].shape[:-1]
    torch.testing.assert_close(rgb.cpu().view(*shp, 3), case["rgb"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(sigma.cpu().view(*shp), case["density"], rtol=1e-4, atol=1e-5)
    wr = (syn.uniforms(g["wr_seed"], *case["rgb"].shape) - 0.5).to(DEV)
    wd = (syn.uniforms(g["wd_seed"], *case["density"].shape) - 0.5).to(DEV)
    ((rgb.view(*shp, 3) * wr).sum() + (sigma.view(*shp) * wd).sum()).backward()
    assert rel_l2(pts.grad.view(*shp, 3), case["d_points"]) < 2e-3
    grads = unflatten(flat.grad.cpu(), p_cpu)
    for k, d in case["grads"].items():
        gk = grads[k].double().flatten()
        assert abs(gk.norm().item() - d["l2"]) <= 2e-3 * max(d["l2"], 1e-12), k
        assert abs(gk.sum().item() - d["sum"]) <= 2e-3 * max(d["abssum"], 1e-12), k


@pytest.mark.parametrize("R,N", [(48, 16), (10, 128), (3, 37)])
def test_nerf_fp32_vs_oracle_with_ray_grads(F, R, N):
    """forward_samples incl. all three gradient routes into ray and the centre route."""
    gen = torch.Generator().manual_seed(R + 31 * N)
    p_cpu = syn.nerf_params(3)
    center = torch.randn(1, R, 3, generator=gen) * 0.1
    ray = torch.randn(1, R, 3, generator=gen) * 0.5 + torch.tensor([0., 0., 1.])
    u = torch.rand(1, R, N, 1, generator=gen)
    depth = ora.stratified_depth(u, N, [1.2, 5.2], "metric")
    q = {k: v.clone().requires_grad_(True) for k, v in p_cpu.items()}
    c_ref, r_ref = center.clone().requires_grad_(True), ray.clone().requires_grad_(True)
    pts, unit = ora.sample_points(c_ref, r_ref, depth)
    rgb_ref, sig_ref = ora.nerf_mlp(q, pts, unit, progress=0.3, c2f=[0.1, 0.5])
    out_ref = ora.composite(r_ref, rgb_ref, sig_ref, depth)
    tgt = torch.rand(1, R, 3, generator=gen)
    loss_ref = ora.mse(out_ref[0], tgt) + 0.1 * out_ref[1].mean()
    loss_ref.backward()

    flat = flat_params(p_cpu).to(DEV).requires_grad_(True)
    c, r = center[0].to(DEV).requires_grad_(True), ray[0].to(DEV).requires_grad_(True)
    d = depth[0, ..., 0].to(DEV)
    prog, c2f = 0.3, [0.1, 0.5]
    rgb_s, sig_s = F.nerf_forward_samples(flat, c, r, d, prog, c2f, "fp32")
    torch.testing.assert_close(rgb_s.cpu(), rgb_ref.detach()[0], rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(sig_s.cpu(), sig_ref.detach()[0], rtol=1e-4, atol=2e-5)
    rgb, dep, op, _ = F.composite(r, rgb_s, sig_s, d)
    assert (rgb.cpu() - out_ref[0].detach()[0]).abs().max() < 1e-3     # north_star forward tolerance
    assert (dep.cpu() - out_ref[1].detach()[0, :, 0]).abs().max() < 1e-3
    assert (op.cpu() - out_ref[2].detach()[0, :, 0]).abs().max() < 1e-3
    loss = ((rgb - tgt[0].to(DEV)) ** 2).mean() + 0.1 * dep.mean()
    loss.backward()
    assert rel_l2(c.grad, c_ref.grad[0]) < 1e-2
    assert rel_l2(r.grad, r_ref.grad[0]) < 1e-2
    grads = unflatten(flat.grad.cpu(), p_cpu)
    for k in q:
        assert rel_l2(grads[k], q[k].grad) < 5e-3, k


def test_mse_gather(F):
    gen = torch.Generator().manual_seed(0)
    B, P, H, W = 3, 50, 12, 16
    image = torch.rand(B, 3, H, W, generator=gen)
    rgb = torch.rand(B, P, 3, generator=gen).requires_grad_(True)
    idx = torch.randperm(H * W, generator=gen)[:P]
    ref = ora.mse(rgb, ora.gather_pixels(image, idx))
    ref.backward()
    x = rgb.detach().to(DEV).requires_grad_(True)
    loss = F.mse_gather(x, image.to(DEV), idx.to(DEV))
    torch.testing.assert_close(loss.cpu(), ref.detach(), rtol=1e-5, atol=1e-7)
    (loss * 3).backward()
    torch.testing.assert_close(x.grad.cpu(), rgb.grad * 3, rtol=1e-5, atol=1e-8)


@pytest.mark.parametrize("B,P,N,with_idx,bg", [(3, 50, 128, True, None), (2, 37, 64, False, 1.0), (4, 5000, 128, True, None),
                                                  (1, 9, 192, True, 0.5)])
def test_composite_with_the_loss_head_in_its_epilogue(F, B, P, N, with_idx, bg):
    """niw_composite_fwd_mse / _bwd_mse (model/nerf.py:276-288 + 458-474 in one forward and one backward
    launch) against the oracle's composite followed by its MSE over the gathered pixels, and against the two-kernel
    form (composite -> mse_gather): same colours (bit for bit), same loss, same gradients with upstream gradients on the
    loss AND on rgb / depth; repeated calls (the ticket of the block-ordered final sum resets itself)."""
    gen = torch.Generator().manual_seed(B * 1000 + P)
    H, W = 60, 100
    R = B * P
    image = torch.rand(B, 3, H, W, generator=gen)
    idx = torch.randperm(H * W, generator=gen)[:P] if with_idx else None
    ray = torch.randn(R, 3, generator=gen)
    rgb_s = torch.rand(R, N, 3, generator=gen)
    sigma = torch.rand(R, N, generator=gen) * 3
    depth_s = (torch.rand(R, N, generator=gen) * 0.1 + 0.01).cumsum(-1)
    w_rgb, w_depth = torch.randn(R, 3, generator=gen), torch.randn(R, generator=gen)

    def total(rgb, depth, loss):
        return 0.7 * loss + (rgb * w_rgb.to(rgb.device)).sum() * 1e-3 + (depth * w_depth.to(rgb.device)).sum() * 1e-3

    # oracle (CPU, fp32)
    a = [t.clone().requires_grad_(True) for t in (ray, rgb_s, sigma)]
    o_rgb, o_depth, o_op, _ = ora.composite(a[0].view(B, P, 3), a[1].view(B, P, N, 3), a[2].view(B, P, N),
                                            depth_s.view(B, P, N, 1), bgcolor=bg)
    gt = ora.gather_pixels(image, idx) if with_idx else image.view(B, 3, H * W).permute(0, 2, 1)[:, :P]
    o_loss = ora.mse(o_rgb, gt)
    total(o_rgb.reshape(R, 3), o_depth.reshape(R), o_loss).backward()

    outs = {}
    for fused in (False, True):
        g = [t.to(DEV).requires_grad_(True) for t in (ray, rgb_s, sigma)]
        for rep in range(2 if fused else 1):
            for t in g:
                t.grad = None
            if fused:
                target = F.MseTarget(image.to(DEV), None if idx is None else idx.to(DEV))
                rgb, depth, op, _, loss = F.composite_mse(g[0], g[1], g[2], depth_s.to(DEV), target, B, P, bgcolor=bg,
                                                          want_prob=False)
            else:
                rgb, depth, op, _ = F.composite(g[0], g[1], g[2], depth_s.to(DEV), bgcolor=bg, want_prob=False)
                loss = F.mse_gather(rgb.view(B, P, 3), image.to(DEV), None if idx is None else idx.to(DEV))
            total(rgb, depth, loss).backward()
            torch.cuda.synchronize()
        outs[fused] = (rgb.detach(), depth.detach(), op.detach(), loss.detach(), [t.grad.clone() for t in g])
    f, u = outs[True], outs[False]
    assert torch.equal(f[0], u[0]) and torch.equal(f[1], u[1]) and torch.equal(f[2], u[2])
    torch.testing.assert_close(f[3], u[3], rtol=2e-6, atol=1e-9)
    torch.testing.assert_close(f[3].cpu(), o_loss.detach(), rtol=1e-5, atol=1e-8)
    for gf, gu, go in zip(f[4], u[4], a):
        assert rel_l2(gf, gu) < 1e-6
        assert rel_l2(gf, go.grad) < 3e-4


def test_sample_pixels_is_a_permutation_prefix(F):
    """niw_sample_pixels: k distinct indices in range, the full draw (k = n) is a permutation, successive calls
    differ (device counter), marginals are uniform to sampling error."""
    n = 48 * 64
    counter = torch.zeros(1, dtype=torch.int64, device=DEV)
    full = F.sample_pixels(n, n, counter, seed=3).cpu()
    assert torch.equal(full.sort().values, torch.arange(n))
    a = F.sample_pixels(n, 64, counter, seed=3).cpu()
    b = F.sample_pixels(n, 64, counter, seed=3).cpu()
    assert int(counter) == 3 and a.unique().numel() == 64 and not torch.equal(a, b)
    assert int(a.min()) >= 0 and int(a.max()) < n
    hits = torch.zeros(8)
    for _ in range(400):
        d = F.sample_pixels(307200, 64, counter, seed=7).cpu()
        assert d.unique().numel() == 64 and int(d.max()) < 307200
        hits += torch.bincount(d * 8 // 307200, minlength=8).float()
    frac = hits / hits.sum()
    assert (frac - 0.125).abs().max() < 0.01, frac      # 25 600 draws: sigma of a bin share ~ 0.002


def test_ops_refuse_cpu_tensors(F):
    with pytest.raises(RuntimeError):
        F.composite(torch.zeros(2, 3), torch.zeros(2, 4, 3), torch.zeros(2, 4), torch.zeros(2, 4))


def test_kabsch_matches_svd_solution(F):
    """Row f1: niw_kabsch (Horn's quaternion form, fp64 Jacobi) against the SVD-based batched Kabsch with the det fix
    (what roma.rigid_points_registration computes) on noisy rigid motions, including a near-reflection case where the
    det fix matters and the reference's own usage (P grid rows + the centre repeated P times)."""
    from neural_invertible_warp_b200 import camera
    gen = torch.Generator().manual_seed(21)
    B, M = 6, 96
    x = torch.randn(B, M, 3, generator=gen)
    x[1, :, 2] *= 1e-3                                                  # nearly planar cloud
    x[2, M // 2:] = x[2, :1]                                            # half of the rows are one repeated point
    w = torch.randn(B, 3, generator=gen) * 0.8
    Rt = camera.lie.so3_to_SO3(w)
    y = x @ Rt.transpose(1, 2) + torch.randn(B, 1, 3, generator=gen) + 0.01 * torch.randn(B, M, 3, generator=gen)
    y[3] = y[3] * torch.tensor([1.0, 1.0, -1.0])                       # mirrored target: the unconstrained optimum is a reflection
    from oracle import reference_port as ora
    R_ref, t_ref = ora.kabsch(x, y)                                    # CPU oracle: SVD with the determinant fix
    with pytest.raises(RuntimeError):                                  # the product has no CPU path
        camera.rigid_points_registration(x, y)
    R, t = F.kabsch(x.to(DEV), y.to(DEV))
    torch.testing.assert_close(R.cpu(), R_ref, rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(t.cpu(), t_ref, rtol=1e-4, atol=2e-5)
    assert torch.allclose(torch.linalg.det(R.cpu()), torch.ones(B), atol=1e-5)
    R2, t2 = camera.rigid_points_registration(x.to(DEV), y.to(DEV))    # the product path dispatches to the kernel
    assert torch.equal(R2, R) and torch.equal(t2, t)
    # the two-step form used under data parallelism (statistics -> [all-reduce] -> solve): same fit, and the statistics
    # are additive over a split of the rows
    st = F.kabsch_stats(x.to(DEV), y.to(DEV))
    torch.testing.assert_close(st.cpu(), ora.kabsch_stats(x, y), rtol=1e-12, atol=1e-12)
    R3, t3 = F.kabsch_solve(st)
    torch.testing.assert_close(R3.cpu(), R_ref, rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(t3.cpu(), t_ref, rtol=1e-4, atol=2e-5)
    st2 = F.kabsch_stats(x[:, :40].to(DEV), y[:, :40].to(DEV)) + F.kabsch_stats(x[:, 40:].to(DEV), y[:, 40:].to(DEV))
    R4, t4 = F.kabsch_solve(st2)
    torch.testing.assert_close(R4, R3, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(t4, t3, rtol=1e-6, atol=1e-6)


def test_ops_reject_empty_and_unsupported_shapes(F):
    """Error behaviour at the edges: the C ABI returns a code (never a silent no-op, never a fallback) and the wrappers
    raise -- empty batches, a sampler shape beyond its shared-memory budget, a non-8x256 parameter vector, and the
    single-ray / single-sample corner still computes."""
    z = lambda *s: torch.zeros(*s, device=DEV)
    with pytest.raises(RuntimeError):
        F.composite(z(0, 3), z(0, 4, 3), z(0, 4), z(0, 4))
    with pytest.raises(RuntimeError):
        F.sample_stratified(None, 0, 8, [1.0, 2.0], "metric", device=DEV)
    with pytest.raises(RuntimeError):
        F.sample_pdf_merge(z(2, 4096), z(2, 4096), 4096, [1.0, 2.0])          # N + Nf beyond the in-kernel sort budget
    with pytest.raises(RuntimeError):
        F.nerf_forward_samples(z(1000), z(1, 3), z(1, 3), z(1, 1), None, None, "fp32", training=False)
    rgb, d, op, prob = F.composite(torch.ones(1, 3, device=DEV), torch.full((1, 1, 3), 0.5, device=DEV), torch.ones(1, 1, device=DEV),
                                   torch.ones(1, 1, device=DEV))
    torch.testing.assert_close(op.cpu(), torch.ones(1))                       # one sample: last interval 1e10 -> opaque
    torch.testing.assert_close(rgb.cpu(), torch.full((1, 3), 0.5))


def test_flat_adam_matches_torch_adam():
    """niw_adam_step (row f2) vs torch.optim.Adam + ExponentialLR (model/nerf.py:33-46) on two groups with odd sizes,
    20 steps, weight decay off/on."""
    from neural_invertible_warp_b200 import engine
    gen = torch.Generator().manual_seed(3)
    shapes_a, shapes_b = [(257, 63), (257,), (5, 3)], [(16, 6), (7,)]
    for wd in (0.0, 1e-2):
        ref_a = [torch.randn(s, generator=gen).requires_grad_(True) for s in shapes_a]
        ref_b = [torch.randn(s, generator=gen).requires_grad_(True) for s in shapes_b]
        ours_a = [torch.nn.Parameter(t.detach().clone().to(DEV)) for t in ref_a]
        ours_b = [torch.nn.Parameter(t.detach().clone().to(DEV)) for t in ref_b]
        o_a = torch.optim.Adam(ref_a, lr=1e-2, weight_decay=wd)
        o_b = torch.optim.Adam(ref_b, lr=3e-3, weight_decay=wd)
        s_a = torch.optim.lr_scheduler.ExponentialLR(o_a, gamma=0.9)
        fa = engine.FlatAdam([dict(params=ours_a, lr=1e-2, gamma=0.9, weight_decay=wd),
                              dict(params=ours_b, lr=3e-3, weight_decay=wd)])
        for it in range(20):
            for r, o in zip(ref_a + ref_b, ours_a + ours_b):
                g = torch.randn(r.shape, generator=gen) * (0.1 + it)
                r.grad = g.clone()
                o.grad.copy_(g.to(DEV))          # gradients are views into the flat bucket
            o_a.step(); o_b.step(); s_a.step()
            fa.step()
        for r, o in zip(ref_a + ref_b, ours_a + ours_b):
            torch.testing.assert_close(o.detach().cpu(), r.detach(), rtol=2e-5, atol=2e-6)
        assert fa.state[:, 0].tolist() == [20.0, 20.0]
        assert ours_a[0].data_ptr() == fa.flat_params.data_ptr()
    # pose-LR warm-up and the BARF progress scalar (model/barf.py:46-60), kept on the device
    ref = [torch.randn(9, 5, generator=gen).requires_grad_(True)]
    ours = [torch.nn.Parameter(ref[0].detach().clone().to(DEV))]
    prog = torch.nn.Parameter(torch.tensor(0.0, device=DEV))
    o = torch.optim.Adam(ref, lr=3e-3)
    fa = engine.FlatAdam([dict(params=ours, lr=3e-3, warmup=4)], progress=[prog], max_iter=200000)
    for it in range(9):
        g = torch.randn(9, 5, generator=gen)
        ref[0].grad = g.clone(); ours[0].grad.copy_(g.to(DEV))
        o.param_groups[0]["lr"] = 3e-3 * min(1, it / 4)
        o.step()
        fa.step()
        assert prog.item() == torch.tensor((it + 1) / 200000).item()        # fp32(it / max_iter), as fill_ stores it
    torch.testing.assert_close(ours[0].detach().cpu(), ref[0].detach(), rtol=2e-5, atol=2e-6)


def test_stratified_depths_with_in_kernel_uniforms(F):
    """niw_sample_stratified_rng (functional.DeviceUniform): the uniforms of Graph.sample_depth (model/nerf.py:334-344) drawn
    inside the kernel.  Every depth lies in its own stratum, the implied uniforms are uniform (mean, variance, no value
    outside [0, 1)), every call draws afresh, the same seed and call number reproduce the draw, and a CUDA-graph replay
    advances the call number on the device."""
    R, N = 4096, 128
    dmin, dmax = 1.2, 5.2
    rng = torch.zeros(2, dtype=torch.int64, device=DEV)
    a = F.sample_stratified(F.DeviceUniform((1, R, N, 1), rng, 7), R, N, [dmin, dmax], "metric")
    b = F.sample_stratified(F.DeviceUniform((1, R, N, 1), rng, 7), R, N, [dmin, dmax], "metric")
    torch.cuda.synchronize()
    assert rng.tolist() == [2, 0]
    k = torch.arange(N, device=DEV, dtype=torch.float64)
    for d in (a, b):
        u = (d.double() - dmin) / (dmax - dmin) * N - k
        assert u.min().item() > -1e-4 and u.max().item() < 1 + 1e-4            # inside the stratum (fp32 rounding of the affine map)
        assert abs(u.mean().item() - 0.5) < 2e-3 and abs(u.var().item() - 1 / 12) < 2e-3
    assert not torch.equal(a, b)
    rng2 = torch.zeros(2, dtype=torch.int64, device=DEV)
    a2 = F.sample_stratified(F.DeviceUniform((1, R, N, 1), rng2, 7), R, N, [dmin, dmax], "metric")
    assert torch.equal(a, a2)                                                  # same seed, same call number
    other = F.sample_stratified(F.DeviceUniform((1, R, N, 1), torch.zeros(2, dtype=torch.int64, device=DEV), 8), R, N, [dmin, dmax], "metric")
    assert not torch.equal(a, other)
    # against the product sampler fed the same uniforms: recover u from the metric depths, feed it back
    inv = F.sample_stratified(F.DeviceUniform((1, R, N, 1), torch.zeros(2, dtype=torch.int64, device=DEV), 7), R, N,
                              torch.tensor([dmin, dmax], device=DEV), "inverse")
    torch.testing.assert_close(inv, 1.0 / (a + 1e-8), rtol=2e-6, atol=0)          # same draw through the inverse parametrisation
    # captured: the call number advances on replay
    static = torch.zeros(2, dtype=torch.int64, device=DEV)
    g = torch.cuda.CUDAGraph()
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        out = F.sample_stratified(F.DeviceUniform((1, R, N, 1), static, 7), R, N, [dmin, dmax], "metric")
    g.replay(); torch.cuda.synchronize()
    first = out.clone()
    g.replay(); torch.cuda.synchronize()
    assert static.tolist() == [2, 0] and not torch.equal(first, out) and torch.equal(first, a)
