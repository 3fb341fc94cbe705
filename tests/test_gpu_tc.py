"""tcgen05 path: descriptor self-test, then the fused BF16 MLP forward against the FP32 CUDA path
and the CPU oracle."""
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


def _bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


@pytest.mark.parametrize("variant", [0, 2])
@pytest.mark.parametrize("N,K", [(256, 256), (128, 64), (64, 32)])
def test_tc_selftest_variants(F, variant, N, K):
    """Variants 0 (K-major) and 2 (MN-major) must be exact BF16 GEMMs.  (Variants 1 / 3 -- LBO and
    SBO exchanged -- were run once on a B200: they fault with an illegal address, which settles
    the descriptor convention; they are not run here because the fault poisons the CUDA context.)"""
    gen = torch.Generator().manual_seed(N + K)
    A = torch.randn(128, K, generator=gen)
    Bm = torch.randn(N, K, generator=gen)
    ref = _bf16(A).double() @ _bf16(Bm).double().t()
    D = F.tc_selftest(A.to(DEV), Bm.to(DEV), variant).cpu().double()
    err = (D - ref).abs().max().item()
    print("tc_selftest variant %d N=%d K=%d max err %.3e" % (variant, N, K, err))
    assert err < 1e-3, "tcgen05 descriptor convention mismatch (variant %d): %.3e" % (variant, err)


@pytest.mark.parametrize("N,K", [(256, 256), (128, 64), (32, 16)])
def test_tc_selftest_cta_pair(F, N, K):
    """cta_group::2: one M = 256 MMA chain over a CTA pair (each CTA stages 128 rows of A and N/2 rows of B)."""
    gen = torch.Generator().manual_seed(7 * N + K)
    A = torch.randn(256, K, generator=gen)
    Bm = torch.randn(N, K, generator=gen)
    ref = _bf16(A).double() @ _bf16(Bm).double().t()
    D = F.tc_selftest(A.to(DEV), Bm.to(DEV), 4).cpu().double()
    err = (D - ref).abs().max().item()
    print("tc_selftest pair N=%d K=%d max err %.3e" % (N, K, err))
    assert err < 1e-3, "cta_group::2 operand split mismatch: %.3e" % err


def _flat(p):
    keys = []
    for i in range(8):
        keys += [f"mlp_feat.{i}.weight", f"mlp_feat.{i}.bias"]
    for i in range(2):
        keys += [f"mlp_rgb.{i}.weight", f"mlp_rgb.{i}.bias"]
    return torch.cat([p[k].reshape(-1) for k in keys])


@pytest.mark.parametrize("R,N", [(4, 128), (37, 16), (64, 192), (1024, 128)])
def test_tc_forward_vs_fp32(F, R, N):
    gen = torch.Generator().manual_seed(R + N)
    flat = _flat(syn.nerf_params(11)).to(DEV)
    center = (torch.randn(R, 3, generator=gen) * 0.1).to(DEV)
    ray = (torch.randn(R, 3, generator=gen) * 0.3 + torch.tensor([0., 0., 1.])).to(DEV)
    u = torch.rand(1, R, N, 1, generator=gen)
    depth = ora.stratified_depth(u, N, [1.2, 5.2], "metric")[0, ..., 0].to(DEV)
    prog, c2f = 0.3, [0.1, 0.5]
    rgb32, sig32 = F.nerf_forward_samples(flat, center, ray, depth, prog, c2f, "fp32", training=False)
    rgb16, sig16 = F.nerf_forward_samples(flat, center, ray, depth, prog, c2f, "bf16", training=False)
    torch.cuda.synchronize()
    e_rgb = (rgb16 - rgb32).abs().max().item()
    e_sig = ((sig16 - sig32).abs() / (1 + sig32.abs())).max().item()
    print("bf16 vs fp32 per-sample: rgb %.3e sigma(rel) %.3e" % (e_rgb, e_sig))
    assert e_rgb < 3e-2 and e_sig < 3e-2
    out32 = F.composite(ray, rgb32, sig32, depth)
    out16 = F.composite(ray, rgb16, sig16, depth)
    for a, b, name in zip(out16[:3], out32[:3], ("rgb", "depth", "opacity")):
        print("composited %s max abs diff %.3e" % (name, (a - b).abs().max().item()))
    assert (out16[0] - out32[0]).abs().max() < 1e-2
    assert (out16[2] - out32[2]).abs().max() < 1e-2


def test_tc_forward_inverse_depth_far_samples(F):
    """LLFF inverse-depth sampling puts the last samples at |x| up to 1e8: composited outputs must
    still agree (transmittance is ~0 there)."""
    R, N = 256, 128
    gen = torch.Generator().manual_seed(5)
    flat = _flat(syn.nerf_params(12)).to(DEV)
    center = (torch.randn(R, 3, generator=gen) * 0.05).to(DEV)
    ray = (torch.randn(R, 3, generator=gen) * 0.2 + torch.tensor([0., 0., 1.])).to(DEV)
    u = torch.rand(1, R, N, 1, generator=gen)
    depth = ora.stratified_depth(u, N, [1, 0], "inverse")[0, ..., 0].to(DEV)
    prog, c2f = 0.3, [0.1, 0.5]
    o32 = F.composite(ray, *F.nerf_forward_samples(flat, center, ray, depth, prog, c2f, "fp32", training=False), depth)
    o16 = F.composite(ray, *F.nerf_forward_samples(flat, center, ray, depth, prog, c2f, "bf16", training=False), depth)
    assert (o16[0] - o32[0]).abs().max() < 1e-2
    assert (o16[2] - o32[2]).abs().max() < 1e-2


def _nerf_keys():
    keys = []
    for i in range(8):
        keys += [f"mlp_feat.{i}.weight", f"mlp_feat.{i}.bias"]
    for i in range(2):
        keys += [f"mlp_rgb.{i}.weight", f"mlp_rgb.{i}.bias"]
    return keys


def _split(flat, like):
    out, off = {}, 0
    for k in _nerf_keys():
        n = like[k].numel()
        out[k] = flat[off:off + n]
        off += n
    return out


@pytest.mark.parametrize("R,N,param,tol", [(1024, 128, "metric", 1e-2), (1024, 128, "inverse", 1e-2),
                                           (37, 16, "metric", 0.3), (33, 192, "metric", 0.3)])
def test_tc_backward_vs_fp32(F, R, N, param, tol):
    """BF16 tensor-core backward (dX chain + dW GEMMs) against the FP32 CUDA-core backward on the same
    inputs, through compositing and an MSE loss against random target colours (the train-step loss).
    Metric = per-tensor relative L2.  At the C2 batch (1 024 rays x 128) every MLP
    weight / bias gradient must be within 1e-2 (north_star: 'gradients within 1e-2 relative for
    BF16 operands').  The small ragged shapes exercise partial tiles and non-uniform warps; with so
    few rays the ReLU-mask flips between a BF16 and an FP32 forward do not average out, so they only
    get a structural bound.  The per-ray input gradients are ill-conditioned at random init
    (SURVEY.md H10 measured ~12% for BF16 operands) and get a loose bound."""
    gen = torch.Generator().manual_seed(R * N)
    p = syn.nerf_params(13)
    flat0 = _flat(p).to(DEV)
    center0 = (torch.randn(R, 3, generator=gen) * 0.1).to(DEV)
    ray0 = (torch.randn(R, 3, generator=gen) * 0.3 + torch.tensor([0., 0., 1.])).to(DEV)
    u = torch.rand(1, R, N, 1, generator=gen)
    rng = [1.2, 5.2] if param == "metric" else [1, 0]
    depth = ora.stratified_depth(u, N, rng, param)[0, ..., 0].to(DEV)
    prog, c2f = 0.3, [0.1, 0.5]
    target = torch.rand(R, 3, generator=gen).to(DEV)
    grads = {}
    for prec in ("fp32", "bf16"):
        flat = flat0.clone().requires_grad_(True)
        c = center0.clone().requires_grad_(True)
        r = ray0.clone().requires_grad_(True)
        rgb_s, sig_s = F.nerf_forward_samples(flat, c, r, depth, prog, c2f, prec, training=True)
        rgb, dep, op, _ = F.composite(r, rgb_s, sig_s, depth)
        ((rgb - target) ** 2).mean().backward()
        torch.cuda.synchronize()
        grads[prec] = (flat.grad.clone(), c.grad.clone(), r.grad.clone())
    g32, g16 = _split(grads["fp32"][0], p), _split(grads["bf16"][0], p)
    for k in _nerf_keys():
        a, b = g16[k].double(), g32[k].double()
        rel = ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
        print("d%-20s rel-L2 %.3e  (|g| %.3e)" % (k, rel, b.norm().item()))
        assert rel < tol, (k, rel)
    for name, i in (("d_center", 1), ("d_ray", 2)):
        a, b = grads["bf16"][i].double(), grads["fp32"][i].double()
        rel = ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
        print("%-21s rel-L2 %.3e" % (name, rel))
        assert rel < 0.5, (name, rel)


def test_tc_full_size_equals_sum_of_chunks(F):
    """C5's per-GPU share (8 192 rays x 128 samples = 8 192 tiles) in ONE call against the same rays in eight calls of
    1 024: samples are independent, so the outputs must be bit-identical whatever tile / CTA-pair / slot a sample lands
    in, and the parameter gradient of the whole batch must equal the sum of the chunks' gradients (fp32 summation order
    of the dW slices is the only difference)."""
    R, N, K = 8192, 128, 8
    gen = torch.Generator().manual_seed(77)
    p = syn.nerf_params(13)
    flat0 = _flat(p).to(DEV)
    center = (torch.randn(R, 3, generator=gen) * 0.1).to(DEV)
    ray = (torch.randn(R, 3, generator=gen) * 0.3 + torch.tensor([0., 0., 1.])).to(DEV)
    depth = (torch.rand(R, N, generator=gen) * 4 + 1).sort(-1).values.to(DEV)
    w_rgb = (torch.rand(R, N, 3, generator=gen) - 0.5).to(DEV)
    w_sig = (torch.rand(R, N, generator=gen) - 0.5).to(DEV)
    prog, c2f = 0.3, [0.1, 0.5]

    def run(lo, hi):
        flat = flat0.clone().requires_grad_(True)
        c, r = center[lo:hi].clone().requires_grad_(True), ray[lo:hi].clone().requires_grad_(True)
        rgb, sig = F.nerf_forward_samples(flat, c, r, depth[lo:hi].contiguous(), prog, c2f, "bf16", training=True)
        ((rgb * w_rgb[lo:hi]).sum() + (sig * w_sig[lo:hi]).sum()).backward()
        return rgb.detach(), sig.detach(), flat.grad, c.grad, r.grad

    whole = run(0, R)
    parts = [run(i * R // K, (i + 1) * R // K) for i in range(K)]
    assert torch.equal(whole[0], torch.cat([q[0] for q in parts]))
    assert torch.equal(whole[1], torch.cat([q[1] for q in parts]))
    assert torch.isfinite(whole[2]).all()
    gsum = torch.stack([q[2] for q in parts]).double().sum(0)
    rel = ((whole[2].double() - gsum).norm() / gsum.norm()).item()
    # the dW accumulators are fp32 (TMEM): summing 1 M samples in one chain vs eight chains of 131 k differs by
    # ~eps sqrt(n) = 7e-5 (measured 6.9e-5)
    assert rel < 5e-4, rel
    for i in (3, 4):      # per-ray input gradients: same atomics per ray in both runs, order within a ray may differ
        a, b = whole[i].double(), torch.cat([q[i] for q in parts]).double()
        assert ((a - b).norm() / b.norm()).item() < 1e-4


# --------------------------------------------------------------------------------------------
# tensor-core paths DIRECTLY against the CPU oracle (reference model/nerf.py:416-474), C2 batch
# --------------------------------------------------------------------------------------------

# tolerances (BASELINE.json north_star): rendered rgb / depth / opacity within 1e-3 max abs for the 1e-3 path (bf16x3:
# split BF16 operands, FP32 accumulate); MLP parameter gradients within 1e-2 relative (per-tensor rel-L2) for BF16
# operands.  Plain BF16 operands do NOT meet 1e-3 on the forward: they are held to the measured level.
# Inverse-depth (LLFF) composited depths reach ~1e1..1e3 (weights x 1/(u+1e-8)): depth is compared as
# |d - d_ref| <= atol + rtol |d_ref|.
OUT_TOL = {"bf16": dict(rgb=5e-3, opacity=1e-3, depth_atol=2e-2, depth_rtol=5e-3),
           "bf16x3": dict(rgb=1e-3, opacity=1e-3, depth_atol=1e-3, depth_rtol=1e-3)}
GRAD_TOL = 1e-2
# The one tensor that sits AT the bar: mlp_feat.0.weight.  Its gradient is G0^T . enc with G0 at the end of the eight-layer
# dX chain, every link of which rounds its G operand to BF16 (~0.35 % relative each, independent: sqrt(8) x 0.35 % ~ 1.0 %).
# SURVEY.md H10 measured 1.3 % for it by emulating BF16 operand rounding in the reference itself, i.e. this is the floor of
# "BF16 operands", not of this implementation; measured here 1.0e-2 (metric depth).  All other tensors and the whole
# parameter vector are held to 1e-2.
GRAD_TOL_FIRST_LAYER = 1.25e-2


def _oracle_c2(p, center, ray, depth, target, prog, c2f):
    q = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    c, r = center.clone().requires_grad_(True), ray.clone().requires_grad_(True)
    d = depth[None, ..., None]
    pts, unit = ora.sample_points(c[None], r[None], d)
    rgb_s, sig_s = ora.nerf_mlp(q, pts, unit, progress=prog, c2f=c2f)
    rgb, dep, op, _ = ora.composite(r[None], rgb_s, sig_s, d)
    loss = ((rgb[0] - target) ** 2).mean()
    loss.backward()
    return dict(rgb=rgb[0].detach(), depth=dep[0, :, 0].detach(), opacity=op[0, :, 0].detach(), loss=loss.detach(),
                grads={k: v.grad for k, v in q.items()}, d_center=c.grad, d_ray=r.grad)


@pytest.mark.parametrize("param", ["metric", "inverse"])
def test_tc_paths_vs_oracle_c2(F, param):
    """BF16 and split-BF16 (bf16x3) tensor-core paths against the CPU oracle at the C2 batch (1 024 rays x 128 samples),
    both depth parametrisations: composited rgb / DEPTH / opacity and every MLP parameter gradient, errors printed."""
    R, N = 1024, 128
    gen = torch.Generator().manual_seed(2024 + (param == "inverse"))
    p = syn.nerf_params(13)
    center = torch.randn(R, 3, generator=gen) * 0.1
    ray = torch.randn(R, 3, generator=gen) * 0.3 + torch.tensor([0., 0., 1.])
    u = torch.rand(1, R, N, 1, generator=gen)
    rng = [1.2, 5.2] if param == "metric" else [1, 0]
    depth = ora.stratified_depth(u, N, rng, param)[0, ..., 0]
    target = torch.rand(R, 3, generator=gen)
    prog, c2f = 0.3, [0.1, 0.5]
    ref = _oracle_c2(p, center, ray, depth, target, prog, c2f)
    for prec in ("bf16", "bf16x3"):
        flat = _flat(p).to(DEV).requires_grad_(True)
        c, r = center.to(DEV).requires_grad_(True), ray.to(DEV).requires_grad_(True)
        d = depth.to(DEV)
        rgb_s, sig_s = F.nerf_forward_samples(flat, c, r, d, prog, c2f, prec, training=True)
        rgb, dep, op, _ = F.composite(r, rgb_s, sig_s, d)
        ((rgb - target.to(DEV)) ** 2).mean().backward()
        torch.cuda.synchronize()
        tol = OUT_TOL[prec]
        e_rgb = (rgb.detach().cpu() - ref["rgb"]).abs().max().item()
        e_op = (op.detach().cpu() - ref["opacity"]).abs().max().item()
        dd = (dep.detach().cpu() - ref["depth"]).abs()
        e_dep = dd.max().item()
        dep_excess = (dd - tol["depth_rtol"] * ref["depth"].abs()).max().item()
        print("[%s %s] vs oracle: rgb %.3e  opacity %.3e  depth abs %.3e (max |depth| %.3e, excess over rtol %.3e)"
              % (prec, param, e_rgb, e_op, e_dep, ref["depth"].abs().max().item(), dep_excess))
        assert e_rgb <= tol["rgb"], (prec, "rgb", e_rgb)
        assert e_op <= tol["opacity"], (prec, "opacity", e_op)
        assert dep_excess <= tol["depth_atol"], (prec, "depth", e_dep, dep_excess)
        g = _split(flat.grad.detach().cpu(), p)
        worst = ("", 0.0)
        for k in _nerf_keys():
            a, b = g[k].double(), ref["grads"][k].double().reshape(-1)
            rel = ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
            worst = max(worst, (k, rel), key=lambda t: t[1])
            assert rel < (GRAD_TOL_FIRST_LAYER if k == "mlp_feat.0.weight" else GRAD_TOL), (prec, k, rel)
        gref = torch.cat([ref["grads"][k].double().reshape(-1) for k in _nerf_keys()])
        rel_all = ((flat.grad.detach().cpu().double() - gref).norm() / gref.norm()).item()
        print("[%s %s] whole MLP parameter gradient rel-L2 %.3e" % (prec, param, rel_all))
        assert rel_all < GRAD_TOL
        rel_c = ((c.grad.cpu().double() - ref["d_center"].double()).norm() / ref["d_center"].double().norm()).item()
        rel_r = ((r.grad.cpu().double() - ref["d_ray"].double()).norm() / ref["d_ray"].double().norm()).item()
        print("[%s %s] gradients vs oracle: worst MLP tensor %s rel-L2 %.3e (bar %.0e); d_center %.3e  d_ray %.3e "
              "(ill-conditioned input path, SURVEY.md H10: reported, bound 0.5)" % (prec, param, worst[0], worst[1], GRAD_TOL, rel_c, rel_r))
        assert rel_c < 0.5 and rel_r < 0.5


@pytest.mark.parametrize("R,N", [(4, 128), (37, 16), (64, 192), (333, 128)])
def test_tc_x3_forward_meets_1e3(F, R, N):
    """The split-precision tensor-core forward (row n1: the 1e-3 path, replacing the CUDA-core FP32 kernels in that role)
    on ragged shapes (partial tiles, a lone CTA pair, N not a multiple of 32), per-sample and composited, inference mode."""
    gen = torch.Generator().manual_seed(3 * R + N)
    p = syn.nerf_params(11)
    flat = _flat(p).to(DEV)
    center = torch.randn(R, 3, generator=gen) * 0.1
    ray = torch.randn(R, 3, generator=gen) * 0.3 + torch.tensor([0., 0., 1.])
    u = torch.rand(1, R, N, 1, generator=gen)
    depth = ora.stratified_depth(u, N, [1.2, 5.2], "metric")[0, ..., 0]
    prog, c2f = 0.3, [0.1, 0.5]
    d = depth[None, ..., None]
    pts, unit = ora.sample_points(center[None], ray[None], d)
    rgb_ref, sig_ref = ora.nerf_mlp(p, pts, unit, progress=prog, c2f=c2f)
    out_ref = ora.composite(ray[None], rgb_ref, sig_ref, d)
    rgb, sig = F.nerf_forward_samples(flat, center.to(DEV), ray.to(DEV), depth.to(DEV), prog, c2f, "bf16x3", training=False)
    e_rgb = (rgb.cpu() - rgb_ref[0]).abs().max().item()
    e_sig = ((sig.cpu() - sig_ref[0]).abs() / (1 + sig_ref[0].abs())).max().item()
    out = F.composite(ray.to(DEV), rgb, sig, depth.to(DEV))
    errs = [(a.cpu().reshape(-1) - b[0].reshape(-1)).abs().max().item() for a, b in zip(out[:3], out_ref[:3])]
    print("bf16x3 vs oracle R=%d N=%d: per-sample rgb %.3e sigma(rel) %.3e | composited rgb %.3e depth %.3e opacity %.3e"
          % (R, N, e_rgb, e_sig, *errs))
    assert e_rgb < 1e-3 and e_sig < 1e-3 and max(errs) < 1e-3


def test_tc_x3_training_records_feed_the_bf16_backward(F):
    """bf16x3 forward in training mode writes the same tile records as the BF16 forward (hi images, ReLU masks): its
    backward is the BF16 backward and must agree with the BF16 path's gradients up to the forward's rounding."""
    R, N = 256, 128
    gen = torch.Generator().manual_seed(99)
    p = syn.nerf_params(13)
    center = (torch.randn(R, 3, generator=gen) * 0.1).to(DEV)
    ray = (torch.randn(R, 3, generator=gen) * 0.3 + torch.tensor([0., 0., 1.])).to(DEV)
    depth = (torch.rand(R, N, generator=gen) * 4 + 1).sort(-1).values.to(DEV)
    w_rgb = (torch.rand(R, N, 3, generator=gen) - 0.5).to(DEV)
    grads = {}
    for prec in ("bf16", "bf16x3"):
        flat = _flat(p).to(DEV).requires_grad_(True)
        rgb, sig = F.nerf_forward_samples(flat, center, ray, depth, 0.3, [0.1, 0.5], prec, training=True)
        ((rgb * w_rgb).sum() + sig.sum() * 0.01).backward()
        grads[prec] = flat.grad.double()
    rel = ((grads["bf16"] - grads["bf16x3"]).norm() / grads["bf16x3"].norm()).item()
    print("bf16 vs bf16x3 parameter gradient rel-L2 %.3e" % rel)
    assert rel < 5e-2          # 256 rays, per-sample loss weights: the two forwards' ReLU masks differ in a few places (measured 2.2e-2)
