"""Streaming-form vs slot-form training forward / backward of the fused MLP: outputs, gradients, timings (CUDA events)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neural_invertible_warp_b200 import functional as F, synthetic as syn

dev = "cuda:0"
keys = []
for i in range(8):
    keys += [f"mlp_feat.{i}.weight", f"mlp_feat.{i}.bias"]
for i in range(2):
    keys += [f"mlp_rgb.{i}.weight", f"mlp_rgb.{i}.bias"]
p = syn.nerf_params(1)
flat0 = torch.cat([p[k].reshape(-1) for k in keys]).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
sizes = [int(a) for a in sys.argv[1:]] or [37, 1024, 8192]
N = 128
for R in sizes:
    g = torch.Generator().manual_seed(R)
    center = (torch.randn(R, 3, generator=g) * 0.1).to(dev)
    ray = (torch.randn(R, 3, generator=g) * 0.3 + torch.tensor([0., 0., 1.])).to(dev)
    depth = (torch.rand(R, N, generator=g) * 4 + 1).sort(-1).values.to(dev)
    w_rgb = (torch.rand(R, N, 3, generator=g) - 0.5).to(dev)
    res = {}
    for mode in ("0", "1"):
        os.environ["NIW_FWD_STREAM"] = mode
        os.environ["NIW_BWD_STREAM"] = mode
        tf, tb = [], []
        for it in range(5):
            fl = flat0.clone().requires_grad_(True)
            c, r = center.clone().requires_grad_(True), ray.clone().requires_grad_(True)
            flush.zero_()
            a, b, e = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            a.record()
            rgb, sig = F.nerf_forward_samples(fl, c, r, depth, 0.3, [0.1, 0.5], "bf16", training=True)
            b.record()
            ((rgb * w_rgb).sum() + sig.sum() * 0.01).backward()
            e.record()
            torch.cuda.synchronize()
            tf.append(a.elapsed_time(b)); tb.append(b.elapsed_time(e))
        res[mode] = (rgb.detach(), sig.detach(), fl.grad.clone(), c.grad.clone(), r.grad.clone())
        S = R * N
        f, bb = sorted(tf[1:])[len(tf) // 2 - 1], sorted(tb[1:])[len(tb) // 2 - 1]
        print("R=%d stream=%s fwd %.3f ms (%.0f TFLOP/s)  bwd+glue %.3f ms (%.0f TFLOP/s)" %
              (R, mode, f, S * 1055744 / f / 1e9, bb, S * 2 * 1055744 / bb / 1e9), flush=True)
    names = ("rgb", "sigma", "d_params", "d_center", "d_ray")
    for n, x, y in zip(names, res["0"], res["1"]):
        d = (x.double() - y.double()).norm() / y.double().norm().clamp_min(1e-30)
        print("   %-9s slot vs stream rel-L2 %.3e  max abs %.3e  equal %s  finite %s" % (n, d.item(), (x - y).abs().max().item(), torch.equal(x, y), bool(torch.isfinite(y).all())))
