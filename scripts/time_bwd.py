"""Times the MLP backward (niw_nerf_bwd: dX chain + dW pass) alone, CUDA events, L2 flushed; env picks the schedule."""
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
N = 128
for R in [int(a) for a in sys.argv[1:]] or [1024]:
    g = torch.Generator().manual_seed(R)
    center = (torch.randn(R, 3, generator=g) * 0.1).to(dev)
    ray = (torch.randn(R, 3, generator=g) * 0.3 + torch.tensor([0., 0., 1.])).to(dev)
    depth = (torch.rand(R, N, generator=g) * 4 + 1).sort(-1).values.to(dev)
    tb = []
    for it in range(8):
        fl = flat0.clone().requires_grad_(True)
        c, r = center.clone().requires_grad_(True), ray.clone().requires_grad_(True)
        rgb, sig = F.nerf_forward_samples(fl, c, r, depth, 0.3, [0.1, 0.5], "bf16", training=True)
        loss = rgb.sum() + sig.sum()
        flush.zero_()
        with F.KernelTimer() as kt:
            loss.backward()
        tb.append(kt.totals()["nerf_bwd"][1])
    tb = sorted(tb[2:])
    b = tb[len(tb) // 2]
    print("R=%d CONCURRENT=%s DW_CTAS=%s: niw_nerf_bwd %.3f ms (%.0f TFLOP/s)" % (
        R, os.environ.get("NIW_BWD_CONCURRENT", "-"), os.environ.get("NIW_BWD_DW_CTAS", "-"), b, R * N * 2 * 1055744 / b / 1e9), flush=True)
