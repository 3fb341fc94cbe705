"""One training forward of the fused MLP with NIW_STREAM_DEBUG=1: prints CTA 0's hand-off time stamps."""
import sys, os
os.environ["NIW_STREAM_DEBUG"] = "1"
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
flat = torch.cat([p[k].reshape(-1) for k in keys]).to(dev).requires_grad_(True)
R, N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, 128
g = torch.Generator().manual_seed(R)
center = (torch.randn(R, 3, generator=g) * 0.1).to(dev)
ray = (torch.randn(R, 3, generator=g) * 0.3 + torch.tensor([0., 0., 1.])).to(dev)
depth = (torch.rand(R, N, generator=g) * 4 + 1).sort(-1).values.to(dev)
for it in range(2):
    print("---- call", it, file=sys.stderr, flush=True)
    rgb, sig = F.nerf_forward_samples(flat, center, ray, depth, 0.3, [0.1, 0.5], "bf16", training=True)
    torch.cuda.synchronize()
