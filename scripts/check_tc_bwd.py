"""Per-tensor comparison of the BF16 tensor-core backward with the FP32 CUDA-core backward."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neural_invertible_warp_b200 import functional as F, synthetic as syn

DEV = "cuda:0"
keys = []
for i in range(8):
    keys += [f"mlp_feat.{i}.weight", f"mlp_feat.{i}.bias"]
for i in range(2):
    keys += [f"mlp_rgb.{i}.weight", f"mlp_rgb.{i}.bias"]
R, N = int(sys.argv[1]) if len(sys.argv) > 1 else 64, int(sys.argv[2]) if len(sys.argv) > 2 else 128
gen = torch.Generator().manual_seed(R * N)
p = syn.nerf_params(13)
flat0 = torch.cat([p[k].reshape(-1) for k in keys]).to(DEV)
center0 = (torch.randn(R, 3, generator=gen) * 0.1).to(DEV)
ray0 = (torch.randn(R, 3, generator=gen) * 0.3 + torch.tensor([0., 0., 1.])).to(DEV)
u = torch.rand(1, R, N, 1, generator=gen)
depth = F.sample_stratified(u.reshape(-1).to(DEV), R, N, [1.2, 5.2], "metric")      # the product sampler (bit-exact with the reference)
prog, c2f = 0.3, [0.1, 0.5]
target = torch.rand(R, 3, generator=gen).to(DEV)
mode = sys.argv[3] if len(sys.argv) > 3 else "mse"
w_rgb = (torch.rand(R, N, 3, generator=gen) - 0.5).to(DEV)
w_sig = (torch.rand(R, N, generator=gen) - 0.5).to(DEV)
grads = {}
for prec in ("fp32", "bf16"):
    flat = flat0.clone().requires_grad_(True)
    c = center0.clone().requires_grad_(True)
    r = ray0.clone().requires_grad_(True)
    rgb_s, sig_s = F.nerf_forward_samples(flat, c, r, depth, prog, c2f, prec, training=True)
    if mode == "mse":
        rgb, dep, op, _ = F.composite(r, rgb_s, sig_s, depth)
        ((rgb - target) ** 2).mean().backward()
    else:
        ((rgb_s * w_rgb).sum() + (sig_s * w_sig).sum()).backward()
    torch.cuda.synchronize()
    grads[prec] = (flat.grad.clone(), c.grad.clone(), r.grad.clone())
off = 0
for k in keys:
    n = p[k].numel()
    a, b = grads["bf16"][0][off:off + n].double(), grads["fp32"][0][off:off + n].double()
    off += n
    rel = ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
    print("d%-20s rel-L2 %.3e  |ref| %.3e |tc| %.3e" % (k, rel, b.norm().item(), a.norm().item()))
for name, i in (("d_center", 1), ("d_ray", 2)):
    a, b = grads["bf16"][i].double(), grads["fp32"][i].double()
    print("%-21s rel-L2 %.3e  |ref| %.3e |tc| %.3e" % (name, ((a - b).norm() / b.norm()).item(), b.norm().item(), a.norm().item()))
