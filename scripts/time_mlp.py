"""Times the fused MLP entry points alone (CUDA events, L2 flushed between launches)."""
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
flat = torch.cat([p[k].reshape(-1) for k in keys]).to(dev)
prog, c2f = 0.3, [0.1, 0.5]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
modes = sys.argv[1:] or ["bf16"]
for prec in modes:
    for R in (1024, 8192):
        N = 128
        g = torch.Generator().manual_seed(0)
        center = (torch.randn(R, 3, generator=g) * 0.1).to(dev)
        ray = (torch.randn(R, 3, generator=g) * 0.3 + torch.tensor([0., 0., 1.])).to(dev)
        depth = (torch.rand(R, N, generator=g) * 4 + 1).sort(-1).values.to(dev)
        for training in (False, True):
            if training:
                fl = flat.clone().requires_grad_(True)
                c, r = center.clone().requires_grad_(True), ray.clone().requires_grad_(True)
            else:
                fl, c, r = flat, center, ray
            tf, tb = [], []
            for it in range(6):
                flush.zero_()
                a, b, e = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                a.record()
                rgb, sig = F.nerf_forward_samples(fl, c, r, depth, prog, c2f, prec, training=training)
                b.record()
                if training:
                    try:
                        (rgb.sum() + sig.sum()).backward()
                    except RuntimeError as ex:
                        print("bwd unsupported:", ex); training = None
                e.record()
                torch.cuda.synchronize()
                tf.append(a.elapsed_time(b)); tb.append(b.elapsed_time(e))
                if training is None:
                    break
            tf, tb = sorted(tf[1:]), sorted(tb[1:])
            S = R * N
            f = tf[len(tf) // 2] if tf else float("nan")
            msg = "%s R=%d training=%s fwd %.3f ms (%.1f TFLOP/s)" % (prec, R, training, f, S * 1055744 / f / 1e9)
            if training:
                bb = tb[len(tb) // 2]
                msg += " bwd(+autograd glue) %.3f ms (%.1f TFLOP/s)" % (bb, S * 2 * 1055744 / bb / 1e9)
            print(msg)
