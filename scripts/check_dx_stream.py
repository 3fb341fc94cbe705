"""Streaming continuation of the dX chain (NIW_DX_STREAM=1) vs the slot-form chain: d_center / d_ray and timing of niw_nerf_bwd_dx."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
from neural_invertible_warp_b200 import functional as F, synthetic as syn, _lib

dev = "cuda:0"
keys = []
for i in range(8):
    keys += [f"mlp_feat.{i}.weight", f"mlp_feat.{i}.bias"]
for i in range(2):
    keys += [f"mlp_rgb.{i}.weight", f"mlp_rgb.{i}.bias"]
p = syn.nerf_params(1)
flat = torch.cat([p[k].reshape(-1) for k in keys]).to(dev)
lib = _lib.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
N = 128
P = lambda t: C.c_void_p(t.data_ptr())
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
for R in [int(a) for a in sys.argv[1:]] or [37, 300, 1024]:
    g = torch.Generator().manual_seed(R)
    center = (torch.randn(R, 3, generator=g) * 0.1).to(dev)
    ray = (torch.randn(R, 3, generator=g) * 0.3 + torch.tensor([0., 0., 1.])).to(dev)
    depth = (torch.rand(R, N, generator=g) * 4 + 1).sort(-1).values.to(dev)
    d_rgb = (torch.rand(R, N, 3, generator=g) - 0.5).to(dev)
    d_sig = (torch.rand(R, N, generator=g) * 0.01).to(dev)
    prog = torch.tensor([0.3], device=dev)
    nb = lib.niw_nerf_workspace_bytes(R, N, 1, 1)
    res = {}
    for mode in ("0", "1"):
        os.environ["NIW_DX_STREAM"] = mode
        ws = torch.zeros(nb, dtype=torch.uint8, device=dev)
        rgb, sig = torch.empty(R, N, 3, device=dev), torch.empty(R, N, device=dev)
        _lib.check(lib.niw_nerf_fwd(P(flat), P(center), P(ray), P(depth), R, N, P(prog), 0.1, 0.5, 1, 1, P(ws), nb, P(rgb), P(sig), st()))
        ts = []
        for it in range(5):
            dP = torch.zeros_like(flat); dc = torch.empty_like(center); dr = torch.empty_like(ray)
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _lib.check(lib.niw_nerf_bwd_dx(P(flat), P(center), P(ray), P(depth), R, N, 1, P(ws), nb, P(d_rgb), P(d_sig), P(dP), P(dc), P(dr), st()))
            b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        res[mode] = (dc.clone(), dr.clone())
        print("R=%d stream=%s niw_nerf_bwd_dx %.3f ms" % (R, mode, sorted(ts[1:])[1]), flush=True)
    for n, x, y in zip(("d_center", "d_ray"), res["0"], res["1"]):
        rel = ((x.double() - y.double()).norm() / x.double().norm()).item()
        print("   %-8s slot vs stream rel-L2 %.3e finite %s" % (n, rel, bool(torch.isfinite(y).all())))
