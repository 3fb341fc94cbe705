"""Standalone timing of the HBM-bound kernels at C5/C4 sizes: stratified sampler, inverse-CDF
sampler + merge, compositor fwd/bwd, pose raygen.  CUDA events on the launching stream, 256 MiB L2 flush
between launches, algorithmic bytes / time against MEASURED_PEAKS.json hbm_gbs.

    python scripts/micro_hbm.py [--json out.json] [--once]     # --once: one launch each (for ncu --set full)
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neural_invertible_warp_b200 import functional as F


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--once", action="store_true")
    ap.add_argument("--rays", type=int, default=65536 * 4)
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    res = measure(a.rays, a.reps, a.once, verbose=True)
    if a.json:
        json.dump(res, open(a.json, "w"), indent=1)


def measure(rays=65536 * 4, reps=10, once=False, verbose=False, dev="cuda:0"):
    """{kernel: {ms, bytes (algorithmic), gbs, frac of the measured HBM peak}} at `rays` rays."""
    class A:
        pass
    a = A()
    a.rays, a.reps, a.once = rays, reps, once
    peak = 6650.0
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["hbm_gbs"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    drain = torch.empty(256 << 20, dtype=torch.uint8, device=dev).zero_()

    def flush_l2():
        """Write a buffer larger than L2, then read another one: the second pass pushes the first one's dirty lines
        out to HBM, so the timed kernel starts on a cold AND clean L2 (otherwise up to 126 MB of write-backs from the
        flush itself land inside the timed kernel and are billed to it)."""
        flush.zero_()
        drain.view(torch.int32).sum()
    g = torch.Generator(device=dev).manual_seed(0)
    R, N, Nc, Nf = a.rays, 128, 64, 128
    S = R * N
    res = {}

    def run(name, fn, nbytes):
        reps = 1 if a.once else a.reps
        if not a.once:
            for _ in range(3):
                fn()
        ts = []
        for _ in range(reps):
            flush_l2()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        res[name] = dict(ms=ms, bytes=nbytes, gbs=nbytes / ms / 1e6, frac=nbytes / ms / 1e6 / peak)
        if verbose:
            print("%-22s %8.3f ms  %7.1f MB  %7.1f GB/s  %.3f of %.0f" % (name, ms, nbytes / 1e6, nbytes / ms / 1e6,
                                                                         nbytes / ms / 1e6 / peak, peak))

    # stratified: 8 B/sample
    u = torch.rand(S, device=dev, generator=g)
    run("stratified", lambda: F.sample_stratified(u, R, N, [1, 0], "inverse"), 8 * S)
    depth = F.sample_stratified(u, R, N, [1.2, 5.2], "metric")
    del u
    # compositor
    ray = torch.randn(R, 3, device=dev, generator=g)
    rgb_s = torch.rand(R, N, 3, device=dev, generator=g).requires_grad_(True)
    sigma = torch.rand(R, N, device=dev, generator=g).requires_grad_(True)
    out = [None]

    def comp_fwd():
        out[0] = F.composite(ray, rgb_s, sigma, depth)
    run("composite_fwd", comp_fwd, 24 * S + 32 * R)
    rgb, dpt, op, prob = out[0]
    go = (torch.ones_like(rgb), torch.ones_like(dpt), torch.ones_like(op))

    def comp_bwd():
        torch.autograd.grad((rgb, dpt, op), (rgb_s, sigma), go, retain_graph=True)
    run("composite_bwd", comp_bwd, 40 * S + 24 * R)
    del out, rgb, dpt, op, go, rgb_s, sigma
    # pdf sampler + merge (C3/C4 shape): 4(3N+Nf) B/ray, + 4 Nf when the fine samples are emitted too
    Rp = R * 2
    uc = torch.rand(Rp * Nc, device=dev, generator=g)
    dc = F.sample_stratified(uc, Rp, Nc, [1.2, 5.2], "metric")
    pdf = torch.rand(Rp, Nc, device=dev, generator=g)
    pdf = pdf / pdf.sum(-1, keepdim=True) * 0.9
    run("pdf_merge", lambda: F.sample_pdf_merge(pdf, dc, Nf, [1.2, 5.2], want_fine=False), 4 * (3 * Nc + Nf) * Rp)
    run("pdf_merge+fine", lambda: F.sample_pdf_merge(pdf, dc, Nf, [1.2, 5.2], want_fine=True),
        4 * (3 * Nc + 2 * Nf) * Rp)
    del uc, dc, pdf
    # raygen: full 480x640 frames, 32 images -> 24 B/ray out
    B, H, W = 32, 480, 640
    pose = torch.eye(3, 4, device=dev).repeat(B, 1, 1)
    intr = torch.tensor([[0.81 * W, 0, W / 2], [0, 0.81 * W, H / 2], [0, 0, 1]], device=dev).repeat(B, 1, 1)
    run("raygen_pose", lambda: F.raygen_pose(pose, intr, H, W), 24 * B * H * W)
    return dict(rays=R, samples_per_ray=N, pdf_shape=[Nc, Nf], peak_gbs=peak, kernels=res)


if __name__ == "__main__":
    main()
