"""The other BASELINE.json configurations on one GPU (bench.py measures configs[1] = C2):

  c3   barf_inn_dtu train step, 300x400 DTU-shaped, 32 images x 64 rays, coarse 64 + fine 128 samples, fwd + bwd
  c4   full-frame eval render 480x640 x (64 + 128) samples (one rank's share with --rows-of K: rows [0, H/K))
  c5   the per-GPU share of the 65 536-ray data-parallel batch: 8 192 rays x 128 samples (same as bench.py --rays 8192)

Each prints one JSON line: rays/s, sample-MLP-evals/s and the MLP kernels' share of the measured dense-BF16 peak
(CUDA events around niw_nerf_fwd / niw_nerf_bwd, L2 flushed between steps).

    python scripts/bench_configs.py c3 c4 [--steps 5] [--rows-of 1]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from neural_invertible_warp_b200 import config as cfgmod, engine, functional as F, synthetic as syn

MLP_FLOP = 2 * 527872
DEV = "cuda:0"
DRAIN = None


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", d["bf16_tflops"]), d["bf16_tflops"], "measured"
    return 1400.0, 1590.0, "fallback"


def timed(fn, steps, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ms = []
    with F.KernelTimer() as kt:
        for _ in range(steps):
            flush.zero_()
            DRAIN.view(torch.int32).sum()      # read pass over another 256 MiB: drains the flush's dirty lines
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        k = kt.totals()
    ms.sort()
    return ms[len(ms) // 2], {n: (c, t / max(c, 1) * (c / steps)) for n, (c, t) in k.items()}


def c3(args, flush):
    B, P, N, Nf, H, W = 32, 64, 64, 128, 300, 400
    opt = cfgmod.builtin_options("barf_inn_dtu", barf_c2f=[0.1, 0.5], device=DEV, data=dict(image_size=[H, W]),
                                 nerf=dict(rand_rays=P * B, sample_intvs=N, fine_sampling=True, sample_intvs_fine=Nf,
                                           depth=dict(range=[1.2, 5.2])),
                                 loss_weight=dict(render_fine=0), arch=dict(mlp_precision="bf16"))
    var0 = engine.synthetic_var(opt, B, 3, dtu=True)
    graph = engine.build_graph(opt, B, initial_poses_w2c=var0.pose.clone())
    graph.nerf.progress.data.fill_(0.3); graph.nerf_fine.progress.data.fill_(0.3)
    graph.pose_net.pose_latent.weight.data = syn.latent_codes(2, B).to(DEV)
    graph.pose_net.pose_embedding.load_state_dict({k: v.to(DEV) for k, v in syn.nvp_params(1).items()})

    def step():
        engine.train_step(opt, graph, cfgmod.AttrDict(var0), 5000)
    ms, k = timed(step, args.steps, flush)
    rays = B * P
    evals = rays * (N + N + Nf)                      # coarse net on 64, fine net on 64 + 128
    out = dict(config="c3", workload="barf_inn_dtu train step, 300x400, %d images x %d rays, %d coarse + %d fine samples, fwd+bwd (eager launches)" % (B, P, N, Nf),
               rays=rays, ms_per_step=ms, rays_per_s=rays / ms * 1e3, mlp_evals_per_s=evals / ms * 1e3,
               mlp_flop=3 * MLP_FLOP * evals, kernels=k)
    # the same step as ONE CUDA-graph replay (device RNG, depth range read on the device, Kabsch fit in a kernel: no host sync)
    try:
        draws = engine.device_ray_draws(DEV, seed=7)

        def step_dev():
            with draws:
                engine.train_step(opt, graph, cfgmod.AttrDict(var0), 5000)
        captured = engine.CapturedStep(step_dev, warmup=2)
        ms_g, _ = timed(captured, args.steps, flush)
        out.update(ms_per_step_graph=ms_g, rays_per_s_graph=rays / ms_g * 1e3)
    except Exception as e:          # report, do not hide
        out.update(graph_capture_error=str(e).splitlines()[0])
    return out


def c4(args, flush):
    B, H, W, N, Nf = 1, 480, 640, 64, 128
    rows = H // args.rows_of
    opt = cfgmod.builtin_options("barf_inn_dtu", barf_c2f=[0.1, 0.5], device=DEV, data=dict(image_size=[H, W]),
                                 nerf=dict(rand_rays=rows * W, sample_intvs=N, fine_sampling=True, sample_intvs_fine=Nf,
                                           depth=dict(range=[1.2, 5.2])),
                                 loss_weight=dict(render_fine=0), arch=dict(mlp_precision="bf16"))
    var0 = engine.synthetic_var(opt, B, 3, dtu=True)
    graph = engine.build_graph(opt, B, initial_poses_w2c=var0.pose.clone())
    graph.nerf.progress.data.fill_(0.3); graph.nerf_fine.progress.data.fill_(0.3)

    def step():
        with torch.no_grad():
            # rank r of K renders pixel rows [r H/K, (r+1) H/K): a contiguous ray_idx range, no collective
            graph._render_pose(opt, var0.pose, intr=var0.intr, mode="eval", depth_range=[1.2, 5.2], idx_start=0,
                               num=rows * W)
    ms, k = timed(step, args.steps, flush)
    rays = rows * W
    evals = rays * (N + N + Nf)
    return dict(config="c4", workload="eval render of pixel rows [0, %d) of one 480x640 frame (1/%d of the frame), %d coarse + %d fine samples, no_grad, one call" % (rows, args.rows_of, N, Nf),
                rays=rays, ms_per_step=ms, rays_per_s=rays / ms * 1e3, mlp_evals_per_s=evals / ms * 1e3,
                mlp_flop=MLP_FLOP * evals, kernels=k)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=["c3", "c4"])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--rows-of", type=int, default=1)
    args = ap.parse_args()
    global DRAIN
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    DRAIN = torch.zeros(256 << 20, dtype=torch.uint8, device=DEV)
    sus, burst, src = peak()
    for name in args.configs:
        r = dict(c3=c3, c4=c4)[name](args, flush)
        mlp_ms = sum(t for n, (c, t) in r["kernels"].items() if n in ("nerf_fwd", "nerf_bwd"))
        r["mlp_ms_per_step"] = mlp_ms
        r["mlp_tflops"] = r.pop("mlp_flop") / (mlp_ms * 1e-3) / 1e12 if mlp_ms > 0 else None
        r["mlp_frac_of_sustained_peak"] = r["mlp_tflops"] / sus if mlp_ms > 0 else None
        r["peak"] = dict(bf16_tflops_sustained=sus, bf16_tflops_burst=burst, source=src)
        r["kernels"] = {n: dict(calls_per_step=c / args.steps, ms_per_step=t) for n, (c, t) in r["kernels"].items()}
        print(json.dumps(r))


if __name__ == "__main__":
    main()
