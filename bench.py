#!/usr/bin/env python
"""Benchmark of the B200-native NeRF ray-render hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--rays R] [--precision bf16|fp32]

Workload (BASELINE.json configs[1], "C2"): one ``barf_inn_llff`` train step -- NVP-warped ray
generation, stratified sampling, fused positional encoding + 8x256 MLP, compositing, MSE, full
backward, Adam -- on 1 024 rays x 128 samples per GPU (16 synthetic 480x640 LLFF-shaped images,
64 rays each, random-init weights).  With N GPUs every rank renders its own 1/N slice of a global
batch of N x 1 024 rays (weak scaling) and the gradients are all-reduced once per step.

Prints ONE JSON line (see DESIGN.md "Measurement" for every field):
  value        train rays/s, inputs resident in HBM, device RNG, per-step CUDA events (max over ranks)
  e2e          same metric through the public API with the step's host-side inputs (camera batch,
               host-drawn ray indices and stratified uniforms) copied from pinned memory every step
               and the loss read back
  roofline     the MLP kernels (the dominant kernels): algorithmic FLOP / CUDA-event time vs the
               measured dense-BF16 peak in MEASURED_PEAKS.json
  cpu_baseline the CPU oracle (oracle/reference_port.py, a restatement of the reference's PyTorch
               code pinned to goldens from the executed reference) on the host cores, bounded sample
``--impl reference`` times that CPU implementation alone (all host threads) and prints the same line shape.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# DRAM bytes of one niw_nerf_fwd + niw_nerf_bwd pair at 1 024 rays x 128 samples, from the committed ncu --set full
# capture profiles/r1_mlp_c2_ncu_full.md (dram__bytes_read.sum + dram__bytes_write.sum of tc_fwd / tc_dx / tc_dw)
MLP_DRAM_BYTES_C2 = int(580.98e6 + 604.24e6 + 1234.36e6)
MLP_FLOP_FWD = 2 * 527872              # per sample (un-padded dims)
MLP_FLOP_TRAIN = 3 * MLP_FLOP_FWD      # fwd + dX + dW
N_SAMPLES = 128
IMAGES = 16
H, W = 480, 640
METRIC = "train rays/sec (fwd+bwd, 128 samp/ray)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rays", type=int, default=1024, help="rays per GPU per step (C2: 1024; C5: 8192)")
    ap.add_argument("--precision", default=None, help="MLP operand precision: bf16 (tcgen05) or fp32 (CUDA cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of CUDA-graph replay")
    ap.add_argument("--no-micro", action="store_true", help="skip the standalone HBM-kernel timings (hbm_kernels)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"],
                    bf16_tflops_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------

class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    power_w_max=max(pw) if pw else None, samples=len(sm), reasons=sorted(reasons))


# ------------------------------------------------------------------------------------------------
# the CPU implementation (oracle port of the reference path)
# ------------------------------------------------------------------------------------------------

def cpu_train_step_factory(rays, seed=0):
    """C2 on the host: returns (step_fn, n_rays).  Everything inside is oracle/reference_port.py."""
    from neural_invertible_warp_b200 import synthetic as syn
    from oracle import reference_port as ora
    B, P = IMAGES, rays // IMAGES
    p = {k: v.requires_grad_(True) for k, v in syn.nerf_params(seed).items()}
    q = {k: v.requires_grad_(True) for k, v in syn.nvp_params(seed + 1).items()}
    code = syn.latent_codes(seed + 2, B).requires_grad_(True)
    intr = syn.intrinsics(B, H, W, 0.81)
    image = syn.images(seed + 3, B, H, W)
    cfg = dict(N=N_SAMPLES, Nf=None, range=[1, 0], param="inverse", L_3D=10, L_view=4, skip=(4,), c2f=[0.1, 0.5])
    opt_a = torch.optim.Adam(list(p.values()), lr=1e-3)
    opt_b = torch.optim.Adam(list(q.values()) + [code], lr=5e-4)

    def step():
        opt_a.zero_grad(); opt_b.zero_grad()
        ray_idx = torch.randperm(H * W)[:P]
        u = torch.rand(B, P, N_SAMPLES, 1)
        ray, center, *_ = ora.warped_rays(q, code, H, W, intr, ray_idx, 0.05)
        out = ora.render_rays(p, center, ray, u, cfg, progress=0.3)
        loss = ora.mse(out["rgb"], ora.gather_pixels(image, ray_idx))
        loss.backward()
        opt_a.step(); opt_b.step()
        return float(loss.detach())
    return step, B * P


def time_cpu(rays, steps, warmup):
    step, n = cpu_train_step_factory(rays)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return n / dt, dt


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (the oracle port: the
    reference is pure Python, it cannot be compiled into oracle/_ref, and /root/reference does not
    exist on the GPU box), all host threads, bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.manual_seed(0)
    steps = max(1, min(args.steps, 5))
    warm = max(1, min(args.warmup, 1))
    rate, dt = time_cpu(args.rays, steps, warm)
    cores = torch.get_num_threads()
    line = dict(impl="reference", metric=METRIC, value=rate, unit="rays/s", n_gpus=args.gpus, steps=steps, warmup=warm,
                ms_per_step=dt * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="fp32",
                data="synthetic", config=workload_config(args, "fp32"),
                cpu_baseline=dict(value=rate, unit="rays/s", cores=cores, kind="port",
                                  sample="%d steps of %d rays x %d samples (the full C2 step), torch CPU fp32, %d threads of %d host cores"
                                         % (steps, args.rays, N_SAMPLES, cores, os.cpu_count() or 0)),
                e2e=dict(value=rate, unit="rays/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def workload_config(args, precision):
    return dict(workload="barf_inn_llff train step (C2): NVP-warped raygen + stratified sampling + PE + 8x256 MLP + "
                         "composite + MSE, fwd+bwd + Adam; %d rays/GPU x %d samples, %d synthetic %dx%d images"
                         % (args.rays, N_SAMPLES, IMAGES, H, W),
                rays_per_gpu=args.rays, samples_per_ray=N_SAMPLES, images=IMAGES, mlp_precision=precision,
                parallelism="dp%d (rays sharded, one gradient all-reduce)" % args.gpus,
                l2="256 MiB written then 256 MiB read between steps (cold, clean L2), outside the per-step CUDA-event pairs",
                launch="one CUDA-graph replay per step (value and e2e)" if not args.no_graph else "eager")


# ------------------------------------------------------------------------------------------------
# ours
# ------------------------------------------------------------------------------------------------

def _leave(world):
    """End of a multi-rank run: NCCL teardown after captured graphs that hold collectives can block for minutes
    (ncclCommDestroy waits on the graphs' communicator references), so the ranks synchronise, flush and exit."""
    if world > 1:
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def run_ours(args):
    import torch.distributed as dist
    from neural_invertible_warp_b200 import _lib, config as cfgmod, engine, synthetic as syn
    from neural_invertible_warp_b200 import functional as F

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    _lib.load()
    precision = args.precision or "bf16"
    if _lib.load().niw_nerf_workspace_bytes(1, 1, F.precision_code(precision), 1) == 0:
        raise SystemExit("precision %s unavailable" % precision)

    rays_global = args.rays * world
    opt = cfgmod.builtin_options("barf_inn_llff", barf_c2f=[0.1, 0.5], device=dev,
                                 nerf=dict(rand_rays=rays_global, sample_intvs=N_SAMPLES),
                                 arch=dict(mlp_precision=precision))
    torch.manual_seed(0)
    graph = engine.build_graph(opt, IMAGES)
    # reference init leaves the warp at the identity (zero-init output layers): perturb like the
    # parity tests do so that every gradient path carries signal
    sd = {k: v.to(dev) for k, v in syn.nvp_params(1).items()}
    graph.warp_mlp.load_state_dict(sd)
    graph.warp_latent.weight.data = syn.latent_codes(2, IMAGES).to(dev)
    graph.nerf.progress.data.fill_(0.3)
    var_dev = engine.synthetic_var(opt, IMAGES, seed=3)
    # the reference's two optimisers (Adam + ExponentialLR on nerf, Adam on warp_mlp + warp_latent) as one flat
    # update kernel per group; the same object is the data-parallel gradient bucket
    adam = engine.FlatAdam(engine.reference_optimizer_groups(opt, graph))
    it = 5000
    P_local = (rays_global // IMAGES + world - 1) // world
    rays_local = P_local * IMAGES

    ray_draws = engine.device_ray_draws(dev, seed=20)      # same seed on every rank: one global draw, sliced per rank

    def step_device():
        v = cfgmod.AttrDict(var_dev)
        with ray_draws:
            loss = engine.train_step(opt, graph, v, it, bucket=adam, rank=rank, world=world)
        adam.step()
        return loss

    # ---- e2e leg: the step's host-side inputs (camera batch, pixel indices, stratified uniforms) wait in pinned
    # memory (a pool of `steps` batches drawn before the timed region, as a prefetching loader would hold them).
    # Every step copies its batch to one of TWO static device buffer sets on a copy stream (overlapping the previous
    # step's kernels), replays the CUDA graph captured over that set, and copies the loss to pinned memory; the host
    # reads each step's loss one step late, so it never drains the GPU ----
    gen = torch.Generator().manual_seed(1234 + rank)
    n_pool = max(args.steps, 1)
    pool_ridx = torch.stack([torch.randperm(H * W, generator=gen)[:P_local] for _ in range(n_pool)]).pin_memory()
    pool_u = torch.rand(n_pool, IMAGES, P_local, N_SAMPLES, 1, generator=gen).pin_memory()
    pin = dict(idx=torch.arange(IMAGES).pin_memory(), intr=var_dev.intr.cpu().pin_memory(),
               pose=var_dev.pose.cpu().pin_memory())

    def make_static():
        return dict(idx=torch.empty(IMAGES, dtype=torch.int64, device=dev), intr=torch.empty_like(var_dev.intr),
                    pose=torch.empty_like(var_dev.pose), ray_idx=torch.empty(P_local, dtype=torch.int64, device=dev),
                    u=torch.empty(IMAGES, P_local, N_SAMPLES, 1, device=dev))
    statics = [make_static(), make_static()]
    h2d = sum(t.numel() * t.element_size() for t in statics[0].values())
    loss_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
    copy_stream = torch.cuda.Stream()
    ready = [torch.cuda.Event() for _ in range(2)]       # the set's copies have landed
    consumed = [torch.cuda.Event() for _ in range(2)]    # the step reading the set (and its loss copy) has finished

    def make_step_static(static):
        def step_static():
            v = cfgmod.AttrDict(idx=static["idx"], intr=static["intr"], pose=static["pose"], image=var_dev.image)
            adam.zero()
            with engine.feed_draws(ray_idx=static["ray_idx"], u=static["u"]):
                v = graph.forward(opt, v, mode="train", iter=it)
            loss = engine.summarize_loss(opt, graph.compute_loss(opt, v, mode="train"))
            if world > 1:
                with engine.overlap_allreduce(graph, adam):
                    (loss.all * (1.0 / world)).backward()
                adam.allreduce()
            else:
                loss.all.backward()
            adam.step()
            return loss.all.detach()
        return step_static

    e2e_body = [make_step_static(statics[0]), make_step_static(statics[1])]

    def run_e2e(steps):
        """`steps` end-to-end steps; returns the losses the host read (all of them, each one step late)."""
        main = torch.cuda.current_stream()
        losses = []
        for i in range(steps):
            b = i & 1
            with torch.cuda.stream(copy_stream):
                if i >= 2:
                    copy_stream.wait_event(consumed[b])          # step i-2 is done with this buffer set
                st = statics[b]
                st["ray_idx"].copy_(pool_ridx[i % n_pool], non_blocking=True)
                st["u"].copy_(pool_u[i % n_pool], non_blocking=True)
                for k in ("idx", "intr", "pose"):
                    st[k].copy_(pin[k], non_blocking=True)
                ready[b].record(copy_stream)
            main.wait_event(ready[b])
            loss = e2e_body[b]()
            loss_host[b].copy_(loss, non_blocking=True)
            consumed[b].record(main)
            if i >= 1:
                consumed[1 - b].synchronize()                    # the user reads the previous step's loss
                losses.append(float(loss_host[1 - b]))
        if steps:
            consumed[(steps - 1) & 1].synchronize()
            losses.append(float(loss_host[(steps - 1) & 1]))
        return losses

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    drain = torch.empty(256 << 20, dtype=torch.uint8, device=dev).zero_()
    sync_token = torch.zeros(1, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps, with_events=True):
        """K steps, each bracketed by CUDA events on the launching stream, L2 flushed in between."""
        evs = []
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            # L2 flush outside the event pair: write 256 MiB, then read another 256 MiB so that the flush's own dirty lines
            # are written back before the step starts (cold and clean L2)
            flush.zero_()
            drain.view(torch.int32).sum()
            if world > 1:
                # the flush is rank-local work outside the timed region: re-align the ranks on the device before the start
                # event, or its jitter shows up inside the step as time spent waiting in the gradient all-reduce
                dist.all_reduce(sync_token)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); step_fn(); b.record()
            evs.append((a, b))
        barrier()
        wall = time.perf_counter() - t0
        ms = sum(a.elapsed_time(b) for a, b in evs)
        return ms, wall

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    # ---- warm-up ----
    for _ in range(max(args.warmup, 3)):
        step_device()
    run_e2e(2)
    barrier()
    step_value, graphed = step_device, False
    if not args.no_graph:
        try:
            step_value = engine.CapturedStep(step_device, warmup=1)
            graphed = True
            for _ in range(2):
                step_value()
        except Exception as e:   # capture refused (e.g. a collective that cannot be captured): stay eager
            if rank == 0:
                print("bench.py: CUDA-graph capture unavailable (%s); timing eager launches" % str(e).splitlines()[0],
                      file=sys.stderr)
            step_value, graphed = step_device, False
            torch.cuda.synchronize()
        if graphed:
            eager_bodies = list(e2e_body)
            try:
                for b in range(2):
                    e2e_body[b] = engine.CapturedStep(eager_bodies[b], warmup=1)
                run_e2e(2)
            except Exception as e:
                if rank == 0:
                    print("bench.py: e2e graph capture unavailable (%s)" % str(e).splitlines()[0], file=sys.stderr)
                e2e_body[:] = eager_bodies
                torch.cuda.synchronize()
    barrier()

    # ---- value: device-resident ----
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms_total, _ = timed(step_value, args.steps)
    ms_total = max_over_ranks(ms_total)
    # per-kernel CUDA-event timing (roofline leg) and the launch count: the same K steps launched eagerly
    # (event records cannot live inside a captured graph)
    n0 = _lib.launch_count()
    with F.KernelTimer() as kt:
        ms_eager, _ = timed(step_device, args.steps)
        kernel_ms = kt.totals()
    launches = _lib.launch_count() - n0
    # ---- e2e: host buffers in, loss out (wall clock around synchronised steps) ----
    barrier()
    t0 = time.perf_counter()
    e2e_losses = run_e2e(args.steps)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    assert len(e2e_losses) == args.steps and all(l == l for l in e2e_losses), "e2e leg: every step's loss must be read"
    clk = clocks.stop() if rank == 0 else None

    ms_per_step = ms_total / args.steps
    rays_per_step = rays_local * world
    value = rays_per_step / (ms_per_step * 1e-3)
    e2e_value = rays_per_step * args.steps / e2e_s

    # ---- roofline of the dominant kernels (the MLP) ----
    pk = peaks()
    mlp_calls = kernel_ms.get("nerf_fwd", (0, 0.0))[0]
    mlp_ms = kernel_ms.get("nerf_fwd", (0, 0.0))[1] + kernel_ms.get("nerf_bwd", (0, 0.0))[1]
    flop_per_step = MLP_FLOP_TRAIN * rays_local * N_SAMPLES
    achieved = flop_per_step * mlp_calls / (mlp_ms * 1e-3) / 1e12 if mlp_ms > 0 else 0.0
    comp_ms = kernel_ms.get("composite_fwd", (0, 0.0))[1] + kernel_ms.get("composite_bwd", (0, 0.0))[1]
    traffic = MLP_DRAM_BYTES_C2 if (rays_local == 1024 and precision == "bf16") else None
    mlp_ms_call = mlp_ms / max(mlp_calls, 1)
    roof = dict(bound="tensor", kernel="niw_nerf_fwd + niw_nerf_bwd (fused PE + 8x256 MLP, fwd + dX + dW)",
                achieved=achieved, peak=pk["bf16_tflops_sustained"], unit="TFLOP/s",
                frac=achieved / pk["bf16_tflops_sustained"], traffic=traffic,
                traffic_note="DRAM bytes per launch pair from profiles/r1_mlp_c2_ncu_full.md (the saved bf16 activation / "
                             "gradient tile images; algorithmic operand bytes are 1.1 MB of weights)",
                hbm_view=(dict(gbs=traffic / (mlp_ms_call * 1e-3) / 1e9, frac=traffic / (mlp_ms_call * 1e-3) / 1e9 / pk["hbm_gbs"],
                               peak=pk["hbm_gbs"]) if traffic and mlp_ms_call > 0 else None),
                peak_source=pk["source"] + " (sustained bf16)",
                flop_per_launch_pair=flop_per_step, mlp_ms_per_step=mlp_ms / max(mlp_calls, 1),
                mlp_share_of_step=mlp_ms / ms_total if ms_total > 0 else None,
                timed="CUDA events around niw_nerf_fwd / niw_nerf_bwd on the launching stream, eager pass of the same "
                      "%d steps (%.3f ms/step eager); the BF16 weight packing (2 kernels, ~11 us) runs earlier on a side "
                      "stream (niw_nerf_pack) and is outside this bracket" % (args.steps, ms_eager / args.steps),
                composite_ms_per_step=comp_ms / max(mlp_calls, 1))

    if rank != 0:
        _leave(world)
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        torch.manual_seed(0)
        rate, dt = time_cpu(args.rays, 3, 1)
        cpu = dict(value=rate, unit="rays/s", cores=torch.get_num_threads(), kind="port",
                   sample="3 steps (after 1 warm-up) of the same C2 step, %d rays x %d samples, oracle/reference_port.py on "
                          "torch CPU fp32, %.2f s/step" % (args.rays, N_SAMPLES, dt))

    hbm = None
    if world == 1 and not args.no_micro:
        # sampler / compositor / raygen alone at 262 144 rays (inputs larger than L2, L2 flushed between launches):
        # algorithmic bytes / CUDA-event time against the measured HBM copy bandwidth
        del flush, drain
        torch.cuda.empty_cache()
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import micro_hbm
        m = micro_hbm.measure(verbose=False, dev=dev)
        hbm = dict(rays=m["rays"], peak_gbs=m["peak_gbs"], peak_source=pk["source"],
                   kernels={k: dict(gbs=round(v["gbs"], 1), frac=round(v["frac"], 3), ms=round(v["ms"], 4), bytes=v["bytes"])
                            for k, v in m["kernels"].items()})

    line = dict(metric=METRIC, value=value, unit="rays/s", n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype=("bf16" if precision == "bf16" else "fp32"), data="synthetic",
                config=workload_config(args, precision), mlp_evals_per_s=value * N_SAMPLES,
                e2e=dict(value=e2e_value, unit="rays/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=4,
                         ms_per_step=e2e_s / args.steps * 1e3,
                         how="public Graph API; host batch in pinned memory -> H2D on a copy stream into one of two static "
                             "buffer sets -> CUDA-graph replay -> loss to pinned memory; the host reads every step's loss, "
                             "one step late; wall clock over all steps"),
                gpu_launches=launches, roofline=roof, hbm_kernels=hbm, cpu_baseline=cpu, clocks=clk)
    print(json.dumps(line))
    _leave(world)


def _protect_stdout():
    """The driver parses ONE JSON line from stdout: libraries that chat on fd 1 (NCCL prints its version there) are sent
    to stderr, and only ``print`` from this script reaches the real stdout."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w", buffering=1)


if __name__ == "__main__":
    _protect_stdout()
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
