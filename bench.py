#!/usr/bin/env python
"""Benchmark of the B200-native NeRF ray-render hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c4|c5]
                    [--rays R] [--precision bf16|bf16x3|fp32]

Workloads (BASELINE.json ``configs``):
  c2 (default, configs[1])  one ``barf_inn_llff`` train step -- NVP-warped ray generation, stratified sampling,
       fused positional encoding + 8x256 MLP, compositing, MSE, full backward, Adam -- on 1 024 rays x 128
       samples per GPU (16 synthetic 480x640 LLFF-shaped images, 64 rays each, random-init weights).  With N
       GPUs every rank renders its own 1/N slice of a global batch of N x 1 024 rays (weak scaling) and the
       gradients are all-reduced once per step.
  c5 (configs[4])  the same step at 8 192 rays per GPU (65 536 rays on 8 GPUs).
  c4 (configs[3])  full-frame ``mode="eval"`` render of one 480x640 view, 64 coarse + 128 fine samples per ray;
       rank r renders pixel rows [r H/N, (r+1) H/N), no collective in the timed region (strong scaling).

Prints ONE JSON line (see DESIGN.md "Measurement" for every field):
  value        rays/s, inputs resident in HBM, device RNG, per-step CUDA events (max over ranks)
  e2e          same metric through the public API with the step's host-side inputs copied from pinned memory every
               step and the result (loss / rendered rows) read back
  roofline     the MLP kernels (the dominant kernels): algorithmic FLOP / CUDA-event time vs the measured dense-BF16
               peak in MEASURED_PEAKS.json (burst figure; the fraction of the sustained figure beside it)
  cpu_baseline the CPU oracle (oracle/reference_port.py, a restatement of the reference's PyTorch code pinned to
               goldens from the executed reference) on the host cores, bounded sample
``--impl reference`` times that CPU implementation alone (all host threads) and prints the same line shape.
"""
import argparse
import contextlib
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

MLP_FLOP_FWD = 2 * 527872              # per sample (un-padded dims)
MLP_FLOP_TRAIN = 3 * MLP_FLOP_FWD      # fwd + dX + dW
N_SAMPLES = 128
IMAGES = 16
H, W = 480, 640
C4_N, C4_NF = 64, 128
METRIC_TRAIN = "train rays/sec (fwd+bwd, 128 samp/ray)"
METRIC_EVAL = "eval rays/sec (full-frame 480x640, 64+128 samp/ray, rows sharded)"
DEFAULT_RAYS = dict(c2=1024, c5=8192)
# DRAM bytes of the MLP kernels per step, parsed from this round's committed ncu --set full capture by
# scripts/ncu_mlp.py (profiles/r2_mlp_traffic.json: {"c2": {"bf16": bytes, ...}, ...}); absent -> traffic null
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r2_mlp_traffic.json")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c4", "c5"])
    ap.add_argument("--rays", type=int, default=None, help="rays per GPU per step (c2: 1024; c5: 8192)")
    ap.add_argument("--precision", default=None,
                    help="MLP operand precision: bf16 (tcgen05), bf16x3 (tcgen05, hi+lo split operands: the 1e-3 path) "
                         "or fp32 (CUDA cores).  Default: bf16 for the train steps, bf16x3 for the eval render")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of CUDA-graph replay")
    ap.add_argument("--no-micro", action="store_true", help="skip the standalone HBM-kernel timings (hbm_kernels)")
    ap.add_argument("--timeline", default=None, metavar="PREFIX",
                    help="after the timed region, record ONE more step under the CUPTI activity tracer and write every rank's "
                         "kernel timeline (stream, start, duration) to PREFIX.rank<r>.txt (not a bench value)")
    ap.add_argument("--no-loss-check", action="store_true", help="N>1: skip the global-loss check against one rank")
    a = ap.parse_args()
    if a.rays is None:
        a.rays = DEFAULT_RAYS.get(a.config, 0)
    return a


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"],
                    bf16_tflops_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


def committed_traffic(config, precision):
    try:
        with open(TRAFFIC_FILE) as f:
            return json.load(f)[config][precision]
    except (OSError, KeyError, ValueError):
        return None


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


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
# workload description (identical in both arms: it names the work, not how an arm runs it)
# ------------------------------------------------------------------------------------------------

def workload_config(args):
    if args.config == "c4":
        return dict(workload="c4: full-frame eval render of one %dx%d view, %d coarse + %d fine samples per ray (two "
                             "8x256 networks, inverse-CDF resampling, compositing), no_grad; pixel rows sharded over the GPUs"
                             % (H, W, C4_N, C4_NF),
                    rays_per_frame=H * W, samples_per_ray=[C4_N, C4_N + C4_NF], images=1,
                    parallelism="rows [r*H/N, (r+1)*H/N) on rank r of N=%d, no collective" % args.gpus,
                    l2="GPU arm: every step streams more than the L2 holds (>= 0.9 GB of per-sample tensors per rank) and "
                       "256 MiB is written then read between steps; CPU arm: not applicable")
    return dict(workload="%s: barf_inn_llff train step: NVP-warped raygen + stratified sampling + PE + 8x256 MLP + composite "
                         "+ MSE, fwd+bwd + Adam; %d rays/GPU x %d samples, %d synthetic %dx%d images"
                         % (args.config, args.rays, N_SAMPLES, IMAGES, H, W),
                rays_per_gpu=args.rays, samples_per_ray=N_SAMPLES, images=IMAGES,
                parallelism="dp%d (rays sharded, gradients all-reduced once per step)" % args.gpus,
                l2="GPU arm: 256 MiB written then 256 MiB read between steps (cold, clean L2), outside the per-step "
                   "CUDA-event pairs; CPU arm: not applicable")


# ------------------------------------------------------------------------------------------------
# the CPU implementation (oracle port of the reference path)
# ------------------------------------------------------------------------------------------------

def cpu_train_step_factory(rays, seed=0):
    """The c2 / c5 step on the host: returns (step_fn, n_rays).  Everything inside is oracle/reference_port.py."""
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


def cpu_eval_step_factory(rows, seed=0):
    """A bounded sample of the c4 render on the host: ``rows`` pixel rows of the frame (same pipeline: raygen from
    the pose, stratified depths, coarse network, compositing, inverse-CDF resampling + merge, fine network)."""
    from neural_invertible_warp_b200 import synthetic as syn
    from oracle import reference_port as ora
    p, pf = syn.nerf_params(seed), syn.nerf_params(seed + 7)
    intr = syn.intrinsics(1, H, W, 0.81)
    pose = syn.dtu_poses(seed + 4, 1)
    cfg = dict(N=C4_N, Nf=C4_NF, range=[1.2, 5.2], param="metric", L_3D=10, L_view=4, skip=(4,), c2f=[0.1, 0.5])
    n = rows * W

    @torch.no_grad()
    def step():
        ray_idx = torch.arange(n)
        center, ray = ora.center_and_ray(H, W, pose, intr, ray_idx)
        u = torch.rand(1, n, C4_N, 1)
        out = ora.render_rays(p, center, ray, u, cfg, progress=0.3, nerf_fine_p=pf)
        return float(out["rgb_fine"].mean())
    return step, n


def time_cpu(make, steps, warmup):
    torch.set_num_threads(host_threads())          # torchrun exports OMP_NUM_THREADS=1: the CPU arm uses every host core
    step, n = make()
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return n / dt, dt


def cpu_sample(args):
    """(factory, description) of the CPU arm's step for this config."""
    if args.config == "c4":
        rows = 4
        return (lambda: cpu_eval_step_factory(rows)), ("%d pixel rows (%d rays x (%d + %d) MLP evals) of the 480x640 frame per "
                                                       "step" % (rows, rows * W, C4_N, C4_N + C4_NF))
    return (lambda: cpu_train_step_factory(args.rays)), ("the full step: %d rays x %d samples, fwd+bwd+Adam" % (args.rays, N_SAMPLES))


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (the oracle port: the reference is pure
    Python, it cannot be compiled into oracle/_ref, and /root/reference does not exist on the GPU box), all host
    threads, exactly --steps timed steps after --warmup untimed ones, each a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.manual_seed(0)
    make, what = cpu_sample(args)
    rate, dt = time_cpu(make, args.steps, args.warmup)
    cores = torch.get_num_threads()
    line = dict(impl="reference", metric=METRIC_EVAL if args.config == "c4" else METRIC_TRAIN, value=rate, unit="rays/s",
                n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=dt * 1e3, higher_is_better=True,
                scaling="strong" if args.config == "c4" else "weak", vs_baseline=None, dtype="fp32", data="synthetic",
                config=workload_config(args), launch="torch CPU eager, fp32, %d threads" % cores,
                cpu_baseline=dict(value=rate, unit="rays/s", cores=cores, kind="port",
                                  sample="%d steps after %d warm-up, each %s; oracle/reference_port.py on torch CPU fp32, %d "
                                         "threads of %d host cores, %.2f s/step"
                                         % (args.steps, args.warmup, what, cores, os.cpu_count() or 0, dt)),
                e2e=dict(value=rate, unit="rays/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# ours: shared helpers
# ------------------------------------------------------------------------------------------------

def _leave(world):
    """End of a multi-rank run: NCCL teardown after captured graphs that hold collectives can block for minutes
    (ncclCommDestroy waits on the graphs' communicator references), so the ranks synchronise, flush and exit."""
    if world > 1:
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


class Harness:
    """Rank / device setup, barriers, L2 flush and the timed loop shared by the train and eval legs."""

    def __init__(self, args):
        import torch.distributed as dist
        self.dist = dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference)")
        torch.cuda.set_device(self.local)
        self.dev = "cuda:%d" % self.local
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device(self.dev))
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)
        self.drain = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev).zero_()
        self.sync_token = torch.zeros(1, device=self.dev)

    def barrier(self):
        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def timed(self, step_fn, steps):
        """K steps, each bracketed by CUDA events on the launching stream, L2 flushed in between."""
        evs = []
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            # L2 flush outside the event pair: write 256 MiB, then read another 256 MiB so that the flush's own dirty lines
            # are written back before the step starts (cold and clean L2)
            self.flush.zero_()
            self.drain.view(torch.int32).sum()
            # ~100 us of spinning on the device in front of the step: the host enqueues the token all-reduce, the start event
            # and the step's graph launch meanwhile, so the GPU never idles between the event and the step's first kernel for
            # as long as the HOST needs to launch the step (20-45 us seen in the N = 2 timeline, profiles/r2_timeline_*),
            # which a training loop that enqueues ahead of the device never pays.  Outside the event pair, at every N alike.
            torch.cuda._sleep(200000)
            if self.world > 1:
                # the flush and the delay are rank-local work outside the timed region (and last a different time on GPUs at
                # different clocks): re-align the ranks on the device right before the start event, or that jitter shows up
                # inside the step as time spent waiting in the gradient sum
                self.dist.all_reduce(self.sync_token)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); step_fn(); b.record()
            evs.append((a, b))
        self.barrier()
        wall = time.perf_counter() - t0
        ms = sum(a.elapsed_time(b) for a, b in evs)
        return ms, wall

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = torch.tensor([x], device=self.dev, dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t)

    def free_flush(self):
        del self.flush, self.drain
        torch.cuda.empty_cache()


def write_timeline(prefix, hs, step_fn):
    """One step under the CUPTI activity tracer (torch.profiler, kernels inside a CUDA-graph replay included): every rank
    writes its kernels as `stream start_us dur_us name`, times relative to the first kernel of the step.  The ranks are
    re-aligned by a token all-reduce right before the step, as in the timed region."""
    from torch.profiler import profile, ProfilerActivity
    for _ in range(2):
        step_fn()
    hs.barrier()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        hs.flush.zero_()
        hs.drain.view(torch.int32).sum()
        if hs.world > 1:
            hs.dist.all_reduce(hs.sync_token)
        marker = torch.cuda.Event(enable_timing=True)
        step_fn()
        torch.cuda.synchronize()
    rows = []
    for e in prof.profiler.kineto_results.events():
        if str(e.device_type()).endswith("CUDA") and e.duration_ns() > 0:
            rows.append((e.start_ns(), e.duration_ns(), e.device_resource_id(), e.name()))
    rows.sort()
    # the step = everything after the flush (a fill, then torch's sum over the drain buffer) and the token all-reduce
    cut = 0
    for i, r in enumerate(rows):
        if "at::native::reduce_kernel" in r[3]:
            cut = i + 1
    if hs.world > 1 and cut < len(rows) and "nccl" in rows[cut][3].lower():
        cut += 1
    rows = rows[cut:]
    t0 = rows[0][0] if rows else 0
    with open("%s.rank%d.txt" % (prefix, hs.rank), "w") as f:
        f.write("# stream start_us dur_us end_us kernel   (rank %d of %d; t = 0 at the step's first kernel)\n" % (hs.rank, hs.world))
        for st, du, sid, name in rows:
            f.write("%3d %9.1f %8.1f %9.1f %s\n" % (sid, (st - t0) / 1e3, du / 1e3, (st - t0 + du) / 1e3, name[:90]))
    hs.barrier()


def mlp_roofline(pk, kernel_ms, flop_per_call_pair, traffic, what, timed_note, ms_total):
    calls = kernel_ms.get("nerf_fwd", (0, 0.0))[0]
    mlp_ms = kernel_ms.get("nerf_fwd", (0, 0.0))[1] + kernel_ms.get("nerf_bwd", (0, 0.0))[1]
    per_call = mlp_ms / max(calls, 1)
    achieved = flop_per_call_pair / (per_call * 1e-3) / 1e12 if per_call > 0 else 0.0
    return dict(bound="tensor", kernel=what, achieved=achieved, peak=pk["bf16_tflops"], unit="TFLOP/s",
                frac=achieved / pk["bf16_tflops"], traffic=traffic,
                peak_source=pk["source"] + " (burst dense bf16: the timed region is milliseconds at full clocks)",
                frac_of_sustained_peak=achieved / pk["bf16_tflops_sustained"], peak_sustained=pk["bf16_tflops_sustained"],
                traffic_note="DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the MLP kernels per launch, from "
                             "the committed ncu --set full capture of this round (profiles/r2_mlp_traffic.json); null when "
                             "no capture of this configuration is committed",
                hbm_view=(dict(gbs=traffic / (per_call * 1e-3) / 1e9, frac=traffic / (per_call * 1e-3) / 1e9 / pk["hbm_gbs"],
                               peak=pk["hbm_gbs"]) if traffic and per_call > 0 else None),
                flop_per_launch=flop_per_call_pair, mlp_calls_per_step=None, mlp_ms_per_step=None,
                mlp_ms_per_launch=per_call, mlp_share_of_step=mlp_ms / ms_total if ms_total > 0 else None, timed=timed_note)


# ------------------------------------------------------------------------------------------------
# ours: train step (c2 / c5)
# ------------------------------------------------------------------------------------------------

def run_train(args):
    from neural_invertible_warp_b200 import _lib, config as cfgmod, engine, synthetic as syn
    from neural_invertible_warp_b200 import functional as F

    hs = Harness(args)
    rank, world, dev, dist = hs.rank, hs.world, hs.dev, hs.dist
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
    engine.use_flat_gradients(graph)            # the kernels accumulate straight into the flat gradient bucket
    # N > 1: the gradient sum runs through our own peer-memory kernels (csrc/p2p.cu) when the ranks share a node
    # (NIW_P2P_ALLREDUCE=0: NCCL all-reduce)
    p2p_on = adam.enable_p2p() if world > 1 else False
    it = 5000
    P_global = rays_global // IMAGES
    P_local = (P_global + world - 1) // world
    rays_local = P_local * IMAGES

    ray_draws = engine.device_ray_draws(dev, seed=20)      # same seed on every rank: one global draw, sliced per rank

    def step_device():
        v = cfgmod.AttrDict(var_dev)
        with ray_draws:
            loss = engine.train_step(opt, graph, v, it, bucket=adam, rank=rank, world=world, optimizer=adam)
        return loss

    # ---- e2e leg: the step's host-side inputs (camera batch, pixel indices, stratified uniforms) are DRAWN ON THE HOST
    # INSIDE the timed region by a loader thread (as a data loader would), written to pinned memory, copied to one of TWO
    # static device buffer sets on a copy stream (overlapping the previous step's kernels); the CUDA graph captured over
    # that set is replayed and the loss copied to pinned memory; the host reads each step's loss one step late ----
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
            with engine.feed_draws(ray_idx=static["ray_idx"], u=static["u"]), \
                    (engine.data_parallel() if world > 1 else contextlib.nullcontext()):
                v = graph.forward(opt, v, mode="train", iter=it)
                loss = engine.summarize_loss(opt, graph.compute_loss(opt, v, mode="train"))
            with engine.backward_schedule(graph):          # dW on a side stream under the pose / warp backward (as train_step)
                if world > 1:
                    with engine.overlap_allreduce(graph, adam):
                        (loss.all * (1.0 / world)).backward()
                    adam.allreduce()
                else:
                    loss.all.backward()
                adam.step(groups=range(1, len(adam.groups)))    # pose / warp groups: final before the side stream is joined
            adam.step(groups=[0])
            return loss.all.detach()
        return step_static

    e2e_body = [make_step_static(statics[0]), make_step_static(statics[1])]

    CHUNK = 2          # steps per staging set

    class HostLoader:
        """The host side of the e2e leg: loader threads draw the pixel indices (k distinct pixels of the frame, what
        ``torch.randperm(H*W)[:k]`` yields, in O(k)) and stratified uniforms of the coming steps into pinned staging sets
        (CHUNK steps per set) while the GPU works on earlier steps -- like a prefetching data loader.  The pinned sets and the
        threads are built once; ``start()`` discards whatever was drawn earlier and lets the threads draw again, ``stop()``
        parks them."""

        def __init__(self, n_sets=4, n_threads=3):
            import queue
            import numpy as np
            self.free, self.ready = queue.Queue(), queue.Queue()
            self.epoch = 0
            for _ in range(n_sets):
                self.free.put(dict(ridx=torch.empty(CHUNK, P_local, dtype=torch.int64).pin_memory(),
                                   u=torch.empty(CHUNK, IMAGES, P_local, N_SAMPLES, 1).pin_memory(), copied=None, epoch=-1))
            self.threads = [threading.Thread(target=self._work, args=(np.random.Generator(np.random.PCG64(1234 + 97 * rank + t)), np),
                                             daemon=True) for t in range(n_threads)]
            for t in self.threads:
                t.start()

        def _work(self, rng, np):
            while True:
                s = self.free.get()
                if s is None:
                    return
                s["epoch"] = self.epoch
                if s["copied"] is not None:
                    s["copied"].synchronize()              # the H2D copies out of this staging set have completed
                    s["copied"] = None
                ridx = s["ridx"].numpy()
                for k in range(CHUNK):
                    ridx[k] = rng.choice(H * W, P_local, replace=False)
                rng.random(out=s["u"].numpy(), dtype=np.float32)          # (numpy releases the GIL while it fills the buffer)
                self.ready.put(s)

        def start(self):
            self.epoch += 1            # sets drawn before this moment are handed back and drawn again (next())

        def next(self):
            while True:
                s = self.ready.get()
                if s["epoch"] == self.epoch:
                    return s
                self.free.put(s)

        def close(self):
            for _ in self.threads:
                self.free.put(None)

    loader = HostLoader()

    def run_e2e(steps):
        """`steps` end-to-end steps; returns (the losses the host read -- all of them, each one step late --, wall seconds
        from the first host->device copy to the last loss read)."""
        main = torch.cuda.current_stream()
        torch.cuda.synchronize()
        loader.start()
        losses = []
        batch = None
        t0 = None
        for i in range(steps):
            b, k = i & 1, i % CHUNK
            if k == 0:
                batch = loader.next()                            # host-side batches of the next CHUNK steps (pinned memory)
            if t0 is None:
                t0 = time.perf_counter()                         # the first set is there: the clock runs from its first copy
            with torch.cuda.stream(copy_stream):
                if i >= 2:
                    copy_stream.wait_event(consumed[b])          # step i-2 is done with this buffer set
                st = statics[b]
                st["ray_idx"].copy_(batch["ridx"][k], non_blocking=True)
                st["u"].copy_(batch["u"][k], non_blocking=True)
                for key in ("idx", "intr", "pose"):
                    st[key].copy_(pin[key], non_blocking=True)
                ready[b].record(copy_stream)
                if k == CHUNK - 1 or i == steps - 1:
                    batch["copied"] = torch.cuda.Event()
                    batch["copied"].record(copy_stream)
                    loader.free.put(batch)
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
        return losses, time.perf_counter() - (t0 or time.perf_counter())

    # ---- warm-up ----
    for _ in range(max(args.warmup, 3)):
        step_device()
    run_e2e(2)
    hs.barrier()
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
    hs.barrier()

    # ---- value: device-resident ----
    clocks = ClockSampler(hs.local)
    if rank == 0:
        clocks.start()
    ms_total, _ = hs.timed(step_value, args.steps)
    ms_total = hs.max_over_ranks(ms_total)
    # per-kernel CUDA-event timing (roofline leg) and the launch count: the same K steps launched eagerly
    # (event records cannot live inside a captured graph)
    n0 = _lib.launch_count()
    with F.KernelTimer() as kt:
        ms_eager, _ = hs.timed(step_device, args.steps)
        kernel_ms = kt.totals()
    launches = _lib.launch_count() - n0
    # ---- e2e: host buffers in, loss out (wall clock around synchronised steps) ----
    hs.barrier()
    e2e_losses, e2e_wall = run_e2e(args.steps)
    e2e_s = hs.max_over_ranks(e2e_wall)
    hs.barrier()
    loader.close()
    assert len(e2e_losses) == args.steps and all(l == l for l in e2e_losses), "e2e leg: every step's loss must be read"
    clk = clocks.stop() if rank == 0 else None

    ms_per_step = ms_total / args.steps
    rays_per_step = rays_local * world
    value = rays_per_step / (ms_per_step * 1e-3)
    e2e_value = rays_per_step * args.steps / e2e_s
    if args.timeline:
        write_timeline(args.timeline, hs, step_value)

    # ---- N > 1: the sharded step's global loss against the same batch rendered by rank 0 alone ----
    loss_check = None
    if world > 1 and not args.no_loss_check:
        g = torch.Generator().manual_seed(4321)
        ridx = torch.randperm(H * W, generator=g)[:P_global].to(dev)
        u = torch.rand(IMAGES, P_global, N_SAMPLES, 1, generator=g).to(dev)
        per = (P_global + world - 1) // world
        with torch.no_grad():
            v = cfgmod.AttrDict(var_dev)
            with engine.feed_draws(ray_idx=ridx, u=u[:, rank * per:(rank + 1) * per].contiguous()), \
                    engine._ShardedRandperm(rank, world, P_global):
                v = graph.forward(opt, v, mode="train", iter=it)
            local = engine.summarize_loss(opt, graph.compute_loss(opt, v, mode="train")).all * (len(v.ray_idx) / float(P_global))
            total = local.detach().clone().double()
            dist.all_reduce(total)
            if rank == 0:
                # the whole global batch on one rank, in slices of the local size (forward only: the loss needs no records)
                acc = torch.zeros((), device=dev, dtype=torch.float64)
                for r in range(world):
                    v1 = cfgmod.AttrDict(var_dev)
                    with engine.feed_draws(ray_idx=ridx, u=u[:, r * per:(r + 1) * per].contiguous()), \
                            engine._ShardedRandperm(r, world, P_global):
                        v1 = graph.forward(opt, v1, mode="train", iter=it)
                    l1 = engine.summarize_loss(opt, graph.compute_loss(opt, v1, mode="train")).all
                    acc += l1.double() * (len(v1.ray_idx) / float(P_global))
                diff = abs(float(total) - float(acc))
                loss_check = dict(global_loss_sharded=float(total), global_loss_one_rank=float(acc), abs_diff=diff,
                                  what="loss of one fed global batch: all-reduced sum of the ranks' scaled shard losses vs "
                                       "the same shards rendered by rank 0 alone")
                assert diff <= 1e-6 * max(1.0, abs(float(acc))), "sharded global loss differs from the one-rank loss: %r" % loss_check

    # ---- roofline of the dominant kernels (the MLP) ----
    pk = peaks()
    flop_per_step = MLP_FLOP_TRAIN * rays_local * N_SAMPLES
    traffic = committed_traffic(args.config, precision) if args.rays == DEFAULT_RAYS.get(args.config) else None
    roof = mlp_roofline(pk, kernel_ms, flop_per_step, traffic,
                        "niw_nerf_fwd + niw_nerf_bwd (fused PE + 8x256 MLP, fwd + dX + dW)",
                        "CUDA events around niw_nerf_fwd / niw_nerf_bwd on the launching stream, eager pass of the same %d "
                        "steps (%.3f ms/step eager); the BF16 weight packing (2 kernels, ~11 us) runs earlier on a side "
                        "stream (niw_nerf_pack) and is outside this bracket" % (args.steps, ms_eager / args.steps), ms_eager)
    roof["mlp_calls_per_step"] = 1
    roof["mlp_ms_per_step"] = roof["mlp_ms_per_launch"]
    comp_ms = kernel_ms.get("composite_fwd", (0, 0.0))[1] + kernel_ms.get("composite_bwd", (0, 0.0))[1]
    roof["composite_ms_per_step"] = comp_ms / max(kernel_ms.get("nerf_fwd", (1, 0.0))[0], 1)

    if rank != 0:
        _leave(world)
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        torch.manual_seed(0)
        make, what = cpu_sample(args)
        n_cpu = 3 if args.rays <= 2048 else 1
        rate, dt = time_cpu(make, n_cpu, 1)
        cpu = dict(value=rate, unit="rays/s", cores=torch.get_num_threads(), kind="port",
                   sample="%d steps (after 1 warm-up), each %s; oracle/reference_port.py on torch CPU fp32, %.2f s/step"
                          % (n_cpu, what, dt))

    hbm = None
    if world == 1 and not args.no_micro:
        # sampler / compositor / raygen alone at 262 144 rays (inputs larger than L2, L2 flushed between launches):
        # algorithmic bytes / CUDA-event time against the measured HBM copy bandwidth
        hs.free_flush()
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import micro_hbm
        m = micro_hbm.measure(verbose=False, dev=dev)
        hbm = dict(rays=m["rays"], peak_gbs=m["peak_gbs"], peak_source=pk["source"],
                   kernels={k: dict(gbs=round(v["gbs"], 1), frac=round(v["frac"], 3), ms=round(v["ms"], 4), bytes=v["bytes"])
                            for k, v in m["kernels"].items()})

    line = dict(metric=METRIC_TRAIN, value=value, unit="rays/s", n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype=("fp32" if precision == "fp32" else "bf16"), data="synthetic",
                config=workload_config(args), mlp_precision=precision,
                launch="one CUDA-graph replay per step (value and e2e)" if graphed else "eager launches",
                mlp_evals_per_s=value * N_SAMPLES,
                e2e=dict(value=e2e_value, unit="rays/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=4,
                         ms_per_step=e2e_s / args.steps * 1e3,
                         how="public Graph API; loader threads draw the pixel indices and stratified uniforms of the coming steps "
                             "into pinned memory (2 steps per staging set; all but the first set are drawn inside the timed "
                             "region, concurrently with the GPU) -> H2D on a copy stream into one of two static buffer "
                             "sets -> CUDA-graph replay -> loss to pinned memory; the host reads every step's loss, one step "
                             "late; wall clock from the first H2D copy to the last loss read"),
                gpu_launches=launches, roofline=roof, hbm_kernels=hbm, cpu_baseline=cpu, clocks=clk, loss_check=loss_check)
    if world > 1:
        line["collective"] = ("gradient sum by this repo's peer-memory kernels over NVLink (csrc/p2p.cu: publish + rank-order "
                              "reduce, one channel per segment); torch.distributed/NCCL for rendezvous and the scalar checks only"
                              if p2p_on else "NCCL all-reduce of the flat gradient bucket (two segments)")
        if p2p_on:
            line["collective_errors"] = [ch.error() for ch in adam._p2p]
    print(json.dumps(line))
    _leave(world)


# ------------------------------------------------------------------------------------------------
# ours: full-frame eval render, rows sharded (c4)
# ------------------------------------------------------------------------------------------------

def run_eval(args):
    from neural_invertible_warp_b200 import _lib, config as cfgmod, engine
    from neural_invertible_warp_b200 import functional as F

    hs = Harness(args)
    rank, world, dev = hs.rank, hs.world, hs.dev
    _lib.load()
    precision = args.precision or "bf16x3"
    if _lib.load().niw_nerf_workspace_bytes(1, 1, F.precision_code(precision), 0) == 0:
        raise SystemExit("precision %s unavailable" % precision)
    if H % world:
        raise SystemExit("c4: %d rows do not split over %d ranks" % (H, world))
    rows = H // world
    row0 = rank * rows
    opt = cfgmod.builtin_options("barf_inn_dtu", barf_c2f=[0.1, 0.5], device=dev, data=dict(image_size=[H, W]),
                                 nerf=dict(rand_rays=rows * W, sample_intvs=C4_N, fine_sampling=True, sample_intvs_fine=C4_NF,
                                           depth=dict(range=[1.2, 5.2])),
                                 loss_weight=dict(render_fine=0), arch=dict(mlp_precision=precision))
    torch.manual_seed(0)
    var0 = engine.synthetic_var(opt, 1, 3, dtu=True)
    graph = engine.build_graph(opt, 1, initial_poses_w2c=var0.pose.clone())
    graph.nerf.progress.data.fill_(0.3); graph.nerf_fine.progress.data.fill_(0.3)
    pose_dev, intr_dev = var0.pose.clone(), var0.intr.clone()

    def step_device():
        return engine.render_rows(opt, graph, pose_dev, intr_dev, row0, row0 + rows, depth_range=[1.2, 5.2])

    # e2e: the view's camera (pose, intrinsics) comes from pinned host memory, the rendered rows (rgb, depth, opacity of
    # the fine pass) go back to pinned host memory, every step
    pin_pose, pin_intr = var0.pose.cpu().pin_memory(), var0.intr.cpu().pin_memory()
    out_host = torch.empty(rows * W, 5).pin_memory()
    h2d = pin_pose.numel() * 4 + pin_intr.numel() * 4
    d2h = out_host.numel() * 4

    def step_e2e():
        pose_dev.copy_(pin_pose, non_blocking=True)
        intr_dev.copy_(pin_intr, non_blocking=True)
        ret = step_device()
        out_host[:, 0:3].copy_(ret.rgb_fine.view(-1, 3), non_blocking=True)
        out_host[:, 3:4].copy_(ret.depth_fine.view(-1, 1), non_blocking=True)
        out_host[:, 4:5].copy_(ret.opacity_fine.view(-1, 1), non_blocking=True)
        torch.cuda.current_stream().synchronize()                 # the caller holds the rows
        return float(out_host[0, 0])

    for _ in range(max(args.warmup, 3)):
        step_device()
    step_e2e()
    hs.barrier()
    clocks = ClockSampler(hs.local)
    if rank == 0:
        clocks.start()
    n0 = _lib.launch_count()
    with F.KernelTimer() as kt:
        ms_total, _ = hs.timed(step_device, args.steps)
        kernel_ms = kt.totals()
    launches = _lib.launch_count() - n0
    ms_total = hs.max_over_ranks(ms_total)
    hs.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    hs.barrier()
    e2e_s = hs.max_over_ranks(time.perf_counter() - t0)
    clk = clocks.stop() if rank == 0 else None

    ms_per_step = ms_total / args.steps
    value = H * W / (ms_per_step * 1e-3)
    e2e_value = H * W * args.steps / e2e_s
    pk = peaks()
    evals = rows * W * (C4_N + C4_N + C4_NF)
    calls = kernel_ms.get("nerf_fwd", (0, 0.0))[0]
    per_step_ms = kernel_ms.get("nerf_fwd", (0, 0.0))[1] / args.steps
    achieved = MLP_FLOP_FWD * evals / (per_step_ms * 1e-3) / 1e12 if per_step_ms > 0 else 0.0
    roof = dict(bound="tensor", kernel="niw_nerf_fwd (fused PE + 8x256 MLP forward; coarse + fine network)", achieved=achieved,
                peak=pk["bf16_tflops_sustained"], unit="TFLOP/s", frac=achieved / pk["bf16_tflops_sustained"],
                frac_of_burst_peak=achieved / pk["bf16_tflops"], peak_burst=pk["bf16_tflops"],
                peak_source=pk["source"] + " (sustained dense bf16: the kernels run back to back for tens of ms per step)",
                traffic=committed_traffic("c4", precision) if world == 1 else None,
                flop_per_step=MLP_FLOP_FWD * evals, mlp_calls_per_step=calls / args.steps, mlp_ms_per_step=per_step_ms,
                mlp_share_of_step=per_step_ms / ms_per_step if ms_per_step > 0 else None,
                note="algorithmic FLOP (2 x 527 872 per MLP evaluation) of this rank's rows / CUDA-event time of its "
                     "niw_nerf_fwd calls; with bf16x3 the tensor cores execute 3 BF16 products per algorithmic one "
                     "(hi*hi + lo*hi + hi*lo), so the executed rate is 3x the figure quoted",
                timed="CUDA events around every niw_nerf_fwd on the launching stream inside the timed steps")
    if rank != 0:
        _leave(world)
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        torch.manual_seed(0)
        make, what = cpu_sample(args)
        rate, dt = time_cpu(make, 3, 1)
        cpu = dict(value=rate, unit="rays/s", cores=torch.get_num_threads(), kind="port",
                   sample="3 steps (after 1 warm-up), each %s; oracle/reference_port.py on torch CPU fp32, %.2f s/step" % (what, dt))
    line = dict(metric=METRIC_EVAL, value=value, unit="rays/s", n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                ms_per_step=ms_per_step, frame_ms=ms_per_step, higher_is_better=True, scaling="strong", vs_baseline=None,
                dtype="bf16" if precision != "fp32" else "fp32", data="synthetic", config=workload_config(args),
                mlp_precision=precision, launch="eager launches (11 kernels per step)",
                mlp_evals_per_s=value * (C4_N + C4_N + C4_NF),
                e2e=dict(value=e2e_value, unit="rays/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                         ms_per_step=e2e_s / args.steps * 1e3,
                         how="engine.render_rows through the public Graph; camera from pinned memory every step, this rank's "
                             "rendered rows (rgb, depth, opacity) copied to pinned memory and awaited every step; wall clock"),
                gpu_launches=launches, roofline=roof, cpu_baseline=cpu, clocks=clk)
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
    elif a.config == "c4":
        run_eval(a)
    else:
        run_train(a)
