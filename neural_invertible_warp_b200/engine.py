"""Host orchestration around the drop-in Graphs: what the reference's ``Model`` classes do to a
graph between construction and ``loss.backward()``, plus the data-parallel hooks the reference
lacks (it asserts a single GPU, options.py:103).

* ``build_graph``      -- ``Model.build_networks`` for the target models (reference model/base.py:34-37,
                          barf.py:38-44, barf_inn_llff.py:25-82, barf_inn_dtu.py:323-336): attaches
                          ``se3_refine`` / ``warp_latent`` + ``warp_mlp`` + ``global_rigid`` / ``pose_net``.
* ``summarize_loss``   -- model/base.py:130-142 without the NaN/Inf host syncs.
* ``train_step``       -- ``graph.forward`` -> ``compute_loss`` -> ``backward`` (model/nerf.py:77-101 minus
                          the optimiser), optionally over a 1/k ray shard with one gradient all-reduce.
* ``test_time_photometric_optim`` -- model/barf.py:153-169.
* ``evaluate_view``    -- model/nerf.py:162-183 per test view: eval render + PSNR / SSIM on the device (8 f3).
* ``GradBucket``       -- flat fp32 bucket of every trainable gradient: ONE NCCL all-reduce per step
                          (SURVEY.md section 8e); rays shard, parameters replicate.
"""
import contextlib
import importlib

import torch
import torch.distributed as dist
from torch import nn

from . import camera
from .config import AttrDict
from .nvp import DeformNetwork


def build_graph(opt, n_images, initial_poses_w2c=None):
    """Instantiate ``neural_invertible_warp_b200.model.<opt.model>.Graph`` on ``opt.device`` with
    the sub-modules the reference's engine attaches."""
    name = opt.model
    mod = importlib.import_module("neural_invertible_warp_b200.model." + name)
    dev = opt.device
    if name == "barf_inn_dtu":
        from .model.pose_models.inn import INNPoseParams
        if initial_poses_w2c is None:
            raise RuntimeError("barf_inn_dtu needs the initial world->camera poses")
        pose_net = INNPoseParams(opt, num_poses=n_images, initial_poses_w2c=initial_poses_w2c.to(dev), device=dev)
        return mod.Graph(opt, pose_net).to(dev)
    graph = mod.Graph(opt).to(dev)
    if name == "barf":
        graph.se3_refine = nn.Embedding(n_images, 6).to(dev)
        nn.init.zeros_(graph.se3_refine.weight)
    elif name == "barf_inn_llff":
        if opt.warp_latent.enc_type != "l2fbarf":
            raise NotImplementedError("warp_latent.enc_type=%r" % (opt.warp_latent.enc_type,))
        graph.warp_latent = nn.Embedding(n_images, opt.warp_latent.embed_dim).to(dev)
        graph.warp_mlp = DeformNetwork(d_feature=opt.warp_latent.embed_dim, d_in=3, d_out_1=1, d_out_2=3, n_blocks=3,
                                       d_hidden=opt.inn.real_nvp.d_hidden, n_layers=1, skip_in=[],
                                       multires=opt.inn.real_nvp.multires, weight_norm=True, actfn=opt.inn.actfn).to(dev)
        if opt.warp_latent.normalize:
            graph.frame_id = (torch.linspace(1, n_images, n_images)[:, None] / n_images).to(dev)
        pose = graph.pose_eye[None].repeat(n_images, 1, 1)
        graph.global_rigid = nn.Embedding(n_images, 12, _weight=pose.reshape(-1, 12).clone()).to(dev)
    return graph


def trainable_parameters(graph):
    """Every parameter an optimiser of the reference touches, in a fixed order (identical on all ranks)."""
    out = []
    for n, p in graph.named_parameters():
        if n.endswith("progress") or n.startswith("global_rigid") or "pose_global" in n:
            continue   # schedules / Kabsch outputs: written through .data, never optimised
        out.append((n, p))
    return out


def summarize_loss(opt, loss):
    """model/base.py:130-142: loss.all = sum_k 10**w_k loss_k."""
    total = None
    for key in list(loss.keys()):
        if key == "all":
            continue
        if opt.loss_weight[key] is not None:
            w = 10 ** float(opt.loss_weight[key])
            term = loss[key] if w == 1.0 else w * loss[key]     # no launch for the common weight 10**0
            total = term if total is None else total + term
    loss.update(all=0. if total is None else total)
    return loss


def use_flat_gradients(graph, enabled=True):
    """Opt the graph's NeRF modules (and warp networks) in to in-kernel gradient accumulation: their backward kernels
    add into the ``.grad`` buffers directly and return None to autograd.  Meant for training loops that own a flat
    gradient bucket (``GradBucket`` / ``FlatAdam``) and call plain ``loss.backward()``; leave it off when gradients are
    consumed through ``torch.autograd.grad`` / ``backward(inputs=...)`` / hooks."""
    from .model._core import NeRFCore
    for m in graph.modules():
        if isinstance(m, (NeRFCore, DeformNetwork)):
            m.accumulate_grads_in_place = bool(enabled)


class GradBucket:
    """One flat fp32 buffer holding every trainable gradient.  ``attach`` points each ``p.grad`` at
    its slice, so backward accumulates straight into the bucket and ``allreduce`` is a single
    collective over ~0.7 M floats (latency-bound on NVLink; nothing to overlap at this size)."""

    def __init__(self, graph):
        self.params = [p for _, p in trainable_parameters(graph)]
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, device=self.params[0].device, dtype=torch.float32)
        self.attach()

    def attach(self):
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def allreduce(self, group=None):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)


class SegmentedAllreduce:
    """All-reduce of a flat gradient buffer made of per-group segments (``self.flat``, ``self.groups`` with
    ``offset`` / ``n``).  A segment whose gradients are final before the backward pass ends can be reduced early
    with ``allreduce_group_async`` (NCCL runs it on its own stream, concurrently with the rest of the backward);
    ``allreduce`` then reduces whatever is left and joins the early ones.  Without early segments it is ONE
    collective over the whole buffer."""

    _pending = None
    _p2p = None

    def enable_p2p(self, group=None):
        """Route the sums through the peer-memory kernels (p2p.P2PChannel, csrc/p2p.cu; one channel per segment, so an
        early segment on the side stream and the rest on the main stream never share flags) when all ranks are GPUs of
        one node; otherwise keep the NCCL collectives.  Collective call: every rank of ``group`` makes it.  Returns
        whether the peer path is on."""
        from . import p2p
        if self._p2p is None and p2p.available(group):
            chans = []
            try:
                for g in self.groups:
                    chans.append(p2p.P2PChannel(g["n"], group))      # raises on every rank alike when any rank failed
                self._p2p = chans
            except RuntimeError as e:
                import warnings
                warnings.warn("niw_b200: peer-memory gradient sum unavailable (%s); using the NCCL all-reduce" % e)
                for c in chans:
                    c.close()
        return self._p2p is not None

    @staticmethod
    def _active(group):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1

    def _segment(self, gi):
        g = self.groups[gi]
        return self.flat[g["offset"]:g["offset"] + g["n"]]

    def allreduce_group_async(self, gi, group=None):
        if not self._active(group):
            return
        if self._pending is None:
            self._pending = {}
        if gi not in self._pending:
            if self._p2p is not None:
                self._p2p[gi].allreduce_(self._segment(gi))       # two launches on the current stream; nothing to wait for
                self._pending[gi] = None
            else:
                self._pending[gi] = dist.all_reduce(self._segment(gi), op=dist.ReduceOp.SUM, group=group, async_op=True)

    def allreduce(self, group=None):
        if not self._active(group):
            return
        pending = self._pending or {}
        if self._p2p is not None:
            for gi in range(len(self.groups)):
                if gi not in pending:
                    self._p2p[gi].allreduce_(self._segment(gi))
            self._pending = {}
            return
        if not pending:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            return
        # what is left goes out as ONE collective per run of adjacent segments (normally a single one: everything
        # behind the early NeRF segment)
        gi, n = 0, len(self.groups)
        while gi < n:
            if gi in pending:
                gi += 1
                continue
            gj = gi
            while gj + 1 < n and gj + 1 not in pending and \
                    self.groups[gj + 1]["offset"] == self.groups[gj]["offset"] + self.groups[gj]["n"]:
                gj += 1
            lo, hi = self.groups[gi]["offset"], self.groups[gj]["offset"] + self.groups[gj]["n"]
            dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=group)
            gi = gj + 1
        for work in pending.values():
            if work is not None:
                work.wait()                   # the current stream waits for the collective; no host block
        self._pending = {}


class overlap_allreduce:
    """``with overlap_allreduce(graph, bucket): loss.backward()`` -- the NeRF MLP gradients are final as soon as the
    fused MLP backward has run, while the pose / warp part of the backward (NVP coupling layers, ~110 us at C2) is
    still to come: their segment of the bucket is all-reduced concurrently with it.  Only when segment 0 of the
    bucket holds exactly the parameters of ``graph.nerf`` (+ ``nerf_fine``), as ``reference_optimizer_groups`` builds it."""

    def __init__(self, graph, bucket, group=None):
        self.mods = [graph.nerf] + ([graph.nerf_fine] if hasattr(graph, "nerf_fine") else [])
        self.bucket, self.group = bucket, group
        self.ok = False
        if isinstance(bucket, SegmentedAllreduce) and SegmentedAllreduce._active(group) and len(bucket.groups) > 1:
            mine = {id(p) for m in self.mods for p in m.mlp_parameters()}
            self.ok = {id(p) for p in bucket.groups[0]["params"]} == mine

    def __enter__(self):
        if self.ok:
            # a module may have been evaluated several times in this step with gradients enabled (slices, several views
            # in one loss): functional._NerfSamples counts those forward calls in ``_pending_backward`` and reports a
            # module only when its LAST backward node has run -- an earlier reduce would race with the later nodes'
            # accumulation into the same segment
            waiting = {id(m) for m in self.mods if getattr(m, "_pending_backward", 0) > 0}
            fire = bool(waiting)

            def ready(m):
                waiting.discard(id(m))
                if fire and not waiting:
                    self.bucket.allreduce_group_async(0, self.group)
            for m in self.mods:
                m._grads_ready = ready
        return self

    def __exit__(self, *exc):
        for m in self.mods:
            m._grads_ready = None
            m._pending_backward = 0      # forward calls whose backward never ran (unused outputs) do not leak into the next step


class FlatAdam(SegmentedAllreduce):
    """The reference's optimisers (``optim`` on ``graph.nerf``: Adam + ExponentialLR, model/nerf.py:33-46;
    ``optim_pose`` on ``warp_mlp`` + ``warp_latent`` / ``se3_refine``: model/barf_inn_llff.py:84-104,
    model/barf.py:46-60) as ONE kernel launch per parameter group (csrc/adam.cu, ``niw_adam_step``).

    ``groups`` = [dict(params=[...], lr=, gamma=1.0, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, warmup=0,
    warmup_params=None, torch_groups=None), ...].  ``warmup_params``: how many leading tensors of ``params`` the linear
    LR warm-up applies to (default: all); ``torch_groups``: tensor counts of the ``param_groups`` the reference's
    torch optimiser has for this group (state_dict interchange, default one group).
    Every group becomes a contiguous fp32 segment of one flat parameter buffer and of one flat gradient buffer
    (the parameters / their ``.grad`` are re-pointed at views, names and shapes untouched), so this object is
    also the data-parallel gradient bucket: ``zero`` / ``allreduce`` have GradBucket's meaning.  The step
    count lives on the device, so ``step`` can be captured in a CUDA graph; ``gamma`` is the per-step
    ExponentialLR factor (lr_t = lr * gamma**(t-1))."""

    def __init__(self, groups, progress=None, max_iter=None):
        """``progress``: the BARF schedule scalars (``nerf.progress`` [, ``nerf_fine.progress``]) that the LAST group's
        kernel sets to step / ``max_iter`` after its update, as the reference's ``train_iteration`` does after the pose
        step (model/barf.py:57-59).  A group dict may carry ``warmup`` (iterations of linear LR warm-up, model/barf.py:48-51)."""
        from . import _lib
        self._lib = _lib.load()
        self.groups = []
        self.progress = list(progress or [])
        self.max_iter = float(max_iter) if max_iter else 0.0
        if len(self.progress) > 2 or (self.progress and not self.max_iter):
            raise ValueError("FlatAdam: at most two progress scalars, and max_iter with them")
        dev = groups[0]["params"][0].device
        sizes = []
        for g in groups:
            n = sum(p.numel() for p in g["params"])
            sizes.append((n + 3) // 4 * 4)                      # 16-byte aligned segments
        total = sum(sizes)
        self.flat_params = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)    # gradients (GradBucket interface)
        self.exp_avg = torch.zeros(total, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(total, device=dev, dtype=torch.float32)
        self.state = torch.zeros(len(groups), 2, device=dev, dtype=torch.float32)
        base = 0
        for gi, (g, size) in enumerate(zip(groups, sizes)):
            off = base
            for p in g["params"]:
                if p.dtype != torch.float32 or not p.is_cuda:
                    raise RuntimeError("niw_b200 FlatAdam: parameters must be CUDA float32 tensors")
                n = p.numel()
                with torch.no_grad():
                    self.flat_params[off:off + n].copy_(p.data.reshape(-1))
                    p.data = self.flat_params[off:off + n].view_as(p)
                p.grad = self.flat[off:off + n].view_as(p)
                off += n
            b1, b2 = g.get("betas", (0.9, 0.999))
            wp = g.get("warmup_params", None)
            warmup_n = size if wp is None else sum(p.numel() for p in g["params"][:wp])
            tg = list(g.get("torch_groups", None) or [len(g["params"])])
            if sum(tg) != len(g["params"]):
                raise ValueError("FlatAdam: torch_groups must partition the group's parameter list")
            self.groups.append(dict(offset=base, n=size, lr=float(g["lr"]), gamma=float(g.get("gamma", 1.0)), b1=float(b1),
                                    b2=float(b2), eps=float(g.get("eps", 1e-8)), wd=float(g.get("weight_decay", 0.0)),
                                    warmup=float(g.get("warmup", 0) or 0), warmup_n=int(warmup_n), torch_groups=tg,
                                    params=list(g["params"])))
            base += size
        for p in self.progress:
            if not p.is_cuda:
                raise RuntimeError("niw_b200 FlatAdam: progress scalars must live on the device")

    def zero(self):
        self.flat.zero_()

    zero_grad = zero

    def step(self, groups=None):
        """One update of every group (``groups``: indices, default all) on the current stream.  The groups are independent,
        so a training step may update the pose / warp groups as soon as their gradients are final and the NeRF group
        after the side-stream weight-gradient pass has been joined (``train_step(optimizer=...)``)."""
        import ctypes
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        for gi, g in enumerate(self.groups):
            if groups is not None and gi not in groups:
                continue
            o = g["offset"] * 4
            ptr = lambda t: ctypes.c_void_p(t.data_ptr() + o)
            prog = self.progress if gi == len(self.groups) - 1 else []
            pp = [ctypes.c_void_p(t.data_ptr()) for t in prog] + [None, None]
            rc = self._lib.niw_adam_step(ptr(self.flat_params), ptr(self.flat), ptr(self.exp_avg), ptr(self.exp_avg_sq),
                                         g["n"], g["lr"], g["gamma"], g["b1"], g["b2"], g["eps"], g["wd"], g["warmup"],
                                         g["warmup_n"], self.max_iter, pp[0], pp[1],
                                         ctypes.c_void_p(self.state.data_ptr() + gi * 8), st)
            if rc:
                raise RuntimeError("niw_adam_step: %s" % self._lib.niw_error_string(rc).decode())

    # -- checkpointing (reference util.py:124-163 persists every optim* / sched* state_dict) ------------------------
    _HYPER = ("offset", "n", "lr", "gamma", "b1", "b2", "eps", "wd", "warmup", "warmup_n")

    def state_dict(self):
        """Both Adam moments, the device step counters (which drive lr * gamma**(t-1), the pose warm-up and the BARF
        ``progress`` scalar) and the group hyper-parameters.  Parameters and ``progress`` belong to the graph's own
        ``state_dict``."""
        return dict(version=1, exp_avg=self.exp_avg.detach().clone(), exp_avg_sq=self.exp_avg_sq.detach().clone(),
                    steps=self.state[:, 0].detach().clone(), max_iter=self.max_iter,
                    groups=[{k: g[k] for k in self._HYPER} for g in self.groups])

    def load_state_dict(self, sd):
        """In place (buffers captured in a CUDA graph stay valid).  The layout must match; learning-rate
        hyper-parameters are taken from the checkpoint, as ``torch.optim.Optimizer.load_state_dict`` does."""
        if len(sd["groups"]) != len(self.groups) or any(a["n"] != b["n"] or a["offset"] != b["offset"]
                                                        for a, b in zip(sd["groups"], self.groups)):
            raise ValueError("FlatAdam.load_state_dict: parameter groups do not match this optimiser")
        with torch.no_grad():
            self.exp_avg.copy_(sd["exp_avg"])
            self.exp_avg_sq.copy_(sd["exp_avg_sq"])
            self.state.zero_()
            self.state[:, 0].copy_(sd["steps"])
        for g, saved in zip(self.groups, sd["groups"]):
            for k in self._HYPER:
                g[k] = saved[k]

    def _group_slices(self, gi):
        g, off, out = self.groups[gi], self.groups[gi]["offset"], []
        for p in g["params"]:
            out.append((p, off, off + p.numel()))
            off += p.numel()
        return out

    def to_torch(self, gi):
        """Group ``gi`` as the reference's own optimiser objects: (torch.optim.Adam, ExponentialLR or None) over the same
        parameter tensors, carrying this optimiser's moments, step count and decayed learning rate -- what
        ``util.save_checkpoint`` (util.py:147-163) stores as ``optim`` / ``sched`` (``optim_pose`` / ``sched_pose``).
        Reads the step counter (one host sync)."""
        g = self.groups[gi]
        t = int(round(float(self.state[gi, 0])))
        ps, k, pgs = g["params"], 0, []
        for cnt in g["torch_groups"]:
            pgs.append(dict(params=ps[k:k + cnt], lr=g["lr"]))
            k += cnt
        optim = torch.optim.Adam(pgs, betas=(g["b1"], g["b2"]), eps=g["eps"], weight_decay=g["wd"])
        if t > 0:
            for p, lo, hi in self._group_slices(gi):
                optim.state[p] = dict(step=torch.tensor(float(t)), exp_avg=self.exp_avg[lo:hi].clone().view_as(p),
                                      exp_avg_sq=self.exp_avg_sq[lo:hi].clone().view_as(p))
        sched = None
        if g["gamma"] != 1.0:
            sched = torch.optim.lr_scheduler.ExponentialLR(optim, gamma=g["gamma"])
            sched.last_epoch = t
            sched._step_count = t + 1
            lr_t = g["lr"] * g["gamma"] ** t
            for pg in optim.param_groups:
                pg["lr"] = lr_t
            sched._last_lr = [lr_t for _ in optim.param_groups]
        return optim, sched

    def load_torch(self, gi, optim_sd, sched_sd=None):
        """Adopt the state of a reference checkpoint's ``optim`` (+ ``sched``) ``state_dict`` for group ``gi``."""
        g = self.groups[gi]
        slices = self._group_slices(gi)
        state = optim_sd["state"]
        t = 0
        with torch.no_grad():
            for i, (p, lo, hi) in enumerate(slices):
                st = state.get(i, None)
                if st is None:
                    self.exp_avg[lo:hi].zero_(); self.exp_avg_sq[lo:hi].zero_()
                    continue
                self.exp_avg[lo:hi].copy_(st["exp_avg"].reshape(-1))
                self.exp_avg_sq[lo:hi].copy_(st["exp_avg_sq"].reshape(-1))
                t = max(t, int(round(float(st["step"]))))
            if sched_sd is not None:
                if int(sched_sd["last_epoch"]) != t and t > 0:
                    raise ValueError("FlatAdam.load_torch: optimiser took %d steps, scheduler %d" % (t, sched_sd["last_epoch"]))
                t = int(sched_sd["last_epoch"])
                g["gamma"] = float(sched_sd["gamma"])
                g["lr"] = float(sched_sd["base_lrs"][0])
            else:
                g["lr"] = float(optim_sd["param_groups"][0].get("initial_lr", optim_sd["param_groups"][0]["lr"]))
            pg0 = optim_sd["param_groups"][0]
            g["b1"], g["b2"] = (float(b) for b in pg0["betas"])
            g["eps"], g["wd"] = float(pg0["eps"]), float(pg0["weight_decay"])
            self.state[gi, 0] = float(t)


def reference_optimizer_groups(opt, graph):
    """The parameter groups of the reference's optimisers for the target models (model/nerf.py:33-46,
    model/barf.py:46-60, model/barf_inn_llff.py:84-104, model/barf_inn_dtu.py:338-351), as FlatAdam group dicts.
    ExponentialLR only when ``optim.sched`` / ``optim.sched_pose`` is set, gamma = (lr_end / lr) ** (1 / max_iter)
    when ``lr_end`` is given (else the YAML's own gamma); the pose warm-up covers the reference's
    ``optim_pose.param_groups[0]`` only (the warp network / ``se3_refine``), not the latent codes added as a second
    param group (model/barf_inn_llff.py:108-111)."""
    def sched(lr, lr_end, sc):
        lr = float(lr)
        if not sc:
            return dict(lr=lr, gamma=1.0)
        if lr_end:
            return dict(lr=lr, gamma=(float(lr_end) / lr) ** (1.0 / opt.max_iter))
        g = sc.get("gamma", None) if hasattr(sc, "get") else None
        return dict(lr=lr, gamma=float(g) if g else 1.0)
    o = opt.optim
    nerf_params = [p for n, p in graph.nerf.named_parameters() if not n.endswith("progress")]
    tg = [len(nerf_params)]
    if opt.nerf.fine_sampling:
        fine = [p for n, p in graph.nerf_fine.named_parameters() if not n.endswith("progress")]
        nerf_params += fine
        tg.append(len(fine))
    groups = [dict(params=nerf_params, torch_groups=tg, **sched(o.lr, o.get("lr_end", None), o.get("sched", None)))]
    first, second = [], []           # optim_pose.param_groups[0] / [1]
    if hasattr(graph, "se3_refine"):
        first += list(graph.se3_refine.parameters())
    if hasattr(graph, "warp_mlp"):
        first += list(graph.warp_mlp.parameters())
        second += list(graph.warp_latent.parameters())
    if hasattr(graph, "pose_net"):
        first += list(graph.pose_net.pose_embedding.parameters())
        second += list(graph.pose_net.pose_latent.parameters())
    if first or second:
        lr = o.get("lr_pose", None) or o.lr
        groups.append(dict(params=first + second, warmup=o.get("warmup_pose", None) or 0, warmup_params=len(first),
                           torch_groups=[n for n in (len(first), len(second)) if n],
                           **sched(lr, o.get("lr_pose_end", None), o.get("sched_pose", None))))
    return groups


def shard_ray_idx(ray_idx, rank, world):
    """Contiguous 1/world slice of the per-image pixel list (same list on every rank)."""
    n = ray_idx.numel()
    per = (n + world - 1) // world
    return ray_idx[rank * per:min((rank + 1) * per, n)]


class _ShardedRandperm:
    """Makes ``torch.randperm`` inside ``Graph.forward`` return this rank's slice of the global
    draw: every rank consumes the same generator state (seeded identically), keeps the first
    ``rand_rays // B`` entries as the global batch and takes its contiguous 1/k of them."""

    def __init__(self, rank, world, global_per_image):
        self.rank, self.world, self.n = rank, world, global_per_image
        self._orig = None

    def __enter__(self):
        self._orig = torch.randperm
        orig, rank, world, n = self._orig, self.rank, self.world, self.n

        def randperm(*a, **k):
            return shard_ray_idx(orig(*a, **k)[:n], rank, world)
        torch.randperm = randperm
        # where this rank's rays sit in the global per-image list (the NVP warp's annealing quirk is keyed on it)
        from . import functional as F
        per = (n + world - 1) // world
        F.ray_shard = (rank * per, n)
        return self

    def __exit__(self, *exc):
        from . import functional as F
        torch.randperm = self._orig
        F.ray_shard = None


class device_ray_draws:
    """Device-side random draws of a training step without torch RNG ops (a captured step then needs no generator-state
    fills in front of every replay).  Makes ``torch.randperm(n, device=cuda)[:k]`` inside ``Graph.forward`` an O(k) kernel: the reference sorts
    H*W random keys to keep ``rand_rays // B`` of them (model/nerf.py:268; five radix-sort passes per step).
    ``torch.randperm`` is replaced by a lazy stand-in whose ``[:k]`` slice launches ``F.sample_pixels`` -- the
    first k entries of a random permutation, drawn from the library's own counter-based stream (capturable:
    replays advance the device counter).  Under data parallelism every rank passes the same seed and counter
    value, so all ranks see the same global draw and ``_ShardedRandperm`` still slices it consistently."""

    class _Lazy:
        def __init__(self, n, counter, seed):
            self.n, self.counter, self.seed = n, counter, seed

        def __getitem__(self, sl):
            if not (isinstance(sl, slice) and sl.start in (None, 0) and sl.step in (None, 1) and sl.stop is not None):
                raise TypeError("device_ray_draws: only randperm(n)[:k] is supported")
            from . import functional as F
            return F.sample_pixels(self.n, min(int(sl.stop), self.n), self.counter, self.seed)

    def __init__(self, device, seed=0):
        self.counter = torch.zeros(1, dtype=torch.int64, device=device)
        self.rng = torch.zeros(2, dtype=torch.int64, device=device)     # stratified uniforms: call number, block ticket
        self.seed = seed

    def __enter__(self):
        self._perm, self._rand = torch.randperm, torch.rand
        me = self

        def randperm(n, **k):
            dev = k.get("device", None)
            if dev is None or torch.device(dev).type != "cuda":
                return me._perm(n, **k)
            return device_ray_draws._Lazy(int(n), me.counter, me.seed)

        def rand(*shape, **k):
            # the stratified jitter torch.rand(B, P, N, 1, device=cuda) of Graph.sample_depth (model/nerf.py:338): drawn inside
            # the sampling kernel instead (functional.DeviceUniform) -- no torch RNG op in the captured step
            dev = k.get("device", None)
            if dev is None or torch.device(dev).type != "cuda" or len(shape) != 4 or shape[-1] != 1 or k.get("generator") is not None:
                return me._rand(*shape, **k)
            from . import functional as F
            return F.DeviceUniform(shape, me.rng, me.seed + 0x9E3779B97F4A7C15)
        torch.randperm, torch.rand = randperm, rand
        return self

    def __exit__(self, *exc):
        torch.randperm, torch.rand = self._perm, self._rand


class feed_draws:
    """Feed host-made random draws to ``Graph.forward``: the next ``torch.randperm`` returns
    ``ray_idx`` and the next ``torch.rand`` returns ``u`` (both already on the device).  This is
    the parity-mode RNG path: uniforms and ray indices come from the caller's
    generator instead of the device generator, e.g. a host data loader or a recorded reference run."""

    def __init__(self, ray_idx=None, u=None):
        self.ray_idx, self.u = ray_idx, u

    def __enter__(self):
        self._rand, self._perm = torch.rand, torch.randperm
        me = self

        def rand(*shape, **k):
            if me.u is None:
                return me._rand(*shape, **k)
            u, me.u = me.u, None
            return u.view(*shape)

        def randperm(n, **k):
            if me.ray_idx is None:
                return me._perm(n, **k)
            r, me.ray_idx = me.ray_idx, None
            return r
        torch.rand, torch.randperm = rand, randperm
        return self

    def __exit__(self, *exc):
        torch.rand, torch.randperm = self._rand, self._perm


def backward_schedule(graph):
    """The graph's functional.BackwardOverlap (made once): opts the network whose backward runs LAST (the coarse NeRF) in
    to the side-stream weight-gradient pass when there is a pose / warp backward to hide it under."""
    from . import functional as F
    ov = getattr(graph, "_backward_overlap", None)
    if ov is None:
        has_warp = hasattr(graph, "warp_mlp") or hasattr(graph, "pose_net") or hasattr(graph, "se3_refine")
        graph.nerf.overlap_weight_gradients = bool(has_warp)
        ov = graph._backward_overlap = F.BackwardOverlap()
    return ov


class data_parallel:
    """``with data_parallel(group):`` -- forward + loss of a step whose per-image ray / point lists are sharded over the
    ranks of ``group`` (None = the default group).  Everything on the path is per-ray except the rigid fit behind the
    ``global_alignment`` loss (reference model/nerf_inn_llff.py:563-572, model/pose_models/inn.py:96-102), which must see an
    image's whole list: inside this context it is computed from rank-summed sufficient statistics."""

    def __init__(self, group=None):
        self.group = group

    def __enter__(self):
        from . import functional as F
        self._old = F.data_parallel_group
        F.data_parallel_group = True if self.group is None else self.group
        return self

    def __exit__(self, *exc):
        from . import functional as F
        F.data_parallel_group = self._old


def train_step(opt, graph, var, it, bucket=None, rank=0, world=1, overlap_dw=True, group=None, optimizer=None):
    """One optimisation step minus the optimiser: forward, loss, backward (+ all-reduce).

    With ``world > 1`` the rank renders a contiguous 1/world slice of the global ray batch; local
    losses are means over the local rays, so they are scaled by ``n_local / n_global`` before the
    summed all-reduce -- the reduced gradient equals the single-GPU gradient of the
    same global batch.  Returns the loss dict (``all`` is the local, scaled value).

    With a flat gradient ``bucket`` (the engine owns every ``.grad``) the backward is scheduled on two streams
    (``overlap_dw``): see functional.BackwardOverlap.  The streams are joined before this function returns.

    ``optimizer`` (a ``FlatAdam`` that is also the ``bucket``): the update is part of the step -- the pose / warp groups are
    updated on the launching stream as soon as their gradients are summed, while the weight-gradient pass (and the sum of
    the NeRF segment behind it) still runs on the side stream; the NeRF group follows after the join.
    """
    B = len(var.idx)
    n_global = opt.nerf.rand_rays // B
    takes_iter = opt.model in ("barf_inn_llff", "nerf_inn_llff", "barf_inn_dtu", "nerf_inn_dtu")
    if bucket is not None:
        bucket.zero()
        if not getattr(graph, "_flat_gradients", False):
            use_flat_gradients(graph)          # the bucket owns every .grad: let the kernels accumulate into it
            graph._flat_gradients = True
    else:
        graph.zero_grad(set_to_none=True)
    if world > 1:
        # (the per-image rigid fit of the global-alignment loss spans the WHOLE list of an image's points, which is now
        # spread over the ranks: data_parallel makes camera.rigid_points_registration all-reduce its 15 sums per image)
        with _ShardedRandperm(rank, world, n_global), data_parallel(group):
            var = graph.forward(opt, var, mode="train", iter=it) if takes_iter else graph.forward(opt, var, mode="train")
            loss = summarize_loss(opt, graph.compute_loss(opt, var, mode="train"))
    else:
        var = graph.forward(opt, var, mode="train", iter=it) if takes_iter else graph.forward(opt, var, mode="train")
        loss = summarize_loss(opt, graph.compute_loss(opt, var, mode="train"))
    scale = 1.0
    if world > 1:
        scale = len(var.ray_idx) / float(n_global)
    if bucket is not None:
        # the weight-gradient pass of the (last) NeRF network runs on a side stream under the pose / warp backward
        sched = backward_schedule(graph) if overlap_dw else contextlib.nullcontext()
        early = optimizer is not None and optimizer is bucket and len(optimizer.groups) > 1
        with sched:
            if world > 1:
                with overlap_allreduce(graph, bucket, group):
                    (loss.all * scale).backward()
                bucket.allreduce(group)
            else:
                loss.all.backward()
            if early:
                optimizer.step(groups=range(1, len(optimizer.groups)))
        if optimizer is not None:
            optimizer.step(groups=[0] if early else None)
    else:
        (loss.all if scale == 1.0 else loss.all * scale).backward()
        if optimizer is not None:
            optimizer.step()
    return loss


class _frozen:
    """``with _frozen(graph):`` -- the graph's parameters do not require gradients inside (test-time pose refinement
    differentiates THROUGH the networks, never into them; the MLP backward then skips its weight-gradient pass)."""

    def __init__(self, graph):
        self.params = [p for p in graph.parameters() if p.requires_grad]

    def __enter__(self):
        for p in self.params:
            p.requires_grad_(False)

    def __exit__(self, *exc):
        for p in self.params:
            p.requires_grad_(True)


def test_time_photometric_optim(opt, graph, var, iters=None, lr=None, on_step=None, captured=None, seed=0):
    """``Model.evaluate_test_time_photometric_optim`` (reference model/barf.py:153-169): absorb the remaining pose
    error of ONE held-out view in a fresh se(3) parameter by ``opt.optim.test_iter`` Adam steps on the photometric
    loss of ``rand_rays`` random pixels per step (``mode="test-optim"``: the pose gradient comes out of the
    ray-generation kernel's backward).  Returns ``var`` with ``se3_refine_test`` / ``pose_refine_test`` set, as the
    reference does.

    ``captured`` (default: whenever no ``on_step`` callback asks to look at every iterate): the iteration -- device-side
    pixel draw, se(3) -> SE(3), ray generation, render, loss, backward into the 6 pose numbers, ``FlatAdam`` update -- is
    captured ONCE in a CUDA graph and replayed; the loop then costs one graph launch per iteration and no host
    synchronisation.  Otherwise the loop runs eagerly with ``torch.optim`` exactly as the reference writes
    it and ``on_step(it, loss, se3)`` is called after every update (tests, logging)."""
    n_iter = opt.optim.test_iter if iters is None else iters
    lr = opt.optim.lr_pose if lr is None else lr
    var.se3_refine_test = torch.nn.Parameter(torch.zeros(1, 6, device=opt.device))
    if captured is None:
        captured = on_step is None and n_iter > 3 and torch.device(opt.device).type == "cuda" \
            and not torch.cuda.is_current_stream_capturing()
    if captured:
        if opt.optim.algo != "Adam":
            raise NotImplementedError("captured test-time refinement implements optim.algo=Adam")
        adam = FlatAdam([dict(params=[var.se3_refine_test], lr=lr)])
        draws = device_ray_draws(opt.device, seed=seed)

        def body():
            adam.zero()
            var.pose_refine_test = camera.lie.se3_to_SE3(var.se3_refine_test)
            with draws:
                v = graph.forward(opt, var, mode="test-optim")
            loss = summarize_loss(opt, graph.compute_loss(opt, v, mode="test-optim"))
            loss.all.backward()
            adam.step()
            return loss.all.detach()
        warm = 2
        with torch.enable_grad(), _frozen(graph):
            step = CapturedStep(body, warmup=warm)        # the warm-up runs are real iterations (they advance the optimiser)
            for _ in range(max(n_iter - warm, 0)):
                step()
        with torch.no_grad():
            var.pose_refine_test = camera.lie.se3_to_SE3(var.se3_refine_test)
        var.test_optim_steps = adam.state[0, 0]           # device scalar: iterations taken (no host read here)
        return var
    optimizer = getattr(torch.optim, opt.optim.algo)
    optim_pose = optimizer([dict(params=[var.se3_refine_test], lr=lr)])
    with torch.enable_grad():
        for it in range(n_iter):
            optim_pose.zero_grad()
            var.pose_refine_test = camera.lie.se3_to_SE3(var.se3_refine_test)
            var = graph.forward(opt, var, mode="test-optim")
            loss = summarize_loss(opt, graph.compute_loss(opt, var, mode="test-optim"))
            loss.all.backward()
            optim_pose.step()
            if on_step is not None:
                on_step(it, loss, var.se3_refine_test)
    return var


@torch.no_grad()
def evaluate_view(opt, graph, var, test_optim=None, fine=None):
    """One iteration of ``Model.evaluate_full`` (reference model/nerf.py:162-183) without the file dumps and LPIPS
    (a library network): optional test-time pose refinement, full-frame ``mode="eval"`` render, PSNR and SSIM of the
    rendered image in one kernel (csrc/metrics.cu) that reads the renderer's [B,HW,3] output directly.  Returns
    ``AttrDict(psnr [B], ssim [B], var)`` with device tensors: nothing is read back to the host here."""
    from . import functional as F
    if test_optim is None:
        # model/nerf.py:172 (barf) and model/nerf_inn_dtu.py:217 ('barf' in opt.model).  model/nerf_inn_llff.py:202 lists
        # model names that do not exist in the reference, so barf_inn_llff never refines there and its next line,
        # get_pose(mode="eval"), then reads the unset var.pose_refine_test (barf_inn_llff.py:395-396): the evident
        # intent -- refine whenever optim.test_photo is set -- is implemented for all three
        test_optim = "barf" in opt.model.lower() and bool(opt.optim.get("test_photo", False))
    if test_optim:
        var = test_time_photometric_optim(opt, graph, var)
    elif opt.optim.get("test_photo", False) and "pose_refine_test" not in var:
        # refinement explicitly skipped although the eval pose composes it: the identity is the unrefined pose
        var.pose_refine_test = torch.eye(3, 4, device=opt.device)[None].repeat(len(var.idx), 1, 1)
    takes_iter = opt.model in ("barf_inn_llff", "nerf_inn_llff", "barf_inn_dtu", "nerf_inn_dtu")
    var = graph.forward(opt, var, mode="eval", iter=None) if takes_iter else graph.forward(opt, var, mode="eval")
    use_fine = opt.nerf.fine_sampling if fine is None else fine
    rgb = var.rgb_fine if (use_fine and "rgb_fine" in var) else var.rgb
    psnr, ssim = F.image_metrics(rgb, var.image.view(-1, 3, opt.H, opt.W), opt.H, opt.W)
    return AttrDict(psnr=psnr, ssim=ssim, var=var)


@torch.no_grad()
def render_rows(opt, graph, pose, intr, row0, row1, depth_range=None, mode="eval"):
    """Pixel rows [row0, row1) of a full-frame render: the slice loop of ``render_by_slices`` (reference
    model/nerf.py:321-332 iterates contiguous ``ray_idx`` ranges of ``rand_rays`` pixels) restricted to one contiguous
    range, which is how an evaluation frame is sharded over k GPUs (rank r takes rows [r H/k, (r+1) H/k); SURVEY.md 8e:
    no collective, the rows are independent).  The range is rendered in chunks of ``opt.nerf.rand_rays`` pixels (one
    chunk when that covers it).  Returns the same edict as ``render`` with [B, (row1-row0) W, K] tensors."""
    start, stop = int(row0) * opt.W, int(row1) * opt.W
    chunk = int(opt.nerf.rand_rays) if opt.nerf.rand_rays else stop - start
    parts = []
    for c in range(start, stop, chunk):
        parts.append(graph._render_pose(opt, pose, intr=intr, mode=mode, depth_range=depth_range, idx_start=c,
                                        num=min(chunk, stop - c)))
    if len(parts) == 1:
        return parts[0]
    return AttrDict({k: torch.cat([p[k] for p in parts], dim=1) for k in parts[0].keys()})


class CapturedStep:
    """A whole training step (forward, loss, backward, all-reduce, optimiser) captured once in a CUDA
    graph and replayed: the ~45 kernel launches of a C2 step cost one graph launch instead of ~1 ms of
    Python / launch overhead.  ``fn`` must be capture-safe: static input tensors, device RNG
    (``torch.rand`` / ``torch.randperm`` advance the graph-registered Philox state on replay), no host
    reads, optimisers built with ``capturable=True``.  Its return value (e.g. the loss dict) refers
    to static tensors that every replay overwrites."""

    def __init__(self, fn, warmup=3):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = fn()

    def __call__(self):
        self.graph.replay()
        return self.out


def synthetic_var(opt, B, seed=0, dtu=False):
    """A ``var`` batch shaped like ``train_data.all`` (data/llff.py:79-92, data/dtu.py:369-380) from
    the deterministic generators in ``synthetic``."""
    from . import synthetic as syn
    dev = opt.device
    var = AttrDict(idx=torch.arange(B, device=dev), image=syn.images(seed, B, opt.H, opt.W).to(dev),
                   intr=syn.intrinsics(B, opt.H, opt.W, 1.8 if dtu else 0.81).to(dev),
                   pose=(syn.dtu_poses(seed + 1, B) if dtu else syn.llff_poses(seed + 1, B)).to(dev))
    if dtu:
        var.depth_range = torch.tensor([[1.2, 5.2]]).repeat(B, 1).to(dev)
    return var
