"""Data-parallel host logic on CPU: two ``gloo`` ranks shard one global ray batch the way
``engine.train_step`` does (engine._ShardedRandperm -> contiguous 1/k slices of the shared pixel
draw, local mean losses scaled by n_local/n_global, ONE flat-bucket all-reduce) and must end up with
the gradient of the single-process step on the whole batch.  The per-rank compute
is the CPU oracle (the product kernels need a GPU); what is under test is the sharding, the loss
scaling and the GradBucket plumbing -- the same objects bench.py drives on NCCL."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

B, P_GLOBAL, N, H, W = 2, 12, 8, 24, 32
ALPHA = 1.0
CFG = dict(N=N, Nf=None, range=[1, 0], param="inverse", L_3D=10, L_view=4, skip=(4,), c2f=[0.1, 0.5])


class _Holder(nn.Module):
    """Parameters under the reference's names, so engine.trainable_parameters / GradBucket see what
    they see on a real Graph (nerf.*, warp_mlp.*, warp_latent.weight; progress is skipped)."""

    def __init__(self):
        super().__init__()
        from neural_invertible_warp_b200 import synthetic as syn
        self.nerf = nn.ParameterDict({k.replace(".", "__"): nn.Parameter(v) for k, v in syn.nerf_params(1).items()})
        self.warp_mlp = nn.ParameterDict({k.replace(".", "__"): nn.Parameter(v) for k, v in syn.nvp_params(2).items()})
        self.warp_latent = nn.Embedding(B, 128, _weight=syn.latent_codes(3, B))
        self.progress = nn.Parameter(torch.tensor(0.3))

    def dicts(self):
        return ({k.replace("__", "."): v for k, v in self.nerf.items()},
                {k.replace("__", "."): v for k, v in self.warp_mlp.items()})


def _local_step(holder, ray_idx, u, image, intr, scale):
    from oracle import reference_port as ora
    p, q = holder.dicts()
    ray, center, *_ = ora.warped_rays(q, holder.warp_latent.weight, H, W, intr, ray_idx, ALPHA)
    out = ora.render_rays(p, center, ray, u, CFG, progress=0.3)
    loss = ora.mse(out["rgb"], ora.gather_pixels(image, ray_idx))
    (loss * scale).backward()
    return loss.detach()


def _inputs():
    from neural_invertible_warp_b200 import synthetic as syn
    gen = torch.Generator().manual_seed(11)
    u = torch.rand(B, P_GLOBAL, N, 1, generator=gen)
    return u, syn.images(4, B, H, W), syn.intrinsics(B, H, W, 0.81)


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from neural_invertible_warp_b200 import engine
    holder = _Holder()
    bucket = engine.GradBucket(holder)
    bucket.zero()
    u, image, intr = _inputs()
    torch.manual_seed(5)                      # every rank draws the same permutation ...
    with engine._ShardedRandperm(rank, world, P_GLOBAL):
        ray_idx = torch.randperm(H * W)       # ... and keeps its contiguous 1/k of the first P_GLOBAL
    per = (P_GLOBAL + world - 1) // world
    u_loc = u[:, rank * per:(rank + 1) * per]
    _local_step(holder, ray_idx, u_loc, image, intr, scale=len(ray_idx) / float(P_GLOBAL))
    bucket.allreduce()
    if rank == 0:
        torch.save(dict(flat=bucket.flat.clone(), n_local=len(ray_idx)), out_path)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.timeout(300)
def test_two_rank_sharded_step_equals_single_process_step(tmp_path):
    out = str(tmp_path / "dp.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    assert got["n_local"] == P_GLOBAL // 2

    from neural_invertible_warp_b200 import engine
    holder = _Holder()
    bucket = engine.GradBucket(holder)
    u, image, intr = _inputs()
    torch.manual_seed(5)
    ray_idx = torch.randperm(H * W)[:P_GLOBAL]
    _local_step(holder, ray_idx, u, image, intr, scale=1.0)
    ref = bucket.flat
    assert ref.abs().max() > 0
    rel = ((got["flat"] - ref).norm() / ref.norm()).item()
    assert rel < 1e-5, rel
    # progress / Kabsch outputs are never part of the bucket
    names = [n for n, _ in engine.trainable_parameters(holder)]
    assert "progress" not in names and len(names) == len(list(holder.parameters())) - 1


def test_shard_ray_idx_partitions_the_global_list():
    from neural_invertible_warp_b200 import engine
    idx = torch.arange(100, 110)
    for world in (1, 2, 3, 4, 8):
        parts = [engine.shard_ray_idx(idx, r, world) for r in range(world)]
        assert torch.equal(torch.cat(parts), idx)
        assert max(len(p) for p in parts) == (10 + world - 1) // world


class _Segments:
    """The bucket interface of engine.FlatAdam (flat gradient buffer + per-group segments) on CPU tensors."""

    def __init__(self, sizes):
        from neural_invertible_warp_b200 import engine
        self.__class__ = type("_SegBucket", (engine.SegmentedAllreduce,), {})
        self.groups, off = [], 0
        for n in sizes:
            self.groups.append(dict(offset=off, n=n, params=[]))
            off += n
        self.flat = torch.zeros(off)


def _seg_worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = {}

    class _Chan:                                 # stands in for p2p.P2PChannel (peer memory needs GPUs): same interface
        calls = 0

        def allreduce_(self, t):
            _Chan.calls += 1
            dist.all_reduce(t)
            return t

    for early in ((), (0,), (0, 1), (1,), ("p2p",), ("p2p", 0)):
        b = _Segments([8, 4, 12])
        if early and early[0] == "p2p":
            assert b.enable_p2p() is False       # gloo / CPU ranks: the peer-memory path is not available, NCCL / gloo stays
            b._p2p = [_Chan() for _ in b.groups]  # the routing with channels in place: one channel call per segment
            early_groups, calls0 = early[1:], _Chan.calls
        else:
            early_groups = early
        b.flat.copy_(torch.arange(24.0) * (rank + 1))
        for gi in early_groups:
            b.allreduce_group_async(gi)          # "this segment's gradients are final": reduced while backward continues
            b.allreduce_group_async(gi)          # idempotent within a step
        b.allreduce()                            # the rest + join
        assert not b._pending
        if early and early[0] == "p2p":
            assert _Chan.calls - calls0 == len(b.groups)
        res[early] = b.flat.clone()
    if rank == 0:
        torch.save(res, out_path)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_early_segment_allreduce_equals_one_collective(tmp_path):
    """engine.SegmentedAllreduce / overlap_allreduce host logic: reducing the NeRF segment early (asynchronously) and
    the rest at the end gives exactly the single whole-buffer all-reduce, every step."""
    out = str(tmp_path / "seg.pt")
    mp.spawn(_seg_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    want = torch.arange(24.0) * 3
    for early, flat in res.items():
        assert torch.equal(flat, want), early


def _kabsch_worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from neural_invertible_warp_b200 import camera, engine, functional as F
    from oracle import reference_port as ora
    # the host path under test is the product's (camera.rigid_points_registration -> F.kabsch_sharded -> all_reduce -> solve);
    # only the two CUDA kernels are stood in for by the oracle, as everywhere in this CPU suite
    F.kabsch_stats, F.kabsch_solve = ora.kabsch_stats, ora.kabsch_from_stats
    x, y = _kabsch_points()
    per = x.shape[1] // world
    xs, ys = x[:, rank * per:(rank + 1) * per], y[:, rank * per:(rank + 1) * per]
    assert F.data_parallel_group is None
    with engine.data_parallel():
        R, t = camera.rigid_points_registration(xs, ys)
    assert F.data_parallel_group is None
    if rank == 1:
        torch.save(dict(R=R, t=t), out_path)
    dist.barrier()
    dist.destroy_process_group()


def _kabsch_points():
    gen = torch.Generator().manual_seed(31)
    Bk, M = 3, 40
    x = torch.randn(Bk, M, 3, generator=gen) + torch.tensor([0.5, -2.0, 4.0])
    w = torch.randn(Bk, 3, generator=gen) * 0.6
    from oracle import reference_port as ora
    Rt = ora.se3_to_SE3(torch.cat([w, torch.zeros(Bk, 3)], dim=-1))[..., :3]
    y = x @ Rt.transpose(1, 2) + torch.randn(Bk, 1, 3, generator=gen) + 0.02 * torch.randn(Bk, M, 3, generator=gen)
    return x, y


@pytest.mark.timeout(300)
def test_sharded_rigid_fit_equals_whole_list_fit(tmp_path):
    """the global-alignment fit (reference model/nerf_inn_llff.py:566-572) spans all of an image's points; with
    the rows sharded over two ranks, the statistics all-reduce inside ``engine.data_parallel`` must give every rank the fit
    of the whole list -- not the fit of its own shard."""
    from oracle import reference_port as ora
    out = str(tmp_path / "kabsch.pt")
    mp.spawn(_kabsch_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    x, y = _kabsch_points()
    R, t = ora.kabsch(x, y)
    torch.testing.assert_close(got["R"], R, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(got["t"], t, rtol=1e-5, atol=1e-5)
    R_shard, _ = ora.kabsch(x[:, 20:], y[:, 20:])                    # what rank 1 would have fitted on its own
    assert (R_shard - R).abs().max() > 1e-4
