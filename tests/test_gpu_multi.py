"""Multi-GPU parity on hardware (needs >= 2 CUDA devices; run with ``gpurun --gpus 2``): a two-rank sharded
``engine.train_step`` with the real kernels and NCCL -- ray shards, early NeRF-segment all-reduce, side-stream dW,
statistics all-reduce of the global-alignment fit -- must reproduce the single-GPU step on the same global batch
(reference loss model/nerf_inn_llff.py:548-573, training command scripts/train_llff.sh:1 uses
--loss_weight.global_alignment=4)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu
B, P_GLOBAL, N, H, W = 4, 64, 32, 48, 64


def _build(dev, precision):
    from neural_invertible_warp_b200 import config as cfgmod, engine, synthetic as syn
    opt = cfgmod.builtin_options("barf_inn_llff", barf_c2f=[0.1, 0.5], device=dev, data=dict(image_size=[H, W]),
                                 nerf=dict(rand_rays=B * P_GLOBAL, sample_intvs=N), loss_weight=dict(global_alignment=4),
                                 arch=dict(mlp_precision=precision))
    torch.manual_seed(0)
    graph = engine.build_graph(opt, B)
    sd = graph.nerf.state_dict()
    graph.nerf.load_state_dict({**sd, **{k: v.to(dev) for k, v in syn.nerf_params(21).items()}})
    graph.nerf.progress.data.fill_(0.3)
    graph.warp_latent.weight.data = syn.latent_codes(22, B).to(dev)
    graph.warp_mlp.load_state_dict({k: v.to(dev) for k, v in syn.nvp_params(23).items()})
    return opt, graph, engine.synthetic_var(opt, B, 24)


def _draws(dev):
    gen = torch.Generator().manual_seed(77)
    return torch.randperm(H * W, generator=gen)[:P_GLOBAL].to(dev), torch.rand(B, P_GLOBAL, N, 1, generator=gen).to(dev)


def _worker(rank, world, port, precision, out_path, collective="nccl"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = "cuda:%d" % rank
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    from neural_invertible_warp_b200 import config as cfgmod, engine
    opt, graph, var = _build(dev, precision)
    adam = engine.FlatAdam(engine.reference_optimizer_groups(opt, graph))
    if collective == "p2p":
        assert adam.enable_p2p(), "peer-memory all-reduce unavailable on this box"
    elif collective == "p2p-fails":
        # one rank cannot set its exchange block up: EVERY rank must fall back to NCCL (no hang, no mixed collectives)
        import warnings
        os.environ["NIW_P2P_TEST_FAIL"] = "1"
        with warnings.catch_warnings(record=True) as caught:
            warnings.simplefilter("always")
            assert adam.enable_p2p() is False and adam._p2p is None
        assert any("NCCL" in str(w.message) for w in caught)
        del os.environ["NIW_P2P_TEST_FAIL"]
    ridx, u = _draws(dev)
    per = (P_GLOBAL + world - 1) // world
    for rep in range(3 if collective == "p2p" else 1):      # p2p: the exchange buffers are double-buffered on a sequence number
        with engine.feed_draws(ray_idx=ridx, u=u[:, rank * per:(rank + 1) * per].contiguous()):
            loss = engine.train_step(opt, graph, cfgmod.AttrDict(var), 5000, bucket=adam, rank=rank, world=world)
    torch.cuda.synchronize()
    if collective == "p2p":
        assert all(ch.error() == 0 for ch in adam._p2p)
    total = (loss.all.detach() * (per / float(P_GLOBAL))).clone()
    dist.all_reduce(total)
    if rank == 0:
        torch.save(dict(flat=adam.flat.cpu(), global_rigid=graph.global_rigid.weight.data.cpu(), loss=float(total),
                        loss_ga=float(loss.global_alignment.detach())), out_path)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.timeout(600)
@pytest.mark.parametrize("precision,collective", [("fp32", "nccl"), ("bf16", "nccl"), ("bf16", "p2p"), ("bf16", "p2p-fails")])
def test_two_gpu_sharded_step_equals_single_gpu_step(tmp_path, precision, collective):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    out = str(tmp_path / "dp.pt")
    mp.spawn(_worker, args=(2, _free_port(), precision, out, collective), nprocs=2, join=True)
    got = torch.load(out)
    from neural_invertible_warp_b200 import config as cfgmod, engine
    dev = "cuda:0"
    opt, graph, var = _build(dev, precision)
    adam = engine.FlatAdam(engine.reference_optimizer_groups(opt, graph))
    ridx, u = _draws(dev)
    with engine.feed_draws(ray_idx=ridx, u=u):
        loss = engine.train_step(opt, graph, cfgmod.AttrDict(var), 5000, bucket=adam)
    torch.cuda.synchronize()
    ref = adam.flat.cpu().double()
    rel = ((got["flat"].double() - ref).norm() / ref.norm()).item()
    print("[2 GPUs, %s, %s] reduced gradient vs single GPU rel-L2 %.3e; loss %.6f vs %.6f" % (precision, collective, rel, got["loss"], float(loss.all)))
    assert rel < (2e-4 if precision == "fp32" else 2e-3), rel
    # the rigid fit saw every image's WHOLE point list (statistics all-reduce), not rank 0's shard
    torch.testing.assert_close(got["global_rigid"], graph.global_rigid.weight.data.cpu(), rtol=1e-4, atol=1e-5)
    assert abs(got["loss"] - float(loss.all.detach())) <= 2e-5 * max(1.0, abs(float(loss.all.detach())))


def _p2p_worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = "cuda:%d" % rank
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    from neural_invertible_warp_b200 import p2p
    assert p2p.available()
    n = 530052
    ch = p2p.P2PChannel(n)
    gen = torch.Generator().manual_seed(5 + rank)
    ok = True
    graph = torch.cuda.CUDAGraph()
    static = torch.zeros(n, device=dev)
    for it in range(6):
        x = torch.randn(n, generator=gen).to(dev)
        ref = x.clone()
        dist.all_reduce(ref)
        if it < 3:                                   # eager calls, then the same channel captured in a CUDA graph and replayed
            got = ch.allreduce_(x.clone())
        else:
            static.copy_(x)
            if it == 3:
                torch.cuda.synchronize()
                with torch.cuda.graph(graph):
                    ch.allreduce_(static)
            graph.replay()
            got = static.clone()
        torch.cuda.synchronize()
        # NCCL's ring / tree order differs from the rank-order sum: equal up to fp32 rounding, and bit-identical across ranks
        ok = ok and bool(torch.allclose(got, ref, rtol=1e-5, atol=1e-5))
        mine = [torch.empty_like(got) for _ in range(world)]
        dist.all_gather(mine, got)
        ok = ok and all(torch.equal(mine[0], m) for m in mine)
    ok = ok and ch.error() == 0
    ch.close()
    if rank == 0:
        torch.save(dict(ok=ok), out_path)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_p2p_allreduce_matches_nccl_and_is_identical_on_all_ranks(tmp_path):
    """csrc/p2p.cu: the peer-memory sum of a gradient-sized vector equals the NCCL all-reduce up to fp32 summation order,
    is bit-identical on every rank (rank-order sum), survives repeated calls on the double-buffered exchange block and
    replays from a CUDA graph."""
    n_dev = torch.cuda.device_count()
    if n_dev < 2:
        pytest.skip("needs two CUDA devices")
    world = min(n_dev, 8)
    out = str(tmp_path / "p2p.pt")
    mp.spawn(_p2p_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert torch.load(out)["ok"]
