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


def _worker(rank, world, port, precision, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = "cuda:%d" % rank
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    from neural_invertible_warp_b200 import config as cfgmod, engine
    opt, graph, var = _build(dev, precision)
    adam = engine.FlatAdam(engine.reference_optimizer_groups(opt, graph))
    ridx, u = _draws(dev)
    per = (P_GLOBAL + world - 1) // world
    with engine.feed_draws(ray_idx=ridx, u=u[:, rank * per:(rank + 1) * per].contiguous()):
        loss = engine.train_step(opt, graph, cfgmod.AttrDict(var), 5000, bucket=adam, rank=rank, world=world)
    torch.cuda.synchronize()
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
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_two_gpu_sharded_step_equals_single_gpu_step(tmp_path, precision):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    out = str(tmp_path / "dp.pt")
    mp.spawn(_worker, args=(2, _free_port(), precision, out), nprocs=2, join=True)
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
    print("[2 GPUs, %s] reduced gradient vs single GPU rel-L2 %.3e; loss %.6f vs %.6f" % (precision, rel, got["loss"], float(loss.all)))
    assert rel < (2e-4 if precision == "fp32" else 2e-3), rel
    # the rigid fit saw every image's WHOLE point list (statistics all-reduce), not rank 0's shard
    torch.testing.assert_close(got["global_rigid"], graph.global_rigid.weight.data.cpu(), rtol=1e-4, atol=1e-5)
    assert abs(got["loss"] - float(loss.all.detach())) <= 2e-5 * max(1.0, abs(float(loss.all.detach())))
