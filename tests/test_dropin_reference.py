"""The drop-in behind the UNMODIFIED reference engine (needs the reference tree: /root/reference here, or
NIW_REFERENCE_ROOT / baseline/_ref on another box; skipped otherwise).

``dropin.install_dropin`` patches the reference's ``model.<name>.Graph / NeRF``, ``model.nvp.nvp_ndr.DeformNetwork`` and
``INNPoseParams``; the reference's own ``Model.build_networks`` (model/base.py:34-37, model/barf_inn_llff.py:25-82,
model/barf_inn_dtu.py:323-336) and ``setup_optimizer`` (model/nerf_inn_llff.py:34-47, model/barf_inn_llff.py:84-104) then
build and own the B200 graph; checkpoints go through the reference's ``util.save_checkpoint`` / ``restore_checkpoint``
per-child loader (util.py:124-163).  With a GPU, the body of ``Model.train_iteration`` runs on the CUDA path."""
import contextlib
import io
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.reference


def _reference_root():
    for cand in (os.environ.get("NIW_REFERENCE_ROOT"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "model")) and os.path.exists(os.path.join(cand, "train.py")):
            return cand
    return None


REF = _reference_root()
needs_reference = pytest.mark.skipif(REF is None, reason="reference tree not available")


class _FakeTrainData:
    """What ``Model.build_networks`` asks of ``train_data``: its length and the (LLFF: unused-by-value) camera poses."""

    def __init__(self, opt, B, var=None):
        self.B, self.opt, self.all = B, opt, var

    def __len__(self):
        return self.B

    def get_all_camera_poses(self, opt):
        from neural_invertible_warp_b200 import synthetic as syn
        return syn.llff_poses(1, self.B)


@pytest.fixture()
def patched(tmp_path):
    os.environ["NIW_REFERENCE_ROOT"] = REF
    from oracle import ref_shim
    ref_shim.install()
    from neural_invertible_warp_b200 import dropin
    with ref_shim.in_reference_dir(), contextlib.redirect_stdout(io.StringIO()):
        mods = dropin.install_dropin(REF)
    yield mods, ref_shim, tmp_path
    dropin.uninstall_dropin()


def _opt(ref_shim, tmp_path, device, B=3, **over):
    opt = ref_shim.load_reference_options("barf_inn_llff", "barf_inn_llff", overrides=dict(
        barf_c2f=[0.1, 0.5], output_path=str(tmp_path), data=dict(image_size=[24, 32]),
        nerf=dict(rand_rays=B * 16, sample_intvs=16), freq=dict(scalar=10 ** 9, vis=10 ** 9), tb=False, visdom=False,
        max_iter=100, **over))
    opt.device = device
    opt.H, opt.W = opt.data.image_size
    return opt


def _build(mods, opt, B):
    with contextlib.redirect_stdout(io.StringIO()):
        m = mods["barf_inn_llff"].Model(opt)
        m.train_data = _FakeTrainData(opt, B)
        m.build_networks(opt)
        m.setup_optimizer(opt)
    return m


@needs_reference
def test_reference_engine_builds_the_b200_graph(patched):
    mods, ref_shim, tmp_path = patched
    from neural_invertible_warp_b200 import nvp
    from neural_invertible_warp_b200.model import barf_inn_llff as ours
    B = 3
    opt = _opt(ref_shim, tmp_path, "cpu", B)
    m = _build(mods, opt, B)
    assert type(m.graph) is ours.Graph and type(m.graph.nerf) is ours.NeRF
    assert type(m.graph.warp_mlp) is nvp.DeformNetwork            # the mandatory swap (model/barf_inn_llff.py:54-55)
    assert m.graph.warp_latent.weight.shape == (B, 128) and m.graph.global_rigid.weight.shape == (B, 12)
    # the reference's optimisers own our parameters: nerf (20 tensors + progress), warp_mlp (36) + warp_latent (1)
    assert sum(len(g["params"]) for g in m.optim.param_groups) == 21
    assert [len(g["params"]) for g in m.optim_pose.param_groups] == [36, 1]
    assert isinstance(m.sched, torch.optim.lr_scheduler.ExponentialLR) and isinstance(m.sched_pose, torch.optim.lr_scheduler.ExponentialLR)
    # state_dict inventory == the reference graph's own (minted in tests/golden/state_dicts.pt from the reference engine)
    inv = torch.load(os.path.join(ROOT, "tests", "golden", "state_dicts.pt"), weights_only=False)["barf_inn_llff"]
    mine = {k: list(v.shape) for k, v in m.graph.state_dict().items()}
    assert set(mine) == set(inv)
    for k, shape in inv.items():
        if not k.startswith(("warp_latent", "global_rigid")):         # per-image tables: the inventory was minted with its own B
            assert mine[k] == list(shape), k


@needs_reference
def test_checkpoint_round_trip_through_reference_util(patched):
    """``util.save_checkpoint`` of an engine that owns the B200 graph, then ``util.restore_checkpoint``
    (per-child ``load_state_dict`` + every optim* / sched* state) into a second, differently initialised engine."""
    mods, ref_shim, tmp_path = patched
    util = ref_shim.import_reference("util")
    B = 3
    opt = _opt(ref_shim, tmp_path, "cpu", B)
    torch.manual_seed(1)
    m1 = _build(mods, opt, B)
    m1.graph.nerf.progress.data.fill_(0.37)
    for p in m1.graph.warp_mlp.parameters():                           # leave the zero-init behind
        p.data.add_(0.01 * torch.randn_like(p))
    # two optimiser steps on synthetic gradients so that optim / sched carry state
    for _ in range(2):
        for o in (m1.optim, m1.optim_pose):
            for g in o.param_groups:
                for p in g["params"]:
                    if p.dim() > 0:                                    # (``progress`` sits in optim too, but never has a gradient)
                        p.grad = torch.randn_like(p) * 1e-3
            o.step()
        m1.sched.step(); m1.sched_pose.step()
    with contextlib.redirect_stdout(io.StringIO()):
        util.save_checkpoint(opt, m1, ep=None, it=2)
    torch.manual_seed(2)
    m2 = _build(mods, opt, B)
    assert not torch.equal(m2.graph.nerf.mlp_feat[0].weight, m1.graph.nerf.mlp_feat[0].weight)
    with contextlib.redirect_stdout(io.StringIO()):
        ep, it = util.restore_checkpoint(opt, m2, resume=True)
    assert (ep, it) == (None, 2)
    sd1, sd2 = m1.graph.state_dict(), m2.graph.state_dict()
    assert set(sd1) == set(sd2)
    for k in sd1:
        assert torch.equal(sd1[k], sd2[k]), k
    assert float(m2.graph.nerf.progress.detach()) == pytest.approx(0.37)
    st1, st2 = m1.optim.state_dict()["state"], m2.optim.state_dict()["state"]
    assert m2.sched.last_epoch == 2 and set(st1) == set(st2) and len(st1) == 20
    for i in st1:
        assert float(st1[i]["step"]) == float(st2[i]["step"]) == 2.0 and torch.equal(st1[i]["exp_avg"], st2[i]["exp_avg"])
    # the kernels' flat parameter view follows the restored values (load_state_dict copies in place)
    flat = m2.graph.nerf.flat_parameters()
    assert torch.equal(flat[:m2.graph.nerf.mlp_feat[0].weight.numel()], m2.graph.nerf.mlp_feat[0].weight.reshape(-1))


@needs_reference
def test_graph_refuses_the_reference_warp_network(patched):
    """Without the DeformNetwork swap the graph must fail loudly, not run the reference's eager warp."""
    mods, ref_shim, tmp_path = patched
    nvp_mod = sys.modules["model.nvp.nvp_ndr"]
    from neural_invertible_warp_b200.model import barf_inn_llff as ours
    opt = _opt(ref_shim, tmp_path, "cpu", 2)
    g = ours.Graph(opt)
    mine, nvp_mod.DeformNetwork = nvp_mod.DeformNetwork, nvp_mod._reference_DeformNetwork     # its __init__ names the class globally
    try:
        g.warp_mlp = nvp_mod.DeformNetwork(d_feature=128, d_in=3, d_out_1=1, d_out_2=3, n_blocks=3, d_hidden=128, n_layers=1,
                                           skip_in=[], multires=6, weight_norm=True, actfn="softplus")
    finally:
        nvp_mod.DeformNetwork = mine
    with pytest.raises(RuntimeError, match="install_dropin"):
        g._warp_network()


@needs_reference
@pytest.mark.gpu
def test_reference_train_iteration_on_the_cuda_path(patched):
    """The body of the reference's ``Model.train_iteration`` (model/barf_inn_llff.py:106-120 ->
    model/nerf_inn_llff.py:80-100) with a fake ``train_data``: forward / loss / backward run on the B200 kernels behind the
    reference's own optimisers, schedulers and ``progress`` update; the loss falls and every optimised tensor moves."""
    mods, ref_shim, tmp_path = patched
    from neural_invertible_warp_b200 import _lib, engine, synthetic as syn
    B = 4
    opt = _opt(ref_shim, tmp_path, "cuda:0", B)
    torch.manual_seed(0)
    m = _build(mods, opt, B)
    m.graph.warp_mlp.load_state_dict({k: v.to(opt.device) for k, v in syn.nvp_params(2).items()})
    var = engine.synthetic_var(opt, B, seed=3)
    EasyDict = sys.modules["easydict"].EasyDict
    m.train_data.all = EasyDict({k: v for k, v in var.items()})
    m.timer = EasyDict(start=0.0, it_mean=None)
    m.ep, m.it = 0, 0
    before = {k: v.detach().clone() for k, v in m.graph.named_parameters()}
    n0 = _lib.launch_count()
    losses = []
    with contextlib.redirect_stdout(io.StringIO()):
        for _ in range(8):
            loss = m.train_iteration(opt, m.train_data.all, range(opt.max_iter))
            m.sched.step()
            losses.append(float(loss.all))
    assert _lib.launch_count() - n0 >= 8 * 10                      # the CUDA library did the work
    assert all(l == l for l in losses) and min(losses[4:]) < losses[0]
    assert m.it == 8 and float(m.graph.nerf.progress) == pytest.approx(8 / opt.max_iter)
    moved = [k for k, v in m.graph.named_parameters() if not torch.equal(v, before[k])]
    assert "nerf.mlp_feat.0.weight" in moved and "warp_latent.weight" in moved and "warp_mlp.lin0_a_1.weight" in moved
