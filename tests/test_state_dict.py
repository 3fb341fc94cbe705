"""Checkpoint compatibility (SURVEY.md 8b "Ownership", 8 f4): ``graph.state_dict()`` of every drop-in Graph has
exactly the keys and shapes of the reference Graph as the reference's engine builds it, so ``util.save_checkpoint`` /
``restore_checkpoint`` files (reference util.py:124-163) interchange.  Inventory minted from the executed reference
(oracle/make_golden.py state_dicts).  Module construction only: no kernel runs, so this is part of the CPU suite."""
import os

import pytest
import torch

from neural_invertible_warp_b200 import config as cfgmod, engine, synthetic as syn

GOLD = os.path.join(os.path.dirname(__file__), "golden", "state_dicts.pt")


def _build(name):
    B = 3
    if name == "barf":
        opt = cfgmod.builtin_options("barf_llff", model="barf", barf_c2f=[0.1, 0.5], device="cpu", data=dict(image_size=[24, 32]))
        return engine.build_graph(opt, B)
    if name == "barf_inn_llff":
        opt = cfgmod.builtin_options("barf_inn_llff", barf_c2f=[0.1, 0.5], device="cpu", data=dict(image_size=[24, 32]))
        return engine.build_graph(opt, B)
    opt = cfgmod.builtin_options("barf_inn_dtu", barf_c2f=[0.1, 0.5], device="cpu", data=dict(image_size=[18, 24]),
                                 nerf=dict(fine_sampling=True))
    return engine.build_graph(opt, B, initial_poses_w2c=syn.dtu_poses(74, B))


@pytest.mark.parametrize("name", ["barf", "barf_inn_llff", "barf_inn_dtu"])
def test_state_dict_inventory_matches_reference(name):
    want = torch.load(GOLD)[name]
    sd = _build(name).state_dict()
    have = {k: list(v.shape) for k, v in sd.items()}
    assert set(have) == set(want), (sorted(set(have) - set(want)), sorted(set(want) - set(have)))
    for k in want:
        assert have[k] == want[k], (k, have[k], want[k])


def test_checkpoint_round_trip(tmp_path):
    """save -> restore through the reference's checkpoint layout (dict(graph=state_dict), util.py:147-156)."""
    a, b = _build("barf_inn_llff"), _build("barf_inn_llff")
    with torch.no_grad():
        for p in a.parameters():
            p.add_(torch.randn_like(p) * 0.01)
    path = tmp_path / "model.ckpt"
    torch.save(dict(epoch=None, iter=123, graph=a.state_dict()), path)
    ck = torch.load(path)
    b.load_state_dict(ck["graph"])
    for (k, x), (_, y) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(x, y), k
    # the flat parameter / gradient views the kernels use follow the loaded values
    assert torch.equal(torch.cat([p.reshape(-1) for p in b.nerf.mlp_parameters()]), b.nerf.flat_parameters().detach().cpu())
