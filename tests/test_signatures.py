"""The drop-in boundary's Python signatures against the inventory minted from the unmodified reference with
``inspect.signature`` (oracle/make_signatures.py -> tests/golden/signatures.json).  Runs without the reference tree.

Rule: every callable keeps the reference's parameters -- same names, order, kinds and defaults -- as a PREFIX of its own
signature; anything it adds must be optional and is pinned in EXTRA below, so a new keyword cannot appear unnoticed."""
import importlib
import inspect
import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "signatures.json")

# optional parameters the B200 classes add (all keyword-defaulted; the reference's callers never pass them)
EXTRA = {
    # the render pipeline asks for the compositing weights only when it resamples from them
    "NeRF.composite": ["want_prob"],
    # workspace with the weight streams already packed on a side stream
    "NeRF.forward_samples": ["prepacked"],
    # DTU-style per-scene depth range also accepted by the LLFF graphs (one shared implementation)
    "Graph.sample_depth": ["depth_range"],
    # position of the given points in the per-image point list (ray shards, shared centre row)
    "model.nvp.nvp_ndr.DeformNetwork.forward": ["index_map"],
    "model.pose_models.inn.INNPoseParams.forward_inn": ["_pts", "_index_map"],
    # only the requested pixels are generated: index tensor or a contiguous range
    "camera.get_center_and_ray": ["ray_idx", "idx_start", "num"],
    "camera.get_unwarped_center_and_ray": ["idx_start", "num"],
    # the reference's own Graph.forward passes ``ind=`` to get_pose in "render_train" mode (model/nerf.py:264) although
    # model/nerf.py:290 does not take it (a TypeError there); accepted and ignored here
    "model.nerf.Graph.get_pose": ["ind"],
}


def _ours(key):
    parts = key.split(".")
    if key.startswith("model.nvp.nvp_ndr.DeformNetwork"):
        from neural_invertible_warp_b200 import nvp
        return getattr(nvp.DeformNetwork, parts[-1])
    if key.startswith("model.pose_models.inn.INNPoseParams"):
        from neural_invertible_warp_b200.model.pose_models import inn
        return getattr(inn.INNPoseParams, parts[-1])
    if key.startswith("camera."):
        from neural_invertible_warp_b200 import camera
        obj = camera
        for p in parts[1:]:
            obj = getattr(obj, p)
        return obj
    mod = importlib.import_module("neural_invertible_warp_b200.model." + parts[1])
    return getattr(getattr(mod, parts[2]), parts[3])


def _extra_for(key):
    if key in EXTRA:
        return EXTRA[key]
    parts = key.split(".")
    return EXTRA.get(".".join(parts[-2:]), [])


def test_signatures_match_reference_inventory():
    with open(GOLDEN) as f:
        inv = json.load(f)
    assert len(inv) >= 130
    problems = []
    for key, ref in sorted(inv.items()):
        try:
            fn = _ours(key)
        except AttributeError as e:
            problems.append("%s: missing (%s)" % (key, e))
            continue
        mine = list(inspect.signature(fn).parameters.values())
        for i, r in enumerate(ref):
            if i >= len(mine):
                problems.append("%s: parameter %r missing" % (key, r["name"]))
                break
            m = mine[i]
            d = None if m.default is inspect.Parameter.empty else repr(m.default)
            if (m.name, m.kind.name, m.default is not inspect.Parameter.empty) != (r["name"], r["kind"], r["has_default"]) \
                    or (r["has_default"] and d != r["default"]):
                problems.append("%s: parameter %d is %s=%s (%s), reference %s=%s (%s)"
                                % (key, i, m.name, d, m.kind.name, r["name"], r["default"], r["kind"]))
        extra = [m for m in mine[len(ref):]]
        names = [m.name for m in extra]
        pinned = [n for n in _extra_for(key) if n not in [r["name"] for r in ref]]   # (DTU graphs take depth_range already)
        if names != pinned:
            problems.append("%s: extra parameters %s, pinned %s" % (key, names, pinned))
        for m in extra:
            if m.default is inspect.Parameter.empty and m.kind.name not in ("VAR_POSITIONAL", "VAR_KEYWORD"):
                problems.append("%s: extra parameter %s has no default" % (key, m.name))
    assert not problems, "\n".join(problems)
