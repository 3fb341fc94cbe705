"""Put the B200 render path behind the UNMODIFIED reference engine.

The reference dispatches by string: ``train.py:20-21`` imports ``model.<opt.model>`` and instantiates ``.Model(opt)``;
``model/base.py:35-37`` imports the same module again and instantiates ``.Graph(opt)`` (DTU: ``Graph(opt, pose_net)``,
``model/barf_inn_dtu.py:334-336``).  Behaviour is additionally keyed on the VALUE of ``opt.model`` (exact-name lists at
``model/nerf.py:172,213-214``, ``model/nerf_inn_llff.py:139-142``; substring tests at ``model/nerf_inn_dtu.py:125,167,217``),
so a true drop-in keeps the reference's model names.  ``install_dropin`` therefore patches the reference's own modules in
place, after importing them from ``reference_root``:

* ``model.<name>.Graph`` / ``.NeRF``  ->  ``neural_invertible_warp_b200.model.<name>.Graph`` / ``.NeRF``
  (the engine classes ``Model`` -- data, optimisers, logging, checkpoints -- stay the reference's),
* ``model.nvp.nvp_ndr.DeformNetwork``  ->  ``neural_invertible_warp_b200.nvp.DeformNetwork``.  MANDATORY: the engine builds
  the warp network itself (``model/barf_inn_llff.py:54-55``, ``model/pose_models/inn.py:23``) and our Graph evaluates it through
  the fused kernels, passing the point-index map the kernels need (see ``nvp.DeformNetwork.forward``),
* ``model.pose_models.inn.INNPoseParams`` (and the name ``model.barf_inn_dtu`` imported it under)  ->  ours,
* ``roma.rigid_points_registration`` is NOT needed any more (the fit is ``niw_kabsch``); the engine's own uses of it
  outside the hot path (pose alignment for evaluation) keep whatever ``roma`` is installed.

    import neural_invertible_warp_b200.dropin as dropin
    dropin.install_dropin("/path/to/neural_invertible_warp")       # then: import train; train.main()

or ``python -m neural_invertible_warp_b200.dropin /path/to/neural_invertible_warp --model=barf_inn_llff --yaml=barf_inn_llff ...``
which runs the reference's ``train.main()`` with the patched modules.
"""
import importlib
import os
import sys

MODELS = ("nerf", "barf", "nerf_inn_llff", "barf_inn_llff", "nerf_inn_dtu", "barf_inn_dtu")

_installed = {}


def install_dropin(reference_root=None, models=MODELS):
    """Patch the reference's model modules (imported from ``reference_root``, which is put on ``sys.path`` if given) so
    that its engine builds and drives the B200 Graphs.  Idempotent; returns {name: reference module}.  The original
    classes stay reachable as ``<module>._reference_Graph`` / ``_reference_NeRF`` / ``_reference_DeformNetwork``."""
    if reference_root is not None:
        reference_root = os.path.abspath(reference_root)
        if not os.path.isdir(os.path.join(reference_root, "model")):
            raise RuntimeError("install_dropin: %s does not look like the reference tree (no model/ directory)" % reference_root)
        if reference_root not in sys.path:
            sys.path.insert(0, reference_root)
    from . import nvp as b200_nvp
    from .model.pose_models import inn as b200_inn

    ref_nvp = importlib.import_module("model.nvp.nvp_ndr")
    if getattr(ref_nvp, "DeformNetwork", None) is not b200_nvp.DeformNetwork:
        ref_nvp._reference_DeformNetwork = ref_nvp.DeformNetwork
        ref_nvp.DeformNetwork = b200_nvp.DeformNetwork
    out = {}
    for name in models:
        ours = importlib.import_module("neural_invertible_warp_b200.model." + name)
        ref = importlib.import_module("model." + name)
        if getattr(ref, "Graph", None) is not ours.Graph:
            ref._reference_Graph, ref._reference_NeRF = ref.Graph, ref.NeRF
            ref.Graph, ref.NeRF = ours.Graph, ours.NeRF
        out[name] = ref
    if "barf_inn_dtu" in models or "nerf_inn_dtu" in models:
        ref_inn = importlib.import_module("model.pose_models.inn")
        if getattr(ref_inn, "INNPoseParams", None) is not b200_inn.INNPoseParams:
            ref_inn._reference_INNPoseParams = ref_inn.INNPoseParams
            ref_inn.INNPoseParams = b200_inn.INNPoseParams
        for name in ("barf_inn_dtu",):
            if name in out and hasattr(out[name], "INNPoseParams"):
                out[name].INNPoseParams = b200_inn.INNPoseParams       # ``from model.pose_models.inn import INNPoseParams``
    _installed.update(out)
    return out


def uninstall_dropin():
    """Restore the reference's classes (tests)."""
    for name, ref in list(_installed.items()):
        if hasattr(ref, "_reference_Graph"):
            ref.Graph, ref.NeRF = ref._reference_Graph, ref._reference_NeRF
            del ref._reference_Graph, ref._reference_NeRF
        _installed.pop(name)
    mods = sys.modules
    nv = mods.get("model.nvp.nvp_ndr")
    if nv is not None and hasattr(nv, "_reference_DeformNetwork"):
        nv.DeformNetwork = nv._reference_DeformNetwork
        del nv._reference_DeformNetwork
    inn = mods.get("model.pose_models.inn")
    if inn is not None and hasattr(inn, "_reference_INNPoseParams"):
        inn.INNPoseParams = inn._reference_INNPoseParams
        if "model.barf_inn_dtu" in mods:
            mods["model.barf_inn_dtu"].INNPoseParams = inn._reference_INNPoseParams
        del inn._reference_INNPoseParams


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0].startswith("-"):
        raise SystemExit("usage: python -m neural_invertible_warp_b200.dropin REFERENCE_ROOT [train.py arguments ...]")
    root = os.path.abspath(argv.pop(0))
    os.chdir(root)                         # the reference resolves options/*.yaml relative to the cwd (options.py:46,59-63)
    install_dropin(root)
    sys.argv = [os.path.join(root, "train.py")] + argv
    train = importlib.import_module("train")
    train.main()


if __name__ == "__main__":
    main()
