"""Mint the public-method signature inventory of the reference's hot-path classes (build container only).

    python oracle/make_signatures.py        # writes tests/golden/signatures.json

``inspect.signature`` of every method SURVEY.md 8b lists as part of the drop-in boundary, taken from the UNMODIFIED
reference imported from /root/reference.  tests/test_signatures.py checks the B200 classes against this file on any box
(the reference tree is not needed there).
"""
import inspect
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

GRAPH_METHODS = ("forward", "compute_loss", "get_pose", "render", "render_local", "render_by_slices", "render_by_slices_local",
                 "sample_depth", "sample_depth_from_pdf", "get_pose_init", "positional_encoding", "L1_loss", "MSE_loss")
NERF_METHODS = ("forward", "forward_samples", "composite", "positional_encoding", "define_network", "tensorflow_init_weights")
MODELS = ("nerf", "barf", "nerf_inn_llff", "barf_inn_llff", "nerf_inn_dtu", "barf_inn_dtu")


def sig(fn):
    out = []
    for p in inspect.signature(fn).parameters.values():
        d = None if p.default is inspect.Parameter.empty else repr(p.default)
        out.append(dict(name=p.name, kind=p.kind.name, has_default=p.default is not inspect.Parameter.empty, default=d))
    return out


def main():
    inv = {}
    for name in MODELS:
        mod = ref_shim.import_reference("model." + name)
        for cls_name, methods in (("Graph", GRAPH_METHODS), ("NeRF", NERF_METHODS)):
            cls = getattr(mod, cls_name)
            inv["model.%s.%s.__init__" % (name, cls_name)] = sig(cls.__init__)
            for m in methods:
                if hasattr(cls, m):
                    inv["model.%s.%s.%s" % (name, cls_name, m)] = sig(getattr(cls, m))
    nvp = ref_shim.import_reference("model.nvp.nvp_ndr")
    inv["model.nvp.nvp_ndr.DeformNetwork.__init__"] = sig(nvp.DeformNetwork.__init__)
    inv["model.nvp.nvp_ndr.DeformNetwork.forward"] = sig(nvp.DeformNetwork.forward)
    inn = ref_shim.import_reference("model.pose_models.inn")
    for m in ("__init__", "get_w2c_poses", "get_warped_rays_in_world", "forward_inn", "solve_for_global_transformation"):
        inv["model.pose_models.inn.INNPoseParams." + m] = sig(getattr(inn.INNPoseParams, m))
    cam = ref_shim.import_reference("camera")
    for f in ("get_center_and_ray", "get_unwarped_center_and_ray", "get_3D_points_from_depth", "convert_NDC", "cam2world",
              "world2cam", "img2cam", "to_hom"):
        inv["camera." + f] = sig(getattr(cam, f))
    for m in ("se3_to_SE3", "SE3_to_se3", "so3_to_SO3", "SO3_to_so3"):
        inv["camera.Lie." + m] = sig(getattr(cam.Lie, m))
    for m in ("__call__", "invert", "compose", "compose_pair"):
        inv["camera.Pose." + m] = sig(getattr(cam.Pose, m))
    path = os.path.join(ROOT, "tests", "golden", "signatures.json")
    with open(path, "w") as f:
        json.dump(inv, f, indent=1, sort_keys=True)
    print("wrote %s (%d callables)" % (path, len(inv)))


if __name__ == "__main__":
    main()
