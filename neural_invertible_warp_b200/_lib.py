"""ctypes binding of csrc/libniw_b200.so (the C ABI declared in include/niw_b200.h).

There is deliberately no CPU or PyTorch fallback: if the library cannot be loaded the import of
any op raises, and every op refuses non-CUDA tensors.
"""
import ctypes
import os

from . import build as _build

_c = ctypes
_P = _c.c_void_p
_LIB = None

NIW_PREC_FP32 = 0
NIW_PREC_BF16 = 1
NIW_PREC_BF16X3 = 2
NIW_NERF_PREPACKED = 2
NIW_NERF_PARAMS = 530052
NIW_NVP_BLOCK_FLOATS = ((128 * 27 + 128 + 1 + 128 * 13 + 3 * 128 + 3) + 3) // 4 * 4   # 5636, include/niw_b200.h
ABI_VERSION = 6

# name -> (restype, argtypes); mirrors include/niw_b200.h one to one
SIGNATURES = {
    "niw_abi_version": (_c.c_int, []),
    "niw_error_string": (_c.c_char_p, [_c.c_int]),
    "niw_launch_count": (_c.c_ulonglong, []),
    "niw_raygen_pose_fwd": (_c.c_int, [_P, _P, _P, _c.c_int64, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _P, _P, _P]),
    "niw_raygen_pose_bwd": (_c.c_int, [_P, _P, _P, _c.c_int64, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _P, _P, _P, _P]),
    "niw_raygen_unwarped": (_c.c_int, [_P, _P, _P, _c.c_int64, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _P, _P]),
    "niw_rays_from_warp_fwd": (_c.c_int, [_P, _c.c_int, _c.c_int, _c.c_int, _P, _P, _P]),
    "niw_rays_from_warp_bwd": (_c.c_int, [_P, _P, _c.c_int, _c.c_int, _c.c_int, _P, _P]),
    "niw_nvp_pack_fwd": (_c.c_int, [_P, _P, _c.c_int, _P, _P, _P, _P]),
    "niw_nvp_pack_bwd": (_c.c_int, [_P, _P, _P, _P, _P, _P, _c.c_int, _P, _P]),
    "niw_nvp_warp_fwd": (_c.c_int, [_P, _P, _P, _c.c_float] + [_c.c_int] * 5 + [_P, _P]),
    "niw_nvp_warp_bwd": (_c.c_int, [_P, _P, _P, _c.c_float] + [_c.c_int] * 5 + [_P, _P, _P, _c.c_int, _P]),
    "niw_nvp_rays_fwd": (_c.c_int, [_P, _P, _P, _P, _P, _c.c_int64, _c.c_float] + [_c.c_int] * 7 + [_P, _P, _P, _P, _P]),
    "niw_sample_pixels": (_c.c_int, [_c.c_int64, _c.c_int, _c.c_uint64, _P, _P, _P]),
    "niw_sample_stratified": (_c.c_int, [_P, _c.c_int64, _c.c_int, _c.c_float, _c.c_float, _c.c_int, _P, _P]),
    "niw_sample_stratified_dev": (_c.c_int, [_P, _c.c_int64, _c.c_int, _P, _c.c_int, _P, _P]),
    "niw_sample_stratified_rng": (_c.c_int, [_c.c_int64, _c.c_int, _c.c_float, _c.c_float, _P, _c.c_int, _c.c_uint64, _P, _P, _P]),
    "niw_sample_pdf_merge": (_c.c_int, [_P, _P, _P, _P, _c.c_int64, _c.c_int, _c.c_int, _P, _P, _P, _P]),
    "niw_composite_fwd": (_c.c_int, [_P, _P, _P, _P, _c.c_int64, _c.c_int, _c.c_float, _P, _P, _P, _P, _P, _P]),
    "niw_composite_bwd": (_c.c_int, [_P, _P, _P, _P, _P, _P, _c.c_int64, _c.c_int, _c.c_float, _P, _P, _P, _P, _P, _P, _P]),
    "niw_composite_mse_scratch_floats": (_c.c_int, []),
    "niw_composite_fwd_mse": (_c.c_int, [_P, _P, _P, _P, _c.c_int64, _c.c_int, _c.c_float, _P, _P, _P, _P, _P,
                                         _P, _P, _c.c_int64, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _P, _P, _P, _P]),
    "niw_composite_bwd_mse": (_c.c_int, [_P, _P, _P, _P, _P, _P, _c.c_int64, _c.c_int, _c.c_float, _P, _P, _P, _P, _P,
                                         _P, _P, _P, _P]),
    "niw_nerf_workspace_bytes": (_c.c_size_t, [_c.c_int64, _c.c_int, _c.c_int, _c.c_int]),
    "niw_nerf_fwd": (_c.c_int, [_P, _P, _P, _P, _c.c_int64, _c.c_int, _P, _c.c_float, _c.c_float, _c.c_int, _c.c_int, _P,
                                _c.c_size_t, _P, _P, _P]),
    "niw_nerf_pack": (_c.c_int, [_P, _P, _c.c_float, _c.c_float, _c.c_int, _c.c_int, _c.c_int64, _c.c_int, _P, _c.c_size_t, _P]),
    "niw_nerf_bwd": (_c.c_int, [_P, _P, _P, _P, _c.c_int64, _c.c_int, _c.c_int, _P, _c.c_size_t, _P, _P, _P, _P, _P, _P]),
    "niw_nerf_bwd_dx": (_c.c_int, [_P, _P, _P, _P, _c.c_int64, _c.c_int, _c.c_int, _P, _c.c_size_t, _P, _P, _P, _P, _P, _P]),
    "niw_nerf_bwd_dw": (_c.c_int, [_c.c_int64, _c.c_int, _c.c_int, _P, _c.c_size_t, _P, _c.c_int, _P]),
    "niw_mse_gather": (_c.c_int, [_P, _P, _P, _c.c_int64, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_float, _P, _P, _P]),
    "niw_mse_gather_needs_zero": (_c.c_int, [_c.c_int, _c.c_int]),
    "niw_image_metrics": (_c.c_int, [_P, _P, _c.c_int, _c.c_int, _c.c_int, _P, _P]),
    "niw_depth_metrics": (_c.c_int, [_P, _P, _P, _c.c_int64, _c.c_float, _P, _P]),
    "niw_p2p_alloc": (_c.c_int, [_c.c_size_t, _c.POINTER(_c.c_void_p), _P]),
    "niw_p2p_open": (_c.c_int, [_P, _c.POINTER(_c.c_void_p)]),
    "niw_p2p_close": (_c.c_int, [_P]),
    "niw_p2p_free": (_c.c_int, [_P]),
    "niw_allreduce_p2p": (_c.c_int, [_P, _c.c_int64, _c.POINTER(_c.c_void_p), _c.c_int, _c.c_int, _c.c_int64, _P]),
    "niw_p2p_error": (_c.c_int, [_P, _c.POINTER(_c.c_uint)]),
    "niw_kabsch": (_c.c_int, [_P, _P, _c.c_int, _c.c_int, _P, _P, _P]),
    "niw_kabsch_stats": (_c.c_int, [_P, _P, _c.c_int, _c.c_int, _P, _P]),
    "niw_kabsch_solve": (_c.c_int, [_P, _c.c_int, _P, _P, _P]),
    "niw_adam_step": (_c.c_int, [_P, _P, _P, _P, _c.c_int64, _c.c_double, _c.c_double] + [_c.c_float] * 5 + [_c.c_int64, _c.c_float, _P, _P, _P, _P]),
    "niw_tc_probe": (_c.c_int, [_c.c_int, _c.c_int, _P, _P]),
    "niw_tc_selftest": (_c.c_int, [_P, _P, _c.c_int, _c.c_int, _c.c_int, _P, _P]),
}


def library_path():
    """csrc/libniw_b200.so; NIW_B200_LIB selects another build of the same sources (kernel-tuning experiments)."""
    return os.environ.get("NIW_B200_LIB") or _build.LIB


def load(build_if_missing=True):
    """Load (building first if the sources are newer and nvcc is available) and type the library."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if build_if_missing and not os.environ.get("NIW_B200_LIB"):
        try:
            if _build.needs_build():
                _build.build()
        except Exception as e:  # no nvcc on this box: fall through to the prebuilt file
            if not os.path.exists(path):
                raise RuntimeError("libniw_b200.so is missing and could not be built: %s" % e)
            # a prebuilt library older than its sources that cannot be rebuilt: usable only if the caller says so
            if not os.environ.get("NIW_B200_ALLOW_STALE"):
                raise RuntimeError("libniw_b200.so is older than csrc/*.cu / include/niw_b200.h and the rebuild failed (%s); "
                                   "fix the build, or set NIW_B200_ALLOW_STALE=1 to load the stale library anyway"
                                   % str(e).splitlines()[0])
            import warnings
            warnings.warn("niw_b200: loading a STALE libniw_b200.so (sources are newer, rebuild failed: %s)"
                          % str(e).splitlines()[0])
    if not os.path.exists(path):
        raise RuntimeError("libniw_b200.so not found at %s -- run `python -m neural_invertible_warp_b200.build`; "
                           "there is no CPU fallback" % path)
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError = ABI mismatch, loud on purpose
        fn.restype = res
        fn.argtypes = args
    if lib.niw_abi_version() != ABI_VERSION:
        raise RuntimeError("libniw_b200.so ABI version %d != %d" % (lib.niw_abi_version(), ABI_VERSION))
    _LIB = lib
    return lib


def launch_count():
    """Kernels launched by the library so far (bench.py: ``gpu_launches``)."""
    return int(load().niw_launch_count())


def check(code):
    if code != 0:
        raise RuntimeError("niw_b200: %s (code %d)" % (load().niw_error_string(code).decode(), code))
