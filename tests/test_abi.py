"""The C-ABI library builds for sm_100a without a GPU, loads, and exports exactly what
include/niw_b200.h declares (no compute calls here)."""
import ctypes
import subprocess

from neural_invertible_warp_b200 import _lib, build, header


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    lib = ctypes.CDLL(path)
    declared = header.declared_symbols()
    assert len(declared) >= 17
    for name in declared:
        assert hasattr(lib, name), name
    assert set(_lib.SIGNATURES) == set(declared), set(_lib.SIGNATURES) ^ set(declared)
    lib.niw_abi_version.restype = ctypes.c_int
    assert lib.niw_abi_version() == _lib.ABI_VERSION
    lib.niw_error_string.restype = ctypes.c_char_p
    assert b"workspace" in lib.niw_error_string(-3)


def test_library_is_blackwell_native():
    """SASS of the fused MLP kernel holds tcgen05 MMAs (UTC*MMA), TMEM loads (LDTM) and bulk async
    copies (UBLKCP) -- not a recompiled mma.sync path."""
    sass = subprocess.run(["cuobjdump", "-sass", build.LIB], capture_output=True, text=True).stdout
    assert "LDTM" in sass
    assert "UBLKCP" in sass
    funcs = {}
    for chunk in sass.split("Function : ")[1:]:
        funcs[chunk.split("\n", 1)[0].strip()] = chunk
    def kernel(name):
        hits = [v for k, v in funcs.items() if name in k]
        assert len(hits) == 1, (name, [k for k in funcs if name in k])
        return hits[0]
    # forward and dX chain: CTA-pair tcgen05 MMAs (cta_group::2) and nothing from the warp-level mma.sync path
    for name in ("tc_fwd_kernel", "tc_dx_kernel"):
        k = kernel(name)
        assert "UTCHMMA.2CTA" in k, name
        assert "HMMA." not in k.replace("UTCHMMA", ""), name
    # weight-gradient pass: tcgen05 for the GEMMs; the thin bias / head products are warp-level m16n8k16 by design
    k = kernel("tc_dw_kernel")
    assert "UTCHMMA" in k and "HMMA.16816" in k


def test_ops_fail_loudly_without_cuda():
    import pytest
    import torch
    from neural_invertible_warp_b200 import functional as F
    with pytest.raises(RuntimeError):
        F.composite(torch.zeros(1, 3), torch.zeros(1, 4, 3), torch.zeros(1, 4), torch.zeros(1, 4))


def test_header_is_plain_c_and_links(tmp_path):
    """include/niw_b200.h is a C header (no C++ or torch types): a C99 host compiles against it with gcc and links
    the library -- the binding a non-Python maintainer would write."""
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "abi.c"
    src.write_text('#include "niw_b200.h"\n'
                   'int main(void) { return (niw_abi_version() == NIW_ABI_VERSION && niw_error_string(0) != 0) ? 0 : 1; }\n')
    exe = tmp_path / "abi"
    libdir = os.path.dirname(build.build())
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"),
                        str(src), "-L", libdir, "-lniw_b200", "-Wl,-rpath," + libdir, "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert subprocess.run([str(exe)]).returncode == 0
