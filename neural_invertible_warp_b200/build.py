"""In-tree build of the CUDA library (sm_100a only).

    python -m neural_invertible_warp_b200.build          # -> csrc/libniw_b200.so

Uses nvcc directly (no torch C++ extension): the library exposes a plain C ABI
(include/niw_b200.h) and is loaded with ctypes, so it has no libtorch dependency and compiles in
about a minute.  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libniw_b200.so")
SOURCES = ["api.cu", "raygen.cu", "nvp.cu", "sampler.cu", "composite.cu", "mlp_fp32.cu", "mlp_tc.cu", "mlp_tc_x3.cu", "mlp_tc_bwd.cu", "tc_selftest.cu", "adam.cu", "metrics.cu", "kabsch.cu", "p2p.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math=false"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


STAMP = LIB + ".srchash"


def source_hash():
    """sha256 over the sources, the header and the compile flags: what the built library is a function of.  (File
    times do not survive the snapshot to the GPU box, contents do.)"""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS + SOURCES).encode())
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")))
    deps.append(os.path.join(os.path.dirname(HERE), "include", "niw_b200.h"))
    for d in deps:
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def needs_build():
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != source_hash()


def build(force=False, verbose=False, extra_flags=(), out=None):
    """``extra_flags`` / ``out``: tuning builds (e.g. -DNIW_NSTAGE=6 into another .so, loaded via NIW_B200_LIB)."""
    if not force and not needs_build() and out is None:
        return LIB
    nvcc = _nvcc()
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")] + list(extra_flags)
    objs = []
    procs = []
    bdir = os.path.join(CSRC, "build" if out is None else "build_" + os.path.basename(out).replace(".so", ""))
    os.makedirs(bdir, exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(bdir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc] + flags + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        log, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, log))
        if verbose and log.strip():
            print(log)
    cmd = [nvcc, "-shared", "-o", out or LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    if out is None and not extra_flags:
        with open(STAMP, "w") as f:
            f.write(source_hash())
    return out or LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
