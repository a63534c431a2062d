"""Build libfhe_b200.so (sm_100a only) in-tree with nvcc.  Used by __graft_entry__.build() and by hand:
    python gpu-homomorphic-encryption_b200/build.py [--force] [--verbose]
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libfhe_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
HOSTCXX = "/usr/bin/g++"

CU_SOURCES = ["ntt_kernels.cu", "ntt_bal.cu", "ntt_fused.cu", "elementwise.cu", "capi.cu", "lincomb.cu", "lincomb_mma.cu", "lincomb_tc.cu", "bfv.cu", "shard.cu", "peaks.cu"]
CPP_SOURCES = ["tables.cpp", "rns_consts.cpp", "wire.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-ccbin", HOSTCXX, "-Xcompiler", "-fPIC,-O2", "--expt-relaxed-constexpr"]


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".hpp", ".h"))] + \
           [os.path.join(HERE, "..", "include", "fhe_b200.h")]


def _stale(out, srcs):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(s) > t for s in srcs)


def _compile(src, verbose):
    obj = os.path.join(OBJ, os.path.basename(src) + ".o")
    if not _stale(obj, [src] + _deps()):
        return obj
    extra = os.environ.get("FHE_B200_NVCC_EXTRA", "").split()          # experiments: extra -D flags (force a rebuild when changing them)
    cmd = [NVCC] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-x", "cu", "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        with open(obj + ".ptxas.log", "w") as f:
            f.write(r.stderr)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = [os.path.join(CSRC, f) for f in CU_SOURCES + CPP_SOURCES if os.path.exists(os.path.join(CSRC, f))]
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), srcs))
    if force or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-ccbin", HOSTCXX, "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


def build_compat_test(force: bool = False) -> str:
    """tests/cpp/test_fhe_compat.bin: the reference's test program re-expressed against include/fhe/*.cuh."""
    root = os.path.dirname(HERE)
    src = os.path.join(root, "tests", "cpp", "test_fhe_compat.cu")
    out = os.path.join(root, "tests", "cpp", "test_fhe_compat.bin")
    hdrs = [os.path.join(root, "include", "fhe", f) for f in os.listdir(os.path.join(root, "include", "fhe"))]
    if force or _stale(out, [src, LIB] + hdrs):
        cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O2", "-std=c++17", "-ccbin", HOSTCXX,
               "-I" + os.path.join(root, "include"), src, "-o", out, "-L" + HERE, "-lfhe_b200",
               "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../../gpu-homomorphic-encryption_b200"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"compat test build failed:\n{r.stdout}\n{r.stderr}")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
    print(build_compat_test(force="--force" in sys.argv))
