"""Static checks of the compiled kernels (no GPU: cuobjdump on the objects build() leaves in gpu-homomorphic-encryption_b200/_obj, nvcc on
tools/sass/balexp.cu): the base conversion really is tcgen05 / TMEM, the tile pass really fetches by bulk copy (TMA), and the four NTT
passes stay within the FMA-heavy-pipe cycles per butterfly that profiles/r02b_ntt_static_accounting.md records -- a regression in
instruction selection (ptxas moving adds and moves onto the binding pipe, a split IMAD.WIDE) shows here, before any GPU time is spent."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "gpu-homomorphic-encryption_b200", "_obj")

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or shutil.which("nvcc") is None, reason="CUDA toolkit not on PATH")


def _sass(path):
    return subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout


def test_base_conversion_is_tcgen05_and_tile_pass_fetches_by_tma():
    if not os.path.exists(os.path.join(OBJ, "lincomb_tc.cu.o")):
        pytest.skip("objects not built (run __graft_entry__.build() first)")
    tc = _sass(os.path.join(OBJ, "lincomb_tc.cu.o"))
    assert "UTCIMMA" in tc and "LDTM" in tc and "UTCBAR" in tc          # tcgen05.mma, tcgen05.ld, tcgen05.commit
    assert not re.search(r"\b[HI]MMA\.", tc)                            # no warp-level mma.sync in the tcgen05 kernel
    bal = _sass(os.path.join(OBJ, "ntt_bal.cu.o"))
    assert "UBLKCP" in bal and "SYNCS" in bal                            # cp.async.bulk + mbarrier


def test_ntt_passes_stay_within_the_recorded_pipe_cycles(tmp_path):
    cubin = str(tmp_path / "balexp.cubin")
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
                    "-DFHE_BAL_EXPERIMENT", "-cubin", "-o", cubin, os.path.join(ROOT, "tools", "sass", "balexp.cu")], check=True,
                   capture_output=True)
    # (kernel, range finder, FMA-heavy cycles per butterfly allowed): final round-2 code 28.6 / 31.4 / 30.4 / 29.7, two per cent of slack
    limits = [("bal_a_kernelILi8ELi16ELb1ELb0", "arange.py", 29.2), ("bal_a_kernelILi8ELi16ELb1ELb1", "arange.py", 32.1),
              ("bal_b_kernelILi8ELi16ELb1ELb0", "brange.py", 31.0), ("bal_b_kernelILi8ELi16ELb1ELb1", "brange.py", 30.3)]
    sd = os.path.join(ROOT, "tools", "sass")
    for kern, finder, limit in limits:
        rng = subprocess.run(["python", os.path.join(sd, finder), cubin, kern], capture_output=True, text=True, check=True).stdout.split()
        assert len(rng) == 2, (kern, rng)
        out = subprocess.run(["python", os.path.join(sd, "count.py"), cubin, kern, "64", rng[0], rng[1]], capture_output=True, text=True,
                             check=True).stdout
        m = re.search(r"heavy_cycles=\d+ \(([\d.]+)/bfly\)", out)
        assert m, out
        assert float(m.group(1)) <= limit, (kern, out)
