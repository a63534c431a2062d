"""The arithmetic identity behind the FOLD form of the tcgen05 base conversion (gpu-homomorphic-encryption_b200/csrc/lincomb_tc.cu), restated with
Python integers and checked against the CPU oracle: splitting z into bytes commutes with the reduction modulo m_k,
    sum_i z_i M_ik  =  sum_(i,a) z_ia F_(i,a),k  (mod m_k),    F_(i,a),k = M_ik 2^(8a) mod m_k,
the byte-weighted partial sums p_b = sum z_ia byte_b(F) stay below 2^25, V = sum p_b 2^(8b) + I c_k stays below 2^83, and ONE fold at 2^60
(2^60 = delta mod m_k for m_k = 2^60 - delta) leaves a value below 2 m_k.  No GPU: this pins the algorithm, the -m gpu tests pin the kernel."""
import numpy as np
import pytest


def _fold_convert(x, src, dst, k):
    """one coefficient, all targets: returns (outputs, largest p_b, largest V, largest folded value / m)"""
    pre, th_hi, th_lo, M, c = k["pre"], k["th_hi"], k["th_lo"], k["M"], k["c"]
    S, T = len(src), len(dst)
    z = [int(x[i]) * int(pre[i]) % int(src[i]) for i in range(S)]
    # rounded integer I: per-term floors of the 128-bit fixed-point products
    f = sum(z[i] * int(th_hi[i]) for i in range(S)) + sum((z[i] * int(th_lo[i])) >> 64 for i in range(S)) + (1 << 63)
    I = f >> 64
    out, pmax, vmax, rmax = [], 0, 0, 0.0
    for t in range(T):
        m = int(dst[t]); delta = (1 << 60) - m
        assert 0 < delta < (1 << 32)
        p = [0] * 8
        for i in range(S):
            for a in range(8):
                za = (z[i] >> (8 * a)) & 0xFF
                F = int(M[i][t]) % m * pow(2, 8 * a, m) % m
                for b in range(8):
                    p[b] += za * ((F >> (8 * b)) & 0xFF)
        V = sum(p[b] << (8 * b) for b in range(8)) + I * int(c[t])
        r = (V & ((1 << 60) - 1)) + (V >> 60) * delta
        pmax = max(pmax, max(p)); vmax = max(vmax, V); rmax = max(rmax, r / m)
        assert r < 2 * m
        out.append(r - m if r >= m else r)
    return out, pmax, vmax, rmax


@pytest.mark.parametrize("S,T", [(24, 25), (25, 24), (8, 24), (4, 5)])
def test_fold_form_equals_the_oracle_conversion(oracle, chain, S, T):
    src, dst = chain[:S], chain[S:S + T]
    olc = oracle.LinComb.conv(src, dst)
    k = olc.constants()
    rng = np.random.default_rng(4200 + S)
    n = 6
    x = np.stack([rng.integers(0, q, n, dtype=np.uint64) for q in src])
    x[:, 0] = 0; x[:, 1] = [q - 1 for q in src]; x[:, 2] = 1
    want = olc.apply(x)
    pm = vm = 0; rm = 0.0
    for j in range(n):
        got, p, v, r = _fold_convert(x[:, j], src, dst, k)
        assert got == [int(w) for w in want[:, j]], j
        pm, vm, rm = max(pm, p), max(vm, v), max(rm, r)
    assert pm < (1 << 25) and vm < (1 << 83) and rm < 2.0
