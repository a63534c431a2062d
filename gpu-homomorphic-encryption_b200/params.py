"""Deterministic parameter selection (host logic): the prime chain of SURVEY 8d and BFV presets for the
BASELINE.json configs.  Intent of fhe::generate_rns_primes / find_ntt_prime / is_prime
(/root/reference/include/rns.cuh:139-149, placeholders in src/rns.cu:199-204) and of the parameter choice the
reference hard-codes in FHEContext::FHEContext (/root/reference/src/fhe.cu:7-24)."""
from __future__ import annotations

_MR_BASES = (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37)


def is_prime(n: int) -> bool:
    """deterministic Miller-Rabin, exact below 3.3e24."""
    if n < 2:
        return False
    for p in _MR_BASES:
        if n % p == 0:
            return n == p
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in _MR_BASES:
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def prime_chain(count: int, bits: int = 60, step_log2: int = 18) -> list[int]:
    """P[k] = k-th largest prime p < 2^bits with p = 1 (mod 2^step_log2); supports every N <= 2^(step_log2-1)."""
    out, step = [], 1 << step_log2
    p = (1 << bits) - step + 1
    while len(out) < count and p > step:
        if is_prime(p):
            out.append(p)
        p -= step
    if len(out) < count:
        raise ValueError("prime chain exhausted")
    return out


def find_ntt_prime(bit_length: int, ntt_size: int) -> int:
    """largest prime below 2^bit_length with q = 1 (mod 2*ntt_size)  (fhe::find_ntt_prime)."""
    step = 2 * ntt_size
    p = ((1 << bit_length) - 1) // step * step + 1
    while p > step:
        if p < (1 << bit_length) and is_prime(p):
            return p
        p -= step
    raise ValueError("no NTT prime of that size")


def generate_rns_primes(bit_length: int, num_primes: int, ntt_size: int) -> list[int]:
    """fhe::generate_rns_primes: distinct descending NTT-friendly primes of the given size."""
    out, step = [], 2 * ntt_size
    p = ((1 << bit_length) - 1) // step * step + 1
    while len(out) < num_primes and p > step:
        if is_prime(p):
            out.append(p)
        p -= step
    if len(out) < num_primes:
        raise ValueError("not enough primes")
    return out


def bfv_preset(name: str) -> dict:
    """BASELINE.json configs.  Q = first L chain primes, auxiliary basis = next R, special modulus = first K of those."""
    presets = {
        # config 2: N=4096, log_q=120 -> L=2 limbs of 60 bits, t=65537 (README params; tests/test_fhe.cu:173-178)
        "c2": dict(n=4096, L=2, R=3, K=1, dnum=2, t=65537, sigma=3.2, hamming_weight=64),
        # config 4: N=2^16, L=24, dnum=3 hybrid key switching (alpha=8, K=8), auxiliary basis 25 limbs
        "c4": dict(n=1 << 16, L=24, R=25, K=8, dnum=3, t=65537, sigma=3.2, hamming_weight=64),
        # reduced shapes for fast parity tests
        "small": dict(n=1024, L=4, R=5, K=2, dnum=2, t=65537, sigma=3.2, hamming_weight=64),
        # smallest ring the limb-sharded path supports (balanced two-pass NTT), limb counts that do not divide evenly by 2 / 4 ranks
        "mid": dict(n=8192, L=6, R=7, K=3, dnum=2, t=65537, sigma=3.2, hamming_weight=64),
    }
    p = dict(presets[name])
    p["primes"] = prime_chain(p["L"] + p["R"])
    return p
