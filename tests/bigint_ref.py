"""Exact big-integer restatement of RNS-BFV used to pin oracle/orc_rns.c and oracle/orc_bfv.c
("a big-integer (Python int) BFV at N <= 1024").  Pure Python; small N only."""
from fractions import Fraction
from math import prod


def crt(residues, moduli):
    """non-negative representative in [0, prod(moduli))."""
    M = prod(moduli)
    x = 0
    for r, m in zip(residues, moduli):
        Mi = M // m
        x += int(r) * Mi * pow(Mi, -1, m)
    return x % M


def centred(x, M):
    """representative in [-M/2, M/2) with the oracle's tie rule: v = floor(x/M + 1/2)."""
    x %= M
    return x - M if 2 * x >= M else x


def round_half_up(num, den):
    """floor(num/den + 1/2) for integers, den > 0."""
    return (2 * num + den) // (2 * den)


def negacyclic_mul(a, b, n):
    out = [0] * n
    for i, ai in enumerate(a):
        if ai == 0:
            continue
        for j, bj in enumerate(b):
            k = i + j
            if k >= n:
                out[k - n] -= ai * bj
            else:
                out[k] += ai * bj
    return out


def poly_from_rns(limbs, moduli, centre=True):
    """limbs: [len(moduli)][n] array -> list of python ints."""
    M = prod(moduli)
    n = len(limbs[0])
    out = []
    for j in range(n):
        x = crt([limbs[i][j] for i in range(len(moduli))], moduli)
        out.append(centred(x, M) if centre else x)
    return out


def poly_to_rns(poly, moduli):
    return [[int(c) % m for c in poly] for m in moduli]
