"""Multi-GPU execution: one process per GPU, torch.distributed for the plumbing (NCCL on GPUs, gloo in CPU tests).

Two ways the path shards (reference prose only: docs/ARCHITECTURE.md:501-512, README.md:320):
  * batch sharding  -- every rank owns its own polynomials / ciphertexts; no data-path collective (bench.py);
  * limb sharding   -- rank g owns a contiguous block of RNS limbs of EVERY polynomial.  NTTs, the tensor product and the
                       key-switch inner product are limb-local; every base conversion needs all source limbs, so the
                       source limbs are all-gathered (uint64 blocks over NVLink) right before it and each rank computes
                       only its own target limbs.  BFV multiply+relinearize needs five such gathers.

The orchestration below is backend-agnostic: `GpuBackend` drives the C ABI; tests inject a CPU backend so that the
sharding / gather logic runs under gloo with world_size 2 in the GPU-less container.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist


def block_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """contiguous block partition; the first (total % world) ranks get one extra item.  -> (begin, count)"""
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, base + (1 if rank < rem else 0)


def allgather_blocks(local: torch.Tensor, counts: list[int], group=None) -> torch.Tensor:
    """local: [P][counts[rank]][N] -> [P][sum(counts)][N].  Blocks may differ by size; padded to the maximum for the
    collective (one all_gather_into_tensor call)."""
    world = len(counts)
    if world == 1:
        return local
    P, _, N = local.shape
    cmax = max(counts)
    pad = torch.zeros((P, cmax, N), dtype=local.dtype, device=local.device)
    pad[:, : local.shape[1]] = local
    flat = torch.empty((world * P, cmax, N), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(flat, pad.contiguous(), group=group)
    out = flat.view(world, P, cmax, N)
    return torch.cat([out[r, :, : counts[r]] for r in range(world)], dim=1).contiguous()


@dataclass
class LimbShard:
    """which limbs of a BFV context (Q = L primes, auxiliary basis = R primes, special = first K auxiliary) a rank owns"""
    L: int
    R: int
    K: int
    rank: int
    world: int

    def __post_init__(self):
        self.qb, self.qc = block_range(self.L, self.rank, self.world)      # Q limbs
        self.rb, self.rc = block_range(self.R, self.rank, self.world)      # auxiliary limbs (index inside the aux basis)
        sb, se = max(self.rb, 0), min(self.rb + self.rc, self.K)           # special limbs = aux[0..K)
        self.sb, self.sc = (sb, se - sb) if se > sb else (0, 0)
        self.q_counts = [block_range(self.L, r, self.world)[1] for r in range(self.world)]
        self.r_counts = [block_range(self.R, r, self.world)[1] for r in range(self.world)]
        self.s_counts = []
        for r in range(self.world):
            b, c = block_range(self.R, r, self.world)
            e = min(b + c, self.K)
            self.s_counts.append(max(0, e - b))


class GpuBackend:
    """compute primitives on int64 CUDA tensors through the C ABI (limb indices are plan indices: Q then auxiliary)"""

    def __init__(self, n, primes, device):
        import fhe_b200
        from fhe_b200 import engine
        self.E = engine
        self.lib = fhe_b200.load_library()
        self.n = n
        self.primes = [int(p) for p in primes]
        self.device = device
        self.plan = fhe_b200.Plan(n, self.primes, device=device)
        self._lc = {}

    def empty(self, *shape):
        return torch.empty(shape, dtype=torch.int64, device=torch.device("cuda", self.device))

    def ntt(self, x, limb_begin, inverse):
        if x.numel() == 0:
            return x
        return (self.plan.inverse if inverse else self.plan.forward)(x, limb_begin=limb_begin, limb_count=x.shape[-2])

    def conv(self, src_primes, dst_primes, x):
        if len(dst_primes) == 0:
            return self.empty(x.shape[0], 0, self.n)
        key = ("c", tuple(src_primes), tuple(dst_primes))
        if key not in self._lc:
            self._lc[key] = self.E.LinComb.conv(src_primes, dst_primes, self.device)
        return self._lc[key].apply(x.contiguous())

    def scale(self, qs, ps, t, targets, x, extra):
        if len(targets) == 0:
            return self.empty(x.shape[0], 0, self.n)
        key = ("s", tuple(qs), tuple(ps), t, tuple(targets))
        if key not in self._lc:
            self._lc[key] = self.E.LinComb.scale(qs, ps, t, targets, True, self.device)
        return self._lc[key].apply(x.contiguous(), extra.contiguous())

    def tensor(self, ext, limb_begin):
        lc = ext.shape[-2]
        out = self.empty(3, lc, self.n)
        if lc:
            self.E.check(self.lib.fhe_b200_bfv_tensor(self.plan.h, out.data_ptr(), ext.contiguous().data_ptr(), 1, limb_begin, lc,
                                                      self.E._stream()))
        return out

    def ks_inner(self, dig, key, limb_begin):
        dnum, lc = dig.shape[0], dig.shape[-2]
        out = self.empty(2, lc, self.n)
        if lc:
            self.E.check(self.lib.fhe_b200_bfv_ks_inner(self.plan.h, out.data_ptr(), dig.contiguous().data_ptr(),
                                                        key.contiguous().data_ptr(), dnum, 1, limb_begin, lc, self.E._stream()))
        return out

    def sub(self, a, b, limb_begin): return self.plan.sub(a.contiguous(), b.contiguous(), limb_begin=limb_begin, limb_count=a.shape[-2])
    def add(self, a, b, limb_begin): return self.plan.add(a.contiguous(), b.contiguous(), limb_begin=limb_begin, limb_count=a.shape[-2])

    def mul_scalar(self, a, scalars, limb_begin):
        return self.plan.mul_scalar(a.contiguous(), scalars, limb_begin=limb_begin, limb_count=a.shape[-2])


class LimbShardedBfv:
    """BFV multiply+relinearize with the limbs of every polynomial sharded over the ranks of `group`.

    Rank-local data: ciphertext [2][qc][N] (this rank's Q limbs, coefficient form); relin key restricted to the
    rank's Q limbs and special limbs.  `backend` supplies ntt / conv / scale / tensor / ks_inner / sub / add / mul_scalar.
    """

    def __init__(self, n, L, R, K, dnum, t, primes, backend, rank=None, world=None, group=None):
        self.n, self.L, self.R, self.K, self.dnum, self.t = n, L, R, K, dnum, int(t)
        self.alpha = L // dnum
        self.primes = [int(p) for p in primes]
        self.Q, self.P = self.primes[:L], self.primes[L:L + R]
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.sh = LimbShard(L, R, K, self.rank, self.world)
        self.be = backend
        sh = self.sh
        self.local_q = self.Q[sh.qb:sh.qb + sh.qc]
        self.local_r = self.P[sh.rb:sh.rb + sh.rc]
        self.local_s = self.P[sh.sb:sh.sb + sh.sc]
        Pprod = 1
        for p in self.P[:K]:
            Pprod *= p
        self.pinv = [pow(Pprod % q, -1, q) for q in self.local_q]

    # ---- helpers -----------------------------------------------------------------------------------------------
    def shard_ciphertext(self, ct_full: torch.Tensor) -> torch.Tensor:
        """[2][L][N] -> this rank's [2][qc][N]"""
        return ct_full[:, self.sh.qb:self.sh.qb + self.sh.qc].contiguous()

    def shard_relin_key(self, rlk_full: torch.Tensor):
        """[dnum][2][L+K][N] -> (keyQ [dnum][2][qc][N], keyP [dnum][2][sc][N])"""
        sh, L = self.sh, self.L
        return (rlk_full[:, :, sh.qb:sh.qb + sh.qc].contiguous(),
                rlk_full[:, :, L + sh.sb:L + sh.sb + sh.sc].contiguous())

    def gather_ciphertext(self, ct_local: torch.Tensor) -> torch.Tensor:
        return allgather_blocks(ct_local, self.sh.q_counts, self.group)

    # ---- the operation -----------------------------------------------------------------------------------------
    def multiply_relin(self, a_loc, b_loc, keyQ, keyP):
        be, sh, L, K = self.be, self.sh, self.L, self.K
        G = self.group
        # 1. all source limbs of the four input polynomials; exact extension Q -> (local) auxiliary limbs
        ext_q = torch.cat([a_loc, b_loc], dim=0).contiguous()                       # [4][qc][N]
        full_in = allgather_blocks(ext_q, sh.q_counts, G)                           # [4][L][N]       gather #1
        ext_r = be.conv(self.Q, self.local_r, full_in)                              # [4][rc][N]
        # 2.-4. NTT, tensor, INTT on local limbs
        ext_q = be.ntt(ext_q, sh.qb, False); ext_r = be.ntt(ext_r, L + sh.rb, False)
        d_q = be.ntt(be.tensor(ext_q, sh.qb), sh.qb, True)                          # [3][qc][N]
        d_r = be.ntt(be.tensor(ext_r, L + sh.rb), L + sh.rb, True)                  # [3][rc][N]
        # 5. round(t/Q .) into the local auxiliary limbs
        d_q_full = allgather_blocks(d_q, sh.q_counts, G)                            # [3][L][N]       gather #2
        s_r = be.scale(self.Q, self.P, self.t, self.local_r, d_q_full, d_r)         # [3][rc][N]
        # 6. exact conversion auxiliary -> local Q limbs
        s_r_full = allgather_blocks(s_r, sh.r_counts, G)                            # [3][R][N]       gather #3
        sc = be.conv(self.P, self.local_q, s_r_full)                                # [3][qc][N]
        # 7. relinearise d2: ModUp each digit into the local limbs of Q u P
        d2_full = allgather_blocks(sc[2:3].contiguous(), sh.q_counts, G)            # [1][L][N]       gather #4
        dig_q, dig_p = [], []
        for dg in range(self.dnum):
            g0, g1 = dg * self.alpha, (dg + 1) * self.alpha
            src = self.Q[g0:g1]
            x = d2_full[:, g0:g1].contiguous()
            # local Q limbs: own-group limbs are copied, the others converted
            tgt_idx = [i for i in range(sh.qb, sh.qb + sh.qc) if not (g0 <= i < g1)]
            conv_q = be.conv(src, [self.Q[i] for i in tgt_idx], x)
            dq = torch.empty_like(sc[2:3])
            k = 0
            for j, i in enumerate(range(sh.qb, sh.qb + sh.qc)):
                if g0 <= i < g1:
                    dq[:, j] = sc[2:3, j]
                else:
                    dq[:, j] = conv_q[:, k]; k += 1
            dig_q.append(dq)
            dig_p.append(be.conv(src, self.local_s, x))                             # [1][sc][N]
        dig_q = be.ntt(torch.cat(dig_q, dim=0).contiguous(), sh.qb, False)          # [dnum][qc][N]
        dig_p = be.ntt(torch.cat(dig_p, dim=0).contiguous(), L + sh.sb, False)      # [dnum][sc][N]
        acc_q = be.ntt(be.ks_inner(dig_q, keyQ, sh.qb), sh.qb, True)                # [2][qc][N]
        acc_p = be.ntt(be.ks_inner(dig_p, keyP, L + sh.sb), L + sh.sb, True)        # [2][sc][N]
        # ModDown: (acc_Q - [acc]_P -> Q) * P^-1, then add d0 / d1
        acc_p_full = allgather_blocks(acc_p, sh.s_counts, G)                        # [2][K][N]       gather #5
        down = be.conv(self.P[:K], self.local_q, acc_p_full)                        # [2][qc][N]
        out = be.mul_scalar(be.sub(acc_q, down, sh.qb), self.pinv, sh.qb)
        return be.add(out, sc[0:2].contiguous(), sh.qb)                              # [2][qc][N]

    def gather_bytes_per_op(self) -> int:
        """uint64 bytes received per rank per multiply (the NVLink traffic the roofline for this op is quoted on)"""
        n8 = self.n * 8
        total = (4 * self.L + 3 * self.L + 3 * self.R + self.L + 2 * self.K) * n8
        mine = (4 * self.sh.qc + 3 * self.sh.qc + 3 * self.sh.rc + self.sh.qc + 2 * self.sh.sc) * n8
        return total - mine
