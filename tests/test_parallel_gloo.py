"""World-size-2 gloo test (CPU) of the limb-sharded BFV multiply+relinearize orchestration
(gpu-homomorphic-encryption_b200/parallel.py): the sharding / all-gather logic runs for real over torch.distributed,
the per-limb arithmetic comes from the CPU oracle through a test-only backend, and the gathered result must equal the
unsharded oracle result word for word."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _t(a):  # uint64 numpy -> int64 torch
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64))


def _n(t):
    return t.numpy().view(np.uint64)


class OracleBackend:
    """test-only compute backend: same interface as parallel.GpuBackend, arithmetic by the CPU oracle"""

    def __init__(self, n, primes):
        import oracle
        self.o, self.n, self.primes = oracle, n, [int(p) for p in primes]

    def ntt(self, x, limb_begin, inverse):
        a = _n(x).copy()
        f = self.o.ntt_inverse if inverse else self.o.ntt_forward
        for p in range(a.shape[0]):
            for l in range(a.shape[1]):
                a[p, l] = f(a[p, l], self.primes[limb_begin + l])
        return _t(a)

    def conv(self, src, dst, x):
        if len(dst) == 0:
            return torch.empty((x.shape[0], 0, self.n), dtype=torch.int64)
        lc = self.o.LinComb.conv(src, dst)
        return _t(np.stack([lc.apply(_n(x[p])) for p in range(x.shape[0])]))

    def scale(self, qs, ps, t, targets, x, extra):
        if len(targets) == 0:
            return torch.empty((x.shape[0], 0, self.n), dtype=torch.int64)
        lc = self.o.LinComb.scale(qs, ps, t, targets, True)
        return _t(np.stack([lc.apply(_n(x[p]), extra=_n(extra[p])) for p in range(x.shape[0])]))

    def _mods(self, limb_begin, lc): return self.primes[limb_begin:limb_begin + lc]

    def tensor(self, ext, limb_begin):
        e = _n(ext); lc = e.shape[1]
        if lc == 0:
            return torch.empty((3, 0, self.n), dtype=torch.int64)
        m = self._mods(limb_begin, lc)
        mul = lambda a, b: self.o.poly_mul(a[None], b[None], m, self.n)[0]
        d1 = self.o.poly_add(mul(e[0], e[3])[None], mul(e[1], e[2])[None], m, self.n)[0]
        return _t(np.stack([mul(e[0], e[2]), d1, mul(e[1], e[3])]))

    def ks_inner(self, dig, key, limb_begin):
        d, k = _n(dig), _n(key); lc = d.shape[1]
        if lc == 0:
            return torch.empty((2, 0, self.n), dtype=torch.int64)
        m = self._mods(limb_begin, lc)
        out = np.zeros((2, lc, self.n), dtype=np.uint64)
        for c in range(2):
            for g in range(d.shape[0]):
                out[c] = self.o.poly_mac(out[c][None], d[g][None], k[g, c][None], m, self.n)[0]
        return _t(out)

    def sub(self, a, b, lb): return _t(self.o.poly_sub(_n(a), _n(b), self._mods(lb, a.shape[1]), self.n))
    def add(self, a, b, lb): return _t(self.o.poly_add(_n(a), _n(b), self._mods(lb, a.shape[1]), self.n))
    def mul_scalar(self, a, s, lb): return _t(self.o.poly_mul_scalar(_n(a), s, self._mods(lb, a.shape[1]), self.n))


def _worker(rank, world, port, cfg, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        import fhe_b200.parallel as par
        n, L, R, K, dnum, t = cfg
        primes = oracle.prime_chain(L + R)
        ctx = oracle.Bfv(n, L, R, K, dnum, t, primes, hw=8)
        _, sk = ctx.secret_keygen(1); pk = ctx.public_keygen(2, sk); rlk = ctx.relin_keygen(3, sk)
        rng = np.random.default_rng(4)
        ca = ctx.encrypt(5, rng.integers(0, t, n, dtype=np.uint64), pk)
        cb = ctx.encrypt(6, rng.integers(0, t, n, dtype=np.uint64), pk)
        expect = ctx.multiply_relin(ca, cb, rlk)
        sb = par.LimbShardedBfv(n, L, R, K, dnum, t, primes, OracleBackend(n, primes))
        kq, kp = sb.shard_relin_key(_t(rlk))
        out_loc = sb.multiply_relin(sb.shard_ciphertext(_t(ca)), sb.shard_ciphertext(_t(cb)), kq, kp)
        full = sb.gather_ciphertext(out_loc)
        ok = bool(np.array_equal(_n(full), expect))
        if rank == 0:
            out_q.put(("ok" if ok else "mismatch", sb.gather_bytes_per_op()))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


@pytest.mark.parametrize("cfg", [(64, 4, 5, 2, 2, 65537), (64, 3, 4, 3, 1, 65537), (32, 6, 7, 2, 3, 786433)])
def test_limb_sharded_multiply_world2_gloo(cfg):
    import oracle
    oracle.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, cfg, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    status, gathered = q.get(timeout=5)
    assert status == "ok"
    assert gathered > 0


def test_block_range_and_shard_bookkeeping():
    import fhe_b200.parallel as par
    for total in (1, 5, 24, 25):
        for world in (1, 2, 3, 4, 8):
            blocks = [par.block_range(total, r, world) for r in range(world)]
            assert sum(c for _, c in blocks) == total
            assert all(blocks[i][0] + blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
    for world in (1, 2, 4, 8):
        tot_s = 0
        for r in range(world):
            sh = par.LimbShard(24, 25, 8, r, world)
            assert 0 <= sh.sc <= sh.rc and (sh.sc == 0 or sh.sb >= sh.rb)
            tot_s += sh.sc
        assert tot_s == 8


def test_single_rank_equals_unsharded():
    """world = 1 (no process group needed): the orchestration itself reproduces the oracle's multiply_relin."""
    import oracle
    import fhe_b200.parallel as par
    n, L, R, K, dnum, t = 64, 4, 5, 2, 2, 65537
    primes = oracle.prime_chain(L + R)
    ctx = oracle.Bfv(n, L, R, K, dnum, t, primes, hw=8)
    _, sk = ctx.secret_keygen(1); pk = ctx.public_keygen(2, sk); rlk = ctx.relin_keygen(3, sk)
    ca = ctx.encrypt(5, ctx.encode([1, 2, 3]), pk); cb = ctx.encrypt(6, ctx.encode([4, 5]), pk)
    sb = par.LimbShardedBfv(n, L, R, K, dnum, t, primes, OracleBackend(n, primes), rank=0, world=1)
    kq, kp = sb.shard_relin_key(_t(rlk))
    out = sb.multiply_relin(_t(ca), _t(cb), kq, kp)
    assert np.array_equal(_n(out), ctx.multiply_relin(ca, cb, rlk))
