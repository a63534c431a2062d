"""Host-side mirror of the reference's operator interface over the C ABI (harness; torch is plumbing).

Mirrors, with the same method names and argument meaning (device buffers in, device buffers out, in-place
transforms, asynchronous on a stream):
    fhe::NTTEngine       /root/reference/include/ntt.cuh:72-103
    fhe::RNS_NTTEngine   /root/reference/include/ntt.cuh:106-137
    fhe::PolynomialOps   /root/reference/include/polynomial.cuh:21-59
    fhe::RNSContext      /root/reference/include/rns.cuh:27-65
    fhe::FHEContext      /root/reference/include/fhe.cuh:78-148   (BfvContext)
Device arrays are torch int64 CUDA tensors used as containers for uint64 words, limb-major [batch][limbs][N].
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import check, load_library, u64p


def _np_u64(x) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(x, dtype=np.uint64))


def _ptr(t: torch.Tensor) -> int:
    assert t.is_cuda and t.dtype == torch.int64 and t.is_contiguous(), "expected a contiguous int64 CUDA tensor"
    return t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def to_device(a: np.ndarray, device=None) -> torch.Tensor:
    """uint64 numpy array -> int64 CUDA tensor (same bits)."""
    a = _np_u64(a)
    return torch.from_numpy(a.view(np.int64)).to(device or "cuda", non_blocking=False)


def to_host(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy().view(np.uint64)


def pinned_empty(shape) -> np.ndarray:
    """page-locked uint64 host buffer (numpy view of a pinned torch tensor, kept alive by the view's base)."""
    t = torch.empty(shape, dtype=torch.int64, pin_memory=True)
    a = t.numpy().view(np.uint64)
    return a


def gaussian_cdt(sigma: float) -> np.ndarray:
    buf = np.zeros(128, dtype=np.uint64)
    ln = load_library().fhe_b200_gaussian_cdt(float(sigma), buf.ctypes.data_as(u64p), 128)
    if ln < 0:
        check(ln)
    return buf[:ln].copy()


WIRE_KINDS = {"ciphertext": 1, "public_key": 2, "secret_key": 3, "switch_key": 4, "plaintext": 5}


def wire_pack(kind: str, words, moduli, ntt_form: bool = False, galois_elt: int = 0) -> bytes:
    """Serialise uint64 words [polys][limbs][n] (host array, or a device tensor which is copied back first) with the 64-byte
    header csrc/wire.cpp documents.  The reference has no serialisation."""
    if isinstance(words, torch.Tensor):
        words = to_host(words)
    w = _np_u64(words)
    m = _np_u64(moduli)
    n, limbs = w.shape[-1], len(m)
    assert w.ndim >= 2 and w.shape[-2] == limbs, "words must be [..., limbs, n]"
    polys = w.size // (n * limbs)
    lib = load_library()
    out = np.empty(lib.fhe_b200_wire_size(n, limbs, polys), dtype=np.uint8)
    check(lib.fhe_b200_wire_pack(WIRE_KINDS[kind], n, limbs, polys, int(bool(ntt_form)), int(galois_elt),
                                 m.ctypes.data_as(u64p), w.ctypes.data_as(u64p), out.ctypes.data))
    return out.tobytes()


def wire_unpack(blob: bytes, moduli=None):
    """-> (header dict, uint64 array [polys][limbs][n]); raises FheB200Error on a bad magic / length / checksum / modulus chain."""
    lib = load_library()
    buf = np.frombuffer(blob, dtype=np.uint8)
    m = _np_u64(moduli) if moduli is not None else None
    mp = m.ctypes.data_as(u64p) if m is not None else None
    f = [C.c_uint32() for _ in range(4)]
    nf, ge = C.c_int(), C.c_uint32()
    refs = [C.byref(x) for x in f] + [C.byref(nf), C.byref(ge)]
    nm = len(m) if m is not None else 0
    check(lib.fhe_b200_wire_unpack(buf.ctypes.data, len(blob), mp, nm, *refs, None))
    kind, n, limbs, polys = (x.value for x in f)
    out = np.empty((polys, limbs, n), dtype=np.uint64)
    check(lib.fhe_b200_wire_unpack(buf.ctypes.data, len(blob), mp, nm, *refs, out.ctypes.data_as(u64p)))
    names = {v: k for k, v in WIRE_KINDS.items()}
    return {"kind": names[kind], "n": n, "limbs": limbs, "polys": polys, "ntt_form": bool(nf.value), "galois_elt": ge.value}, out


class Plan:
    """fhe_b200_plan: ring degree N, RNS moduli, device twiddle tables."""

    def __init__(self, n: int, moduli, device: int = 0):
        self.lib = load_library()
        self.n = int(n)
        self.moduli = [int(m) for m in moduli]
        self.device = device
        m = _np_u64(self.moduli)
        h = C.c_void_p()
        check(self.lib.fhe_b200_plan_create(self.n, m.ctypes.data_as(u64p), len(self.moduli), device, C.byref(h)))
        self.h = h
        self.torch_device = torch.device("cuda", device)

    def close(self):
        if getattr(self, "h", None):
            self.lib.fhe_b200_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- shape helpers: tensors are [batch, limb_count, N] (batch may be omitted)
    def _dims(self, t: torch.Tensor, limb_count):
        assert t.shape[-1] == self.n, f"last dimension must be N={self.n}"
        total = t.numel() // self.n
        lc = limb_count if limb_count is not None else (t.shape[-2] if t.dim() >= 2 else 1)
        assert total % lc == 0
        return total // lc, lc

    def tables(self, limb: int):
        out = [np.zeros(self.n, dtype=np.uint64) for _ in range(4)]
        check(self.lib.fhe_b200_plan_tables(self.h, limb, *[o.ctypes.data_as(u64p) for o in out]))
        return out

    def forward(self, x: torch.Tensor, out: torch.Tensor | None = None, limb_begin=0, limb_count=None):
        out = x if out is None else out
        b, lc = self._dims(x, limb_count)
        if x.numel() == 0:
            return out
        check(self.lib.fhe_b200_ntt_forward(self.h, _ptr(out), _ptr(x), b, limb_begin, lc, _stream()))
        return out

    def inverse(self, x: torch.Tensor, out: torch.Tensor | None = None, limb_begin=0, limb_count=None):
        out = x if out is None else out
        b, lc = self._dims(x, limb_count)
        if x.numel() == 0:
            return out
        check(self.lib.fhe_b200_ntt_inverse(self.h, _ptr(out), _ptr(x), b, limb_begin, lc, _stream()))
        return out

    def negacyclic_mul(self, a, b, out=None, limb_begin=0, limb_count=None):
        out = torch.empty_like(a) if out is None else out
        bt, lc = self._dims(a, limb_count)
        check(self.lib.fhe_b200_negacyclic_mul(self.h, _ptr(out), _ptr(a), _ptr(b), bt, limb_begin, lc, _stream()))
        return out

    def bitrev_permute(self, x):
        out = torch.empty_like(x)
        check(self.lib.fhe_b200_bitrev_permute(self.h, _ptr(out), _ptr(x), x.numel() // self.n, _stream()))
        return out

    def _ew(self, fn, a, b, out, limb_begin, limb_count):
        out = torch.empty_like(a) if out is None else out
        bt, lc = self._dims(a, limb_count)
        check(getattr(self.lib, fn)(self.h, _ptr(out), _ptr(a), _ptr(b), bt, limb_begin, lc, _stream()))
        return out

    def add(self, a, b, out=None, limb_begin=0, limb_count=None): return self._ew("fhe_b200_poly_add", a, b, out, limb_begin, limb_count)
    def sub(self, a, b, out=None, limb_begin=0, limb_count=None): return self._ew("fhe_b200_poly_sub", a, b, out, limb_begin, limb_count)
    def mul(self, a, b, out=None, limb_begin=0, limb_count=None): return self._ew("fhe_b200_poly_mul", a, b, out, limb_begin, limb_count)

    def mac(self, acc, a, b, out=None, limb_begin=0, limb_count=None):
        out = torch.empty_like(a) if out is None else out
        bt, lc = self._dims(a, limb_count)
        check(self.lib.fhe_b200_poly_mac(self.h, _ptr(out), _ptr(acc), _ptr(a), _ptr(b), bt, limb_begin, lc, _stream()))
        return out

    def _scalar(self, fn, a, scalars, out, limb_begin, limb_count):
        out = torch.empty_like(a) if out is None else out
        bt, lc = self._dims(a, limb_count)
        s = _np_u64([int(v) % self.moduli[limb_begin + i] for i, v in enumerate(scalars)])
        assert len(s) == lc
        check(getattr(self.lib, fn)(self.h, _ptr(out), _ptr(a), s.ctypes.data_as(u64p), bt, limb_begin, lc, _stream()))
        return out

    def mul_scalar(self, a, scalars, out=None, limb_begin=0, limb_count=None):
        return self._scalar("fhe_b200_poly_mul_scalar", a, scalars, out, limb_begin, limb_count)

    def add_scalar(self, a, scalars, out=None, limb_begin=0, limb_count=None):
        return self._scalar("fhe_b200_poly_add_scalar", a, scalars, out, limb_begin, limb_count)

    def negate(self, a, out=None, limb_begin=0, limb_count=None):
        out = torch.empty_like(a) if out is None else out
        bt, lc = self._dims(a, limb_count)
        check(self.lib.fhe_b200_poly_negate(self.h, _ptr(out), _ptr(a), bt, limb_begin, lc, _stream()))
        return out

    def modswitch_drop_last(self, x, limb_begin=0, limb_count=None):
        bt, lc = self._dims(x, limb_count)
        out = torch.empty(x.shape[:-2] + (lc - 1, self.n), dtype=torch.int64, device=x.device)
        check(self.lib.fhe_b200_modswitch_drop_last(self.h, _ptr(out), _ptr(x), bt, limb_begin, lc, _stream()))
        return out

    def ntt_host(self, h: np.ndarray, direction: int, limb_begin=0, limb_count=None):
        """host-buffer entry point (e2e path): h is uint64 [batch][limb_count][N], transformed in place."""
        assert h.dtype == np.uint64 and h.flags.c_contiguous
        lc = limb_count if limb_count is not None else h.shape[-2]
        bt = h.size // (lc * self.n)
        check(self.lib.fhe_b200_ntt_host(self.h, h.ctypes.data_as(C.c_void_p), bt, limb_begin, lc, direction))
        return h


# ---- fhe::uint256_t edge --------------------------------------------------------------------------------------
def unpack_u256(d_u256: torch.Tensor) -> torch.Tensor:
    """[count][4] int64 (uint256_t words) -> [count] int64 (limbs[0])."""
    count = d_u256.numel() // 4
    out = torch.empty(count, dtype=torch.int64, device=d_u256.device)
    check(load_library().fhe_b200_unpack_u256(_ptr(out), _ptr(d_u256), count, _stream()))
    return out


def pack_u256(d: torch.Tensor) -> torch.Tensor:
    out = torch.empty(d.numel(), 4, dtype=torch.int64, device=d.device)
    check(load_library().fhe_b200_pack_u256(_ptr(out), _ptr(d), d.numel(), _stream()))
    return out


class NTTEngine:
    """fhe::NTTEngine(polynomial_degree, modulus): single-modulus transforms on device arrays of N uint64."""

    def __init__(self, polynomial_degree: int, modulus: int, device: int = 0):
        self.plan = Plan(polynomial_degree, [modulus], device)
        self.n = polynomial_degree

    def forward(self, d_data: torch.Tensor): return self.plan.forward(d_data.view(-1, 1, self.n)).view_as(d_data)
    def inverse(self, d_data: torch.Tensor): return self.plan.inverse(d_data.view(-1, 1, self.n)).view_as(d_data)

    def multiply(self, d_result, d_a, d_b):
        self.plan.negacyclic_mul(d_a.view(-1, 1, self.n), d_b.view(-1, 1, self.n), out=d_result.view(-1, 1, self.n))
        return d_result

    def forward_batch(self, d_data, batch_size):
        assert d_data.numel() == batch_size * self.n
        return self.forward(d_data)

    def inverse_batch(self, d_data, batch_size):
        assert d_data.numel() == batch_size * self.n
        return self.inverse(d_data)


class RNS_NTTEngine:
    """fhe::RNS_NTTEngine(polynomial_degree, rns_moduli, num_primes): limb-major d_rns_data[i*N + j]."""

    def __init__(self, polynomial_degree: int, rns_moduli, device: int = 0):
        self.plan = Plan(polynomial_degree, rns_moduli, device)
        self.n = polynomial_degree
        self.num_primes = len(self.plan.moduli)

    def _v(self, t): return t.view(-1, self.num_primes, self.n)
    def forward_rns(self, d_rns_data): self.plan.forward(self._v(d_rns_data)); return d_rns_data
    def inverse_rns(self, d_rns_data): self.plan.inverse(self._v(d_rns_data)); return d_rns_data

    def multiply_rns(self, d_result, d_a, d_b):
        self.plan.negacyclic_mul(self._v(d_a), self._v(d_b), out=self._v(d_result))
        return d_result

    def to_rns(self, d_u256: torch.Tensor) -> torch.Tensor:
        """d_u256: [N][4] words of 256-bit coefficients -> [num_primes][N] residues."""
        count = d_u256.numel() // 4
        out = torch.empty(self.num_primes, count, dtype=torch.int64, device=d_u256.device)
        check(self.plan.lib.fhe_b200_to_rns_u256(self.plan.h, _ptr(out), _ptr(d_u256), count, 0, self.num_primes, _stream()))
        return out


    def from_rns(self, d_rns: torch.Tensor) -> torch.Tensor:
        """[num_primes][N] residues -> [N][4] words of the CRT value in [0, Q)  (Q < 2^256)."""
        count = d_rns.numel() // self.num_primes
        out = torch.empty(count, 4, dtype=torch.int64, device=d_rns.device)
        check(self.plan.lib.fhe_b200_from_rns_u256(self.plan.h, _ptr(out), _ptr(d_rns), count, 0, self.num_primes, _stream()))
        return out


class PolynomialOps:
    """fhe::PolynomialOps over one plan (any number of limbs)."""

    def __init__(self, plan: Plan):
        self.plan = plan

    def add(self, result, a, b): return self.plan.add(a, b, out=result)
    def sub(self, result, a, b): return self.plan.sub(a, b, out=result)
    def mul_ntt(self, result, a, b): return self.plan.negacyclic_mul(a, b, out=result)
    mul = mul_ntt
    mul_negacyclic = mul_ntt
    def mul_scalar(self, result, a, scalar): return self.plan.mul_scalar(a, [scalar] * a.shape[-2], out=result)
    def add_scalar(self, result, a, scalar): return self.plan.add_scalar(a, [scalar] * a.shape[-2], out=result)


class LinComb:
    """fhe_b200_lincomb: exact RNS base conversion / scale-and-round."""

    def __init__(self, handle, S, T):
        self.h, self.S, self.T = handle, S, T
        self.lib = load_library()

    @classmethod
    def conv(cls, src, dst, device=0):
        s, d = _np_u64(src), _np_u64(dst)
        h = C.c_void_p()
        check(load_library().fhe_b200_lincomb_create_conv(s.ctypes.data_as(u64p), len(s), d.ctypes.data_as(u64p), len(d),
                                                          device, C.byref(h)))
        return cls(h, len(s), len(d))

    @classmethod
    def scale(cls, qs, ps, t, targets, with_extra, device=0):
        q, tg = _np_u64(qs), _np_u64(targets)
        p = _np_u64(ps) if len(ps) else np.zeros(1, dtype=np.uint64)
        h = C.c_void_p()
        check(load_library().fhe_b200_lincomb_create_scale(q.ctypes.data_as(u64p), len(q), p.ctypes.data_as(u64p), len(ps),
                                                           int(t), tg.ctypes.data_as(u64p), len(tg), int(with_extra),
                                                           device, C.byref(h)))
        return cls(h, len(q), len(tg))

    def constants(self):
        S, T = self.S, self.T
        arr = dict(pre=np.zeros(S, np.uint64), th_hi=np.zeros(S, np.uint64), th_lo=np.zeros(S, np.uint64),
                   M=np.zeros(S * T, np.uint64), c=np.zeros(T, np.uint64), lam=np.zeros(T, np.uint64))
        check(self.lib.fhe_b200_lincomb_constants(self.h, *[a.ctypes.data_as(u64p) for a in arr.values()]))
        arr["M"] = arr["M"].reshape(S, T)
        return arr

    def apply(self, x: torch.Tensor, extra: torch.Tensor | None = None) -> torch.Tensor:
        n = x.shape[-1]
        batch = x.numel() // (self.S * n)
        out = torch.empty(x.shape[:-2] + (self.T, n), dtype=torch.int64, device=x.device)
        check(self.lib.fhe_b200_lincomb_apply(self.h, _ptr(out), _ptr(x), _ptr(extra) if extra is not None else None,
                                              n, batch, _stream()))
        return out

    def __del__(self):
        try:
            self.lib.fhe_b200_lincomb_destroy(self.h)
        except Exception:
            pass


class RNSContext:
    """fhe::RNSContext(primes): RNS arithmetic on limb-major residues [num_primes][count]."""

    def __init__(self, primes, n: int, device: int = 0):
        self.plan = Plan(n, primes, device)
        self.primes = self.plan.moduli

    def num_primes(self): return len(self.primes)
    def add_rns(self, result, a, b): return self.plan.add(a, b, out=result)
    def sub_rns(self, result, a, b): return self.plan.sub(a, b, out=result)
    def mul_rns(self, result, a, b): return self.plan.mul(a, b, out=result)
    def mod_switch_rns(self, x): return self.plan.modswitch_drop_last(x)

    def base_extend(self, x, target_primes):
        lc = LinComb.conv(self.primes, target_primes, self.plan.device)
        return lc.apply(x)


class BfvContext:
    """fhe::FHEContext on RNS (BFV).  primes = Q (L) then auxiliary basis (R); P = first K auxiliary primes."""

    def __init__(self, n, L, R, K, dnum, t, primes, sigma=3.2, hamming_weight=64, device=0):
        self.lib = load_library()
        self.n, self.L, self.R, self.K, self.dnum, self.t = n, L, R, K, dnum, t
        p = _np_u64(primes)
        assert len(p) == L + R
        h = C.c_void_p()
        check(self.lib.fhe_b200_bfv_create(n, L, R, K, dnum, int(t), p.ctypes.data_as(u64p), float(sigma),
                                           int(hamming_weight), device, C.byref(h)))
        self.h = h
        self.dev = torch.device("cuda", device)

    def _empty(self, *shape): return torch.empty(shape, dtype=torch.int64, device=self.dev)

    def set_rng_key(self, key32: bytes | None):
        """256-bit key -> samplers draw from ChaCha20 (production); None -> the reproducible splitmix generator (tests)"""
        assert key32 is None or len(key32) == 32
        check(self.lib.fhe_b200_bfv_set_rng_key(self.h, key32))

    def keygen(self, seed_sk, seed_pk):
        sk = self._empty(self.L + self.R, self.n); pk = self._empty(2, self.L, self.n)
        check(self.lib.fhe_b200_bfv_keygen(self.h, seed_sk, seed_pk, _ptr(sk), _ptr(pk), _stream()))
        return sk, pk

    def relinkey_gen(self, seed, sk):
        rlk = self._empty(self.dnum, 2, self.L + self.K, self.n)
        check(self.lib.fhe_b200_bfv_relinkeygen(self.h, seed, _ptr(sk), _ptr(rlk), _stream()))
        return rlk

    def encode(self, values) -> torch.Tensor:
        pt = np.zeros(self.n, dtype=np.uint64)
        v = np.asarray(values, dtype=np.uint64)[: self.n]
        pt[: v.size] = v % np.uint64(self.t)
        return to_device(pt, self.dev)

    def decode(self, pt: torch.Tensor): return to_host(pt)

    def encrypt(self, seed, pt, pk):
        batch = pt.numel() // self.n
        ct = self._empty(batch, 2, self.L, self.n)
        check(self.lib.fhe_b200_bfv_encrypt(self.h, seed, _ptr(pt), _ptr(pk), _ptr(ct), batch, _stream()))
        return ct

    def decrypt(self, ct, sk):
        batch = ct.numel() // (2 * self.L * self.n)
        pt = self._empty(batch, self.n)
        check(self.lib.fhe_b200_bfv_decrypt(self.h, _ptr(ct), _ptr(sk), _ptr(pt), batch, _stream()))
        return pt

    def add(self, a, b):
        out = torch.empty_like(a)
        check(self.lib.fhe_b200_bfv_add(self.h, _ptr(a), _ptr(b), _ptr(out), a.numel() // (2 * self.L * self.n), _stream()))
        return out

    def sub(self, a, b):
        out = torch.empty_like(a)
        check(self.lib.fhe_b200_bfv_sub(self.h, _ptr(a), _ptr(b), _ptr(out), a.numel() // (2 * self.L * self.n), _stream()))
        return out

    def add_plain(self, ct, pt, subtract=False):
        out = torch.empty_like(ct)
        check(self.lib.fhe_b200_bfv_add_plain(self.h, _ptr(ct), _ptr(pt), _ptr(out), ct.numel() // (2 * self.L * self.n),
                                              int(subtract), _stream()))
        return out

    def multiply_plain(self, ct, pt):
        out = torch.empty_like(ct)
        check(self.lib.fhe_b200_bfv_multiply_plain(self.h, _ptr(ct), _ptr(pt), _ptr(out), ct.numel() // (2 * self.L * self.n),
                                                   _stream()))
        return out

    # Galois automorphisms / rotations and the modulus chain (include/fhe.cuh:59-61,86,109-116)
    def galoiskey_gen(self, seed, galois_elt, sk):
        gk = self._empty(self.dnum, 2, self.L + self.K, self.n)
        check(self.lib.fhe_b200_bfv_galoiskeygen(self.h, C.c_uint64(seed), C.c_uint32(galois_elt), _ptr(sk), _ptr(gk), _stream()))
        return gk

    def apply_galois(self, ct, galois_elt, gk):
        out = torch.empty_like(ct)
        check(self.lib.fhe_b200_bfv_apply_galois(self.h, _ptr(ct), C.c_uint32(galois_elt), _ptr(gk), _ptr(out),
                                                 ct.numel() // (2 * self.L * self.n), _stream()))
        return out

    def rotate_rows(self, ct, steps, gk): return self.apply_galois(ct, pow(3, steps % (self.n // 2), 2 * self.n), gk)
    def rotate_columns(self, ct, gk): return self.apply_galois(ct, 2 * self.n - 1, gk)

    def mod_switch_to_next(self, ct):
        b = ct.numel() // (2 * self.L * self.n)
        out = self._empty(b, 2, self.L - 1, self.n)
        check(self.lib.fhe_b200_bfv_mod_switch_to_next(self.h, _ptr(ct), _ptr(out), b, _stream()))
        return out

    def mod_switch_to_level(self, ct, drop):
        """drop the last `drop` limbs of Q (FHEContext::mod_switch_to_level, include/fhe.cuh:110); result [B][2][L-drop][N]"""
        b = ct.numel() // (2 * self.L * self.n)
        out = self._empty(b, 2, self.L - drop, self.n)
        check(self.lib.fhe_b200_bfv_mod_switch_to_level(self.h, _ptr(ct), _ptr(out), b, drop, _stream()))
        return out

    # SIMD slot encoding (fhe::BatchEncoder, include/fhe.cuh:151-166): slot i is the value of the plaintext polynomial at
    # the evaluation point the engine's NTT puts at position i; needs t = 1 (mod 2N)
    def _slot_plan(self):
        if getattr(self, "_tplan", None) is None:
            self._tplan = Plan(self.n, [self.t], self.dev.index or 0)
        return self._tplan

    def batch_encode(self, values) -> torch.Tensor:
        v = np.zeros(self.n, dtype=np.uint64)
        a = np.asarray(values, dtype=np.uint64)[: self.n]
        v[: a.size] = a % np.uint64(self.t)
        d = to_device(v.reshape(1, 1, self.n), self.dev)
        return self._slot_plan().inverse(d).view(self.n)

    def batch_decode(self, pt: torch.Tensor) -> np.ndarray:
        d = pt.clone().view(-1, 1, self.n)
        return to_host(self._slot_plan().forward(d)).reshape(-1, self.n)

    def multiply(self, a, b, rlk, want_scaled=False, out=None):
        batch = a.numel() // (2 * self.L * self.n)
        out = torch.empty_like(a) if out is None else out
        sc = self._empty(batch, 3, self.L, self.n) if want_scaled else None
        check(self.lib.fhe_b200_bfv_multiply_relin(self.h, _ptr(a), _ptr(b), _ptr(rlk), _ptr(out),
                                                   _ptr(sc) if want_scaled else None, batch, _stream()))
        return (out, sc) if want_scaled else out

    def noise_budget(self, ct, sk) -> np.ndarray:
        """invariant noise budget in bits per ciphertext (FHEContext::estimate_noise_budget, include/fhe.cuh:142); synchronises."""
        batch = ct.numel() // (2 * self.L * self.n)
        bits = np.zeros(batch, dtype=np.float64)
        check(self.lib.fhe_b200_bfv_noise_budget(self.h, _ptr(ct), _ptr(sk), batch, bits.ctypes.data_as(C.POINTER(C.c_double)), _stream()))
        return bits

    def multiply_no_relin(self, a, b, out=None):
        """3-component product [B][3][L][N] (FHEContext::multiply before its relinearize call, src/fhe.cu:198-219); a is b (the same
        tensor) takes the squaring path."""
        batch = a.numel() // (2 * self.L * self.n)
        out = out if out is not None else self._empty(batch, 3, self.L, self.n)
        check(self.lib.fhe_b200_bfv_multiply(self.h, _ptr(a), _ptr(b), _ptr(out), batch, _stream()))
        return out

    def relinearize(self, ct3, rlk, out=None):
        """[B][3][L][N] -> [B][2][L][N] by hybrid key switching of the third component (FHEContext::relinearize, src/fhe.cu:226-235)."""
        batch = ct3.numel() // (3 * self.L * self.n)
        out = out if out is not None else self._empty(batch, 2, self.L, self.n)
        check(self.lib.fhe_b200_bfv_relinearize(self.h, _ptr(ct3), _ptr(rlk), _ptr(out), batch, _stream()))
        return out

    def add3(self, a3, b3):
        """sum of two 3-component ciphertexts (lazy relinearisation: relinearize the sum once)."""
        out = torch.empty_like(a3)
        batch = a3.numel() // (self.L * self.n)
        check(self.lib.fhe_b200_poly_add(self.lib.fhe_b200_bfv_plan(self.h), _ptr(out), _ptr(a3), _ptr(b3), batch, 0, self.L, _stream()))
        return out

    def multiply_host(self, h_a: np.ndarray, h_b: np.ndarray, rlk, h_out: np.ndarray):
        batch = h_a.size // (2 * self.L * self.n)
        check(self.lib.fhe_b200_bfv_multiply_relin_host(self.h, h_a.ctypes.data_as(C.c_void_p), h_b.ctypes.data_as(C.c_void_p),
                                                        _ptr(rlk), h_out.ctypes.data_as(C.c_void_p), batch))
        return h_out

    def close(self):
        if getattr(self, "h", None):
            self.lib.fhe_b200_bfv_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


SHARD_HANDLE_BYTES = 128


def shard_partition(total: int, rank: int, world: int) -> tuple[int, int]:
    """the limb block partition of the sharded path (host only): -> (begin, count)"""
    b, c = C.c_uint32(), C.c_uint32()
    check(load_library().fhe_b200_shard_partition(total, rank, world, C.byref(b), C.byref(c)))
    return b.value, c.value


class BfvShard:
    """One rank of a limb-sharded BFV group (fhe_b200_shard, csrc/shard.cu): NTT-domain work limb-sharded, base conversions
    coefficient-sharded, the transpositions done by the producing kernels' stores into the peers' buffers over NVLink.

    Set-up is collective: every rank creates its object, the 128-byte handles are carried to every rank in rank order (pass
    `handles=` explicitly, or let `connect()` all-gather them through torch.distributed), then `connect`.
    Ciphertexts are sharded by coefficient block: [batch][2][L][N/world]; keys by limb: [dnum][2][cW][N]."""

    def __init__(self, ctx: "BfvContext", rank: int, world: int, max_batch: int = 1):
        self.lib = load_library()
        self.ctx, self.rank, self.world, self.max_batch = ctx, int(rank), int(world), int(max_batch)
        h = C.c_void_p()
        check(self.lib.fhe_b200_shard_create(ctx.h, self.rank, self.world, self.max_batch, C.byref(h)))
        self.h = h
        v = [C.c_uint32() for _ in range(6)]
        nb = C.c_uint64()
        check(self.lib.fhe_b200_shard_info(self.h, *[C.byref(x) for x in v], C.byref(nb)))
        (self.coeff_begin, self.coeff_count, self.key_limb_begin, self.key_limb_count, self.ext_limb_begin,
         self.ext_limb_count) = (x.value for x in v)

    def handle(self) -> bytes:
        buf = C.create_string_buffer(SHARD_HANDLE_BYTES)
        check(self.lib.fhe_b200_shard_handle(self.h, buf))
        return buf.raw

    def connect(self, handles=None, group=None):
        if handles is None:
            import torch.distributed as dist
            handles = [None] * self.world
            dist.all_gather_object(handles, self.handle(), group=group)
        assert len(handles) == self.world and all(len(x) == SHARD_HANDLE_BYTES for x in handles)
        blob = C.create_string_buffer(b"".join(handles), SHARD_HANDLE_BYTES * self.world)
        check(self.lib.fhe_b200_shard_connect(self.h, blob))
        nb = C.c_uint64()
        check(self.lib.fhe_b200_shard_info(self.h, None, None, None, None, None, None, C.byref(nb)))
        self.nvlink_bytes_per_op = nb.value
        return self

    def shard_ct(self, ct: torch.Tensor) -> torch.Tensor:
        """full ciphertexts [B][2][L][N] -> this rank's coefficient block [B][2][L][N/world]"""
        return ct[..., self.coeff_begin:self.coeff_begin + self.coeff_count].contiguous()

    def slice_key(self, key: torch.Tensor) -> torch.Tensor:
        out = torch.empty((self.ctx.dnum, 2, self.key_limb_count, self.ctx.n), dtype=torch.int64, device=key.device)
        check(self.lib.fhe_b200_shard_slice_key(self.h, _ptr(key), _ptr(out), _stream()))
        return out

    def multiply(self, a: torch.Tensor, b: torch.Tensor, key_slice: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        batch = a.numel() // (2 * self.ctx.L * self.coeff_count)
        out = torch.empty_like(a) if out is None else out
        check(self.lib.fhe_b200_bfv_multiply_relin_sharded(self.h, _ptr(a), _ptr(b), _ptr(key_slice), _ptr(out), batch, _stream()))
        return out

    def multiply_stage(self, stage: int, a, b, key_slice, out):
        """one stage (0..4) of multiply -- for driving several ranks on one device: stage k for every rank before stage k+1"""
        batch = a.numel() // (2 * self.ctx.L * self.coeff_count)
        check(self.lib.fhe_b200_bfv_multiply_relin_sharded_stage(self.h, stage, _ptr(a), _ptr(b), _ptr(key_slice), _ptr(out), batch, _stream()))
        return out

    def check(self):
        check(self.lib.fhe_b200_shard_check(self.h, _stream()))

    def close(self):
        if getattr(self, "h", None):
            self.lib.fhe_b200_shard_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
