"""CPU oracle (TEST INFRASTRUCTURE ONLY) -- ctypes binding of oracle/liboracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this package.  The product
(gpu-homomorphic-encryption_b200/) never does.

Every function restates a piece of the reference's RNS-NTT / BFV path; the C
sources cite the reference file:line each one follows.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

u64p = C.POINTER(C.c_uint64)
i64p = C.POINTER(C.c_int64)
u32p = C.POINTER(C.c_uint32)


def build(force: bool = False) -> str:
    """Compile liboracle.so with gcc (oracle/Makefile)."""
    srcs = [os.path.join(_HERE, f) for f in ("orc_ntt.c", "orc_rns.c", "orc_bfv.c", "orc_math.h", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_NATIVE_LIB = os.path.join(_HERE, "_native", "liboracle_native.so")


def build_native() -> str | None:
    """The same sources compiled with -O3 -march=native ON THE MACHINE THAT RUNS THEM (bench.py's CPU baseline legs call this on
    the GPU box; the file is neither committed nor shipped, since the build container's CPU may differ).  None if it fails."""
    try:
        os.makedirs(os.path.dirname(_NATIVE_LIB), exist_ok=True)
        srcs = [os.path.join(_HERE, f) for f in ("orc_ntt.c", "orc_rns.c", "orc_bfv.c")]
        subprocess.check_call(["/usr/bin/gcc", "-O3", "-march=native", "-funroll-loops", "-fopenmp", "-fPIC", "-std=gnu11", "-shared",
                               "-o", _NATIVE_LIB] + srcs + ["-lm"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        return _NATIVE_LIB
    except Exception:
        return None


def use_native() -> bool:
    """switch this process to the -march=native build (call before the first oracle call); False if it could not be built"""
    global _LIB_PATH, _lib
    path = build_native()
    if not path:
        return False
    _LIB_PATH, _lib = path, None
    return True


_REF_ROOT = "/root/reference"
_REF_LIB = os.path.join(_HERE, "_ref", "libref_kernels.so")


def build_ref() -> str | None:
    """Compile the REFERENCE's own complete device functions (add_mod / sub_mod / mul_mod_montgomery through its batch kernels,
    src/bigint.cu:171-215) from the sources where they lie into oracle/_ref/ (git-ignored).  Only where /root/reference exists
    (the build container); the GPU box uses the prebuilt file that travels with the snapshot."""
    if os.path.isdir(_REF_ROOT):
        subprocess.check_call(["make", "-C", _HERE, "-s", "_ref"], stdout=subprocess.DEVNULL)
    return _REF_LIB if os.path.exists(_REF_LIB) else None


_REF_NTT_LIB = os.path.join(_HERE, "_ref", "libref_ntt.so")
_ref_ntt_lib = None


def ref_time_poly(what: str, q: int, count: int, reps: int = 20):
    """milliseconds per call of the REFERENCE's own kernel on cuda:0 (oracle/ref_ntt_kernels.cu): "pointwise_mul"
    (ntt_pointwise_mul_kernel), "poly_add" (poly_add_kernel) over `count` coefficients, or "ntt_forward" (NTTEngine::forward's two
    launches, count <= 1024).  None when oracle/_ref was never built or the launch is illegal in the reference."""
    global _ref_ntt_lib
    if _ref_ntt_lib is None:
        if not os.path.exists(_REF_NTT_LIB):
            return None
        _ref_ntt_lib = C.CDLL(_REF_NTT_LIB)
        _ref_ntt_lib.ref_time_poly_kernel.restype = C.c_int
        _ref_ntt_lib.ref_time_poly_kernel.argtypes = [C.c_int, C.c_uint64, C.c_uint32, C.c_int, C.POINTER(C.c_float)]
    ms = C.c_float()
    rc = _ref_ntt_lib.ref_time_poly_kernel({"pointwise_mul": 0, "poly_add": 1, "ntt_forward": 2}[what], C.c_uint64(q), count, reps, C.byref(ms))
    return float(ms.value) if rc == 0 else None


_ref_lib = None


def ref_kernels():
    """ctypes handle on oracle/_ref/libref_kernels.so (needs a GPU to call), or None when it was never built."""
    global _ref_lib
    if _ref_lib is None and os.path.exists(_REF_LIB):
        L = C.CDLL(_REF_LIB)
        for f in ("ref_batch_mod_add", "ref_batch_mod_sub", "ref_batch_mod_mul_montgomery"):
            getattr(L, f).restype = C.c_int
            getattr(L, f).argtypes = [u64p, u64p, C.c_uint64, u64p, C.c_uint32, u64p]
        _ref_lib = L
    return _ref_lib


def ref_batch(op: str, a, b, q: int):
    """run the reference's batch_mod_{add,sub,mul}_kernel on cuda:0 -> (uint64 results, True if every result fitted 64 bits)"""
    L = ref_kernels()
    a = np.ascontiguousarray(a, dtype=np.uint64); b = np.ascontiguousarray(b, dtype=np.uint64)
    out = np.zeros_like(a); hi = np.zeros(1, dtype=np.uint64)
    fn = {"add": L.ref_batch_mod_add, "sub": L.ref_batch_mod_sub, "mul_montgomery": L.ref_batch_mod_mul_montgomery}[op]
    rc = fn(a.ctypes.data_as(u64p), b.ctypes.data_as(u64p), C.c_uint64(q), out.ctypes.data_as(u64p), a.size, hi.ctypes.data_as(u64p))
    if rc:
        raise RuntimeError(f"reference kernel failed (rc={rc})")
    return out, int(hi[0]) == 0


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        L = _lib
        L.orc_find_psi.restype = C.c_uint64
        L.orc_find_psi.argtypes = [C.c_uint64, C.c_uint32]
        for f in ("orc_add_mod", "orc_sub_mod", "orc_mul_mod", "orc_pow_mod"):
            getattr(L, f).restype = C.c_uint64
            getattr(L, f).argtypes = [C.c_uint64] * 3
        L.orc_inv_mod.restype = C.c_uint64
        L.orc_inv_mod.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_inv_general.restype = C.c_uint64
        L.orc_inv_general.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_rng.restype = C.c_uint64
        L.orc_rng.argtypes = [C.c_uint64] * 3
        for f in ("orc_rng_key_of", "orc_item_seed_of"):
            getattr(L, f).restype = C.c_uint64
            getattr(L, f).argtypes = [C.c_uint64] * 2
        L.orc_is_prime.argtypes = [C.c_uint64]
        L.orc_prime_chain.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, u64p]
        L.orc_ntt_tables.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, u64p, u64p]
        L.orc_ntt_tables.restype = None
        L.orc_rns_tables.argtypes = [u64p, u64p, C.c_uint32, C.c_uint32]
        L.orc_rns_tables.restype = None
        L.orc_rns_ntt_batch.argtypes = [u64p, u64p, u64p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int]
        L.orc_lc_make_conv.restype = C.c_void_p
        L.orc_lc_make_conv.argtypes = [u64p, C.c_uint32, u64p, C.c_uint32]
        L.orc_lc_make_scale.restype = C.c_void_p
        L.orc_lc_make_scale.argtypes = [u64p, C.c_uint32, u64p, C.c_uint32, C.c_uint64, u64p, C.c_uint32, C.c_int]
        L.orc_lc_apply.argtypes = [C.c_void_p, u64p, u64p, u64p, C.c_uint32]
        L.orc_lc_apply.restype = None
        L.orc_lc_free.argtypes = [C.c_void_p]
        L.orc_lc_free.restype = None
        L.orc_lc_S.argtypes = [C.c_void_p]
        L.orc_lc_T.argtypes = [C.c_void_p]
        L.orc_lc_get.argtypes = [C.c_void_p] + [u64p] * 6
        L.orc_lc_get.restype = None
        L.orc_bfv_create.restype = C.c_void_p
        L.orc_bfv_create.argtypes = [C.c_uint32] * 5 + [C.c_uint64, u64p]
        L.orc_bfv_destroy.argtypes = [C.c_void_p]
        L.orc_bfv_destroy.restype = None
        L.orc_gaussian_cdt.argtypes = [C.c_double, u64p, C.c_uint32]
        L.orc_gaussian_cdt.restype = C.c_uint32
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags.c_contiguous
    return a.ctypes.data_as(u64p)


def _pi(a: np.ndarray):
    assert a.dtype == np.int64 and a.flags.c_contiguous
    return a.ctypes.data_as(i64p)


def _u64(x) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(x, dtype=np.uint64))


# ---------------------------------------------------------------- scalars / primes
def add_mod(a, b, q): return int(lib().orc_add_mod(a, b, q))
def sub_mod(a, b, q): return int(lib().orc_sub_mod(a, b, q))
def mul_mod(a, b, q): return int(lib().orc_mul_mod(a, b, q))
def pow_mod(a, e, q): return int(lib().orc_pow_mod(a, e, q))
def inv_mod(a, q): return int(lib().orc_inv_mod(a, q))
def rng64(seed, stream, idx): return int(lib().orc_rng(seed, stream, idx))
def is_prime(n): return bool(lib().orc_is_prime(n))
def find_psi(q, n): return int(lib().orc_find_psi(q, n))
def max_threads(): return int(lib().orc_max_threads())


def prime_chain(count: int, bits: int = 60, step_log2: int = 18) -> list[int]:
    """k-th largest primes p < 2^bits with p = 1 (mod 2^step_log2)."""
    out = np.zeros(count, dtype=np.uint64)
    rc = lib().orc_prime_chain(bits, step_log2, count, _p(out))
    if rc:
        raise ValueError(f"prime chain exhausted (rc={rc})")
    return [int(x) for x in out]


def ntt_tables(q: int, n: int):
    f = np.zeros(n, dtype=np.uint64); i = np.zeros(n, dtype=np.uint64)
    lib().orc_ntt_tables(q, n, find_psi(q, n), _p(f), _p(i))
    return f, i


# ---------------------------------------------------------------- transforms
def ntt_forward(a, q: int) -> np.ndarray:
    a = _u64(a).copy()
    if lib().orc_ntt_forward(_p(a), a.size, C.c_uint64(q)):
        raise ValueError("q is not NTT friendly for this N")
    return a


def ntt_inverse(a, q: int) -> np.ndarray:
    a = _u64(a).copy()
    if lib().orc_ntt_inverse(_p(a), a.size, C.c_uint64(q)):
        raise ValueError("q is not NTT friendly for this N")
    return a


def negacyclic_dft_def(a, q: int) -> np.ndarray:
    a = _u64(a); out = np.zeros_like(a)
    lib().orc_negacyclic_dft_def(_p(out), _p(a), a.size, C.c_uint64(q), C.c_uint64(find_psi(q, a.size)))
    return out


def schoolbook_negacyclic(a, b, q: int) -> np.ndarray:
    a = _u64(a); b = _u64(b); out = np.zeros_like(a)
    lib().orc_schoolbook_negacyclic(_p(out), _p(a), _p(b), a.size, C.c_uint64(q))
    return out


def negacyclic_mul_ntt(a, b, q: int) -> np.ndarray:
    a = _u64(a); b = _u64(b); out = np.zeros_like(a)
    if lib().orc_negacyclic_mul_ntt(_p(out), _p(a), _p(b), a.size, C.c_uint64(q)):
        raise ValueError("q is not NTT friendly for this N")
    return out


def bitrev_perm(n: int) -> np.ndarray:
    lg = n.bit_length() - 1
    idx = np.arange(n, dtype=np.uint32)
    r = np.zeros(n, dtype=np.uint32)
    for b in range(lg):
        r |= ((idx >> b) & 1) << (lg - 1 - b)
    return r


def set_rng_key(key32: bytes | None):
    """mirror of fhe_b200_bfv_set_rng_key: 32-byte key -> every sampler draws from ChaCha20; None -> the splitmix generator"""
    L = lib()
    L.orc_set_rng_key.argtypes = [C.c_char_p]
    L.orc_set_rng_key.restype = None
    assert key32 is None or len(key32) == 32
    L.orc_set_rng_key(key32)


def chacha20_block(key_words, counter, nonce_words):
    """RFC 8439 block function -> 16 uint32 words (known-answer test of the restatement)"""
    L = lib()
    k = (C.c_uint32 * 8)(*key_words); nn = (C.c_uint32 * 3)(*nonce_words); out = (C.c_uint32 * 16)()
    L.orc_chacha20_block_raw.restype = None
    L.orc_chacha20_block_raw(k, C.c_uint32(counter), nn, out)
    return [int(v) for v in out]


def rng_key(seed: int, stream: int) -> int:
    return int(lib().orc_rng_key_of(C.c_uint64(seed & (2**64 - 1)), C.c_uint64(stream & (2**64 - 1))))


def item_seed(seed: int, item: int) -> int:
    """seed of batch item `item` of a call seeded with `seed` (specification shared with the CUDA samplers)"""
    return int(lib().orc_item_seed_of(C.c_uint64(seed & (2**64 - 1)), C.c_uint64(item)))


class RnsNtt:
    """Batched RNS transforms over [batch][limbs][N] (RNS_NTTEngine::forward_rns/inverse_rns)."""

    def __init__(self, n: int, moduli):
        self.n = n
        self.moduli = _u64(moduli)
        self.tables = np.zeros(len(self.moduli) * 4 * n, dtype=np.uint64)
        lib().orc_rns_tables(_p(self.tables), _p(self.moduli), len(self.moduli), n)

    def _run(self, data, inverse, threads):
        data = _u64(data).copy()
        flat = data.reshape(-1)
        limbs = len(self.moduli)
        batch = flat.size // (limbs * self.n)
        assert batch * limbs * self.n == flat.size
        used = lib().orc_rns_ntt_batch(_p(flat), _p(self.tables), _p(self.moduli), limbs, self.n, batch,
                                       int(inverse), int(threads))
        self.threads_used = used
        return data

    def forward(self, data, threads=0): return self._run(data, 0, threads)
    def inverse(self, data, threads=0): return self._run(data, 1, threads)

    def run_inplace(self, flat: np.ndarray, batch: int, inverse: bool, threads: int = 0) -> int:
        return lib().orc_rns_ntt_batch(_p(flat), _p(self.tables), _p(self.moduli), len(self.moduli), self.n, batch,
                                       int(inverse), int(threads))


def _elementwise(fn, a, b, moduli, n, extra=None):
    a = _u64(a); moduli = _u64(moduli)
    limbs = len(moduli); batch = a.size // (limbs * n)
    out = np.zeros_like(a)
    if extra is not None:
        getattr(lib(), fn)(_p(out.reshape(-1)), _p(_u64(extra).reshape(-1)), _p(a.reshape(-1)), _p(_u64(b).reshape(-1)),
                           n, _p(moduli), limbs, batch)
    else:
        getattr(lib(), fn)(_p(out.reshape(-1)), _p(a.reshape(-1)), _p(_u64(b).reshape(-1)), n, _p(moduli), limbs, batch)
    return out


def poly_add(a, b, moduli, n): return _elementwise("orc_poly_add", a, b, moduli, n)
def poly_sub(a, b, moduli, n): return _elementwise("orc_poly_sub", a, b, moduli, n)
def poly_mul(a, b, moduli, n): return _elementwise("orc_poly_mul", a, b, moduli, n)
def poly_mac(acc, a, b, moduli, n): return _elementwise("orc_poly_mac", a, b, moduli, n, extra=acc)
def poly_mul_scalar(a, scalars, moduli, n): return _elementwise("orc_poly_mul_scalar", a, _u64(scalars), moduli, n)


def to_rns(values, moduli) -> np.ndarray:
    values = _u64(values); moduli = _u64(moduli)
    out = np.zeros((len(moduli), values.size), dtype=np.uint64)
    lib().orc_to_rns(_p(out.reshape(-1)), _p(values), values.size, _p(moduli), len(moduli))
    return out


def apply_galois_poly(a, g, q) -> np.ndarray:
    """a(x) -> a(x^g) in Z_q[x]/(x^n+1), g odd."""
    a = _u64(a); out = np.zeros_like(a)
    lib().orc_apply_galois_poly(_p(a), _p(out), C.c_uint32(a.size), C.c_uint32(g), C.c_uint64(q))
    return out


def modswitch_drop_last(x, moduli) -> np.ndarray:
    x = _u64(x); moduli = _u64(moduli)
    limbs = len(moduli); n = x.size // limbs
    out = np.zeros((limbs - 1, n), dtype=np.uint64)
    lib().orc_modswitch_drop_last(_p(out.reshape(-1)), _p(x.reshape(-1)), n, _p(moduli), limbs)
    return out


# ---------------------------------------------------------------- RNS linear combination
class LinComb:
    """out_k = sum_i z_i M[i][k] + (I mod m_k) c_k + extra_k lam_k (see oracle/orc_rns.c)."""

    def __init__(self, handle, src, dst):
        self.h = handle
        self.src = _u64(src); self.dst = _u64(dst)
        self.S = len(self.src); self.T = len(self.dst)

    @classmethod
    def conv(cls, src, dst):
        s = _u64(src); d = _u64(dst)
        return cls(lib().orc_lc_make_conv(_p(s), len(s), _p(d), len(d)), s, d)

    @classmethod
    def scale(cls, qs, ps, t, targets, with_extra):
        q = _u64(qs); p = _u64(ps) if len(ps) else np.zeros(1, dtype=np.uint64); tg = _u64(targets)
        return cls(lib().orc_lc_make_scale(_p(q), len(q), _p(p), len(ps), C.c_uint64(t), _p(tg), len(tg),
                                           int(with_extra)), q, tg)

    def constants(self):
        S, T = self.S, self.T
        pre = np.zeros(S, np.uint64); th = np.zeros(S, np.uint64); tl = np.zeros(S, np.uint64)
        M = np.zeros(S * T, np.uint64); c = np.zeros(T, np.uint64); lam = np.zeros(T, np.uint64)
        lib().orc_lc_get(self.h, _p(pre), _p(th), _p(tl), _p(M), _p(c), _p(lam))
        return dict(pre=pre, th_hi=th, th_lo=tl, M=M.reshape(S, T), c=c, lam=lam)

    def apply(self, x, extra=None) -> np.ndarray:
        x = _u64(x); n = x.size // self.S
        out = np.zeros((self.T, n), dtype=np.uint64)
        ex = _p(_u64(extra).reshape(-1)) if extra is not None else None
        lib().orc_lc_apply(self.h, _p(out.reshape(-1)), _p(x.reshape(-1)), ex, n)
        return out

    def __del__(self):
        try:
            lib().orc_lc_free(self.h)
        except Exception:
            pass


# ---------------------------------------------------------------- BFV
def gaussian_cdt(sigma: float) -> np.ndarray:
    buf = np.zeros(128, dtype=np.uint64)
    ln = lib().orc_gaussian_cdt(float(sigma), _p(buf), 128)
    return buf[:ln].copy()


def sample_ternary(n, seed, stream, thr=1 << 31):
    out = np.zeros(n, dtype=np.int64)
    lib().orc_sample_ternary(_pi(out), n, C.c_uint64(seed), C.c_uint64(stream), C.c_uint32(thr))
    return out


def sample_ternary_hw(n, seed, stream, hw):
    out = np.zeros(n, dtype=np.int64)
    lib().orc_sample_ternary_hw(_pi(out), n, C.c_uint64(seed), C.c_uint64(stream), C.c_uint32(hw))
    return out


def sample_gaussian(n, seed, stream, cdt):
    out = np.zeros(n, dtype=np.int64); cdt = _u64(cdt)
    lib().orc_sample_gaussian(_pi(out), n, C.c_uint64(seed), C.c_uint64(stream), _p(cdt), len(cdt))
    return out


def sample_uniform(n, q, seed, stream):
    out = np.zeros(n, dtype=np.uint64)
    lib().orc_sample_uniform(_p(out), n, C.c_uint64(q), C.c_uint64(seed), C.c_uint64(stream))
    return out


class Bfv:
    """Host RNS-BFV.  primes = Q (L) followed by the auxiliary basis (R); P = first K aux primes."""

    def __init__(self, n, L, R, K, dnum, t, primes, sigma=3.2, hw=64, thr=1 << 31):
        self.n, self.L, self.R, self.K, self.dnum, self.t = n, L, R, K, dnum, t
        self.alpha = L // dnum
        self.primes = _u64(primes)
        assert len(self.primes) == L + R
        self.h = lib().orc_bfv_create(n, L, R, K, dnum, C.c_uint64(t), _p(self.primes))
        if not self.h:
            raise ValueError("bad BFV parameters")
        self.cdt = gaussian_cdt(sigma)
        self.hw, self.thr = hw, thr

    def consts(self):
        d = np.zeros(self.L, np.uint64); p = np.zeros(self.L, np.uint64); pi = np.zeros(self.L, np.uint64)
        lib().orc_bfv_get_consts(C.c_void_p(self.h), _p(d), _p(p), _p(pi))
        return dict(delta=d, p_mod_q=p, pinv_mod_q=pi)

    def secret_keygen(self, seed):
        s = np.zeros(self.n, np.int64); sk = np.zeros((self.L + self.R, self.n), np.uint64)
        lib().orc_bfv_secret_keygen(C.c_void_p(self.h), C.c_uint64(seed), C.c_uint32(self.hw), C.c_uint32(self.thr),
                                    _pi(s), _p(sk.reshape(-1)))
        return s, sk

    def public_keygen(self, seed, sk):
        pk = np.zeros((2, self.L, self.n), np.uint64)
        lib().orc_bfv_public_keygen(C.c_void_p(self.h), C.c_uint64(seed), _p(sk.reshape(-1)), _p(self.cdt),
                                    len(self.cdt), _p(pk.reshape(-1)))
        return pk

    def relin_keygen(self, seed, sk):
        rlk = np.zeros((self.dnum, 2, self.L + self.K, self.n), np.uint64)
        lib().orc_bfv_relin_keygen(C.c_void_p(self.h), C.c_uint64(seed), _p(sk.reshape(-1)), _p(self.cdt),
                                   len(self.cdt), _p(rlk.reshape(-1)))
        return rlk

    def encode(self, values):
        pt = np.zeros(self.n, np.uint64)
        v = np.asarray(values, dtype=np.uint64)[: self.n]
        pt[: v.size] = v % np.uint64(self.t)
        return pt

    def encrypt(self, seed, pt, pk, item=0):
        """batch item `item` of an encrypt call seeded with `seed` (fhe_b200_bfv_encrypt: ciphertext b of the batch is item b)"""
        ct = np.zeros((2, self.L, self.n), np.uint64)
        lib().orc_bfv_encrypt(C.c_void_p(self.h), C.c_uint64(item_seed(seed, item)), _p(_u64(pt)), _p(pk.reshape(-1)),
                              C.c_uint32(self.thr), _p(self.cdt), len(self.cdt), _p(ct.reshape(-1)))
        return ct

    def decrypt(self, ct, sk):
        pt = np.zeros(self.n, np.uint64)
        lib().orc_bfv_decrypt(C.c_void_p(self.h), _p(_u64(ct).reshape(-1)), _p(sk.reshape(-1)), _p(pt))
        return pt

    def add(self, a, b):
        out = np.zeros((2, self.L, self.n), np.uint64)
        lib().orc_bfv_add(C.c_void_p(self.h), _p(_u64(a).reshape(-1)), _p(_u64(b).reshape(-1)), _p(out.reshape(-1)))
        return out

    def sub(self, a, b):
        out = np.zeros((2, self.L, self.n), np.uint64)
        lib().orc_bfv_sub(C.c_void_p(self.h), _p(_u64(a).reshape(-1)), _p(_u64(b).reshape(-1)), _p(out.reshape(-1)))
        return out

    def add_plain(self, ct, pt, subtract=False):
        out = np.zeros((2, self.L, self.n), np.uint64)
        lib().orc_bfv_add_plain(C.c_void_p(self.h), _p(_u64(ct).reshape(-1)), _p(_u64(pt)), _p(out.reshape(-1)), int(subtract))
        return out

    def multiply_plain(self, ct, pt):
        out = np.zeros((2, self.L, self.n), np.uint64)
        lib().orc_bfv_multiply_plain(C.c_void_p(self.h), _p(_u64(ct).reshape(-1)), _p(_u64(pt)), _p(out.reshape(-1)))
        return out

    # SIMD slot encoding (fhe::BatchEncoder): slot i = value of the plaintext polynomial at the evaluation point the
    # negacyclic NTT modulo t leaves at position i (bit-reversed order); needs t = 1 (mod 2N)
    def batch_encode(self, values):
        v = np.zeros(self.n, np.uint64)
        a = np.asarray(values, dtype=np.uint64)[: self.n]
        v[: a.size] = a % np.uint64(self.t)
        return ntt_inverse(v, self.t)

    def batch_decode(self, pt):
        return ntt_forward(pt, self.t)

    # Galois automorphisms / rotations and the modulus chain (declared only in the reference: include/fhe.cuh:59-61,86,109-116)
    def galois_keygen(self, seed, g, sk):
        gk = np.zeros((self.dnum, 2, self.L + self.K, self.n), np.uint64)
        lib().orc_bfv_galois_keygen(C.c_void_p(self.h), C.c_uint64(seed), C.c_uint32(g), _p(sk.reshape(-1)), _p(self.cdt),
                                    len(self.cdt), _p(gk.reshape(-1)))
        return gk

    def apply_galois(self, ct, g, gk):
        out = np.zeros((2, self.L, self.n), np.uint64)
        lib().orc_bfv_apply_galois(C.c_void_p(self.h), _p(_u64(ct).reshape(-1)), C.c_uint32(g), _p(gk.reshape(-1)), _p(out.reshape(-1)))
        return out

    def mod_switch_to_next(self, ct):
        out = np.zeros((2, self.L - 1, self.n), np.uint64)
        lib().orc_bfv_mod_switch_to_next(C.c_void_p(self.h), _p(_u64(ct).reshape(-1)), _p(out.reshape(-1)))
        return out

    def mod_switch_to_level(self, ct, drop):
        out = np.zeros((2, self.L - drop, self.n), np.uint64)
        lib().orc_bfv_mod_switch_to_level(C.c_void_p(self.h), _p(_u64(ct).reshape(-1)), C.c_uint32(drop), _p(out.reshape(-1)))
        return out

    def multiply(self, a, b):
        """3-component product [3][L][n] (no relinearisation)."""
        out = np.zeros((3, self.L, self.n), np.uint64)
        lib().orc_bfv_multiply(C.c_void_p(self.h), _p(_u64(a).reshape(-1)), _p(_u64(b).reshape(-1)), _p(out.reshape(-1)))
        return out

    def relinearize(self, ct3, rlk):
        out = np.zeros((2, self.L, self.n), np.uint64)
        lib().orc_bfv_relinearize(C.c_void_p(self.h), _p(_u64(ct3).reshape(-1)), _p(rlk.reshape(-1)), _p(out.reshape(-1)))
        return out

    def multiply_relin(self, a, b, rlk, want_scaled=False):
        out = np.zeros((2, self.L, self.n), np.uint64)
        sc = np.zeros((3, self.L, self.n), np.uint64) if want_scaled else None
        lib().orc_bfv_multiply_relin(C.c_void_p(self.h), _p(_u64(a).reshape(-1)), _p(_u64(b).reshape(-1)),
                                     _p(rlk.reshape(-1)), _p(out.reshape(-1)),
                                     _p(sc.reshape(-1)) if want_scaled else None)
        return (out, sc) if want_scaled else out

    def __del__(self):
        try:
            lib().orc_bfv_destroy(C.c_void_p(self.h))
        except Exception:
            pass
