"""gpu-homomorphic-encryption_b200 -- B200-native RNS-NTT polynomial-arithmetic engine.

Python here is harness only: a ctypes binding of the C ABI in include/fhe_b200.h (libfhe_b200.so, built in-tree
by build.py) plus thin classes that mirror the reference's operator interface (fhe::NTTEngine, RNS_NTTEngine,
PolynomialOps, RNSContext, FHEContext -- /root/reference/include/*.cuh) over torch CUDA tensors.  There is no
CPU fallback: importing works anywhere (so the library's exports can be checked), every compute call needs the
CUDA library and a B200.

The directory name contains '-', so import it through the repo-root shim:  ``import fhe_b200``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfhe_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "fhe_b200.h")

u64p = C.POINTER(C.c_uint64)
_vp = C.c_void_p


class FheB200Error(RuntimeError):
    pass


_lib = None


def load_library() -> C.CDLL:
    """dlopen libfhe_b200.so.  Fails loudly if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FheB200Error(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.fhe_b200_last_error.restype = C.c_char_p
    lib.fhe_b200_launch_count.restype = C.c_uint64
    lib.fhe_b200_profile_enable.argtypes = [C.c_int]
    lib.fhe_b200_profile_read.argtypes = [C.c_int, u64p, C.POINTER(C.c_double), u64p]
    lib.fhe_b200_plan_n.restype = C.c_uint32
    lib.fhe_b200_plan_limbs.restype = C.c_uint32
    sig = {
        "fhe_b200_plan_create": [C.c_uint32, u64p, C.c_uint32, C.c_int, C.POINTER(_vp)],
        "fhe_b200_plan_destroy": [_vp],
        "fhe_b200_plan_moduli": [_vp, u64p],
        "fhe_b200_plan_tables": [_vp, C.c_uint32, u64p, u64p, u64p, u64p],
        "fhe_b200_ntt_forward": [_vp, _vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, _vp],
        "fhe_b200_ntt_inverse": [_vp, _vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, _vp],
        "fhe_b200_negacyclic_mul": [_vp, _vp, _vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, _vp],
        "fhe_b200_bitrev_permute": [_vp, _vp, _vp, C.c_uint32, _vp],
        "fhe_b200_ntt_host": [_vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int],
        "fhe_b200_poly_add": [_vp, _vp, _vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, _vp],
        "fhe_b200_poly_sub": [_vp, _vp, _vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, _vp],
        "fhe_b200_poly_mul": [_vp, _vp, _vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, _vp],
        "fhe_b200_poly_mac": [_vp, _vp, _vp, _vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, _vp],
        "fhe_b200_poly_mul_scalar": [_vp, _vp, _vp, u64p, C.c_uint32, C.c_uint32, C.c_uint32, _vp],
        "fhe_b200_poly_add_scalar": [_vp, _vp, _vp, u64p, C.c_uint32, C.c_uint32, C.c_uint32, _vp],
        "fhe_b200_poly_negate": [_vp, _vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, _vp],
        "fhe_b200_unpack_u256": [_vp, _vp, C.c_size_t, _vp],
        "fhe_b200_pack_u256": [_vp, _vp, C.c_size_t, _vp],
        "fhe_b200_to_rns_u256": [_vp, _vp, _vp, C.c_size_t, C.c_uint32, C.c_uint32, _vp],
        "fhe_b200_from_rns_u256": [_vp, _vp, _vp, C.c_size_t, C.c_uint32, C.c_uint32, _vp],
        "fhe_b200_lincomb_create_conv": [u64p, C.c_uint32, u64p, C.c_uint32, C.c_int, C.POINTER(_vp)],
        "fhe_b200_lincomb_create_scale": [u64p, C.c_uint32, u64p, C.c_uint32, C.c_uint64, u64p, C.c_uint32, C.c_int,
                                          C.c_int, C.POINTER(_vp)],
        "fhe_b200_lincomb_destroy": [_vp],
        "fhe_b200_lincomb_apply": [_vp, _vp, _vp, _vp, C.c_uint32, C.c_uint32, _vp],
        "fhe_b200_lincomb_constants": [_vp, u64p, u64p, u64p, u64p, u64p, u64p],
        "fhe_b200_modswitch_drop_last": [_vp, _vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, _vp],
        "fhe_b200_bfv_create": [C.c_uint32] * 5 + [C.c_uint64, u64p, C.c_float, C.c_uint32, C.c_int, C.POINTER(_vp)],
        "fhe_b200_bfv_destroy": [_vp],
        "fhe_b200_bfv_set_rng_key": [_vp, C.c_char_p],
        "fhe_b200_bfv_keygen": [_vp, C.c_uint64, C.c_uint64, _vp, _vp, _vp],
        "fhe_b200_bfv_relinkeygen": [_vp, C.c_uint64, _vp, _vp, _vp],
        "fhe_b200_bfv_encrypt": [_vp, C.c_uint64, _vp, _vp, _vp, C.c_uint32, _vp],
        "fhe_b200_bfv_decrypt": [_vp, _vp, _vp, _vp, C.c_uint32, _vp],
        "fhe_b200_bfv_add": [_vp, _vp, _vp, _vp, C.c_uint32, _vp],
        "fhe_b200_bfv_multiply_relin": [_vp, _vp, _vp, _vp, _vp, _vp, C.c_uint32, _vp],
        "fhe_b200_bfv_multiply": [_vp, _vp, _vp, _vp, C.c_uint32, _vp],
        "fhe_b200_bfv_noise_budget": [_vp, _vp, _vp, C.c_uint32, C.POINTER(C.c_double), _vp],
        "fhe_b200_bfv_relinearize": [_vp, _vp, _vp, _vp, C.c_uint32, _vp],
        "fhe_b200_bfv_sub": [_vp, _vp, _vp, _vp, C.c_uint32, _vp],
        "fhe_b200_bfv_add_plain": [_vp, _vp, _vp, _vp, C.c_uint32, C.c_int, _vp],
        "fhe_b200_bfv_multiply_plain": [_vp, _vp, _vp, _vp, C.c_uint32, _vp],
        "fhe_b200_bfv_galoiskeygen": [_vp, C.c_uint64, C.c_uint32, _vp, _vp, _vp],
        "fhe_b200_bfv_apply_galois": [_vp, _vp, C.c_uint32, _vp, _vp, C.c_uint32, _vp],
        "fhe_b200_bfv_mod_switch_to_next": [_vp, _vp, _vp, C.c_uint32, _vp],
        "fhe_b200_bfv_mod_switch_to_level": [_vp, _vp, _vp, C.c_uint32, C.c_uint32, _vp],
        "fhe_b200_bfv_tensor": [_vp, _vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, _vp],
        "fhe_b200_bfv_ks_inner": [_vp, _vp, _vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _vp],
        "fhe_b200_bfv_multiply_relin_host": [_vp, _vp, _vp, _vp, _vp, C.c_uint32],
        "fhe_b200_measure_int_peaks": [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int],
        "fhe_b200_shard_create": [_vp, C.c_int, C.c_int, C.c_uint32, C.POINTER(_vp)],
        "fhe_b200_shard_destroy": [_vp],
        "fhe_b200_shard_handle": [_vp, _vp],
        "fhe_b200_shard_connect": [_vp, _vp],
        "fhe_b200_shard_partition": [C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)],
        "fhe_b200_shard_info": [_vp] + [C.POINTER(C.c_uint32)] * 6 + [u64p],
        "fhe_b200_shard_slice_key": [_vp, _vp, _vp, _vp],
        "fhe_b200_bfv_multiply_relin_sharded": [_vp, _vp, _vp, _vp, _vp, C.c_uint32, _vp],
        "fhe_b200_bfv_multiply_relin_sharded_stage": [_vp, C.c_int, _vp, _vp, _vp, _vp, C.c_uint32, _vp],
        "fhe_b200_shard_check": [_vp, _vp],
        "fhe_b200_bfv_info": [_vp] + [C.POINTER(C.c_uint32)] * 5 + [u64p],
        "fhe_b200_bfv_plan": [_vp],
        "fhe_b200_gaussian_cdt": [C.c_double, u64p, C.c_uint32],
        "fhe_b200_wire_pack": [C.c_uint32] * 4 + [C.c_int, C.c_uint32, u64p, u64p, _vp],
        "fhe_b200_wire_unpack": [_vp, C.c_size_t, u64p, C.c_uint32] + [C.POINTER(C.c_uint32)] * 4 + [C.POINTER(C.c_int),
                                 C.POINTER(C.c_uint32), u64p],
    }
    lib.fhe_b200_wire_size.argtypes = [C.c_uint32] * 3
    lib.fhe_b200_wire_size.restype = C.c_size_t
    for name, argtypes in sig.items():
        fn = getattr(lib, name, None)
        if fn is None:
            continue          # tests/test_cabi_exports.py reports any symbol the header declares but the .so lacks
        fn.argtypes = argtypes
        if name != "fhe_b200_bfv_plan":
            fn.restype = C.c_int
        else:
            fn.restype = _vp
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load_library().fhe_b200_last_error()
        raise FheB200Error(f"libfhe_b200 error {rc}: {msg.decode() if msg else ''}")


from .engine import (BfvContext, BfvShard, LinComb, NTTEngine, Plan, PolynomialOps, RNSContext, RNS_NTTEngine,  # noqa: E402
                     gaussian_cdt, pinned_empty, wire_pack, wire_unpack, WIRE_KINDS)

__all__ = ["load_library", "check", "FheB200Error", "Plan", "NTTEngine", "RNS_NTTEngine", "PolynomialOps",
           "RNSContext", "LinComb", "BfvContext", "BfvShard", "gaussian_cdt", "pinned_empty", "wire_pack", "wire_unpack", "WIRE_KINDS", "LIB_PATH", "HEADER_PATH"]
