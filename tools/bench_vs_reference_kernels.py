"""The reference's own element-wise kernels (batch_mod_add / sub / mul_kernel, src/bigint.cu:171-215, 256-bit operands, compiled
from /root/reference into oracle/_ref) against the C-ABI element-wise ops (64-bit residues) on the same number of coefficients, same
box, device-resident operands, CUDA events.  The only part of the reference that launches legally and computes something defined
(SURVEY 8d); its NTT does not.  Prints one JSON object."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fhe_b200, oracle
L = oracle.ref_kernels()
assert L is not None, "oracle/_ref/libref_kernels.so missing (run __graft_entry__.build() where /root/reference exists)"
L.ref_time_batch_kernel.restype = C.c_int
L.ref_time_batch_kernel.argtypes = [C.c_int, C.c_uint64, C.c_uint32, C.c_int, C.POINTER(C.c_float)]
n, batch, reps = 1 << 16, 256, 20
count = n * batch
q = oracle.prime_chain(1)[0]
plan = fhe_b200.Plan(n, [q])
a = torch.randint(0, q, (batch, 1, n), dtype=torch.int64, device="cuda"); b = torch.randint(0, q, (batch, 1, n), dtype=torch.int64, device="cuda")
out = torch.empty_like(a)
res = {"coefficients": count, "modulus": q}
for op, name, fn in ((0, "add", plan.add), (1, "sub", plan.sub), (2, "mul", plan.mul)):
    ms = C.c_float()
    assert L.ref_time_batch_kernel(op, q, count, reps, C.byref(ms)) == 0
    fn(a, b, out=out); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn(a, b, out=out)
    e1.record(); torch.cuda.synchronize()
    ours = e0.elapsed_time(e1) / reps
    res[name] = {"reference_ms": round(ms.value, 4), "reference_Gcoef_s": round(count / ms.value / 1e6, 2), "reference_GBs": round(96 * count / ms.value / 1e6, 1),
                 "ours_ms": round(ours, 4), "ours_Gcoef_s": round(count / ours / 1e6, 2), "ours_GBs": round(24 * count / ours / 1e6, 1),
                 "speedup": round(ms.value / ours, 2)}
print(json.dumps(res))
