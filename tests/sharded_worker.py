"""One rank of the multi-process limb-sharded multiply check (launched by tests/test_gpu_sharded.py and tools/ through
`python -m torch.distributed.run --nproc-per-node G tests/sharded_worker.py [--preset c4] [--batch B]`).
Every rank computes the plain single-GPU multiply of the same ciphertexts as the expected value, runs the sharded multiply with
the other ranks (CUDA IPC peer buffers, stores over NVLink) and compares its coefficient block word for word.  Exit code 0 = equal."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", default="mid")
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    import fhe_b200
    from fhe_b200.engine import to_device
    from fhe_b200.params import bfv_preset

    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    p = bfv_preset(a.preset)
    g = fhe_b200.BfvContext(p["n"], p["L"], p["R"], p["K"], p["dnum"], p["t"], p["primes"], p["sigma"], p["hamming_weight"], device=lr)
    sk, pk = g.keygen(1, 2)
    rlk = g.relinkey_gen(3, sk)
    rng = np.random.default_rng(77)
    B = a.batch
    m1 = rng.integers(0, p["t"], (B, p["n"]), dtype=np.uint64); m2 = rng.integers(0, p["t"], (B, p["n"]), dtype=np.uint64)
    ca = g.encrypt(10, to_device(m1, f"cuda:{lr}"), pk); cb = g.encrypt(100, to_device(m2, f"cuda:{lr}"), pk)
    want = g.multiply(ca, cb, rlk)
    sh = fhe_b200.BfvShard(g, rank, world, max_batch=B).connect()
    ks = sh.slice_key(rlk)
    sa, sb = sh.shard_ct(ca), sh.shard_ct(cb)
    ok = True
    out = None
    for _ in range(a.reps):                     # back to back, no barrier: exercises the buffer-reuse ordering (default stream: kernel by kernel)
        out = sh.multiply(sa, sb, ks, out=out)
    sh.check()
    ok = ok and bool(torch.equal(out, sh.shard_ct(want)))
    torch.cuda.synchronize()
    with torch.cuda.stream(torch.cuda.Stream()):        # a real stream: eager, captured into a CUDA graph, replayed twice
        out2 = torch.zeros_like(out)
        for _ in range(4):
            out2.zero_()
            sh.multiply(sa, sb, ks, out=out2)
        sh.check()
        ok = ok and bool(torch.equal(out2, sh.shard_ct(want)))
    sq = sh.multiply(sa, sa, ks)                # squaring path
    sh.check()
    ok = ok and bool(torch.equal(sq, sh.shard_ct(g.multiply(ca, ca, rlk))))
    flag = torch.tensor([0 if ok else 1], device=f"cuda:{lr}")
    dist.all_reduce(flag)
    if rank == 0:
        print(f"sharded multiply x{world} ranks, preset {a.preset}, batch {B}: {'bit-exact' if int(flag) == 0 else 'MISMATCH'}", flush=True)
    dist.barrier()
    sh.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 0 else 1)


if __name__ == "__main__":
    main()
