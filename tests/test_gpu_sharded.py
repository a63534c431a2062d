"""Limb-sharded BFV multiply + relinearize (csrc/shard.cu, BASELINE.json config 4): sharded == single-GPU, word for word.

* virtual ranks: `world` shard objects on ONE device in one process, driven stage by stage -- the same kernels, scatter tables,
  epoch flags and peer stores as across GPUs (the "peer" buffers are in the same memory), so the driver's single-GPU test tier
  covers the logic without ever making one launch spin on a flag that a later launch on the same GPU writes;
* processes: one rank per GPU under torch.distributed.run (CUDA IPC, stores over NVLink), skipped below 2 GPUs."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def fhe():
    import fhe_b200
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    fhe_b200.load_library()
    return fhe_b200


def _setup(fhe, preset, world, B, seed=7):
    from fhe_b200.engine import to_device
    from fhe_b200.params import bfv_preset
    p = bfv_preset(preset)
    ctxs = [fhe.BfvContext(p["n"], p["L"], p["R"], p["K"], p["dnum"], p["t"], p["primes"], p["sigma"], p["hamming_weight"])
            for _ in range(world)]
    g = ctxs[0]
    sk, pk = g.keygen(1, 2)
    rlk = g.relinkey_gen(3, sk)
    rng = np.random.default_rng(seed)
    m1 = rng.integers(0, p["t"], (B, p["n"]), dtype=np.uint64); m2 = rng.integers(0, p["t"], (B, p["n"]), dtype=np.uint64)
    ca = g.encrypt(10, to_device(m1), pk); cb = g.encrypt(100, to_device(m2), pk)
    shards = [fhe.BfvShard(ctxs[r], r, world, max_batch=B) for r in range(world)]
    handles = [s.handle() for s in shards]
    for s in shards:
        s.connect(handles)
    return p, ctxs, shards, rlk, ca, cb


def _run(shards, a, b, keys, streams=None, reps=1):
    """several ranks on ONE device: kernels that wait for one another must not be separate launches on one GPU (nothing guarantees
    that they run at the same time), so stage k is issued for every rank before stage k+1 for any rank, all on one stream --
    every wait finds its flags set.  The kernels, scatter tables, peer stores and epoch logic are those of the multi-GPU run."""
    outs = [torch.empty_like(x) for x in a]
    for _ in range(reps):
        for stage in range(5):
            for r, s in enumerate(shards):
                s.multiply_stage(stage, a[r], b[r], keys[r], outs[r])
    for s in shards:
        s.check()
    return outs


@pytest.mark.parametrize("world,fused", [(1, "0"), (2, "0"), (4, "0"), (2, "1")])
def test_virtual_ranks_equal_single_gpu(fhe, world, fused, monkeypatch):
    monkeypatch.setenv("FHE_B200_SHARD_TIMEOUT_MS", "3000")
    monkeypatch.setenv("FHE_B200_FUSED_TILE", fused)
    B = 2
    p, ctxs, shards, rlk, ca, cb = _setup(fhe, "mid", world, B)
    want = ctxs[0].multiply(ca, cb, rlk)
    want_sq = ctxs[0].multiply(ca, ca, rlk)
    torch.cuda.synchronize()
    streams = None
    keys = [s.slice_key(rlk) for s in shards]
    a = [s.shard_ct(ca) for s in shards]; b = [s.shard_ct(cb) for s in shards]
    torch.cuda.synchronize()
    assert sum(s.key_limb_count for s in shards) == p["L"] + p["K"] and sum(s.ext_limb_count for s in shards) == p["L"] + p["R"]
    outs = _run(shards, a, b, keys, streams, reps=3)        # back to back without a barrier: buffer-reuse ordering
    got = torch.cat(outs, dim=-1)
    assert torch.equal(got, want)
    outs = _run(shards, a, a, keys, streams)                # squaring path (same tensor for both operands)
    assert torch.equal(torch.cat(outs, dim=-1), want_sq)
    # one ciphertext of the batch only (batch below max_batch)
    outs = _run(shards, [x[:1].contiguous() for x in a], [x[:1].contiguous() for x in b], keys, streams)
    assert torch.equal(torch.cat(outs, dim=-1), want[:1])
    for s in shards:
        s.close()


def test_sharded_rejects_bad_arguments(fhe):
    from fhe_b200.params import bfv_preset
    p = bfv_preset("mid")
    g = fhe.BfvContext(p["n"], p["L"], p["R"], p["K"], p["dnum"], p["t"], p["primes"], p["sigma"], p["hamming_weight"])
    with pytest.raises(fhe.FheB200Error):
        fhe.BfvShard(g, 0, 3)                   # not a power of two
    with pytest.raises(fhe.FheB200Error):
        fhe.BfvShard(g, 2, 2)                   # rank outside the group
    with pytest.raises(fhe.FheB200Error):
        fhe.BfvShard(g, 0, 16)                  # more ranks than key limbs
    s = fhe.BfvShard(g, 0, 1)
    x = torch.zeros((1, 2, p["L"], p["n"]), dtype=torch.int64, device="cuda")
    k = torch.zeros((p["dnum"], 2, p["L"] + p["K"], p["n"]), dtype=torch.int64, device="cuda")
    with pytest.raises(fhe.FheB200Error):
        s.multiply(x, x, k)                     # not connected
    s.connect([s.handle()])
    with pytest.raises(fhe.FheB200Error):
        s.multiply(torch.cat([x, x]), torch.cat([x, x]), k)        # batch above max_batch
    c2 = bfv_preset("c2")
    g2 = fhe.BfvContext(c2["n"], c2["L"], c2["R"], c2["K"], c2["dnum"], c2["t"], c2["primes"], c2["sigma"], c2["hamming_weight"])
    with pytest.raises(fhe.FheB200Error):
        fhe.BfvShard(g2, 0, 1)                  # N = 4096 has no two-pass transform to scatter from


def test_sharded_timeout_is_reported(fhe, monkeypatch):
    """a rank whose peer never issues the operation gives up and says so (instead of hanging the GPU)"""
    monkeypatch.setenv("FHE_B200_SHARD_TIMEOUT_MS", "200")
    p, ctxs, shards, rlk, ca, cb = _setup(fhe, "mid", 2, 1)
    s = shards[0]
    k = s.slice_key(rlk)
    s.multiply(s.shard_ct(ca), s.shard_ct(cb), k)           # rank 1 never runs
    with pytest.raises(fhe.FheB200Error, match="timed out"):
        s.check()


def test_config4_world1_equals_plain(fhe):
    """the sharded code path at full config-4 size on one GPU (world = 1: every store is local)"""
    p, ctxs, shards, rlk, ca, cb = _setup(fhe, "c4", 1, 1)
    want = ctxs[0].multiply(ca, cb, rlk)
    s = shards[0]
    out = s.multiply(s.shard_ct(ca), s.shard_ct(cb), s.slice_key(rlk))
    s.check()
    assert torch.equal(out, want)


@pytest.mark.parametrize("preset,batch", [("mid", 2), ("c4", 1)])
def test_processes_one_rank_per_gpu(fhe, preset, batch):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (gpurun --gpus 2)")
    world = 1 << (n.bit_length() - 1)
    if preset == "mid":
        world = min(world, 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29571", os.path.join(ROOT, "tests", "sharded_worker.py"), "--preset", preset, "--batch", str(batch)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "bit-exact" in r.stdout
