"""Host-side logic of the limb-sharded multiply that needs no GPU: the block partition of limbs over ranks (C ABI, csrc/shard.cu)
and, under gloo with world size 2, the rank-ordered exchange of the 128-byte handles that `BfvShard.connect()` performs."""
import os
import socket

import pytest

import fhe_b200
from fhe_b200.engine import SHARD_HANDLE_BYTES, shard_partition


@pytest.mark.parametrize("total,world", [(49, 8), (32, 8), (49, 2), (13, 4), (9, 4), (6, 16), (1, 1), (24, 16)])
def test_partition_is_contiguous_balanced_and_complete(total, world):
    pos = 0
    sizes = []
    for r in range(world):
        b, c = shard_partition(total, r, world)
        assert b == pos
        pos += c
        sizes.append(c)
    assert pos == total and max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def test_partition_rejects_bad_ranks():
    with pytest.raises(fhe_b200.FheB200Error):
        shard_partition(10, 4, 4)
    with pytest.raises(fhe_b200.FheB200Error):
        shard_partition(10, 0, 17)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = bytes([rank]) * SHARD_HANDLE_BYTES                # stands in for fhe_b200_shard_handle's blob
    handles = [None] * world
    dist.all_gather_object(handles, mine)                    # what BfvShard.connect() does before fhe_b200_shard_connect
    ok = all(len(h) == SHARD_HANDLE_BYTES and h[0] == r for r, h in enumerate(handles))
    limbs = [shard_partition(49, r, world) for r in range(world)]
    q.put((rank, ok, limbs))
    dist.destroy_process_group()


def test_handle_exchange_in_rank_order_gloo_world2():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
    assert [r[1] for r in res] == [True, True]
    assert res[0][2] == res[1][2] == [(0, 25), (25, 24)]
