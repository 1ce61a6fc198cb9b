"""world_size-2 gloo test of the multi-GPU plumbing (block sharding + single final gather)."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, nblocks, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from temp_fhe_transciphering_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ct = bytes(range(256)) * ((nblocks * 16 + 255) // 256)
    ct = ct[: nblocks * 16]
    mine, b0, b1 = sharding.shard_ciphertext(ct, rank, world)
    assert len(mine) == 16 * (b1 - b0)
    # stand-in for the per-rank transcipher: result row = f(block bytes), so order errors show
    local = np.zeros((b1 - b0, 128, 2049), dtype=np.uint64)
    for i in range(b1 - b0):
        local[i, :, :] = np.uint64(int.from_bytes(mine[16 * i:16 * i + 8], "little"))
        local[i, 0, 0] = np.uint64(b0 + i)
    full = sharding.gather_results(local, nblocks, rank, world, dist)
    if rank == 0:
        ok = full.shape == (nblocks, 128, 2049) and all(int(full[b, 0, 0]) == b for b in range(nblocks))
        ok = ok and all(int(full[b, 5, 7]) == int.from_bytes(ct[16 * b:16 * b + 8], "little") for b in range(nblocks))
        q.put(bool(ok))
    else:
        assert full is None
    dist.destroy_process_group()


def test_block_sharding_and_gather_world2():
    from temp_fhe_transciphering_b200 import sharding
    assert [sharding.block_range(8, r, 2) for r in range(2)] == [(0, 4), (4, 8)]
    assert [sharding.block_range(5, r, 2) for r in range(2)] == [(0, 2), (2, 5)]
    assert [sharding.block_range(1, r, 4) for r in range(4)] == [(0, 0), (0, 0), (0, 0), (0, 1)]
    ctx = mp.get_context("spawn")
    for nblocks in (5, 8):
        q = ctx.Queue()
        procs = [ctx.Process(target=_worker, args=(r, 2, 29517 + nblocks, nblocks, q)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        assert q.get(timeout=5) is True
