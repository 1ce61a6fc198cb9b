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


def _mini_worker(rank, world, port, nvals, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import temp_fhe_transciphering_b200 as cbs
    from temp_fhe_transciphering_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    vals = np.random.default_rng(nvals).integers(0, 65536, nvals, dtype=np.uint16)
    # stand-in ciphertexts: row [value*16 + bit] carries the cleartext bit in word 0, the value index in word 1
    rows = np.zeros((nvals * 16, 2049), dtype=np.uint64)
    for v in range(nvals):
        for i in range(16):
            rows[v * 16 + i, 0] = (int(vals[v]) >> (15 - i)) & 1
            rows[v * 16 + i, 1] = v

    def decode(r):
        r = r.reshape(-1, 16, 2049)
        return np.array([sum(int(r[v, i, 0]) << (15 - i) for i in range(16)) for v in range(r.shape[0])], dtype=np.uint16)

    # inner product: this rank's (x slice, y slice) -> partial through the same circuit plan the GPU executes
    mine = sharding.shard_inner_product_values(rows, rank, world)
    p0, p1 = sharding.pair_range(nvals // 2, rank, world)
    assert mine.shape[0] == 2 * (p1 - p0) * 16
    assert list(mine.reshape(-1, 16, 2049)[:, 0, 1]) == list(range(p0, p1)) + list(range(nvals // 2 + p0, nvals // 2 + p1))
    part_ip = cbs.inner_product_plan_check(decode(mine))[0] if p1 > p0 else 0
    v0, v1 = sharding.value_range(nvals, rank, world)
    part_max = int(vals[v0:v1].max()) if v1 > v0 else 0
    t = torch.tensor([part_ip, part_max], dtype=torch.int64)
    outs = [torch.zeros_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, outs, dst=0)
    if rank == 0:
        ips = np.array([int(o[0]) for o in outs], dtype=np.uint16)
        maxs = np.array([int(o[1]) for o in outs], dtype=np.uint16)
        h = nvals // 2
        want_ip = sum((int(x) * int(y)) % 65536 for x, y in zip(vals[:h], vals[h:])) % 65536
        # rank 0 finishes with the sum / max circuits over the partial results (cbs_sum_u16 / cbs_max_u16 on the GPU)
        q.put(bool(int(ips.astype(np.uint32).sum() % 65536) == want_ip and cbs.max_plan_check(maxs)[0] == int(vals.max())))
    dist.destroy_process_group()


def test_mini_workload_sharding_world2():
    ctx = mp.get_context("spawn")
    for nvals in (16, 50):
        q = ctx.Queue()
        procs = [ctx.Process(target=_mini_worker, args=(r, 2, 29617 + nvals, nvals, q)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        assert q.get(timeout=5) is True
