"""Block sharding across GPUs: every AES block is independent through all ten
rounds, so blocks are partitioned contiguously across ranks, keys are replicated per GPU, there is
no collective in the data path, and results are gathered once at the end.  The max mini-workload
reduces locally per rank and finishes with one tiny gather (16 LWE per rank)."""
import numpy as np


def block_range(nblocks, rank, world):
    """Contiguous [b0, b1) of `nblocks` owned by `rank` (same split the stage executable uses)."""
    return nblocks * rank // world, nblocks * (rank + 1) // world


def shard_ciphertext(ct_bytes, rank, world):
    nblocks = len(ct_bytes) // 16
    b0, b1 = block_range(nblocks, rank, world)
    return bytes(ct_bytes[16 * b0:16 * b1]), b0, b1


def gather_results(local, nblocks, rank, world, dist=None, device=None):
    """Gather per-rank [b1-b0][128][2049] uint64 results to rank 0 -> [nblocks][128][2049] (None elsewhere).

    `dist` is torch.distributed (gloo on CPU tests, nccl on GPUs); with world == 1 it is not used."""
    local = np.ascontiguousarray(local, dtype=np.uint64)
    if world == 1:
        return local
    import torch
    sizes = [block_range(nblocks, r, world) for r in range(world)]
    maxb = max(b1 - b0 for b0, b1 in sizes)
    buf = torch.zeros((maxb, 128, 2049), dtype=torch.int64, device=device)
    if local.shape[0]:
        buf[: local.shape[0]] = torch.from_numpy(local.view(np.int64)).to(buf.device)
    outs = [torch.zeros_like(buf) for _ in range(world)] if rank == 0 else None
    if dist.get_backend() == "nccl":
        allb = [torch.zeros_like(buf) for _ in range(world)]
        dist.all_gather(allb, buf)
        outs = allb if rank == 0 else None
    else:
        dist.gather(buf, outs, dst=0)
    if rank != 0:
        return None
    parts = [outs[r][: sizes[r][1] - sizes[r][0]].cpu().numpy().view(np.uint64) for r in range(world)]
    return np.concatenate(parts, axis=0)


def pair_range(npairs, rank, world):
    """Contiguous [p0, p1) of the inner product's `npairs` (x_i, y_i) pairs owned by `rank`."""
    return npairs * rank // world, npairs * (rank + 1) // world


def shard_inner_product_values(lwe_bits, rank, world):
    """Rows of [nvals*16][2049] bit ciphertexts this rank needs for its partial inner product: its slice of the
    first half (x) followed by the same slice of the second half (y), i.e. again a (first half, second half) input."""
    a = np.asarray(lwe_bits, dtype=np.uint64).reshape(-1, 16, 2049)
    npairs = a.shape[0] // 2
    p0, p1 = pair_range(npairs, rank, world)
    return np.concatenate([a[p0:p1], a[npairs + p0:npairs + p1]], axis=0).reshape(-1, 2049)


def value_range(nvals, rank, world):
    """Contiguous [v0, v1) of the max workload's values owned by `rank`."""
    return nvals * rank // world, nvals * (rank + 1) // world
