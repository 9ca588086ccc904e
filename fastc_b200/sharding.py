"""Multi-process sharding of one texture over ranks (one process per GPU).

The reference shards a job's raster block range over <= 256 pthreads
(reference Core/src/ThreadGroup.cpp:133-192: ceil(nBlocks / nThreads) contiguous blocks per
thread).  Across GPUs the same idea is applied to whole BLOCK ROWS, so each rank's input is
one contiguous slab of the image and its output one contiguous byte range.
Two things cross ranks, both tiny next to the encode:
  * the BC7 watermark chain: the word a solid-colour block carries is the number of solid
    blocks before it in raster order over the WHOLE texture (reference
    BPTCEncoder/src/Compressor.cpp:135-140,1457) -> one all-gather of per-rank counts and an
    exclusive prefix sum;
  * the gather of the compressed slabs to rank 0 (NCCL over NVLink on GPUs; gloo in the CPU
    tests).
torch.distributed is only the transport here; nothing in this module touches pixels.
"""
from __future__ import annotations


def shard_block_rows(block_rows: int, rank: int, world: int) -> tuple[int, int]:
    """Half-open range of block rows owned by `rank`: contiguous, disjoint, covering
    [0, block_rows), sizes differing by at most one row."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return block_rows * rank // world, block_rows * (rank + 1) // world


def slab_geometry(width: int, height: int, rank: int, world: int, block_bytes: int) -> dict:
    """Everything a rank needs to encode its slab with the device API: pixel rows, block
    count, the raster index of its first block (keys the per-block RNG streams) and its
    byte range in the gathered output."""
    if width % 4 or height % 4:
        raise ValueError("image dimensions must be multiples of the 4x4 block")
    bx = width // 4
    r0, r1 = shard_block_rows(height // 4, rank, world)
    return {"row0": r0 * 4, "rows": (r1 - r0) * 4, "num_blocks": (r1 - r0) * bx, "block_index_base": r0 * bx,
            "out_offset": r0 * bx * block_bytes, "out_bytes": (r1 - r0) * bx * block_bytes}


def watermark_base(my_solid_blocks: int, rank: int, world: int, device=None, group=None) -> int:
    """Number of solid-colour blocks owned by lower ranks (exclusive prefix over ranks)."""
    if world == 1:
        return 0
    import torch
    import torch.distributed as dist
    counts = torch.zeros(world, dtype=torch.int64, device=device)
    counts[rank] = int(my_solid_blocks)
    dist.all_reduce(counts, group=group)  # an all-gather of one integer per rank
    return int(counts[:rank].sum().item())


def gather_slabs(local, rank: int, world: int, sizes: list[int], gather_list=None, group=None):
    """Gathers every rank's compressed slab (uint8 tensors of the given byte sizes, which may
    differ by one block row) on rank 0.  Returns the list of slabs on rank 0, None elsewhere.
    `gather_list` lets the caller pre-allocate the receive buffers (bench.py does)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return [local]
    if rank == 0:
        if gather_list is None:
            gather_list = [torch.empty(s, dtype=torch.uint8, device=local.device) for s in sizes]
        gather_list[0].copy_(local)
        # one batch: the receives are independent transfers (NCCL runs them as one group; issued one by
        # one they are serialised on the process group: 7 x 8 MB took 0.44 ms of a 24 ms step at N = 8)
        ops = [dist.P2POp(dist.irecv, gather_list[r], r, group) for r in range(1, world)]
        for q in dist.batch_isend_irecv(ops):
            q.wait()
        return gather_list
    for q in dist.batch_isend_irecv([dist.P2POp(dist.isend, local, 0, group)]):
        q.wait()
    return None
