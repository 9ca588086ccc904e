"""CPU, world_size 2 and 3 over gloo: the N > 1 host logic of bench.py / fastc_b200.sharding
(block-row slabs, watermark chain across ranks, gather of compressed slabs).  Each rank
"encodes" its slab with the CPU oracle -- keyed RNG, wm_base, block_index_base exactly as the
GPU ranks pass them to fastc_gpu_compress_device -- and the gathered bytes must equal a
single-process encode of the whole texture."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fastc_b200.sharding import gather_slabs, shard_block_rows, slab_geometry, watermark_base
from fastc_b200.synth import synth_rgba

W, H, Q, SEED = 128, 192, 2, 77


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _image():
    # window of the BASELINE generator with solid tiles on both sides of the rank boundaries
    return np.ascontiguousarray(synth_rgba(512, 512, 1)[16:16 + H, 160:160 + W])


def _solid_count(img):
    h, w = img.shape[:2]
    b = img.reshape(h // 4, 4, w // 4, 4, 4).transpose(0, 2, 1, 3, 4).reshape(-1, 16, 4)
    return int((b == b[:, :1]).all((1, 2)).sum())


def _worker(rank, world, port, fmt, out_path):
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent))
    import ctypes as C
    from _checkers import BLOCK_BYTES, Oracle, _p
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        o = Oracle()
        geo = slab_geometry(W, H, rank, world, BLOCK_BYTES[fmt])
        slab = np.ascontiguousarray(_image()[geo["row0"]:geo["row0"] + geo["rows"]])
        wm = watermark_base(_solid_count(slab), rank, world)
        out = np.zeros(geo["out_bytes"], dtype=np.uint8)
        if fmt == "BPTC":
            o.lib.fastc_oracle_bc7_keyed.argtypes = [C.POINTER(C.c_uint8), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                                     C.POINTER(C.c_uint8), C.c_int, C.c_uint64, C.c_uint32, C.c_uint32]
            o.lib.fastc_oracle_bc7_keyed.restype = None
            o.lib.fastc_oracle_bc7_keyed(_p(slab), W, geo["rows"], 0, geo["num_blocks"], _p(out), Q, SEED, wm,
                                         geo["block_index_base"])
        else:
            out, _ = o.compress(fmt, slab)
        sizes = [slab_geometry(W, H, r, world, BLOCK_BYTES[fmt])["out_bytes"] for r in range(world)]
        got = gather_slabs(torch.from_numpy(out), rank, world, sizes)
        if rank == 0:
            np.save(out_path, torch.cat(got).numpy())
        else:
            assert got is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_shard_block_rows_partition():
    for rows in (1, 2, 7, 48, 2048):
        for world in (1, 2, 3, 8):
            spans = [shard_block_rows(rows, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == rows
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_block_rows(8, 2, 2)
    g = slab_geometry(8192, 8192, 3, 8, 16)
    assert g == {"row0": 3072, "rows": 1024, "num_blocks": 524288, "block_index_base": 3 * 524288,
                 "out_offset": 3 * 8388608, "out_bytes": 8388608}


@pytest.mark.parametrize("world,fmt", [(2, "BPTC"), (3, "BPTC"), (2, "DXT5"), (2, "ETC1")])
def test_sharded_encode_equals_single_process(oracle, tmp_path, world, fmt):
    out_path = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(world, _free_port(), fmt, out_path), nprocs=world, join=True)
    got = np.load(out_path)
    img = _image()
    assert _solid_count(img) >= 16  # the watermark chain is actually exercised
    want, _ = oracle.compress(fmt, img, quality=Q, rng_mode=1, seed=SEED)
    assert got.size == want.size and (got == want).all()
