#!/usr/bin/env python3
"""End-to-end time of fastc_gpu_compress(BPTC -q 50, pinned host buffers) on an 8192-wide slab of the bench
texture for different host-side chunkings: auto (capi.cu plan_chunks), one chunk, two halves.
usage: time_e2e_chunks.py [rows ...]   (default: 8192 4096 2048 1024 = the slabs of 1 / 2 / 4 / 8 GPUs)"""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from fastc_b200 import ECompressionFormat as F, lib  # noqa: E402
from fastc_b200.synth import synth_rgba_torch  # noqa: E402

g = lib()
W = 8192
for rows in [int(a) for a in sys.argv[1:]] or [8192, 4096, 2048, 1024]:
    d = synth_rgba_torch(W, rows, 1, full_height=8192, device="cuda")
    pin = torch.empty(d.shape, dtype=torch.uint8, pin_memory=True)
    pin.copy_(d)
    torch.cuda.synchronize()
    nblk = (W // 4) * (rows // 4)
    out = torch.empty(nblk * 16, dtype=torch.uint8, pin_memory=True).numpy()
    for name, cb in (("auto", 0), ("single", 1 << 30), ("half", nblk // 2), ("auto", 0)):
        for _ in range(3):
            g.compress(F.BPTC, pin.numpy(), out, quality=50, seed=1, chunk_blocks=cb)
        ts = []
        for _ in range(4):
            t0 = time.perf_counter()
            _, tm = g.compress(F.BPTC, pin.numpy(), out, quality=50, seed=1, chunk_blocks=cb)
            ts.append((time.perf_counter() - t0) * 1e3)
        print(f"rows {rows:5d} {name:6s} ms {min(ts):8.2f} (all {[round(t, 1) for t in ts]}) kernel_ms "
              f"{tm['kernel_ms']:.1f} launches {tm['kernel_launches']}", flush=True)
    del d, pin, out
