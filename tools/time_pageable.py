#!/usr/bin/env python3
"""Pinned vs pageable host memory through the C ABI (fastc_gpu_compress / _batch): what the staging of
pageable buffers (capi.cu: upload / download) costs.  Prints one JSON line per case."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from fastc_b200 import ECompressionFormat as F, lib  # noqa: E402
from fastc_b200.synth import synth_rgba_torch  # noqa: E402


def best(fn, reps=4):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts)


def main():
    g = lib()
    pin = lambda t: torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t)
    for fmt, size, q in (("DXT1", 8192, 0), ("ETC1", 8192, 0), ("DXT1", 4096, 0), ("BPTC", 2048, 50)):
        d = synth_rgba_torch(size, size, 1, opaque=(fmt == "ETC1"))
        hp = pin(d).numpy(); hg = np.array(hp)
        nb = (size // 4) ** 2 * (8 if fmt != "BPTC" else 16)
        op = pin(torch.zeros(nb, dtype=torch.uint8)).numpy(); og = np.empty_like(op)
        a = best(lambda: g.compress(F[fmt], hp, op, quality=q, seed=1))
        b = best(lambda: g.compress(F[fmt], hg, og, quality=q, seed=1))
        print(json.dumps({"case": f"{fmt} {size}^2", "pinned_ms": a, "pageable_ms": b, "ratio": b / a,
                          "equal": bool((op == og).all())}))
    n, size = 256, 1024
    texs = torch.stack([synth_rgba_torch(size, size, k + 1) for k in range(n)])
    hp = pin(texs).numpy(); hg = np.array(hp)
    for fmt, bsz in (("DXT1", 8), ("DXT5", 16)):
        op = pin(torch.zeros((n, (size // 4) ** 2 * bsz), dtype=torch.uint8)).numpy(); og = np.empty_like(op)
        a = best(lambda: g.compress_batch(F[fmt], [hp[k] for k in range(n)], outs=[op[k] for k in range(n)]))
        b = best(lambda: g.compress_batch(F[fmt], [hg[k] for k in range(n)], outs=[og[k] for k in range(n)]))
        print(json.dumps({"case": f"{fmt} batch {n} x {size}^2", "pinned_ms": a, "pageable_ms": b, "ratio": b / a,
                          "equal": bool((op == og).all())}))


if __name__ == "__main__":
    main()
