#!/usr/bin/env python3
"""Short single-GPU workload for ncu: `python tools/profile_run.py FORMAT SIZE QUALITY REPS`
runs REPS device-resident compressions of a SIZE^2 synthetic texture (the bench generator)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from fastc_b200 import ECompressionFormat as F, lib  # noqa: E402
from fastc_b200.synth import synth_rgba_torch  # noqa: E402

fmt = sys.argv[1] if len(sys.argv) > 1 else "BPTC"
size = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
q = int(sys.argv[3]) if len(sys.argv) > 3 else 50
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
g = lib()
img = synth_rgba_torch(size, size, 1, opaque=(fmt == "ETC1"))
out = torch.zeros((size // 4) ** 2 * 16, dtype=torch.uint8, device="cuda")
for _ in range(reps):
    g.compress_device(F[fmt], img, out, width=size, height=size, quality=q, seed=1)
torch.cuda.synchronize()
print("done", fmt, size, q, reps)
