#!/usr/bin/env python3
"""A/B of library builds on ETC1 (device-resident, one quality level): output hash + kernel time per variant.
    python tools/ab_etc1.py variants/a.so variants/b.so --size 4096 --quality 2"""
import argparse, hashlib, shutil, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "fastc_b200" / "libfastc_gpu.so"
CHILD = r"""
import sys, hashlib, torch
sys.path.insert(0, %r)
from fastc_b200 import ECompressionFormat as F, lib
from fastc_b200.synth import synth_rgba_torch
size, q = int(sys.argv[1]), int(sys.argv[2])
g = lib()
img = synth_rgba_torch(size, size, 1, opaque=True)
out = torch.zeros((size // 4) ** 2 * 8, dtype=torch.uint8, device="cuda")
g.compress_device(F.ETC1, img, out, width=size, height=size, etc1_quality=q)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    g.compress_device(F.ETC1, img, out, width=size, height=size, etc1_quality=q)
b.record(); torch.cuda.synchronize()
print(hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest()[:16], a.elapsed_time(b) / 3)
""" % str(ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("libs", nargs="+"); ap.add_argument("--size", type=int, default=4096); ap.add_argument("--quality", type=int, default=2)
a = ap.parse_args()
backup = LIB.with_suffix(".so.orig"); shutil.copy2(LIB, backup)
try:
    for lib in a.libs:
        shutil.copy2(lib, LIB)
        r = subprocess.run([sys.executable, "-c", CHILD, str(a.size), str(a.quality)], capture_output=True, text=True)
        print(f"{Path(lib).name:16s} q{a.quality} {a.size}^2: {r.stdout.strip() or r.stderr[-300:]}", flush=True)
finally:
    shutil.copy2(backup, LIB); backup.unlink()
