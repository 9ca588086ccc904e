#!/usr/bin/env python3
"""Bytes-equality of library builds: for each given libfastc_gpu.so variant, copy it over
fastc_b200/libfastc_gpu.so and hash the BC7 output of the bench texture in a fresh process.
    python tools/compare_libs.py variants/a.so variants/b.so --size 2048 --quality 50"""
import argparse
import hashlib
import shutil
import subprocess
import sys
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
img = synth_rgba_torch(size, size, 1)
out = torch.zeros((size // 4) ** 2 * 16, dtype=torch.uint8, device="cuda")
g.compress_device(F.BPTC, img, out, width=size, height=size, quality=q, seed=1)
torch.cuda.synchronize()
print(hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest())
""" % str(ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("libs", nargs="+")
    ap.add_argument("--size", type=int, nargs="+", default=[2048])
    ap.add_argument("--quality", type=int, nargs="+", default=[50])
    a = ap.parse_args()
    backup = LIB.with_suffix(".so.orig")
    shutil.copy2(LIB, backup)
    ok = True
    try:
        for size in a.size:
            for q in a.quality:
                hashes = []
                for lib in a.libs:
                    shutil.copy2(lib, LIB)
                    r = subprocess.run([sys.executable, "-c", CHILD, str(size), str(q)], capture_output=True, text=True)
                    h = r.stdout.strip().splitlines()[-1] if r.returncode == 0 and r.stdout.strip() else "FAILED " + r.stderr[-300:]
                    hashes.append(h)
                    print(f"size {size} q {q} {Path(lib).name:24s} {h}", flush=True)
                same = len(set(hashes)) == 1
                ok = ok and same
                print(f"size {size} q {q}: {'IDENTICAL' if same else 'DIFFERENT'}", flush=True)
    finally:
        shutil.copy2(backup, LIB)
        backup.unlink()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
