#!/usr/bin/env python3
"""Measured work of the GPU's own annealing schedule ("the kernel must export its own PBE counter"):
runs ONE BC7 compression of the bench texture with a -DFASTC_GPU_COUNTERS build of the library and writes
profiles/r02_work_counters.json (read by tools/collect_counts.py -> `roofline.work_counters` of the bench line).

    tools/build_variant.sh counters -DFASTC_GPU_COUNTERS          # on the CPU box; travels with the snapshot
    python tools/measure_counters.py variants/counters.so [--size 8192] [--quality 50]     # on the GPU box

The counters: QuantizedError calls (the first evaluation of every fitted chain in bc7_setup + every annealing
step of bc7_anneal / bc7_anneal_tail, the tail's discarded speculative evaluations included) and pixels
evaluated by the annealing steps (each against its two candidate buckets: the reference's "pixel-bucket
evaluations" are 1.9 per pixel)."""
import argparse
import ctypes as C
import json
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "fastc_b200" / "libfastc_gpu.so"

CHILD = r"""
import ctypes as C, json, sys, torch
sys.path.insert(0, %r)
from fastc_b200 import ECompressionFormat as F, lib
from fastc_b200.synth import synth_rgba_torch
size, q = int(sys.argv[1]), int(sys.argv[2])
g = lib()
img = synth_rgba_torch(size, size, 1)
out = torch.zeros((size // 4) ** 2 * 16, dtype=torch.uint8, device="cuda")
g.compress_device(F.BPTC, img, out, width=size, height=size, quality=q, seed=1)
torch.cuda.synchronize()
a, b = C.c_uint64(0), C.c_uint64(0)
assert g.cdll.fastc_gpu_bc7_counters(C.byref(a), C.byref(b)) == 0
print(json.dumps({"qe_calls": a.value, "pixels_evaluated": b.value}))
""" % str(ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("lib")
    ap.add_argument("--size", type=int, default=8192)
    ap.add_argument("--quality", type=int, default=50)
    a = ap.parse_args()
    backup = LIB.with_suffix(".so.orig")
    shutil.copy2(LIB, backup)
    try:
        shutil.copy2(a.lib, LIB)
        r = subprocess.run([sys.executable, "-c", CHILD, str(a.size), str(a.quality)], capture_output=True, text=True)
    finally:
        shutil.copy2(backup, LIB)
        backup.unlink()
    if r.returncode != 0:
        sys.exit(r.stderr[-800:])
    d = json.loads(r.stdout.strip().splitlines()[-1])
    blocks = (a.size // 4) ** 2
    d.update({
        "what": f"FASTC_GPU_COUNTERS build, one BC7 -q {a.quality} compression of the synthetic {a.size}^2 texture (seed 1); "
                "last chunk of the submission (the counters are reset per chunk; one chunk up to 8192^2)",
        "blocks": blocks,
        "qe_calls_per_block": d["qe_calls"] / blocks,
        "pixels_evaluated_per_block": d["pixels_evaluated"] / blocks,
        "pixel_bucket_evals_per_block": 2.0 * d["pixels_evaluated"] / blocks,
        "pixels_per_call": d["pixels_evaluated"] / max(d["qe_calls"], 1),
        "reference_model": "SURVEY 8d: 1,307 QE calls / block x 11.0 px x 1.91 buckets = 27.4 k PBE / block at -q 50",
    })
    dst = ROOT / "gpurun_out" / "r02_work_counters.json"
    dst.parent.mkdir(exist_ok=True)
    dst.write_text(json.dumps(d, indent=1) + "\n")
    print(json.dumps(d))


if __name__ == "__main__":
    main()
