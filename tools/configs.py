#!/usr/bin/env python3
"""One JSON line per BASELINE config that `bench.py` does not cover (bench.py is configs[2]):

  config 1  BPTC -q 0, 256x256: GPU vs the reference (-t 1), identical blocks, mode histogram
  config 2  BPTC -q 50, 2048x2048: GPU (device-resident and C ABI) vs the reference with every host
            thread (static split and -j 64), PSNR of both by the reference's decoder / formula,
            bit-identical block fraction
  config 5  ETC1, 4096x4096 RGB: kernel-only and C-ABI time on 1..N GPUs of this box (one process,
            block-row slabs per GPU), identical to the single-GPU bytes

usage: python tools/configs.py [--gpus N] [--skip-reference] [--only 1 2 5]
The reference arm needs oracle/_ref/libfastc_ref.so (built where /root/reference exists; it travels).
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from fastc_b200 import ECompressionFormat as F, lib  # noqa: E402
from fastc_b200.synth import synth_rgba_torch  # noqa: E402
from _checkers import Reference  # noqa: E402


def pinned(t):
    p = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    p.copy_(t)
    torch.cuda.synchronize()
    return p.numpy()


def device_ms(g, fmt, d_in, d_out, size, reps=5, **kw):
    for _ in range(2):
        g.compress_device(fmt, d_in, d_out, width=size, height=size, **kw)
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.compress_device(fmt, d_in, d_out, width=size, height=size, **kw)
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return min(ms)


def host_ms(g, fmt, h_in, h_out, reps=5, **kw):
    for _ in range(2):
        g.compress(fmt, h_in, h_out, **kw)
    ms = []
    for _ in range(reps):
        t0 = time.perf_counter()
        g.compress(fmt, h_in, h_out, **kw)
        ms.append((time.perf_counter() - t0) * 1e3)
    return min(ms)


def mode_histogram(blocks):
    first = [int(b) for b in blocks.reshape(-1, 16)[:, 0]]
    modes = np.array([((b & -b).bit_length() - 1) if b else 8 for b in first])  # unary prefix: mode = index of the lowest set bit
    return [int((modes == m).sum()) for m in range(9)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--skip-reference", action="store_true")
    ap.add_argument("--only", type=int, nargs="*", default=[1, 2, 5], help="configs to run")
    args = ap.parse_args()
    g = lib()
    ref = Reference() if (Reference.available() and not args.skip_reference) else None
    threads = min(os.cpu_count() or 1, 256)

    if 1 in args.only:
        config1(g, ref)
    if 2 in args.only:
        config2(g, ref, threads)
    if 5 in args.only:
        config5(g, ref, threads, args.gpus)


def config1(g, ref):
    size = 256
    d_in = synth_rgba_torch(size, size, 1, device="cuda")
    img = np.ascontiguousarray(d_in.cpu().numpy())
    got, _ = g.compress(F.BPTC, img, quality=0)
    line = {"config": 1, "workload": "BPTC -q 0, 256x256 RGBA", "mode_histogram_0_7_other": mode_histogram(got)}
    if ref:
        want, ms = ref.compress("BPTC", img, quality=0, threads=1, seed=1)
        line.update({"reference_ms_t1": ms, "bit_identical_blocks": int((got.reshape(-1, 16) == want.reshape(-1, 16)).all(1).sum()),
                     "blocks": got.size // 16})
    print(json.dumps(line), flush=True)


def config2(g, ref, threads):
    size = 2048
    d_in = synth_rgba_torch(size, size, 1, device="cuda")
    d_out = torch.zeros((size // 4) ** 2 * 16, dtype=torch.uint8, device="cuda")
    k_ms = device_ms(g, F.BPTC, d_in, d_out, size, quality=50, seed=1)
    h_in = pinned(d_in)
    h_out = torch.empty(d_out.numel(), dtype=torch.uint8, pin_memory=True).numpy()
    e_ms = host_ms(g, F.BPTC, h_in, h_out, quality=50, seed=1)
    mpix = size * size / 1e6
    line = {"config": 2, "workload": "BPTC -q 50, 2048x2048 RGBA", "gpu_device_ms": k_ms, "gpu_device_mpix_s": mpix / (k_ms / 1e3),
            "gpu_e2e_ms": e_ms, "gpu_e2e_mpix_s": mpix / (e_ms / 1e3)}
    if ref:
        want, ms_static = ref.compress("BPTC", h_in, quality=50, threads=threads, seed=None)
        _, ms_j64 = ref.compress("BPTC", h_in, quality=50, threads=threads, job_size=64, seed=None)
        p_ref = ref.psnr(h_in, ref.decode("BPTC", want, size, size))
        p_gpu = ref.psnr(h_in, ref.decode("BPTC", h_out, size, size))
        line.update({"reference_threads": threads, "reference_ms_static_split": ms_static, "reference_ms_j64": ms_j64,
                     "reference_mpix_s": mpix / (min(ms_static, ms_j64) / 1e3),
                     "speedup_e2e_vs_reference": min(ms_static, ms_j64) / e_ms,
                     "psnr_gpu_db": p_gpu, "psnr_ref_db": p_ref, "delta_db": p_gpu - p_ref,
                     "bit_identical_block_fraction": float((h_out.reshape(-1, 16) == want.reshape(-1, 16)).all(1).mean())})
    print(json.dumps(line), flush=True)


def config5(g, ref, threads, gpus):
    size = 4096
    d_in = synth_rgba_torch(size, size, 1, opaque=True, device="cuda")
    d_out = torch.zeros((size // 4) ** 2 * 8, dtype=torch.uint8, device="cuda")
    k_ms = device_ms(g, F.ETC1, d_in, d_out, size)
    h_in = pinned(d_in)
    outs = {}
    mpix = size * size / 1e6
    line = {"config": 5, "workload": "ETC1, 4096x4096 RGB", "gpu_kernel_ms": k_ms, "gpu_kernel_mpix_s": mpix / (k_ms / 1e3), "e2e": []}
    n = 1
    while n <= gpus:
        h_out = torch.empty(d_out.numel(), dtype=torch.uint8, pin_memory=True).numpy()
        e_ms = host_ms(g, F.ETC1, h_in, h_out, num_gpus=n)
        outs[n] = h_out.copy()
        line["e2e"].append({"gpus": n, "ms": e_ms, "mpix_s": mpix / (e_ms / 1e3), "identical_to_1_gpu": bool((outs[n] == outs[1]).all())})
        n *= 2
    if ref:
        rows = 512  # bounded slab of the same texture
        want, ms = ref.compress("ETC1", np.ascontiguousarray(h_in[:rows]), threads=threads, seed=None)
        line.update({"reference_threads": threads, "reference_mpix_s": size * rows / 1e6 / (ms / 1e3),
                     "reference_sample": f"top 4096x{rows} slab", "bit_identical_on_sample": bool((outs[1][:want.size] == want).all())})
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
