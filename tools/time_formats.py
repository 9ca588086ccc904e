#!/usr/bin/env python3
"""Device-resident kernel timing of the streaming formats (DXT1 / DXT5 / ETC1) at the
BASELINE config sizes: CUDA events on the launching stream, L2 flushed between
launches.  Prints one JSON line per format (Mpix/s, algorithmic GB/s vs the measured
HBM peak)."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from fastc_b200 import ECompressionFormat as F, lib  # noqa: E402
from fastc_b200.synth import synth_rgba_torch  # noqa: E402


def main():
    g = lib()
    peaks = ROOT / "MEASURED_PEAKS.json"
    hbm = json.loads(peaks.read_text())["hbm_gbs"] if peaks.exists() else 6650.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for fmt, size, bpb in (("DXT1", 4096, 8), ("DXT5", 4096, 16), ("ETC1", 4096, 8)):
        img = synth_rgba_torch(size, size, 1, opaque=(fmt == "ETC1"))
        nblk = (size // 4) ** 2
        out = torch.zeros(nblk * bpb, dtype=torch.uint8, device="cuda")
        for _ in range(3):
            g.compress_device(F[fmt], img, out, width=size, height=size)
        torch.cuda.synchronize()
        ms = []
        for k in range(10):
            flush.fill_(k)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            g.compress_device(F[fmt], img, out, width=size, height=size)
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        t = sorted(ms)[len(ms) // 2]
        gbs = nblk * (64 + bpb) / (t / 1e3) / 1e9
        print(json.dumps({"format": fmt, "size": size, "ms": t, "mpix_s": size * size / 1e6 / (t / 1e3),
                          "algo_gbs": gbs, "hbm_peak_gbs": hbm, "hbm_frac": gbs / hbm}))
    if "--quick" not in sys.argv:
        batch_config4(g, hbm)


def batch_config4(g, hbm):
    """BASELINE configs[3]: 256 textures of 1024^2 in one batch submission (host buffers, PCIe inside
    the timed region) next to the kernel-only time of the same 256 launches on resident data."""
    import numpy as np
    import time
    n, size = 256, 1024
    dev = [synth_rgba_torch(size, size, seed, device="cuda") for seed in range(1, n + 1)]
    pinned = [torch.empty((size, size, 4), dtype=torch.uint8, pin_memory=True) for _ in range(n)]
    for p_, t in zip(pinned, dev):
        p_.copy_(t)
    torch.cuda.synchronize()
    host = [p_.numpy() for p_ in pinned]
    for fmt, bpb in (("DXT1", 8), ("DXT5", 16)):
        nblk = (size // 4) ** 2
        out = torch.zeros(nblk * bpb, dtype=torch.uint8, device="cuda")
        for t in dev[:3]:
            g.compress_device(F[fmt], t, out, width=size, height=size)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for t in dev:   # 1 GiB of input: larger than L2
            g.compress_device(F[fmt], t, out, width=size, height=size)
        b.record()
        torch.cuda.synchronize()
        k_ms = a.elapsed_time(b)
        pouts = [torch.empty(nblk * bpb, dtype=torch.uint8, pin_memory=True).numpy() for _ in range(n)]
        g.compress_batch(F[fmt], host[:8], outs=pouts[:8])
        t0 = time.perf_counter()
        _, tm = g.compress_batch(F[fmt], host, outs=pouts)
        e_ms = (time.perf_counter() - t0) * 1e3
        mpix = n * size * size / 1e6
        gbs = n * nblk * (64 + bpb) / (k_ms / 1e3) / 1e9
        print(json.dumps({"config": "batch of 256 x 1024^2", "format": fmt, "kernel_ms": k_ms,
                          "kernel_mpix_s": mpix / (k_ms / 1e3), "algo_gbs": gbs, "hbm_peak_gbs": hbm,
                          "hbm_frac": gbs / hbm, "e2e_ms": e_ms, "e2e_mpix_s": mpix / (e_ms / 1e3),
                          "e2e_pcie_gbs": (tm["h2d_bytes"] + tm["d2h_bytes"]) / (e_ms / 1e3) / 1e9,
                          "host_memory": "pinned"}))


if __name__ == "__main__":
    main()
