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


if __name__ == "__main__":
    main()
