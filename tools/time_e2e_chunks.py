import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from fastc_b200 import ECompressionFormat as F, lib
from fastc_b200.synth import synth_rgba_torch
g = lib()
size = 8192
d = synth_rgba_torch(size, size, 1, device="cuda")
pin = torch.empty(d.shape, dtype=torch.uint8, pin_memory=True); pin.copy_(d); torch.cuda.synchronize()
out = torch.empty((size // 4) ** 2 * 16, dtype=torch.uint8, pin_memory=True).numpy()
for name, cb in (("auto", 0), ("single", 1 << 30), ("auto", 0), ("half", (size // 4) ** 2 // 2)):
    for _ in range(3): g.compress(F.BPTC, pin.numpy(), out, quality=50, seed=1, chunk_blocks=cb)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); _, tm = g.compress(F.BPTC, pin.numpy(), out, quality=50, seed=1, chunk_blocks=cb); ts.append((time.perf_counter() - t0) * 1e3)
    print(name, [round(t, 1) for t in ts], "kernel_ms", round(tm["kernel_ms"], 1), "launches", tm["kernel_launches"])
