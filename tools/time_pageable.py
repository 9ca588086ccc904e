import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from fastc_b200 import ECompressionFormat as F, lib
from fastc_b200.synth import synth_rgba_torch
g = lib()
for fmt, size, q in ((F.BPTC, 8192, 50), (F.DXT1, 8192, 0), (F.ETC1, 8192, 0)):
    d = synth_rgba_torch(size, size, 1, opaque=(fmt == F.ETC1), device="cuda")
    pin = torch.empty(d.shape, dtype=torch.uint8, pin_memory=True); pin.copy_(d); torch.cuda.synchronize()
    pag = d.cpu().numpy().copy()
    nb = (size // 4) ** 2 * (8 if fmt != F.BPTC else 16)
    out_pin = torch.empty(nb, dtype=torch.uint8, pin_memory=True).numpy()
    out_pag = np.zeros(nb, dtype=np.uint8)
    for name, i, o in (("pinned", pin.numpy(), out_pin), ("pageable", pag, out_pag)):
        for _ in range(2): g.compress(fmt, i, o, quality=q, seed=1)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); g.compress(fmt, i, o, quality=q, seed=1); ts.append((time.perf_counter() - t0) * 1e3)
        print(fmt.name, size, name, "ms", round(min(ts), 2), "Mpix/s", round(size * size / 1e6 / (min(ts) / 1e3), 1))
    assert (out_pin == out_pag).all()
