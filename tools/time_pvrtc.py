import sys, time, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from fastc_b200 import ECompressionFormat as F, lib
from fastc_b200.synth import synth_rgba
from _checkers import Reference
g = lib(); r = Reference()
for size in (256, 1024, 2048):
    img = synth_rgba(size, size, 1)
    d_in = torch.from_numpy(img).cuda(); d_out = torch.zeros((size//4)**2*8, dtype=torch.uint8, device='cuda')
    for _ in range(2): g.compress_device(F.PVRTC4, d_in, d_out, width=size, height=size)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.compress_device(F.PVRTC4, d_in, d_out, width=size, height=size); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    t0 = time.perf_counter(); want, rms = r.compress("PVRTC4", img, seed=None); 
    print(json.dumps({"size": size, "gpu_ms": ms, "ref_ms": rms, "equal": bool(np.array_equal(d_out.cpu().numpy(), want))}))

# a batch: the textures are encoded side by side (grid.y = texture)
for size, n in ((512, 64), (1024, 32)):
    imgs = [synth_rgba(size, size, k + 1) for k in range(n)]
    g.compress_batch(F.PVRTC4, imgs[:2])
    t0 = time.perf_counter(); outs, tm = g.compress_batch(F.PVRTC4, imgs); dt = (time.perf_counter() - t0) * 1e3
    want, rms = r.compress("PVRTC4", imgs[-1], seed=None)
    print(json.dumps({"batch": n, "size": size, "gpu_total_ms": dt, "gpu_kernel_ms": tm["kernel_ms"],
                      "gpu_mpix_s": n * size * size / dt / 1e3, "ref_ms_per_texture": rms,
                      "ref_mpix_s_one_core": size * size / rms / 1e3, "last_equal": bool(np.array_equal(outs[-1], want))}))
