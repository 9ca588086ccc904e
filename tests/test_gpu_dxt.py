"""GPU parity: DXT1/DXT5 CUDA path vs the CPU oracle (and the compiled
reference when its prebuilt .so travelled with the snapshot).  Bit-exact."""
import numpy as np
import pytest

from _checkers import BLOCK_BYTES, Reference
from fastc_b200 import ECompressionFormat as F
from fastc_b200.synth import synth_rgba

pytestmark = pytest.mark.gpu

CASES = [
    (256, 256, 1, {}),
    (512, 128, 2, {"noise_mask": 63}),
    (64, 1024, 3, {"opaque": True}),
    (4, 4, 4, {}),
    (1028, 12, 5, {}),
]


def _mismatch(a, b, fmt):
    a = a.reshape(-1, BLOCK_BYTES[fmt]); b = b.reshape(-1, BLOCK_BYTES[fmt])
    return int((a != b).any(1).sum())


@pytest.mark.parametrize("fmt", ["DXT1", "DXT5"])
@pytest.mark.parametrize("w,h,seed,kw", CASES)
def test_dxt_matches_oracle(gpu, oracle, fmt, w, h, seed, kw):
    img = synth_rgba(w, h, seed, **kw)
    got, tm = gpu.compress(F[fmt], img)
    want, _ = oracle.compress(fmt, img)
    assert _mismatch(got, want, fmt) == 0
    assert tm["kernel_launches"] >= 1


@pytest.mark.parametrize("fmt", ["DXT1", "DXT5"])
def test_dxt_random_and_edge_blocks(gpu, oracle, fmt):
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, (128, 128, 4), dtype=np.uint8)
    img[:4, :4] = 0                      # solid black, alpha 0
    img[:4, 4:8] = 255                   # solid white
    img[4:8, :4] = (10, 200, 30, 77)     # solid colour with alpha
    img[8:12, :8, :3] = 128              # constant RGB, varying alpha (T13: not "constant")
    img[12:16, :4] = img[12, 0]          # constant
    img[12:16, 0:4, 0] = np.arange(4)    # tiny gradient (magn < 4 -> luminance axis)
    got, _ = gpu.compress(F[fmt], img)
    want, _ = oracle.compress(fmt, img)
    assert _mismatch(got, want, fmt) == 0


@pytest.mark.parametrize("fmt", ["DXT1", "DXT5"])
def test_dxt_matches_reference_so(gpu, fmt):
    if not Reference.available():
        pytest.skip("prebuilt reference .so not shipped")
    ref = Reference()
    img = synth_rgba(512, 512, 11)
    got, _ = gpu.compress(F[fmt], img)
    want, _ = ref.compress(fmt, img)
    assert _mismatch(got, want, fmt) == 0


@pytest.mark.parametrize("fmt", ["DXT1", "DXT5"])
def test_dxt_block_range_and_chunking(gpu, oracle, fmt):
    """CompressionJob semantics: only [first, first+n) is written; chunked
    pipeline == one submission."""
    img = synth_rgba(256, 128, 9)
    full, _ = gpu.compress(F[fmt], img)
    bs = BLOCK_BYTES[fmt]
    out = np.full(full.size, 0xEE, dtype=np.uint8)
    gpu.compress(F[fmt], img, out, first_block=70, num_blocks=300)
    assert (out[:70 * bs] == 0xEE).all() and (out[370 * bs:] == 0xEE).all()
    assert (out[70 * bs:370 * bs] == full[70 * bs:370 * bs]).all()
    chunked, tm = gpu.compress(F[fmt], img, chunk_blocks=64 * 3)
    assert (chunked == full).all()
    assert tm["kernel_launches"] > 1


def test_dxt_device_path_large_roundtrip(gpu, oracle):
    """BASELINE config 4 shape (1024^2 textures): device-resident path, and a
    size-independent property -- decode(encode(x)) stays close to x."""
    import torch
    img = synth_rgba(1024, 1024, 21, opaque=True)
    d_in = torch.from_numpy(img).cuda()
    for fmt in ("DXT1", "DXT5"):
        d_out = torch.zeros(65536 * BLOCK_BYTES[fmt], dtype=torch.uint8, device="cuda")
        n = gpu.compress_device(F[fmt], d_in, d_out, width=1024, height=1024)
        torch.cuda.synchronize()
        assert n == 1
        got = d_out.cpu().numpy()
        host, _ = gpu.compress(F[fmt], img)
        assert (got == host).all()
        dec = oracle.decode(fmt, got, 1024, 1024)
        assert oracle.psnr(img, dec) > 30.0


def test_dxt_config4_batch_of_1024_textures(gpu, oracle):
    """BASELINE configs[3]: DXT1 and DXT5 on a batch of 256 synthetic 1024x1024 textures (seeds 1..256),
    ONE batch submission (fastc_gpu_compress_batch, textures pipelined through the staging slots).
    Every texture of the batch equals its own single-texture submission; the first 48 are also
    compared with the CPU oracle bit for bit (the oracle needs ~75 ms per texture and format)."""
    import torch
    from fastc_b200.synth import synth_rgba_torch
    imgs = [np.ascontiguousarray(synth_rgba_torch(1024, 1024, seed, device="cuda").cpu().numpy())
            for seed in range(1, 257)]
    assert (imgs[0] == synth_rgba(1024, 1024, 1)).all()   # the device generator is the numpy generator
    for fmt in ("DXT1", "DXT5"):
        outs, tm = gpu.compress_batch(F[fmt], imgs)
        assert len(outs) == 256 and tm["kernel_launches"] >= 256
        assert tm["h2d_bytes"] == 256 * 1024 * 1024 * 4
        for k in range(0, 256, 5):
            single, _ = gpu.compress(F[fmt], imgs[k])
            assert _mismatch(outs[k], single, fmt) == 0, (fmt, k)
        for k in range(48):
            want, _ = oracle.compress(fmt, imgs[k])
            assert _mismatch(outs[k], want, fmt) == 0, (fmt, k)


@pytest.mark.parametrize("fmt", ["DXT1", "BPTC"])
def test_concurrent_ranged_submissions_like_threadgroup(gpu, oracle, fmt):
    """FasTC's ThreadGroup calls a CompressionFunc from N threads at once on disjoint block ranges of
    the same buffers (Core/src/ThreadGroup.cpp:146-188): concurrent fastc_gpu_compress calls on one
    device must serialise safely and give the bytes of one whole-image call."""
    import threading
    img = synth_rgba(256, 256, 9)
    nblk = 64 * 64
    whole, _ = gpu.compress(F[fmt], img, quality=0)
    out = np.zeros_like(whole)
    nthreads = 8
    per = (nblk + nthreads - 1) // nthreads          # ceil split, as PrepareThreads does
    errs = []

    def work(t):
        try:
            first = t * per
            gpu.compress(F[fmt], img, out, quality=0, first_block=first, num_blocks=min(per, nblk - first))
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(t,)) for t in range(nthreads)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    assert _mismatch(out, whole, fmt) == 0


@pytest.mark.parametrize("fmt,q", [("DXT5", 0), ("ETC1", 0), ("BPTC", 3)])
def test_pageable_and_pinned_host_buffers_give_the_same_bytes(gpu, fmt, q):
    """Pageable caller memory (what FasTC passes) is staged through the library's pinned ring /
    per-slot pinned download buffers; pinned caller memory goes straight to the copy engines.
    Same bytes either way, also when only one side is pinned and across several chunks."""
    import torch
    img = synth_rgba(1024, 2048, 5, opaque=(fmt == "ETC1"))          # 8 MiB: several DXT / ETC1 chunks
    pin_in = torch.empty(img.shape, dtype=torch.uint8, pin_memory=True)
    pin_in.numpy()[...] = img
    nbytes = (1024 // 4) * (2048 // 4) * BLOCK_BYTES[fmt]
    pin_out = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    a, _ = gpu.compress(F[fmt], img, quality=q, seed=4)                                   # pageable -> pageable
    b, _ = gpu.compress(F[fmt], pin_in.numpy(), pin_out.numpy(), quality=q, seed=4)       # pinned -> pinned
    c, _ = gpu.compress(F[fmt], img, pin_out.numpy().copy() * 0, quality=q, seed=4, chunk_blocks=256 * 7)
    d, _ = gpu.compress(F[fmt], pin_in.numpy(), np.zeros(nbytes, np.uint8), quality=q, seed=4)  # pinned -> pageable
    assert (a == b).all() and (a == c).all() and (a == d).all()
