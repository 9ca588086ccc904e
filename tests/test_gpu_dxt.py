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
