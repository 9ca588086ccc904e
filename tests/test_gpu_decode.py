"""GPU parity of the decoders and the PSNR reduction against the
oracle's restatement of the reference decoders (itself pinned to the compiled reference
and to the golden fixtures): bit-exact pixels on encoder output AND on random bit
patterns (every BC7 mode / partition / rotation, invalid ETC1 deltas, both DXT1 orders)."""
from pathlib import Path

import numpy as np
import pytest

from _checkers import BLOCK_BYTES
from fastc_b200 import ECompressionFormat as F
from fastc_b200.synth import synth_rgba

pytestmark = pytest.mark.gpu

GOLDEN = sorted((Path(__file__).resolve().parent / "golden").glob("*.npz"))


@pytest.mark.parametrize("path", GOLDEN, ids=[p.stem for p in GOLDEN])
@pytest.mark.parametrize("fmt,key", [("DXT1", "DXT1"), ("DXT5", "DXT5"), ("ETC1", "ETC1"), ("BPTC", "BPTC_q0")])
def test_decode_matches_golden(gpu, path, fmt, key):
    g = np.load(path)
    h, w = g["image"].shape[:2]
    dec = gpu.decompress(F[fmt], g[key], w, h)
    assert (dec == g[f"{key}_decoded"]).all()
    psnr = gpu.psnr(g["image"], dec)
    assert abs(psnr - float(g[f"{key}_psnr"])) < 1e-9


@pytest.mark.parametrize("fmt", ["DXT1", "DXT5", "ETC1", "BPTC"])
def test_decode_random_bit_patterns(gpu, oracle, fmt):
    rng = np.random.default_rng(3)
    w, h = 256, 128
    rnd = rng.integers(0, 256, (w // 4) * (h // 4) * BLOCK_BYTES[fmt], dtype=np.uint8)
    if fmt == "BPTC":  # uniform over modes 0-7 plus a few reserved (all-zero prefix) blocks
        blk = rnd.reshape(-1, 16)
        m = rng.integers(0, 8, len(blk))
        blk[:, 0] = (blk[:, 0] & ~((1 << (m + 1)) - 1).astype(np.uint8)) | (1 << m).astype(np.uint8)
        blk[:7, 0] = 0
    got = gpu.decompress(F[fmt], rnd, w, h)
    want = oracle.decode(fmt, rnd, w, h)
    bad = np.nonzero((got != want).any(-1))
    assert (got == want).all(), f"{len(bad[0])} pixels differ, first at {bad[0][:4]}, {bad[1][:4]}"


def test_psnr_matches_oracle_and_identical_is_inf(gpu, oracle):
    a = synth_rgba(512, 256, 1)
    b = synth_rgba(512, 256, 2, noise_mask=63)
    assert abs(gpu.psnr(a, b) - oracle.psnr(a, b)) < 1e-9
    assert gpu.psnr(a, a) == float("inf")


def test_device_roundtrip_encode_decode_psnr(gpu, oracle):
    """encode -> decode -> PSNR without leaving the device (what `tc` does per image)."""
    import torch
    img = synth_rgba(1024, 1024, 5)
    d_in = torch.from_numpy(img).cuda()
    for fmt in ("DXT5", "BPTC"):
        d_cmp = torch.zeros(256 * 256 * 16, dtype=torch.uint8, device="cuda")
        d_dec = torch.zeros_like(d_in)
        gpu.compress_device(F[fmt], d_in, d_cmp, width=1024, height=1024, quality=1, seed=1)
        gpu.decompress_device(F[fmt], d_cmp, d_dec, width=1024, height=1024)
        psnr = gpu.psnr_device(d_in, d_dec, width=1024, height=1024)
        want = oracle.decode(fmt, d_cmp.cpu().numpy(), 1024, 1024)
        assert (d_dec.cpu().numpy() == want).all()
        assert abs(psnr - oracle.psnr(img, want)) < 1e-9
        assert psnr > 30.0
