"""CPU (-m "not gpu"): pins the oracle restatement against the UNMODIFIED
reference compiled into oracle/_ref (skipped where that .so is absent), and
against the committed golden fixtures (tests/golden, generated from the
reference by tools/make_golden.py)."""
import numpy as np
import pytest

from _checkers import BLOCK_BYTES
from fastc_b200.synth import synth_rgba


def _bad(a, b, fmt):
    return int((a.reshape(-1, BLOCK_BYTES[fmt]) != b.reshape(-1, BLOCK_BYTES[fmt])).any(1).sum())


@pytest.mark.parametrize("fmt", ["DXT1", "DXT5"])
@pytest.mark.parametrize("w,h,seed,kw", [(256, 256, 1, {}), (128, 512, 2, {"noise_mask": 63}),
                                         (256, 64, 3, {"opaque": True})])
def test_dxt_oracle_equals_reference(oracle, reference, fmt, w, h, seed, kw):
    img = synth_rgba(w, h, seed, **kw)
    a, _ = oracle.compress(fmt, img)
    b, _ = reference.compress(fmt, img)
    assert _bad(a, b, fmt) == 0


@pytest.mark.parametrize("fmt", ["DXT1", "DXT5"])
def test_dxt_oracle_equals_reference_random(oracle, reference, fmt):
    img = np.random.default_rng(5).integers(0, 256, (128, 128, 4), dtype=np.uint8)
    a, _ = oracle.compress(fmt, img)
    b, _ = reference.compress(fmt, img)
    assert _bad(a, b, fmt) == 0


@pytest.mark.parametrize("w,h,seed,kw", [(256, 256, 1, {}), (128, 512, 2, {"noise_mask": 63}),
                                         (256, 64, 3, {"opaque": True})])
def test_etc1_oracle_equals_reference(oracle, reference, w, h, seed, kw):
    img = synth_rgba(w, h, seed, **kw)
    a, _ = oracle.compress("ETC1", img)
    b, _ = reference.compress("ETC1", img)
    assert _bad(a, b, "ETC1") == 0


def test_etc1_oracle_equals_reference_random_lowvariance_solid(oracle, reference):
    rng = np.random.default_rng(5)
    noise = rng.integers(0, 256, (64, 256, 4), dtype=np.uint8)
    base = rng.integers(0, 256, (16, 64, 1, 1, 4))
    low = (base + rng.integers(-6, 7, (16, 64, 4, 4, 4))).clip(0, 255).astype(np.uint8)
    low = low.transpose(0, 2, 1, 3, 4).reshape(64, 256, 4)
    cols = rng.integers(0, 256, (64, 64, 4), dtype=np.uint8)
    cols[0, :8, :3] = 0
    cols[0, 8:16, :3] = 255
    solid = np.repeat(np.repeat(cols, 4, 0), 4, 1)
    img = np.ascontiguousarray(np.concatenate([noise, low, solid], 0))
    a, _ = oracle.compress("ETC1", img)
    b, _ = reference.compress("ETC1", img)
    assert _bad(a, b, "ETC1") == 0


def test_bc7_oracle_equals_reference_q0_config1(oracle, reference):
    """BASELINE config 1: 256x256 RGBA, -q 0, single thread: bit-identical incl. the
    watermark sequence of the solid-colour blocks (T1)."""
    img = synth_rgba(256, 256, 1)
    a, _ = oracle.compress("BPTC", img, quality=0, rng_mode=0)
    b, _ = reference.compress("BPTC", img, quality=0)
    assert _bad(a, b, "BPTC") == 0


@pytest.mark.parametrize("q,seed", [(1, 7), (2, 99), (5, 12345), (50, 1), (200, 31337)])
def test_bc7_oracle_equals_reference_annealing_pinned_lcg(oracle, reference, q, seed):
    """quality > 0 with the reference's process-global LCG pinned to `seed` (single thread):
    the restatement draws the same random numbers in the same order -> identical bytes and
    identical LCG state afterwards."""
    img = np.ascontiguousarray(synth_rgba(256, 256, 1)[96:128, 32:160])  # alpha, opaque and solid tiles
    if q >= 50:
        img = img[:16, :64]
    a, lcg = oracle.compress("BPTC", img, quality=q, rng_mode=0, lcg_state=seed)
    b, _ = reference.compress("BPTC", img, quality=q, seed=seed)
    assert _bad(a, b, "BPTC") == 0
    assert lcg == reference.get_seed()


@pytest.mark.parametrize("fmt", ["DXT1", "DXT5", "ETC1", "BPTC"])
def test_decoders_and_psnr_equal_reference(oracle, reference, fmt):
    img = synth_rgba(128, 128, 4)
    cmp, _ = reference.compress(fmt, img, quality=0)
    a = oracle.decode(fmt, cmp, 128, 128)
    b = reference.decode(fmt, cmp, 128, 128)
    assert (a == b).all()
    assert oracle.psnr(img, a) == reference.psnr(img, b)
    rng = np.random.default_rng(3)
    rnd = rng.integers(0, 256, cmp.size, dtype=np.uint8)
    if fmt == "BPTC":  # make every block start with a valid unary mode prefix
        blk = rnd.reshape(-1, 16)
        m = rng.integers(0, 8, len(blk))
        blk[:, 0] = (blk[:, 0] & ~((1 << (m + 1)) - 1).astype(np.uint8)) | (1 << m).astype(np.uint8)
    assert (oracle.decode(fmt, rnd, 128, 128) == reference.decode(fmt, rnd, 128, 128)).all()


def _styled_blocks(rng, by, bx):
    """An image of by x bx blocks, each drawn in a random 'style' that targets a branch of the encoders."""
    img = np.zeros((by * 4, bx * 4, 4), dtype=np.uint8)
    for j in range(by):
        for i in range(bx):
            style = rng.integers(0, 10)
            blk = np.zeros((4, 4, 4), dtype=np.int64)
            base = rng.integers(0, 256, 4)
            if style == 0:                                   # noise
                blk = rng.integers(0, 256, (4, 4, 4))
            elif style == 1:                                 # low variance around a colour
                blk = base + rng.integers(-3, 4, (4, 4, 4))
            elif style == 2:                                 # two colours (collinear points, T7)
                other = rng.integers(0, 256, 4)
                pick = rng.integers(0, 2, (4, 4, 1))
                blk = np.where(pick == 1, base, other)
            elif style == 3:                                 # gradient along x with a little noise
                ramp = np.arange(4).reshape(1, 4, 1) * rng.integers(1, 40)
                blk = base + ramp + rng.integers(0, 2, (4, 4, 4))
            elif style == 4:                                 # opaque-ish alpha around the 250 threshold (T11, T18)
                blk = rng.integers(0, 256, (4, 4, 4))
                blk[..., 3] = rng.integers(248, 256, (4, 4))
            elif style == 5:                                 # alpha ramp over smooth colour (modes 4 / 5)
                blk = base + rng.integers(-8, 9, (4, 4, 4))
                blk[..., 3] = np.linspace(rng.integers(0, 128), rng.integers(128, 256), 16).reshape(4, 4)
            elif style == 6:                                 # fully transparent, RGB varies
                blk = rng.integers(0, 256, (4, 4, 4))
                blk[..., 3] = 0
            elif style == 7:                                 # solid colour (BC7 watermark, DXT / ETC1 solid paths)
                blk = np.broadcast_to(base, (4, 4, 4)).copy()
            elif style == 8:                                 # constant colour, varying alpha (T13)
                blk = np.broadcast_to(base, (4, 4, 4)).copy()
                blk[..., 3] = rng.integers(0, 256, (4, 4))
            else:                                            # saturated extremes
                blk = rng.choice([0, 255], (4, 4, 4))
            img[4 * j:4 * j + 4, 4 * i:4 * i + 4] = np.clip(blk, 0, 255)
    return img


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_oracle_equals_reference_on_styled_random_blocks(oracle, reference, seed):
    """Every encoder, on blocks built to hit the special branches (collinear points, alpha near the
    opaque threshold, transparent / solid / constant-colour blocks, saturated values): the restatement
    and the compiled reference agree byte for byte -- BC7 at -q 0 and, LCG pinned, at -q 3."""
    rng = np.random.default_rng(seed)
    img = _styled_blocks(rng, 12, 16)
    for fmt in ("DXT1", "DXT5", "ETC1"):
        a, _ = oracle.compress(fmt, img)
        b, _ = reference.compress(fmt, img)
        assert _bad(a, b, fmt) == 0, fmt
    a, _ = oracle.compress("BPTC", img, quality=0, rng_mode=0)
    b, _ = reference.compress("BPTC", img, quality=0)
    assert _bad(a, b, "BPTC") == 0
    a, lcg = oracle.compress("BPTC", img, quality=3, rng_mode=0, lcg_state=seed)
    b, _ = reference.compress("BPTC", img, quality=3, seed=seed)
    assert _bad(a, b, "BPTC") == 0 and lcg == reference.get_seed()
    dec_o, dec_r = oracle.decode("BPTC", b, 64, 48), reference.decode("BPTC", b, 64, 48)
    assert (dec_o == dec_r).all()


# BPTCC::CompressionSettings beyond the annealing steps (reference BPTCCompressor.h:123-158): the
# oracle against the reference's own BPTCC::Compress(job, settings).
# Masks that leave every block at least one mode: with an empty set the reference runs off the end of its
# mode table (an assert compiled out, Compressor.cpp:1858 / :1811) and may crash.  A mask is safe when it
# meets each set BoxSelection can return: {1,3,7}, {0,2}, {4,5,6,7}, {0,1,2,3,6,7}.
SETTINGS_MASKS = [0x4F, 0xA5, 0x8D, 0x66, 0xB1, 0x1B]


@pytest.mark.parametrize("mask", SETTINGS_MASKS)
def test_bc7_oracle_block_modes_equal_reference_q0(oracle, reference, mask):
    img = synth_rgba(128, 128, 4)
    a, _ = oracle.compress("BPTC", img, quality=0, rng_mode=0, block_modes=mask)
    b = reference.compress_bptc_settings(img, quality=0, block_modes=mask)
    assert a.reshape(-1, 16).any(1).all()
    assert (a == b).all()
    b0 = a.reshape(-1, 16)[:, 0].astype(np.int64)
    modes = np.log2(b0 & -b0).astype(int)
    blocks = img.reshape(32, 4, 32, 4, 4).transpose(0, 2, 1, 3, 4).reshape(-1, 16, 4)
    normal = ~(blocks == blocks[:, :1]).all((1, 2)) & ~(blocks[..., 3] == 0).all(1)
    assert ((mask >> modes[normal]) & 1).all()


@pytest.mark.parametrize("mask,q", [(0xFF, 0), (0xFF, 3), (0x4F, 0), (0xB1, 2)])
def test_bc7_oracle_nonuniform_metric_equals_reference(oracle, reference, mask, q):
    img = synth_rgba(128, 128, 6, noise_mask=63)
    a, st = oracle.compress("BPTC", img, quality=q, rng_mode=0, lcg_state=77, block_modes=mask, error_metric=1)
    b = reference.compress_bptc_settings(img, quality=q, block_modes=mask, error_metric=1, seed=77)
    assert (a == b).all()
    assert st == reference.get_seed()


@pytest.mark.parametrize("quality", [0, 1, 2])
def test_etc1_oracle_quality_levels_equal_reference(oracle, reference, quality):
    """rg_etc1's three quality levels (FasTC only ever passes cLowQuality; BASELINE configs[4] names
    the high one): the oracle against rg_etc1::pack_etc1_block itself."""
    rng = np.random.default_rng(17)
    h = 64 if quality == 2 else 128
    smooth = synth_rgba(128, h, 3, opaque=True)
    noise = rng.integers(0, 256, (32, 128, 4), dtype=np.uint8)
    base = rng.integers(0, 256, (8, 32, 1, 1, 4))
    low = (base + rng.integers(-6, 7, (8, 32, 4, 4, 4))).clip(0, 255).astype(np.uint8)
    low = low.transpose(0, 2, 1, 3, 4).reshape(32, 128, 4)
    # half-solid blocks: one 2x4 / 4x2 subblock of a single colour (the constrained solid path)
    half = rng.integers(0, 256, (32, 128, 4), dtype=np.uint8)
    cols = rng.integers(0, 256, (8, 32, 4), dtype=np.uint8)
    for by in range(8):
        for bx in range(32):
            k = (by * 32 + bx) % 4
            ys, xs = [(slice(0, 2), slice(0, 4)), (slice(2, 4), slice(0, 4)), (slice(0, 4), slice(0, 2)),
                      (slice(0, 4), slice(2, 4))][k]
            blk = half[by * 4:by * 4 + 4, bx * 4:bx * 4 + 4]
            blk[ys, xs] = cols[by, bx]
            if (by + bx) % 3 == 0:   # both halves solid, close colours
                blk[...] = cols[by, bx]
                blk[ys, xs] = (cols[by, bx].astype(int) + rng.integers(-9, 10, 4)).clip(0, 255)
    img = np.ascontiguousarray(np.concatenate([smooth, noise, low, half], 0))
    img[..., 3] = 255
    a, _ = oracle.compress("ETC1", img, etc1_quality=quality)
    b = reference.compress_etc1_quality(img, quality)
    bad = np.flatnonzero((a.reshape(-1, 8) != b.reshape(-1, 8)).any(1))
    assert bad.size == 0, (quality, bad[:10])
