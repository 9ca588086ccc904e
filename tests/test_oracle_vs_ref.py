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
