"""Golden fixtures (tests/golden/*.npz) = outputs of the UNMODIFIED reference,
generated in the build container by tools/make_golden.py (the reference's own
tests hold no vectors for this path).

CPU (-m "not gpu"): the oracle restatement reproduces every fixture bit for bit
-- including BC7 at quality 8 / 50 with the reference's global LCG pinned, and
the LCG state after the run (same number of draws).
GPU (-m gpu): the CUDA path reproduces the deterministic fixtures bit for bit
(DXT1, DXT5, ETC1, BC7 quality 0) and meets the PSNR tolerance at quality 50.
"""
from pathlib import Path

import numpy as np
import pytest

from _checkers import BLOCK_BYTES
from fastc_b200 import ECompressionFormat as F

GOLDEN = sorted((Path(__file__).resolve().parent / "golden").glob("*.npz"))
IDS = [p.stem for p in GOLDEN]


def _bad(a, b, fmt):
    bs = BLOCK_BYTES[fmt]
    return np.nonzero((a.reshape(-1, bs) != b.reshape(-1, bs)).any(1))[0]


def test_golden_fixtures_present():
    assert len(GOLDEN) >= 4


@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
@pytest.mark.parametrize("fmt", ["DXT1", "DXT5", "ETC1"])
def test_oracle_reproduces_golden_dxt_etc1(oracle, path, fmt):
    g = np.load(path)
    img = g["image"]
    h, w = img.shape[:2]
    got, _ = oracle.compress(fmt, img)
    bad = _bad(got, g[fmt], fmt)
    assert len(bad) == 0, f"{len(bad)} blocks differ, first {bad[:8]}"
    dec = oracle.decode(fmt, g[fmt], w, h)
    assert (dec == g[f"{fmt}_decoded"]).all()
    assert oracle.psnr(img, dec) == float(g[f"{fmt}_psnr"])


@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
@pytest.mark.parametrize("q", [0, 8, 50])
def test_oracle_reproduces_golden_bc7(oracle, path, q):
    g = np.load(path)
    img = g["image"]
    h, w = img.shape[:2]
    got, lcg = oracle.compress("BPTC", img, quality=q, rng_mode=0, lcg_state=int(g["lcg_state"]))
    bad = _bad(got, g[f"BPTC_q{q}"], "BPTC")
    assert len(bad) == 0, f"q={q}: {len(bad)} blocks differ, first {bad[:8]}"
    assert lcg == int(g[f"BPTC_q{q}_lcg_after"])
    dec = oracle.decode("BPTC", got, w, h)
    assert oracle.psnr(img, dec) == float(g[f"BPTC_q{q}_psnr"])
    if q == 0:
        assert (dec == g["BPTC_q0_decoded"]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
@pytest.mark.parametrize("fmt,key", [("DXT1", "DXT1"), ("DXT5", "DXT5"), ("ETC1", "ETC1"), ("BPTC", "BPTC_q0")])
def test_gpu_reproduces_golden(gpu, path, fmt, key):
    g = np.load(path)
    got, _ = gpu.compress(F[fmt], g["image"], quality=0)
    bad = _bad(got, g[key], fmt)
    assert len(bad) == 0, f"{len(bad)} blocks differ, first {bad[:8]}"


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_gpu_bc7_q50_psnr_vs_golden(gpu, oracle, path):
    """north_star: PSNR(GPU) >= PSNR(reference) - 0.05 dB at quality > 0 (reference decoder + formula)."""
    g = np.load(path)
    img = g["image"]
    h, w = img.shape[:2]
    got, _ = gpu.compress(F.BPTC, img, quality=50, seed=1)
    psnr = oracle.psnr(img, oracle.decode("BPTC", got, w, h))
    assert psnr >= float(g["BPTC_q50_psnr"]) - 0.05, (psnr, float(g["BPTC_q50_psnr"]))
