#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the UNMODIFIED reference compiled into
oracle/_ref/libfastc_ref.so (GammaUNC/FasTC @ 0f8cef65; `make -C oracle ref`).

The reference's own tests hold no golden vectors for BPTC / DXT / ETC1 / Core
(SURVEY.md §4, §8c), so parity is pinned on outputs of the reference itself run
in the build container.  Each fixture stores the input image, the reference's
compressed bytes per format (BC7 at quality 0, and at quality 8 / 50 with the
reference's LCG pinned to a known state, single thread, fresh watermark
counter), the reference decoder's output for those bytes and the reference's
PSNR.  Re-run:  python tools/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from _checkers import Reference  # noqa: E402
from fastc_b200.synth import synth_rgba  # noqa: E402

LCG_STATE = 12345


def special_blocks() -> np.ndarray:
    """64x64 RGBA: one 4x4 block per interesting case, the rest seeded noise of
    several amplitudes."""
    rng = np.random.default_rng(2024)
    img = np.zeros((64, 64, 4), dtype=np.uint8)
    blocks = img.reshape(16, 4, 16, 4, 4).transpose(0, 2, 1, 3, 4)  # view [by, bx, y, x, c]
    for by in range(16):
        for bx in range(16):
            amp = (1, 4, 16, 64, 128)[(by + bx) % 5]
            base = rng.integers(0, 256, 4)
            blk = (base + rng.integers(-amp, amp + 1, (4, 4, 4))).clip(0, 255)
            if by % 2 == 0:
                blk[..., 3] = 255
            blocks[by, bx] = blk
    k = iter(range(256))

    def put(blk):
        i = next(k)
        blocks[i // 16, i % 16] = np.asarray(blk, dtype=np.int64).clip(0, 255)

    ramp = np.arange(16).reshape(4, 4)
    put(np.broadcast_to([0, 0, 0, 0], (4, 4, 4)))                # solid black transparent
    put(np.broadcast_to([255, 255, 255, 255], (4, 4, 4)))        # solid white
    put(np.broadcast_to([10, 200, 30, 77], (4, 4, 4)))           # solid with alpha
    put(np.broadcast_to([1, 254, 128, 255], (4, 4, 4)))          # solid, odd values
    b = rng.integers(0, 256, (4, 4, 4)); b[..., 3] = 0; put(b)   # all-transparent, rgb varies
    b = np.zeros((4, 4, 4), int); b[..., :3] = (50, 60, 70); b[:2, :, :3] = (200, 10, 30); b[..., 3] = 255; put(b)  # 2 colours
    b = np.zeros((4, 4, 4), int); b[..., :3] = ramp[..., None] * 16; b[..., 3] = 255; put(b)   # collinear grey ramp
    b = np.zeros((4, 4, 4), int); b[..., 0] = ramp * 17; b[..., 1] = 255 - ramp * 17; b[..., 2] = 128; b[..., 3] = 255; put(b)
    b = np.zeros((4, 4, 4), int); b[..., :3] = (100, 100, 100); b[..., 3] = 252; b[0, 0] = (0, 255, 0, 250); put(b)  # alpha in [250,255)
    b = np.zeros((4, 4, 4), int); b[..., :3] = (90, 120, 30); b[..., 3] = ramp * 17; put(b)    # alpha ramp, constant rgb
    b = rng.integers(0, 256, (4, 4, 4)); b[..., 3] = 128; put(b)                               # constant alpha < 250
    b = np.zeros((4, 4, 4), int); b[...] = (1, 2, 3, 255); b[0, 0] = (1, 2, 4, 255); put(b)    # near-solid
    b = np.zeros((4, 4, 4), int); b[..., :3] = 128; b[..., 3] = rng.integers(0, 256, (4, 4)); put(b)  # same rgb, varying alpha (T13)
    b = np.zeros((4, 4, 4), int); b[:, :2] = (255, 0, 0, 255); b[:, 2:] = (0, 0, 255, 255); put(b)   # left/right split
    b = np.zeros((4, 4, 4), int); b[:2] = (0, 255, 0, 255); b[2:] = (255, 255, 0, 255); put(b)       # top/bottom split
    b = np.zeros((4, 4, 4), int); b[...] = (3, 2, 1, 255); b[1, 1] = (4, 2, 1, 255); put(b)          # tiny gradient
    return np.ascontiguousarray(img)


def make(name: str, img: np.ndarray, ref: Reference):
    h, w = img.shape[:2]
    out = {"image": img}
    for fmt in ("DXT1", "DXT5", "ETC1"):
        cmp, _ = ref.compress(fmt, img)
        out[f"{fmt}"] = cmp
        dec = ref.decode(fmt, cmp, w, h)
        out[f"{fmt}_decoded"] = dec
        out[f"{fmt}_psnr"] = np.float64(ref.psnr(img, dec))
    for q in (0, 8, 50):
        cmp, _ = ref.compress("BPTC", img, quality=q, seed=LCG_STATE)
        out[f"BPTC_q{q}"] = cmp
        out[f"BPTC_q{q}_lcg_after"] = np.uint32(ref.get_seed())
        dec = ref.decode("BPTC", cmp, w, h)
        out[f"BPTC_q{q}_psnr"] = np.float64(ref.psnr(img, dec))
        if q == 0:
            out["BPTC_q0_decoded"] = dec
    out["lcg_state"] = np.uint32(LCG_STATE)
    path = ROOT / "tests" / "golden" / f"{name}.npz"
    np.savez_compressed(path, **out)
    print(f"{path.relative_to(ROOT)}: {path.stat().st_size} bytes, {w}x{h}, "
          f"PSNR bc7 q0/q8/q50 = {out['BPTC_q0_psnr']:.3f}/{out['BPTC_q8_psnr']:.3f}/{out['BPTC_q50_psnr']:.3f}")


def main():
    ref = Reference()
    make("special_64", special_blocks(), ref)
    make("synth_64x48_seed1", synth_rgba(64, 48, 1, full_height=256, y0=40), ref)
    # a window of the BASELINE generator that contains a solid tile, a transparent tile and alpha
    make("synth_128x64_tiles", np.ascontiguousarray(synth_rgba(512, 512, 1)[128:192, 40:168]), ref)
    make("noise63_64", synth_rgba(64, 64, 3, noise_mask=63), ref)


if __name__ == "__main__":
    main()
