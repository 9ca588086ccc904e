"""Integer-only procedural RGBA texture generator.

The same formula is used for every config in BASELINE.json so that the GPU
path, the oracle and the compiled reference all see byte-identical inputs.
Everything is exact integer arithmetic (numpy uint32/int64), so the C++ twin in
fastc_b200/core/synth.h produces the same bytes.
"""
from __future__ import annotations

import numpy as np

_M32 = np.uint64(0xFFFFFFFF)


def fmix32(h: np.ndarray) -> np.ndarray:
    """murmur3 finaliser on uint32 (computed in uint64 and masked)."""
    h = h.astype(np.uint64) & _M32
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x85EBCA6B)) & _M32
    h ^= h >> np.uint64(13)
    h = (h * np.uint64(0xC2B2AE35)) & _M32
    h ^= h >> np.uint64(16)
    return h


def _tri(t: np.ndarray, p: int) -> np.ndarray:
    return np.abs(((t % p) * 510) // p - 255)


def synth_rgba(width: int, height: int, seed: int = 1, *, opaque: bool = False,
               noise_mask: int = 15, y0: int = 0, full_height: int | None = None) -> np.ndarray:
    """Return a (height, width, 4) uint8 RGBA image.

    `y0`/`full_height` generate rows [y0, y0+height) of a taller image with the
    same coordinates (used to shard one texture across ranks and for the
    "top slab" CPU-baseline sample).
    """
    H = full_height if full_height is not None else height
    x = np.arange(width, dtype=np.int64)[None, :]
    y = (np.arange(height, dtype=np.int64) + y0)[:, None]
    half = (noise_mask + 1) // 2

    def nz(c: int) -> np.ndarray:
        key = (x + 8192 * y + seed * 0x9E3779B9 + c * 0x85EBCA6B) & 0xFFFFFFFF
        return (fmix32(key.astype(np.uint64)).astype(np.int64) & noise_mask) - half

    clamp = lambda v: np.clip(v, 0, 255)
    r = clamp(x * 255 // max(width - 1, 1) + nz(0))
    g = clamp(y * 255 // max(H - 1, 1) + nz(1))
    b = clamp(_tri(x + y, 97) + nz(2))
    tx, ty = x // 64, y // 64
    a = np.where((tx + ty) % 4 == 0, clamp(_tri(y + 0 * x, 61) + nz(3)), 255)

    solid = (7 * tx + 13 * ty) % 29 == 5
    sc = fmix32(((131 * tx + 977 * ty + seed) & 0xFFFFFFFF).astype(np.uint64)).astype(np.int64)
    r = np.where(solid, sc & 0xFF, r)
    g = np.where(solid, (sc >> 8) & 0xFF, g)
    b = np.where(solid, (sc >> 16) & 0xFF, b)
    a = np.where(solid, 255, a)
    a = np.where((5 * tx + 11 * ty) % 31 == 7, 0, a)
    if opaque:
        a = np.full_like(a, 255)

    out = np.empty((height, width, 4), dtype=np.uint8)
    out[..., 0] = np.broadcast_to(r, (height, width))
    out[..., 1] = np.broadcast_to(g, (height, width))
    out[..., 2] = np.broadcast_to(b, (height, width))
    out[..., 3] = np.broadcast_to(a, (height, width))
    return out


def synth_rgba_torch(width: int, height: int, seed: int = 1, *, opaque: bool = False, noise_mask: int = 15,
                     y0: int = 0, full_height: int | None = None, device="cuda"):
    """Same image as synth_rgba(), generated on `device` with torch integer ops
    (used by bench.py: 8192^2 takes ~25 s in numpy).  Returns (height, width, 4) uint8."""
    import torch
    H = full_height if full_height is not None else height
    i64 = torch.int64
    x = torch.arange(width, dtype=i64, device=device)[None, :]
    y = (torch.arange(height, dtype=i64, device=device) + y0)[:, None]
    half = (noise_mask + 1) // 2
    M = 0xFFFFFFFF

    def fmix(h):
        h = h & M
        h = h ^ (h >> 16)
        h = (h * 0x85EBCA6B) & M
        h = h ^ (h >> 13)
        h = (h * 0xC2B2AE35) & M
        h = h ^ (h >> 16)
        return h

    def nz(c):
        return (fmix((x + 8192 * y + seed * 0x9E3779B9 + c * 0x85EBCA6B) & M) & noise_mask) - half

    def tri(t, p):
        return torch.abs(((t % p) * 510) // p - 255)

    clamp = lambda v: torch.clamp(v, 0, 255)
    r = clamp(x * 255 // max(width - 1, 1) + nz(0))
    g = clamp(y * 255 // max(H - 1, 1) + nz(1))
    b = clamp(tri(x + y, 97) + nz(2))
    tx, ty = x // 64, y // 64
    a = torch.where((tx + ty) % 4 == 0, clamp(tri(y + 0 * x, 61) + nz(3)), torch.full_like(r, 255))
    solid = (7 * tx + 13 * ty) % 29 == 5
    sc = fmix((131 * tx + 977 * ty + seed) & M)
    r = torch.where(solid, sc & 0xFF, r)
    g = torch.where(solid, (sc >> 8) & 0xFF, g.expand(height, width))
    b = torch.where(solid, (sc >> 16) & 0xFF, b)
    a = torch.where(solid, torch.full_like(a, 255), a)
    a = torch.where((5 * tx + 11 * ty) % 31 == 7, torch.zeros_like(a), a)
    if opaque:
        a = torch.full_like(a, 255)
    out = torch.stack([r.expand(height, width), g.expand(height, width), b.expand(height, width),
                       a.expand(height, width)], dim=-1).to(torch.uint8)
    return out.contiguous()
