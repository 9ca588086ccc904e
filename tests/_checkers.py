"""ctypes bindings for the CHECKERS: the CPU oracle (oracle/libfastc_oracle.so)
and the compiled reference (oracle/_ref/libfastc_ref.so).

Test infrastructure only.  Nothing under fastc_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_SO = ROOT / "oracle" / "libfastc_oracle.so"
REF_SO = ROOT / "oracle" / "_ref" / "libfastc_ref.so"

FMT = {"DXT1": 0, "DXT5": 1, "ETC1": 2, "BPTC": 3, "PVRTC4": 4}
BLOCK_BYTES = {"DXT1": 8, "DXT5": 16, "ETC1": 8, "BPTC": 16, "PVRTC4": 8}

_u8p = C.POINTER(C.c_uint8)


def _p(a: np.ndarray):
    return a.ctypes.data_as(_u8p)


def nblocks(w: int, h: int) -> int:
    return (w // 4) * (h // 4)


class Oracle:
    def __init__(self):
        if not ORACLE_SO.exists():
            raise RuntimeError(f"{ORACLE_SO} missing: run `make -C oracle oracle` (or __graft_entry__.build())")
        L = self.lib = C.CDLL(str(ORACLE_SO))
        L.fastc_oracle_dxt.argtypes = [C.c_int, _u8p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _u8p]
        L.fastc_oracle_dxt.restype = None
        if hasattr(L, "fastc_oracle_etc1"):
            L.fastc_oracle_etc1.argtypes = [_u8p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _u8p]
            L.fastc_oracle_etc1.restype = None
        if hasattr(L, "fastc_oracle_bc7"):
            L.fastc_oracle_bc7.argtypes = [_u8p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _u8p,
                                           C.c_int, C.c_int, C.POINTER(C.c_uint32), C.c_uint64, C.c_uint32]
            L.fastc_oracle_bc7.restype = None
        L.fastc_oracle_etc1_quality.argtypes = [_u8p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _u8p, C.c_int]
        L.fastc_oracle_etc1_quality.restype = None
        L.fastc_oracle_bc7_settings.argtypes = [_u8p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _u8p,
                                                C.c_int, C.c_int, C.POINTER(C.c_uint32), C.c_uint64, C.c_uint32,
                                                C.c_uint32, C.c_uint32, C.c_int]
        L.fastc_oracle_bc7_settings.restype = None
        if hasattr(L, "fastc_oracle_decode"):
            L.fastc_oracle_decode.argtypes = [C.c_int, _u8p, C.c_uint32, C.c_uint32, _u8p]
            L.fastc_oracle_decode.restype = None
            L.fastc_oracle_psnr.argtypes = [_u8p, _u8p, C.c_uint32, C.c_uint32]
            L.fastc_oracle_psnr.restype = C.c_double

    def compress(self, fmt: str, img: np.ndarray, *, quality: int = 50, first_block: int = 0,
                 num_blocks: int | None = None, rng_mode: int = 1, lcg_state: int = 1,
                 seed: int = 0, wm_base: int = 0, block_index_base: int = 0, block_modes: int = 0xFF,
                 error_metric: int = 0, etc1_quality: int = 0):
        """Returns (bytes array for the WHOLE image (untouched blocks zero), final lcg state)."""
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w = img.shape[:2]
        nb = nblocks(w, h)
        if num_blocks is None:
            num_blocks = nb - first_block
        out = np.zeros(nb * BLOCK_BYTES[fmt], dtype=np.uint8)
        st = C.c_uint32(lcg_state)
        if fmt in ("DXT1", "DXT5"):
            self.lib.fastc_oracle_dxt(int(fmt == "DXT5"), _p(img), w, h, first_block, num_blocks, _p(out))
        elif fmt == "ETC1":
            self.lib.fastc_oracle_etc1_quality(_p(img), w, h, first_block, num_blocks, _p(out), etc1_quality)
        else:
            self.lib.fastc_oracle_bc7_settings(_p(img), w, h, first_block, num_blocks, _p(out), quality, rng_mode,
                                               C.byref(st), seed, wm_base, block_index_base, block_modes,
                                               error_metric)
        return out, st.value

    def decode(self, fmt: str, cmp: np.ndarray, w: int, h: int) -> np.ndarray:
        cmp = np.ascontiguousarray(cmp, dtype=np.uint8)
        out = np.zeros((h, w, 4), dtype=np.uint8)
        self.lib.fastc_oracle_decode(FMT[fmt], _p(cmp), w, h, _p(out))
        return out

    def psnr(self, a: np.ndarray, b: np.ndarray) -> float:
        a = np.ascontiguousarray(a, dtype=np.uint8)
        b = np.ascontiguousarray(b, dtype=np.uint8)
        h, w = a.shape[:2]
        return float(self.lib.fastc_oracle_psnr(_p(a), _p(b), w, h))


class Reference:
    """The UNMODIFIED reference compiled into oracle/_ref (see oracle/Makefile)."""

    def __init__(self):
        if not REF_SO.exists():
            raise RuntimeError(f"{REF_SO} missing (built only where /root/reference exists)")
        L = self.lib = C.CDLL(str(REF_SO))
        L.fastc_ref_compress.argtypes = [C.c_int, _u8p, C.c_uint32, C.c_uint32, _u8p, C.c_uint32,
                                         C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.fastc_ref_compress.restype = C.c_int
        L.fastc_ref_decompress.argtypes = [C.c_int, _u8p, C.c_uint32, C.c_uint32, _u8p]
        L.fastc_ref_decompress.restype = C.c_int
        L.fastc_ref_psnr.argtypes = [_u8p, _u8p, C.c_uint32, C.c_uint32]
        L.fastc_ref_psnr.restype = C.c_double
        L.fastc_ref_set_state.argtypes = [C.c_uint32, C.c_uint32]
        L.fastc_ref_set_state.restype = None
        L.fastc_ref_get_seed.restype = C.c_uint32
        L.fastc_ref_bptc_compress_settings.argtypes = [_u8p, C.c_uint32, C.c_uint32, _u8p, C.c_int, C.c_uint32,
                                                       C.c_int]
        L.fastc_ref_bptc_compress_settings.restype = C.c_int
        L.fastc_ref_etc1_compress_quality.argtypes = [_u8p, C.c_uint32, C.c_uint32, _u8p, C.c_int]
        L.fastc_ref_etc1_compress_quality.restype = C.c_int

    @staticmethod
    def available() -> bool:
        return REF_SO.exists()

    def set_state(self, seed: int, wm_count: int = 0):
        self.lib.fastc_ref_set_state(seed & 0xFFFFFFFF, wm_count)

    def get_seed(self) -> int:
        return int(self.lib.fastc_ref_get_seed())

    def compress(self, fmt: str, img: np.ndarray, *, quality: int = 50, threads: int = 1,
                 job_size: int = 0, seed: int | None = 1, wm_count: int = 0):
        """seed=None leaves the reference's RNG/watermark globals alone (timing runs)."""
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w = img.shape[:2]
        out = np.zeros(nblocks(w, h) * BLOCK_BYTES[fmt], dtype=np.uint8)
        if seed is not None:
            self.set_state(seed, wm_count)
        ms = C.c_double(0)
        rc = self.lib.fastc_ref_compress(FMT[fmt], _p(img), w, h, _p(out), out.size, quality, threads,
                                         job_size, C.byref(ms))
        if rc != 0:
            raise RuntimeError("reference CompressImageData failed")
        return out, ms.value

    def write_ktx(self, fmt: str, cmp: np.ndarray, w: int, h: int, path) -> None:
        """The reference's ImageWriterKTX through ImageFile::Write."""
        cmp = np.ascontiguousarray(cmp, dtype=np.uint8)
        self.lib.fastc_ref_write_ktx.argtypes = [C.c_int, _u8p, C.c_uint32, C.c_uint32, C.c_char_p]
        if self.lib.fastc_ref_write_ktx(FMT[fmt], _p(cmp), w, h, str(path).encode()) != 0:
            raise RuntimeError("reference KTX write failed")

    def compress_etc1_quality(self, img: np.ndarray, quality: int) -> np.ndarray:
        """rg_etc1::pack_etc1_block over the image at cLow / cMedium / cHigh quality (0 / 1 / 2)."""
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w = img.shape[:2]
        out = np.zeros(nblocks(w, h) * 8, dtype=np.uint8)
        self.lib.fastc_ref_etc1_compress_quality(_p(img), w, h, _p(out), quality)
        return out

    def compress_bptc_settings(self, img: np.ndarray, *, quality: int = 0, block_modes: int = 0xFF,
                               error_metric: int = 0, seed: int = 1, wm_count: int = 0) -> np.ndarray:
        """BPTCC::Compress(job, settings) with m_BlockModes / m_ErrorMetric (single thread)."""
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w = img.shape[:2]
        out = np.zeros(nblocks(w, h) * 16, dtype=np.uint8)
        self.set_state(seed, wm_count)
        if self.lib.fastc_ref_bptc_compress_settings(_p(img), w, h, _p(out), quality, block_modes, error_metric) != 0:
            raise RuntimeError("reference BPTCC::Compress failed")
        return out

    def decode(self, fmt: str, cmp: np.ndarray, w: int, h: int) -> np.ndarray:
        cmp = np.ascontiguousarray(cmp, dtype=np.uint8)
        out = np.zeros((h, w, 4), dtype=np.uint8)
        if self.lib.fastc_ref_decompress(FMT[fmt], _p(cmp), w, h, _p(out)) != 0:
            raise RuntimeError("reference decode failed")
        return out

    def psnr(self, a: np.ndarray, b: np.ndarray) -> float:
        a = np.ascontiguousarray(a, dtype=np.uint8)
        b = np.ascontiguousarray(b, dtype=np.uint8)
        h, w = a.shape[:2]
        return float(self.lib.fastc_ref_psnr(_p(a), _p(b), w, h))
