"""ctypes mirror of the reference's Core API over libfastc_gpu.so.

Names, argument meaning and error behaviour follow the reference:
  * ECompressionFormat      reference/Base/include/FasTC/CompressionFormat.h
  * SCompressionSettings    reference/Core/include/FasTC/TexComp.h:30-76
  * CompressImageData       reference/Core/include/FasTC/TexComp.h:82-89,
                            reference/Core/src/TexComp.cpp:427-525
  * CompressedImage.GetCompressedSize  reference/Core/src/CompressedImage.cpp:136-149
The heavy lifting is the C ABI declared in include/fastc_gpu.h.
"""
from __future__ import annotations

import ctypes as C
import enum
import sys
from dataclasses import dataclass
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libfastc_gpu.so"


class FastcGpuError(RuntimeError):
    pass


class ECompressionFormat(enum.IntEnum):
    """Subset of FasTC::ECompressionFormat that has a GPU encoder; values are the
    C ABI's `fastc_gpu_format`."""
    DXT1 = 0
    DXT5 = 1
    ETC1 = 2
    BPTC = 3
    PVRTC4 = 4  # image-level: square power-of-two textures, whole texture per call, Morton block order


BLOCK_BYTES = {ECompressionFormat.DXT1: 8, ECompressionFormat.DXT5: 16,
               ECompressionFormat.ETC1: 8, ECompressionFormat.BPTC: 16, ECompressionFormat.PVRTC4: 8}


@dataclass
class SCompressionSettings:
    """Field-for-field mirror of the reference struct (TexComp.h:30-76); every
    field is initialised (the reference's ctor leaves five of them unset, D5)."""
    format: ECompressionFormat = ECompressionFormat.BPTC
    bUseSIMD: bool = False
    iNumThreads: int = 1
    iQuality: int = 50
    iNumCompressions: int = 1
    iJobSize: int = 0
    bUseAtomics: bool = False
    bUsePVRTexLib: bool = False
    bUseNVTT: bool = False
    logStream: object = None
    # extensions (not in the reference): GPU count (0 = all visible) and RNG seed
    iNumGPUs: int = 1
    seed: int = 0


class _Timing(C.Structure):
    _fields_ = [("kernel_ms", C.c_double), ("total_ms", C.c_double),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("kernel_launches", C.c_uint32)]


class _Options(C.Structure):
    """fastc_gpu_options (include/fastc_gpu.h): BPTCC::CompressionSettings' m_BlockModes /
    m_ErrorMetric and rg_etc1's quality level."""
    _fields_ = [("struct_size", C.c_uint32), ("bptc_block_modes", C.c_uint32),
                ("bptc_error_metric", C.c_int32), ("etc1_quality", C.c_int32),
                ("bptc_block_stats", C.c_void_p)]


def _options(block_modes: int = 0xFF, error_metric: int = 0, etc1_quality: int = 0, block_stats=None):
    return _Options(C.sizeof(_Options), block_modes, error_metric, etc1_quality,
                    block_stats.ctypes.data if block_stats is not None else None)


class _Job(C.Structure):
    _fields_ = [("rgba_host", C.c_void_p), ("out_host", C.c_void_p),
                ("width", C.c_uint32), ("height", C.c_uint32)]


class GpuLibrary:
    """Loads libfastc_gpu.so and types every symbol include/fastc_gpu.h declares."""

    SYMBOLS = [
        "fastc_gpu_device_count", "fastc_gpu_init", "fastc_gpu_shutdown", "fastc_gpu_block_bytes",
        "fastc_gpu_compressed_size", "fastc_gpu_compress", "fastc_gpu_compress_batch",
        "fastc_gpu_compress_device", "fastc_gpu_count_solid_device", "fastc_gpu_bc7_counters",
        "fastc_gpu_compress_opt", "fastc_gpu_compress_batch_opt", "fastc_gpu_compress_device_opt",
        "fastc_gpu_debug_bc7_dump", "fastc_gpu_bc7_stage_ms",
        "fastc_gpu_decompress", "fastc_gpu_decompress_device", "fastc_gpu_psnr", "fastc_gpu_psnr_device",
        "fastc_gpu_last_error",
    ]

    def __init__(self, path: Path = LIB_PATH):
        if not path.exists():
            raise FastcGpuError(
                f"{path} is missing: build it with `make gpu` (or __graft_entry__.build()). "
                "There is no CPU fallback.")
        L = self.cdll = C.CDLL(str(path))
        vp, u32, u64, i = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
        L.fastc_gpu_device_count.restype = i
        L.fastc_gpu_init.argtypes = [i]
        L.fastc_gpu_shutdown.restype = None
        L.fastc_gpu_block_bytes.argtypes = [i]
        L.fastc_gpu_block_bytes.restype = u32
        L.fastc_gpu_compressed_size.argtypes = [i, u32, u32]
        L.fastc_gpu_compressed_size.restype = u64
        L.fastc_gpu_compress.argtypes = [i, vp, u32, u32, u32, u32, vp, i, u64, u32, i, C.POINTER(_Timing)]
        L.fastc_gpu_compress_batch.argtypes = [i, C.POINTER(_Job), u32, i, u64, i, C.POINTER(_Timing)]
        L.fastc_gpu_compress_device.argtypes = [i, vp, u32, u32, u32, u32, vp, i, u64, u32, u32, vp,
                                                C.POINTER(u32)]
        po = C.POINTER(_Options)
        L.fastc_gpu_compress_opt.argtypes = L.fastc_gpu_compress.argtypes + [po]
        L.fastc_gpu_compress_batch_opt.argtypes = L.fastc_gpu_compress_batch.argtypes + [po]
        L.fastc_gpu_compress_device_opt.argtypes = L.fastc_gpu_compress_device.argtypes + [po]
        L.fastc_gpu_count_solid_device.argtypes = [vp, u32, u32, u32, u32, vp, C.POINTER(u32)]
        L.fastc_gpu_bc7_counters.argtypes = [C.POINTER(u64), C.POINTER(u64)]
        L.fastc_gpu_debug_bc7_dump.argtypes = [u32, vp, vp]
        L.fastc_gpu_bc7_stage_ms.argtypes = [i, C.POINTER(C.c_double)]
        L.fastc_gpu_decompress.argtypes = [i, vp, u32, u32, vp, C.POINTER(_Timing)]
        L.fastc_gpu_decompress_device.argtypes = [i, vp, u32, u32, vp, vp]
        L.fastc_gpu_psnr.argtypes = [vp, vp, u32, u32, C.POINTER(C.c_double)]
        L.fastc_gpu_psnr_device.argtypes = [vp, vp, u32, u32, vp, C.POINTER(C.c_double)]
        L.fastc_gpu_last_error.restype = C.c_char_p

    def error(self) -> str:
        return (self.cdll.fastc_gpu_last_error() or b"").decode()

    def check(self, rc: int):
        if rc != 0:
            raise FastcGpuError(self.error())

    # ---- host -> host -------------------------------------------------------
    @staticmethod
    def _check_image(rgba) -> tuple[int, int]:
        if (not isinstance(rgba, np.ndarray) or rgba.dtype != np.uint8 or rgba.ndim != 3 or rgba.shape[2] != 4
                or not rgba.flags.c_contiguous):
            raise FastcGpuError("rgba must be a C-contiguous (H, W, 4) uint8 array")
        return rgba.shape[0], rgba.shape[1]

    @staticmethod
    def _check_output(out, size: int):
        if not isinstance(out, np.ndarray) or out.dtype != np.uint8 or not out.flags.c_contiguous:
            raise FastcGpuError("the output must be a C-contiguous uint8 array")
        if out.nbytes < size:
            # reference: "Not enough space for compressed data!" (TexComp.cpp:493-496)
            raise FastcGpuError("Not enough space for compressed data!")

    def compress(self, fmt: int, rgba: np.ndarray, out: np.ndarray | None = None, *, quality: int = 50,
                 seed: int = 0, first_block: int = 0, num_blocks: int = 0, chunk_blocks: int = 0,
                 num_gpus: int = 1, block_modes: int = 0xFF, error_metric: int = 0, etc1_quality: int = 0,
                 block_stats: np.ndarray | None = None):
        """rgba: (H, W, 4) uint8, C-contiguous (pinned or pageable host memory).
        block_stats (BPTC): optional (blocks, 10) float64 array receiving per block the mode, the path and
        the error of each mode tried (fastc_gpu_bptc_block_stat).  Returns (out bytes, timing dict)."""
        h, w = self._check_image(rgba)
        if block_stats is not None and (block_stats.dtype != np.float64 or not block_stats.flags.c_contiguous
                                        or block_stats.shape != ((w // 4) * (h // 4), 10)):
            raise FastcGpuError("block_stats must be a C-contiguous (blocks, 10) float64 array")
        size = int(self.cdll.fastc_gpu_compressed_size(int(fmt), w, h))
        if out is None:
            out = np.zeros(size, dtype=np.uint8)
        else:
            self._check_output(out, size)
        tm = _Timing()
        opt = _options(block_modes, error_metric, etc1_quality, block_stats)
        self.check(self.cdll.fastc_gpu_compress_opt(int(fmt), rgba.ctypes.data, w, h, first_block, num_blocks,
                                                    out.ctypes.data, quality, seed, chunk_blocks, num_gpus,
                                                    C.byref(tm), C.byref(opt)))
        return out, {"kernel_ms": tm.kernel_ms, "total_ms": tm.total_ms, "h2d_bytes": tm.h2d_bytes,
                     "d2h_bytes": tm.d2h_bytes, "kernel_launches": tm.kernel_launches}

    def compress_batch(self, fmt: int, images: list[np.ndarray], *, quality: int = 50, seed: int = 0,
                       num_gpus: int = 1, outs: list[np.ndarray] | None = None, block_modes: int = 0xFF,
                       error_metric: int = 0, etc1_quality: int = 0):
        """One submission for a list of textures (fastc_gpu_compress_batch).  `outs`: optional
        preallocated output arrays (e.g. pinned memory), one per texture."""
        jobs = (_Job * len(images))()
        given, outs = outs, []
        if given is not None and len(given) != len(images):
            raise FastcGpuError("one output array per texture")
        for k, im in enumerate(images):
            h, w = self._check_image(im)
            size = int(self.cdll.fastc_gpu_compressed_size(int(fmt), w, h))
            if given is not None:
                o = given[k]
                self._check_output(o, size)
            else:
                o = np.zeros(size, dtype=np.uint8)
            outs.append(o)
            jobs[k] = _Job(im.ctypes.data, o.ctypes.data, w, h)
        tm = _Timing()
        opt = _options(block_modes, error_metric, etc1_quality)
        self.check(self.cdll.fastc_gpu_compress_batch_opt(int(fmt), jobs, len(images), quality, seed, num_gpus,
                                                          C.byref(tm), C.byref(opt)))
        return outs, {"kernel_ms": tm.kernel_ms, "total_ms": tm.total_ms, "h2d_bytes": tm.h2d_bytes,
                      "d2h_bytes": tm.d2h_bytes, "kernel_launches": tm.kernel_launches}

    # ---- device -> device (torch tensors are only the memory/stream plumbing) --
    def compress_device(self, fmt: int, rgba_dev, out_dev, *, width: int, height: int, quality: int = 50,
                        seed: int = 0, first_block: int = 0, num_blocks: int = 0, wm_base: int = 0,
                        block_index_base: int = 0, stream: int | None = None, block_modes: int = 0xFF,
                        error_metric: int = 0, etc1_quality: int = 0) -> int:
        """rgba_dev / out_dev: CUDA torch uint8 tensors on the current device.
        Asynchronous on `stream` (raw cudaStream_t; default torch's current)."""
        import torch
        if stream is None:
            stream = torch.cuda.current_stream().cuda_stream
        n = C.c_uint32(0)
        opt = _options(block_modes, error_metric, etc1_quality)
        self.check(self.cdll.fastc_gpu_compress_device_opt(int(fmt), rgba_dev.data_ptr(), width, height, first_block,
                                                           num_blocks, out_dev.data_ptr(), quality, seed, wm_base,
                                                           block_index_base, stream, C.byref(n), C.byref(opt)))
        return n.value

    def count_solid_device(self, rgba_dev, *, width: int, height: int, first_block: int = 0,
                           num_blocks: int = 0, stream: int | None = None) -> int:
        import torch
        if stream is None:
            stream = torch.cuda.current_stream().cuda_stream
        if num_blocks == 0:
            num_blocks = (width // 4) * (height // 4) - first_block
        n = C.c_uint32(0)
        self.check(self.cdll.fastc_gpu_count_solid_device(rgba_dev.data_ptr(), width, height, first_block,
                                                          num_blocks, stream, C.byref(n)))
        return n.value


    # ---- decoders + PSNR (the step after the encode path) -----------------------
    def decompress(self, fmt: int, cmp: np.ndarray, width: int, height: int) -> np.ndarray:
        cmp = np.ascontiguousarray(cmp, dtype=np.uint8)
        need = int(self.cdll.fastc_gpu_compressed_size(int(fmt), width, height))
        if cmp.nbytes < need:
            raise FastcGpuError("compressed buffer smaller than the image needs")
        out = np.empty((height, width, 4), dtype=np.uint8)
        self.check(self.cdll.fastc_gpu_decompress(int(fmt), cmp.ctypes.data, width, height, out.ctypes.data, None))
        return out

    def decompress_device(self, fmt: int, cmp_dev, rgba_dev, *, width: int, height: int, stream: int | None = None):
        import torch
        if stream is None:
            stream = torch.cuda.current_stream().cuda_stream
        self.check(self.cdll.fastc_gpu_decompress_device(int(fmt), cmp_dev.data_ptr(), width, height,
                                                         rgba_dev.data_ptr(), stream))

    def psnr(self, a: np.ndarray, b: np.ndarray) -> float:
        a = np.ascontiguousarray(a, dtype=np.uint8)
        b = np.ascontiguousarray(b, dtype=np.uint8)
        if a.shape != b.shape:
            raise FastcGpuError("PSNR needs two images of the same size")
        h, w = a.shape[:2]
        r = C.c_double(0)
        self.check(self.cdll.fastc_gpu_psnr(a.ctypes.data, b.ctypes.data, w, h, C.byref(r)))
        return r.value

    def psnr_device(self, a_dev, b_dev, *, width: int, height: int, stream: int | None = None) -> float:
        import torch
        if stream is None:
            stream = torch.cuda.current_stream().cuda_stream
        r = C.c_double(0)
        self.check(self.cdll.fastc_gpu_psnr_device(a_dev.data_ptr(), b_dev.data_ptr(), width, height, stream,
                                                   C.byref(r)))
        return r.value

    def bc7_stage_ms(self, enable: bool = True, read: bool = True):
        """Arms / reads the per-stage CUDA-event timing of the BC7 pipeline (ms):
        {classify+scan, select, setup(+sort), anneal, pack, total} of the last device-API call."""
        arr = (C.c_double * 6)()
        self.check(self.cdll.fastc_gpu_bc7_stage_ms(int(enable), arr if read else None))
        return dict(zip(("classify", "select", "setup", "anneal", "pack", "total"), arr)) if read else None


_lib: GpuLibrary | None = None


def lib() -> GpuLibrary:
    global _lib
    if _lib is None:
        _lib = GpuLibrary()
    return _lib


class CompressedImage:
    """Owns compressed bytes (reference/Core/include/FasTC/CompressedImage.h)."""

    def __init__(self, width: int, height: int, fmt: ECompressionFormat, data: np.ndarray):
        self.width, self.height, self.format, self.data = width, height, fmt, data

    @staticmethod
    def GetCompressedSize(width: int, height: int, fmt: ECompressionFormat) -> int:
        return ((width + 3) // 4) * ((height + 3) // 4) * BLOCK_BYTES[ECompressionFormat(fmt)]


def _report_error(msg: str):
    # reference: fprintf(stderr, "TexComp -- %s\n", msg) (TexComp.cpp:157-159)
    print(f"TexComp -- {msg}", file=sys.stderr)


def CompressImageData(data: np.ndarray, width: int, height: int, cmpData: np.ndarray, cmpDataSz: int,
                      settings: SCompressionSettings) -> bool:
    """Same contract as the reference: returns False (after printing
    `TexComp -- <msg>` to stderr) on failure, prints `Compression time: %0.3f ms`
    to stdout on success.  `data`: width*height*4 RGBA bytes; `cmpData`: output."""
    if settings.bUseSIMD:
        _report_error("Platform does not support SIMD!\n")  # TexComp.cpp:440-445 (D7)
        return False
    try:
        fmt = ECompressionFormat(settings.format)
    except ValueError:
        _report_error("Unknown compression format")
        return False
    if width % 4 or height % 4 or width == 0 or height == 0:
        _report_error("ERROR - CompressImageData: width or height is not multiple of block dimension")  # TexComp.cpp:472-476
        return False
    if cmpDataSz < CompressedImage.GetCompressedSize(width, height, fmt):
        _report_error("Not enough space for compressed data!")  # TexComp.cpp:493-496
        return False
    img = np.ascontiguousarray(data, dtype=np.uint8).reshape(height, width, 4)
    n = max(1, settings.iNumCompressions)
    total = 0.0
    try:
        for _ in range(n):
            _, tm = lib().compress(fmt, img, cmpData, quality=max(0, settings.iQuality), seed=settings.seed,
                                   chunk_blocks=max(0, settings.iJobSize), num_gpus=max(0, settings.iNumGPUs))
            total += tm["total_ms"]
    except FastcGpuError as e:
        _report_error(str(e))
        return False
    print("Compression time: %0.3f ms" % (total / n))  # TexComp.cpp:517
    return True
