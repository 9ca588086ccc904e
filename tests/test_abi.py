"""CPU (-m "not gpu"): the C-ABI library loads and exports every symbol that
include/fastc_gpu.h declares; no compute call is made (no GPU here)."""
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    from fastc_b200.api import GpuLibrary
    header = (ROOT / "include" / "fastc_gpu.h").read_text()
    declared = set(re.findall(r"\b(fastc_gpu_[a-z0-9_]+)\s*\(", header))
    declared -= {"fastc_gpu_format", "fastc_gpu_job", "fastc_gpu_timing"}
    assert declared, "no declarations parsed"
    g = GpuLibrary()
    for name in sorted(declared):
        assert hasattr(g.cdll, name), f"libfastc_gpu.so does not export {name}"
    assert set(GpuLibrary.SYMBOLS) == declared


def test_sizes_without_gpu():
    from fastc_b200.api import GpuLibrary, CompressedImage, ECompressionFormat as F
    g = GpuLibrary()
    for fmt, bs in ((F.DXT1, 8), (F.DXT5, 16), (F.ETC1, 8), (F.BPTC, 16)):
        assert g.cdll.fastc_gpu_block_bytes(int(fmt)) == bs
        assert g.cdll.fastc_gpu_compressed_size(int(fmt), 256, 128) == 64 * 32 * bs
        assert CompressedImage.GetCompressedSize(256, 128, fmt) == 64 * 32 * bs


def test_compress_image_data_error_paths(capsys):
    """Error behaviour of the reference's CompressImageData (TexComp.cpp:427-525)
    that needs no GPU: SIMD request, bad dimensions, short output buffer."""
    from fastc_b200 import CompressImageData, SCompressionSettings, ECompressionFormat as F
    img = np.zeros((8, 8, 4), dtype=np.uint8)
    out = np.zeros(64, dtype=np.uint8)
    s = SCompressionSettings(format=F.DXT1, bUseSIMD=True)
    assert CompressImageData(img, 8, 8, out, out.size, s) is False
    assert "TexComp -- Platform does not support SIMD!" in capsys.readouterr().err
    s = SCompressionSettings(format=F.DXT1)
    assert CompressImageData(img[:, :6], 6, 8, out, out.size, s) is False
    assert "width or height is not multiple of block dimension" in capsys.readouterr().err
    assert CompressImageData(img, 8, 8, out[:8], 8, s) is False
    assert "Not enough space" in capsys.readouterr().err


def test_no_gpu_means_loud_failure():
    """No CPU fallback: on a box without CUDA the compute entry points fail."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fastc_b200.api import GpuLibrary, FastcGpuError, ECompressionFormat as F
    g = GpuLibrary()
    with pytest.raises(FastcGpuError):
        g.compress(F.DXT1, np.zeros((8, 8, 4), dtype=np.uint8))
