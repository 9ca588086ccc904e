"""The C++ host layer (include/FasTC/*.h, fastc_b200/core/): FasTC's Core API and the `tc`
CLI over the C ABI.  CPU tests cover job arithmetic, the job list, error behaviour and the
CLI's argument handling; GPU tests drive CompressImageData / the per-format CompressionFunc
entry points / CompressImageList / CompressedImage / `tc` and compare with the oracle."""
import os
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest

from _checkers import BLOCK_BYTES
from fastc_b200.synth import synth_rgba

ROOT = Path(__file__).resolve().parent.parent
CORE = ROOT / "fastc_b200" / "core"
TC, SELFTEST = CORE / "tc", CORE / "core_selftest"


def _run(cmd, **kw):
    return subprocess.run([str(c) for c in cmd], capture_output=True, text=True, timeout=600, **kw)


def write_tga(path: Path, img: np.ndarray):
    """32-bit uncompressed TGA, rows bottom-up (what the reference's loader flips back)."""
    h, w = img.shape[:2]
    hdr = struct.pack("<BBBHHBHHHHBB", 0, 0, 2, 0, 0, 0, 0, 0, w, h, 32, 8)
    bgra = img[::-1, :, [2, 1, 0, 3]]
    path.write_bytes(hdr + np.ascontiguousarray(bgra).tobytes())


def test_binaries_built():
    assert TC.exists() and SELFTEST.exists(), "run `make core` (or __graft_entry__.build())"


def test_core_api_selftest():
    r = _run([SELFTEST, "api"])
    assert r.returncode == 0, r.stdout + r.stderr
    assert "api: ok" in r.stdout
    for msg in ("TexComp -- Platform does not support SIMD!",
                "TexComp -- ERROR - CompressImageData: width or height is not multiple of block dimension",
                "TexComp -- Not enough space for compressed data!",
                "TexComp -- No data sent to compress!",
                "TexComp -- Could not find adequate compression function for specified settings"):
        assert msg in r.stderr


def test_tc_cli_argument_handling(tmp_path):
    r = _run([TC])
    assert r.returncode == 1 and "Usage: tc [OPTIONS] imagefile" in r.stderr
    assert _run([TC, "-h"]).returncode == 0
    for bad in (["-q"], ["-q", "-3", "x.tga"], ["-t", "0", "x.tga"], ["-n", "-1", "x.tga"], ["-f"], ["-d"],
                ["-q", "5"]):
        r = _run([TC] + bad)
        assert r.returncode == 1 and "Usage" in r.stderr, bad
    assert _run([TC, tmp_path / "missing.tga"]).returncode == 1
    img = synth_rgba(16, 8, 1)
    write_tga(tmp_path / "a.tga", img)
    r = _run([TC, "-f", "PVRTC", tmp_path / "a.tga"])
    assert r.returncode == 1 and "not supported on the GPU path" in r.stderr
    r = _run([TC, "-simd", "-nd", tmp_path / "a.tga"])  # rejected before any GPU work
    assert r.returncode == 1 and "Platform does not support SIMD!" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("fmt,q", [("DXT1", 0), ("DXT5", 0), ("ETC1", 0), ("BPTC", 0), ("BPTC", 4)])
def test_core_api_on_gpu_matches_oracle(gpu, oracle, tmp_path, fmt, q):
    w, h = 128, 64
    img = synth_rgba(512, 512, 1)[128:192, 40:168] if fmt == "BPTC" else synth_rgba(w, h, 3)
    img = np.ascontiguousarray(img)
    (tmp_path / "in.raw").write_bytes(img.tobytes())
    r = _run([SELFTEST, "gpu", tmp_path / "in.raw", w, h, fmt, q, 0, tmp_path / "o"])
    assert r.returncode == 0 and "gpu: ok" in r.stdout, r.stdout + r.stderr
    assert r.stdout.count("Compression time: ") == 2  # CompressImageData + CompressImageList
    want, _ = oracle.compress(fmt, img, quality=q, rng_mode=1, seed=0)
    whole = np.fromfile(tmp_path / "o.whole", dtype=np.uint8)
    assert (whole == want).all()
    assert (np.fromfile(tmp_path / "o.split", dtype=np.uint8) == want).all()   # CompressionFunc x3 == one call
    assert (np.fromfile(tmp_path / "o.list0", dtype=np.uint8) == want).all()
    if q == 0:
        top, _ = oracle.compress(fmt, np.ascontiguousarray(img[:h // 2]), quality=0)
        assert (np.fromfile(tmp_path / "o.list1", dtype=np.uint8) == top).all()
    dec = oracle.decode(fmt, want, w, h)
    assert (np.fromfile(tmp_path / "o.dec", dtype=np.uint8).reshape(h, w, 4) == dec).all()
    psnr = float(r.stdout.split("PSNR: ")[1].split()[0])
    assert abs(psnr - oracle.psnr(img, dec)) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["BPTC", "DXT1", "DXT5", "ETC1"])
def test_tc_cli_end_to_end(gpu, oracle, tmp_path, fmt):
    """`tc -f FMT -q 0 -d out.ktx in.tga`: the two stdout lines of the reference CLI, and the
    compressed payload at byte 96 of the KTX (reference IO/src/ImageWriterKTX.cpp:69-160)."""
    img = synth_rgba(64, 48, 1, full_height=256, y0=40)
    write_tga(tmp_path / "img.tga", img)
    r = _run([TC, "-f", fmt, "-q", "0", "-t", "8", "-j", "32", "-n", "2", "-d", tmp_path / "out.ktx",
              tmp_path / "img.tga"])
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().splitlines()
    assert lines[0].startswith("Compression time: ") and lines[0].endswith(" ms")
    want, _ = oracle.compress(fmt, img, quality=0)
    psnr = oracle.psnr(img, oracle.decode(fmt, want, 64, 48))
    assert lines[1] == "PSNR: %.3f" % psnr
    ktx = (tmp_path / "out.ktx").read_bytes()
    assert ktx[:12] == bytes([0xAB, 0x4B, 0x54, 0x58, 0x20, 0x31, 0x31, 0xBB, 0x0D, 0x0A, 0x1A, 0x0A])
    assert struct.unpack_from("<I", ktx, 92)[0] == want.size
    assert ktx[96:96 + want.size] == want.tobytes()
    # default output name: <basename>-<fmt>.png in the working directory, decoded pixels
    r = _run([TC, "-f", fmt, "-q", "0", tmp_path / "img.tga"], cwd=tmp_path)
    assert r.returncode == 0
    png = tmp_path / f"img-{fmt.lower()}.png"
    assert png.exists() and png.read_bytes()[:8] == b"\x89PNG\r\n\x1a\n"
    # a compressed KTX loads back and decodes to the same pixels (written as TGA)
    r = _run([TC, "-f", "DXT1", "-q", "0", "-d", tmp_path / "again.tga", tmp_path / "out.ktx"])
    assert r.returncode == 0, r.stdout + r.stderr
