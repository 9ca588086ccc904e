"""The C++ host layer (include/FasTC/*.h, fastc_b200/core/): FasTC's Core API and the `tc`
CLI over the C ABI.  CPU tests cover job arithmetic, the job list, error behaviour and the
CLI's argument handling; GPU tests drive CompressImageData / the per-format CompressionFunc
entry points / CompressImageList / CompressedImage / `tc` and compare with the oracle."""
import os
import sys
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
    r = _run([TC, "-f", "PVRTCLib", tmp_path / "a.tga"])
    assert r.returncode == 1 and "not supported on the GPU path" in r.stderr
    r = _run([TC, "-f", "PVRTC", "-nd", tmp_path / "a.tga"])  # 16 x 8: the reference's own size check
    assert r.returncode == 1 and "PVRTC4 images must be square and power-of-two" in r.stderr
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
    # PNG input (BASELINE configs[0] names a PNG): same payload as from the TGA
    (tmp_path / "img.png").write_bytes(_png_bytes(64, 48, 6, img.reshape(48, 64 * 4).tolist(), [4, 1, 2]))
    r = _run([TC, "-f", fmt, "-q", "0", "-d", tmp_path / "frompng.ktx", tmp_path / "img.png"])
    assert r.returncode == 0, r.stdout + r.stderr
    assert (tmp_path / "frompng.ktx").read_bytes()[96:96 + want.size] == want.tobytes()


def _png_bytes(w, h, ctype, rows, filters, plte=None, interlace=0, depth=8):
    """Hand-rolled PNG writer (per-row filter types chosen by the test)."""
    import zlib
    ch = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]

    def paeth(a, b, c):
        p = a + b - c
        pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
        return a if pa <= pb and pa <= pc else (b if pb <= pc else c)

    raw = bytearray()
    prev = [0] * (w * ch)
    for j in range(h):
        cur = list(rows[j])
        f = filters[j % len(filters)]
        raw.append(f)
        for i, v in enumerate(cur):
            a = cur[i - ch] if i >= ch else 0
            b = prev[i]
            c = prev[i - ch] if i >= ch else 0
            pred = (0, a, b, (a + b) >> 1, paeth(a, b, c))[f]
            raw.append((v - pred) & 0xFF)
        prev = cur

    def chunk(tag, body):
        return struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body) & 0xFFFFFFFF)

    z = zlib.compress(bytes(raw), 6)
    out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, interlace))
    if plte is not None:
        out += chunk(b"PLTE", bytes(plte))
    half = len(z) // 2   # two IDAT chunks: the loader must concatenate them
    return out + chunk(b"IDAT", z[:half]) + chunk(b"IDAT", z[half:]) + chunk(b"IEND", b"")


def _read_tga(path):
    d = path.read_bytes()
    w, h = struct.unpack_from("<HH", d, 12)
    px = np.frombuffer(d, dtype=np.uint8, count=w * h * 4, offset=18 + d[0]).reshape(h, w, 4)
    return px[::-1, :, [2, 1, 0, 3]]     # rows bottom-up, BGRA


@pytest.mark.parametrize("ctype", [0, 2, 3, 4, 6])
def test_png_loader_colour_types_and_filters(tmp_path, ctype):
    """PNG input (BASELINE configs[0] names a PNG): 8-bit grey / RGB / palette / grey+alpha / RGBA,
    every filter type, split IDAT -- the coverage of the reference's libpng loader
    (IO/src/ImageLoaderPNG.cpp:58-260).  Checked through ImageFile::Load -> ImageFile::Write (TGA)."""
    rng = np.random.default_rng(ctype)
    w, h = 20, 12
    ch = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]
    data = rng.integers(0, 256, (h, w * ch), dtype=np.uint8)
    plte = None
    if ctype == 3:
        data %= 7
        plte = rng.integers(0, 256, 7 * 3, dtype=np.uint8)
    src = tmp_path / "in.png"
    src.write_bytes(_png_bytes(w, h, ctype, data.tolist(), [0, 1, 2, 3, 4], plte=None if plte is None else plte.tolist()))
    r = _run([SELFTEST, "convert", src, tmp_path / "out.tga"])
    assert r.returncode == 0, r.stdout + r.stderr
    got = _read_tga(tmp_path / "out.tga")
    d = data.reshape(h, w, ch).astype(np.uint8)
    want = np.full((h, w, 4), 255, dtype=np.uint8)
    if ctype == 0:
        want[..., :3] = d
    elif ctype == 2:
        want[..., :3] = d
    elif ctype == 3:
        want[..., :3] = plte.reshape(7, 3)[d[..., 0]]
    elif ctype == 4:
        want[..., :3] = d[..., :1]
        want[..., 3] = d[..., 1]
    else:
        want = d
    assert (got == want).all()


def test_png_loader_rejects_what_the_reference_rejects(tmp_path):
    bad16 = tmp_path / "d16.png"
    bad16.write_bytes(_png_bytes(4, 4, 0, [[0] * 4] * 4, [0], depth=16))
    r = _run([SELFTEST, "convert", bad16, tmp_path / "o.tga"])
    assert r.returncode != 0 and "Only 8-bit images currently supported." in r.stderr
    notpng = tmp_path / "x.png"
    notpng.write_bytes(b"not a png at all")
    r = _run([SELFTEST, "convert", notpng, tmp_path / "o.tga"])
    assert r.returncode != 0 and "Incorrect PNG signature" in r.stderr
    inter = tmp_path / "i.png"
    inter.write_bytes(_png_bytes(4, 4, 2, [[0] * 12] * 4, [0], interlace=1))
    r = _run([SELFTEST, "convert", inter, tmp_path / "o.tga"])
    assert r.returncode != 0


def test_png_write_then_load_roundtrip(tmp_path):
    img = synth_rgba(64, 32, 3)
    write_tga(tmp_path / "a.tga", img)
    assert _run([SELFTEST, "convert", tmp_path / "a.tga", tmp_path / "b.png"]).returncode == 0
    assert _run([SELFTEST, "convert", tmp_path / "b.png", tmp_path / "c.tga"]).returncode == 0
    assert (_read_tga(tmp_path / "c.tga") == img).all()
    try:
        from PIL import Image as PILImage   # an independent decoder agrees with the writer
    except ImportError:
        return
    assert (np.asarray(PILImage.open(tmp_path / "b.png").convert("RGBA")) == img).all()


@pytest.mark.parametrize("fmt", ["BPTC", "DXT1", "DXT5", "PVRTC4"])
def test_ktx_file_is_byte_identical_to_the_reference_writer(tmp_path, fmt):
    """a whole .ktx written by our ImageFile (what `tc -d out.ktx` calls) equals, byte
    for byte, the file the reference's ImageWriterKTX (IO/src/ImageWriterKTX.cpp:69-160) writes for
    the same payload: identifier, endianness, GL enums, dimensions, the KTXorientation key/value block
    with its padding, imageSize and the payload.  CPU only."""
    from _checkers import Reference, BLOCK_BYTES
    if not Reference.available():
        pytest.skip("oracle/_ref/libfastc_ref.so not built")
    w, h = (64, 64) if fmt == "PVRTC4" else (64, 32)
    payload = np.random.default_rng(3).integers(0, 256, (w // 4) * (h // 4) * BLOCK_BYTES[fmt], dtype=np.uint8)
    (tmp_path / "p.bin").write_bytes(payload.tobytes())
    r = subprocess.run([str(SELFTEST), "writektx", fmt, str(w), str(h), str(tmp_path / "p.bin"), str(tmp_path / "ours.ktx")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # (a fresh process per file: the reference writer caches imageSize in function-local statics)
    code = ("import sys, numpy as np; sys.path.insert(0, sys.argv[1]); from _checkers import Reference; "
            "Reference().write_ktx(sys.argv[2], np.fromfile(sys.argv[3], dtype=np.uint8), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6])")
    subprocess.run([sys.executable, "-c", code, str(ROOT / "tests"), fmt, str(tmp_path / "p.bin"), str(w), str(h),
                    str(tmp_path / "ref.ktx")], check=True)
    ours, ref = (tmp_path / "ours.ktx").read_bytes(), (tmp_path / "ref.ktx").read_bytes()
    assert len(ours) == 96 + payload.size
    assert ours == ref


@pytest.mark.gpu
def test_tc_l_writes_the_per_block_log(gpu, oracle, tmp_path):
    """`tc -l`: "<basename>.log" with the reference's per-block statistics lines
    ("<block>: BlockStat_Mode -- m", ... reference BPTCEncoder/src/Compressor.cpp:106-131, :1947-1992);
    the compressed output is the same as without -l."""
    img = synth_rgba(64, 48, 1, full_height=256, y0=40)
    write_tga(tmp_path / "img.tga", img)
    r = _run([TC, "-f", "BPTC", "-q", "0", "-l", "-d", "out.ktx", "img.tga"], cwd=tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    want, _ = oracle.compress("BPTC", img, quality=0)
    assert (tmp_path / "out.ktx").read_bytes()[96:96 + want.size] == want.tobytes()
    log = (tmp_path / "img.log").read_text().splitlines()
    nblk = 16 * 12
    assert len(log) == nblk * 18
    b0 = want.reshape(-1, 16)[:, 0].astype(np.int64)
    modes = np.log2(b0 & -b0).astype(int)
    for i in (0, 17, nblk - 1):
        blk = [ln for ln in log if ln.startswith(f"{i}: ")]
        assert len(blk) == 18
        assert blk[1] == f"{i}: BlockStat_Mode -- {modes[i]}"
        assert blk[0].startswith(f"{i}: BlockStat_Path -- ")
        assert blk[2] == f"{i}: BlockStat_ModeZeroEstimate -- -1"


@pytest.mark.gpu
def test_tc_cli_pvrtc_end_to_end(gpu, tmp_path):
    """tc -f PVRTC: compress, decode, PSNR line, decoded PNG written (the decode equals the reference's decoder
    on the reference's blocks, tests/test_gpu_pvrtc.py)."""
    img = synth_rgba(128, 128, 2)
    write_tga(tmp_path / "p.tga", img)
    out = tmp_path / "p-dec.png"
    r = _run([TC, "-f", "PVRTC", "-d", out, tmp_path / "p.tga"])
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Compression time: " in r.stdout and "PSNR: " in r.stdout and out.exists()
    psnr = float(r.stdout.split("PSNR: ")[1].split()[0])
    assert 15.0 < psnr < 60.0
