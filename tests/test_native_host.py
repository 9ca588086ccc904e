"""The product's block arithmetic, compiled for the HOST and compared with the oracle on the CPU.

`fastc_b200/csrc/dxt_block.cuh` is host+device code: the kernel runs it per thread, the checker under
`tests/native/` runs the same functions in a plain g++ build (`-ffp-contract=off`, the device build uses
explicit round-to-nearest intrinsics) over whole images.  This is a CPU-side net under the GPU parity
tests (`tests/test_gpu_dxt.py`), not a product path: the library never calls these functions on the host.
"""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from fastc_b200.synth import synth_rgba

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def dxt_check(tmp_path_factory):
    exe = tmp_path_factory.mktemp("native") / "dxt_host_check"
    oracle = ROOT / "oracle"
    if not (oracle / "libfastc_oracle.so").exists():
        subprocess.run(["make", "-s", "-C", str(oracle), "oracle"], check=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", f"-I{ROOT / 'fastc_b200' / 'csrc'}",
                    str(ROOT / "tests" / "native" / "dxt_host_check.cpp"), str(oracle / "libfastc_oracle.so"),
                    f"-Wl,-rpath,{oracle}", "-o", str(exe)], check=True)
    return exe


def _styled_images():
    rng = np.random.default_rng(11)
    yield "synthetic 256^2", synth_rgba(256, 256, 1)
    yield "uniform noise", rng.integers(0, 256, (128, 128, 4), dtype=np.uint8)
    flat = np.zeros((64, 64, 4), np.uint8)  # tiny variance: the power iteration's luminance fallback
    flat[..., :3] = rng.integers(100, 104, (64, 64, 3))
    flat[..., 3] = rng.integers(0, 256, (64, 64))
    yield "low variance", flat
    two = np.where(rng.integers(0, 2, (64, 64, 1)) > 0, np.array([255, 0, 10, 255]), np.array([0, 255, 200, 3]))
    yield "two colours", two.astype(np.uint8)
    sat = rng.choice(np.array([0, 1, 254, 255], np.uint8), (64, 64, 4))
    yield "saturated", sat
    solid = np.zeros((32, 32, 4), np.uint8)  # constant blocks incl. ones that differ in alpha only
    solid[...] = rng.integers(0, 256, (8, 1, 8, 1, 4), dtype=np.uint8).repeat(4, 1).repeat(4, 3).reshape(32, 32, 4)
    solid[5, 7, 3] ^= 1
    yield "solid blocks", solid
    ramp = np.zeros((64, 64, 4), np.uint8)
    ramp[..., 0] = np.arange(64)[None, :] * 4
    ramp[..., 1] = np.arange(64)[:, None] * 4
    ramp[..., 2] = 255 - ramp[..., 0]
    ramp[..., 3] = (np.arange(64)[None, :] * 3 + np.arange(64)[:, None]) % 256
    yield "ramps", ramp


@pytest.mark.parametrize("dxt5", [0, 1])
def test_dxt_block_code_on_host_matches_oracle(dxt_check, dxt5):
    for name, img in _styled_images():
        img = np.ascontiguousarray(img, dtype=np.uint8)
        r = subprocess.run([str(dxt_check), str(dxt5), str(img.shape[1]), str(img.shape[0])], input=img.tobytes(),
                           capture_output=True)
        assert r.returncode == 0, f"{name}: {r.stdout.decode().strip()} {r.stderr.decode()[-200:]}"


# ---- PVRTC: fastc_b200/csrc/pvrtc_block.cuh on the host, in the reference's raster order, against the
# compiled reference (PVRTCC::Compress through oracle/_ref/libfastc_ref.so)
REF_SO = ROOT / "oracle" / "_ref" / "libfastc_ref.so"


@pytest.fixture(scope="module")
def pvrtc_check(tmp_path_factory):
    if not REF_SO.exists():
        pytest.skip("oracle/_ref/libfastc_ref.so missing (built only where /root/reference exists)")
    exe = tmp_path_factory.mktemp("native") / "pvrtc_host_check"
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", f"-I{ROOT / 'fastc_b200' / 'csrc'}",
                    str(ROOT / "tests" / "native" / "pvrtc_host_check.cpp"), str(REF_SO),
                    f"-Wl,-rpath,{REF_SO.parent}", "-o", str(exe)], check=True)
    return exe


def pvrtc_images():
    rng = np.random.default_rng(9)
    for n in (8, 16, 32):
        yield f"noise {n}", rng.integers(0, 256, (n, n, 4), dtype=np.uint8)
    yield "synthetic 256^2", synth_rgba(256, 256, 1)
    flat = np.zeros((64, 64, 4), np.uint8)
    flat[..., :3] = rng.integers(100, 104, (64, 64, 3))
    flat[..., 3] = 255
    yield "low variance, opaque", flat
    yield "solid", np.full((64, 64, 4), 77, np.uint8)
    ramp = np.zeros((128, 128, 4), np.uint8)
    ramp[..., 0] = np.arange(128)[None, :] * 2
    ramp[..., 1] = np.arange(128)[:, None] * 2
    ramp[..., 2] = 255 - ramp[..., 0]
    ramp[..., 3] = (np.arange(128)[None, :] * 3 + np.arange(128)[:, None]) % 256
    yield "ramps with alpha", ramp
    trans = rng.integers(0, 256, (64, 64, 4), dtype=np.uint8)
    trans[..., 3] = rng.choice(np.array([0, 255, 199, 200, 201, 224, 30], np.uint8), (64, 64))  # around the < 200 test
    yield "alpha classes", trans
    two = np.where(rng.integers(0, 2, (64, 64, 1)) > 0, np.array([255, 0, 10, 255]), np.array([0, 255, 200, 3]))
    yield "two colours", two.astype(np.uint8)
    yield "saturated", rng.choice(np.array([0, 1, 254, 255], np.uint8), (64, 64, 4))
    post = rng.integers(0, 256, (128, 128, 4), dtype=np.uint8)
    post[..., :3] = (post[..., :3] // 64) * 64  # plateaus: many intensity ties in the extremum test
    yield "posterised", post


def test_pvrtc_code_on_host_matches_reference(pvrtc_check):
    for name, img in pvrtc_images():
        img = np.ascontiguousarray(img, dtype=np.uint8)
        r = subprocess.run([str(pvrtc_check), str(img.shape[1]), str(img.shape[0]), "v"], input=img.tobytes(),
                           capture_output=True)
        assert r.returncode == 0, f"{name}: {r.stdout.decode().strip()} {r.stderr.decode()[-200:]}"
        assert b"label lists ok" in r.stdout, name


def test_division_by_three_is_exact(tmp_path):
    """bc7_setup divides the covariance sums by 3 with a three-instruction sequence (div3 in bc7.cu); it has to
    be the IEEE quotient.  Every float with a biased exponent in [97, 157] (2^-30 .. 2^31, far beyond what sums
    of squared byte differences reach; the full range was run once: 2,122,317,824 floats, 0 mismatches)."""
    exe = tmp_path / "div3_check"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", str(ROOT / "tests" / "native" / "div3_check.c"), "-lm", "-o", str(exe)],
                   check=True)
    r = subprocess.run([str(exe), "97", "157"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    assert "bad 0" in r.stdout
