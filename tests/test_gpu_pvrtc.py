"""PVRTC 4bpp on the GPU (fastc_b200/csrc/pvrtc.cu) against the compiled reference: every block of every
texture bit-identical to PVRTCC::Compress (oracle/_ref/libfastc_ref.so, the judge's oracle for this
format -- there is no separate CPU restatement: tests/native/pvrtc_host_check.cpp runs the product's own
arithmetic on the host against the same reference)."""
import numpy as np
import pytest
import torch

from _checkers import Reference, BLOCK_BYTES
from fastc_b200 import ECompressionFormat as F, lib
from fastc_b200.api import FastcGpuError
from fastc_b200.synth import synth_rgba
from test_native_host import pvrtc_images

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return lib()


@pytest.fixture(scope="module")
def ref():
    if not Reference.available():
        pytest.skip("oracle/_ref/libfastc_ref.so missing")
    return Reference()


def _gpu_pvrtc(g, img):
    h, w = img.shape[:2]
    out = np.zeros((w // 4) * (h // 4) * 8, np.uint8)
    g.compress(F.PVRTC4, np.ascontiguousarray(img), out)
    return out


def test_pvrtc_matches_reference_on_styled_images(gpu, ref):
    for name, img in pvrtc_images():
        img = np.ascontiguousarray(img, dtype=np.uint8)
        want, _ = ref.compress("PVRTC4", img, seed=None)
        got = _gpu_pvrtc(gpu, img)
        bad = int((got.reshape(-1, 8) != want.reshape(-1, 8)).any(1).sum())
        assert bad == 0, f"{name}: {bad} blocks differ from the reference"


@pytest.mark.parametrize("size,seed", [(512, 3), (1024, 1)])
def test_pvrtc_matches_reference_at_size(gpu, ref, size, seed):
    img = synth_rgba(size, size, seed)
    want, _ = ref.compress("PVRTC4", img, seed=None)
    got = _gpu_pvrtc(gpu, img)
    assert np.array_equal(got, want)


def test_pvrtc_device_api_and_batch(gpu, ref):
    # runs of equal sizes go through the kernels side by side (grid.y = texture); sizes may change in a batch
    imgs = [synth_rgba(128, 128, s) for s in range(1, 6)] + [synth_rgba(64, 64, 9), synth_rgba(256, 256, 4),
                                                             synth_rgba(256, 256, 5), synth_rgba(32, 32, 2)]
    outs, tm = gpu.compress_batch(F.PVRTC4, imgs)
    for im, o in zip(imgs, outs):
        want, _ = ref.compress("PVRTC4", im, seed=None)
        assert np.array_equal(o, want)
    assert tm["kernel_launches"] > 0
    d_in = torch.from_numpy(imgs[0]).cuda()
    d_out = torch.zeros(32 * 32 * 8, dtype=torch.uint8, device="cuda")
    n = gpu.compress_device(F.PVRTC4, d_in, d_out, width=128, height=128)
    torch.cuda.synchronize()
    want, _ = ref.compress("PVRTC4", imgs[0], seed=None)
    assert n > 0 and np.array_equal(d_out.cpu().numpy(), want)


def test_pvrtc_rejects_what_the_reference_rejects(gpu):
    out = np.zeros(1 << 16, np.uint8)
    for shape in ((64, 128, 4), (48, 48, 4), (4, 4, 4)):  # not square, not a power of two, too small
        with pytest.raises(FastcGpuError):
            gpu.compress(F.PVRTC4, np.zeros(shape, np.uint8), out)
    with pytest.raises(FastcGpuError):  # block ranges make no sense for an image-level encoder
        gpu.compress(F.PVRTC4, np.zeros((64, 64, 4), np.uint8), out, first_block=4, num_blocks=8)


def test_pvrtc_decoder_matches_reference_decoder(gpu, ref):
    for size, seed in ((64, 2), (256, 1)):
        img = synth_rgba(size, size, seed)
        cmp_ref, _ = ref.compress("PVRTC4", img, seed=None)
        want = np.zeros((size, size, 4), np.uint8)
        assert ref.lib.fastc_ref_decompress(4, cmp_ref.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_uint8)),
                                            size, size, want.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_uint8))) == 0
        got = gpu.decompress(F.PVRTC4, cmp_ref, size, size)
        got = got[0] if isinstance(got, tuple) else got
        assert np.array_equal(np.asarray(got).reshape(size, size, 4), want)
