// PVRTC 4bpp encoder for sm_100a.
//
// Behavioural contract: bit-identical to PVRTCC::Compress(job, eWrapMode_Wrap), the call FasTC's Core
// makes (reference/Core/src/TexComp.cpp:64-66, reference/PVRTCEncoder/src/Compressor.cpp:861-944);
// the arithmetic lives in pvrtc_block.cuh with the reference lines it follows.
//
// Unlike the block formats this encoder is image-level: two raster scans label every pixel with the
// nearby local intensity extrema (a distance-limited dilation), then each block's two colours are
// averages over the labelled extrema and the modulation bits come from the bilinearly upscaled
// colour images.  How the reference's scan ORDER maps to the GPU:
//   * intensities, extremum classes, block colours and modulation bits are per-pixel / per-block
//     work: plain data-parallel kernels;
//   * the forward scan reads the up and left neighbours: pixels of one anti-diagonal are
//     independent, so it runs as one launch per anti-diagonal over the (h + 3) x w visits (stream
//     order is the barrier; for the square images the reference accepts this order gives every
//     visit exactly the neighbour states the raster scan gives, wrap-around reads of unvisited
//     pixels included);
//   * the backward scan reads the right neighbour AND, at the start of a row, the last pixel of the
//     row processed before (through the wrap-around): it is one serial chain over all pixels in the
//     reference, and it is one here -- one thread per label kind (the high and the low labels never
//     mix).  It bounds the encoder: PVRTC scales over textures (batches), not inside one.
#include "kernels.h"
#include "pvrtc_block.cuh"

namespace fastc {
namespace {

using namespace pvr;

__global__ void pvr_intensity(const uint32_t *__restrict__ img, uint32_t n, float *__restrict__ intensity,
                              uint8_t *__restrict__ ibyte) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float v = intensity_of(img[i]);
    intensity[i] = v;
    ibyte[i] = (uint8_t)intensity_byte(v);
  }
}

__global__ void pvr_classify(const uint8_t *__restrict__ ibyte, uint32_t w, uint32_t h, uint8_t *__restrict__ cls) {
  const uint32_t n = w * h;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    cls[i] = (uint8_t)classify_extremum(ibyte, w, h, i % w, i / w);
}

// One anti-diagonal t = x + yy of LabelImageForward's (h + 3) x w visits.
__global__ void pvr_forward_diag(PixelLabels *labels, const uint8_t *__restrict__ cls, uint32_t w, uint32_t h, uint32_t t,
                                 uint32_t yy0, uint32_t count, uint32_t *overflow) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const uint32_t yy = yy0 + k, x = t - yy;
  if (!forward_pixel(labels, w, h, x, yy, cls[wrap((int32_t)yy, h) * w + x])) atomicOr(overflow, 1u);
}

// LabelImageBackward: the serial chain, block 0 = high labels, block 1 = low labels.
__global__ void pvr_backward(PixelLabels *labels, uint32_t w, uint32_t h, uint32_t *overflow) {
  if (threadIdx.x != 0) return;
  const bool high = blockIdx.x == 0;
  bool ok = true;
  for (int32_t j = (int32_t)h + 2; j >= 0; j--) {
    const uint32_t r = wrap(j, h) * w, ra = wrap(j - 1, h) * w, rb = wrap(j + 1, h) * w;
    for (int32_t i = (int32_t)w - 1; i >= 0; i--) {
      const uint32_t xr = wrap(i + 1, w), xl = wrap(i - 1, w);
      PixelLabels &l = labels[r + (uint32_t)i];
      Label &me = high ? l.high : l.low;
      if (me.distance == 1) continue;
      const PixelLabels *nb[5] = {&labels[ra + xr], &labels[r + xr], &labels[rb + xr], &labels[rb + (uint32_t)i], &labels[rb + xl]};
      const Label *n5[5];
      for (int q = 0; q < 5; q++) n5[q] = high ? &nb[q]->high : &nb[q]->low;
      ok = dilate_backward(me, n5) && ok;
    }
  }
  if (!ok) atomicOr(overflow, 1u);
}

__global__ void pvr_low_high(const PixelLabels *__restrict__ labels, const float *__restrict__ intensity,
                             const uint32_t *__restrict__ img, uint32_t w, uint32_t h, uint32_t *__restrict__ fields) {
  const uint32_t bw = w >> 2, nb = bw * (h >> 2);
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  fields[b] = low_high_block(labels, intensity, img, w, h, b % bw, b / bw);
}

__global__ void pvr_modulate(const uint32_t *__restrict__ fields, const uint32_t *__restrict__ img, uint32_t w, uint32_t h,
                             uint2 *__restrict__ out) {
  const uint32_t bw = w >> 2, nb = bw * (h >> 2);
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  const uint32_t bx = b % bw, by = b / bw;
  // the block's 64 bits: modulation in the low word, colour fields in the high word; blocks are
  // stored in the reference's interleaved (Morton) order
  out[block_index(bx, by)] = make_uint2(modulation_block(fields, img, w, h, bx, by), fields[b]);
}

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

void pvrtc_free_workspace(PvrtcWorkspace &ws) {
  if (ws.base) cudaFree(ws.base);
  if (ws.host_flag) cudaFreeHost(ws.host_flag);
  ws.base = nullptr; ws.bytes = 0; ws.host_flag = nullptr;
}

cudaError_t launch_pvrtc(PvrtcWorkspace &ws, const void *rgba_dev, uint32_t width, uint32_t height, void *out_dev,
                         cudaStream_t stream, uint32_t *launches) {
  const uint32_t w = width, h = height, n = w * h, nb = (w >> 2) * (h >> 2);
  const size_t o_int = 0, o_ib = o_int + align256((size_t)n * 4), o_cls = o_ib + align256(n), o_fields = o_cls + align256(n),
               o_flag = o_fields + align256((size_t)nb * 4), o_labels = o_flag + 256,
               total = o_labels + (size_t)n * sizeof(pvr::PixelLabels);
  if (ws.bytes < total) {
    cudaError_t e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return e;
    if (ws.base) cudaFree(ws.base);
    ws.base = nullptr; ws.bytes = 0;
    e = cudaMalloc(&ws.base, total);
    if (e != cudaSuccess) return e;
    ws.bytes = total;
  }
  uint8_t *base = static_cast<uint8_t *>(ws.base);
  float *intensity = reinterpret_cast<float *>(base + o_int);
  uint8_t *ibyte = base + o_ib, *cls = base + o_cls;
  uint32_t *fields = reinterpret_cast<uint32_t *>(base + o_fields), *flag = reinterpret_cast<uint32_t *>(base + o_flag);
  pvr::PixelLabels *labels = reinterpret_cast<pvr::PixelLabels *>(base + o_labels);
  const uint32_t *img = static_cast<const uint32_t *>(rgba_dev);
  cudaError_t e = cudaMemsetAsync(base + o_flag, 0, 256 + (size_t)n * sizeof(pvr::PixelLabels), stream);  // calloc'ed labels
  if (e != cudaSuccess) return e;
  const uint32_t grid = std::min<uint32_t>((n + 255) / 256, 148 * 16);
  pvr_intensity<<<grid, 256, 0, stream>>>(img, n, intensity, ibyte);
  pvr_classify<<<grid, 256, 0, stream>>>(ibyte, w, h, cls);
  uint32_t nl = 2;
  for (uint32_t t = 0; t <= (w - 1) + (h + 2); t++) {
    const uint32_t yy0 = t > w - 1 ? t - (w - 1) : 0, yy1 = std::min(h + 2, t);
    const uint32_t count = yy1 - yy0 + 1;
    pvr_forward_diag<<<(count + 127) / 128, 128, 0, stream>>>(labels, cls, w, h, t, yy0, count, flag);
    nl++;
  }
  pvr_backward<<<2, 32, 0, stream>>>(labels, w, h, flag);
  pvr_low_high<<<(nb + 127) / 128, 128, 0, stream>>>(labels, intensity, img, w, h, fields);
  pvr_modulate<<<(nb + 127) / 128, 128, 0, stream>>>(fields, img, w, h, static_cast<uint2 *>(out_dev));
  nl += 3;
  if (launches) *launches += nl;
  return cudaGetLastError();
}

}  // namespace fastc
