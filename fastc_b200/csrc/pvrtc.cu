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
//     independent.  One CTA per label kind walks the anti-diagonals of a band of rows with a barrier
//     per step (pvr_forward_rows); for the square images the reference accepts this order gives
//     every visit exactly the neighbour states the raster scan gives, wrap-around reads of
//     unvisited pixels included;
//   * the backward scan reads the right neighbour AND, at the start of a row, the last pixel of the
//     row processed before (through the wrap-around): one serial chain over all pixels in the
//     reference.  Rows stay sequential here (one CTA per label kind walks them), but inside a row
//     the distances are a scan over maps on {0..4} and the lists follow in three rounds (see
//     pvr_backward_rows).
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

// (ntex textures of w x h, one after the other)
__global__ void pvr_classify(const uint8_t *__restrict__ ibyte, uint32_t w, uint32_t h, uint32_t ntex, uint8_t *__restrict__ cls) {
  const uint32_t per = w * h, n = per * ntex;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t t = i / per, k = i - t * per;
    cls[i] = (uint8_t)classify_extremum(ibyte + (size_t)t * per, w, h, k % w, k / w);
  }
}

// LabelImageForward in one launch: one CTA per label kind, one thread per visited row of a band of up
// to 1,024 rows; at step s thread t handles column s - t of its row, so its up neighbour (thread
// t - 1, step s - 1) and left neighbour (itself, step s - 1) are done: the anti-diagonal order of
// anti-diagonal order of the scan with a CTA barrier per diagonal.  Bands follow each other.
__global__ void __launch_bounds__(1024)
pvr_forward_rows(PixelLabels *labels, const uint8_t *__restrict__ cls, uint32_t w, uint32_t h, uint32_t *overflow) {
  const bool high = blockIdx.x == 0;
  labels += (size_t)blockIdx.y * w * h;  // blockIdx.y: the texture of a batch
  cls += (size_t)blockIdx.y * w * h;
  const uint32_t rows = h + 3, tid = threadIdx.x, nt = blockDim.x;
  bool ok = true;
  for (uint32_t band = 0; band < rows; band += nt) {
    const uint32_t nrows = min(nt, rows - band), yy = band + tid;
    const uint32_t crow = wrap((int32_t)yy, h) * w;
    for (uint32_t s = 0; s < w + nrows - 1; s++) {
      const uint32_t x = s - tid;  // (wraps for s < tid: then x >= w)
      if (tid < nrows && x < w) ok = forward_pixel_kind(labels, w, h, x, yy, cls[crow + x], high) && ok;
      __syncthreads();
    }
  }
  if (!ok) atomicOr(overflow, 1u);
}

// ---- LabelImageBackward.  In the reference this is one serial chain over all pixels: a pixel reads its
// RIGHT neighbour's new label, and the first pixel of a row reads the last pixel of the row before
// (wrap-around).  Rows therefore stay sequential here, but a row is parallel:
//   * the new DISTANCE of a pixel is a function of its right neighbour's new distance alone once the
//     other four neighbours (row below: final, row above: untouched) are known -- a map on {0..4}.
//     The row's distances are the running composition of these maps from the row's start: a scan;
//   * a label LIST taken over from the right neighbour means distance = the neighbour's + 1, and
//     distances stop at 4, so list dependencies are chains of at most three pixels: the lists are
//     written in three rounds, new distance 2, then 3, then 4; what a pixel reads from its right
//     neighbour in its round is final.
// One CTA per label kind (the high and the low labels never mix), one thread per pixel (per group of
// pixels for rows wider than the CTA).
constexpr int kBackThreads = 1024;
constexpr int kBackMaxPer = 16;  // pixels per thread: rows up to 16384 wide

// a map on {0..4} as five 3-bit fields
__device__ __forceinline__ uint32_t fn_apply(uint32_t f, uint32_t x) { return (f >> (3 * x)) & 7u; }
__device__ __forceinline__ uint32_t fn_after(uint32_t g, uint32_t f) {  // x -> g(f(x))
  uint32_t r = 0;
#pragma unroll
  for (int x = 0; x < 5; x++) r |= fn_apply(g, fn_apply(f, x)) << (3 * x);
  return r;
}
constexpr uint32_t kFnIdentity = 0u | (1u << 3) | (2u << 6) | (3u << 9) | (4u << 12);

// DilateLabelBackward's distance rule (:358-388) for own distance c, the smallest positive distance m
// among the other four neighbours (5: none) and the right neighbour's distance x.  upd: the label is
// rewritten (distance and list).
__device__ __forceinline__ uint32_t back_distance(uint32_t c, uint32_t m, uint32_t x, bool &upd) {
  upd = false;
  if (c == 1) return 1;
  const uint32_t md = (x > 0 && x < m) ? x : m;
  const uint32_t nd = md + 1;
  if ((c != 0 && c < nd) || nd > 4) return c;
  upd = true;
  return nd;
}

__global__ void __launch_bounds__(kBackThreads)
pvr_backward_rows(PixelLabels *labels, uint32_t w, uint32_t h, uint32_t *overflow) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_prev[kBackThreads];  // composition of everything before thread t's pixels
  __shared__ Label s_first_right;            // column 0's label as the row's first pixel sees it
  const bool high = blockIdx.x == 0;
  labels += (size_t)blockIdx.y * w * h;  // blockIdx.y: the texture of a batch
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nthreads = blockDim.x;
  // pixels per thread (rows narrower than a warp leave the upper lanes without pixels: identity maps)
  const uint32_t per = (uint32_t)tid < w ? (w >= (uint32_t)nthreads ? w / (uint32_t)nthreads : 1u) : 0u;
  const uint32_t k0 = (uint32_t)tid * (w >= (uint32_t)nthreads ? w / (uint32_t)nthreads : 1u);
  auto lab = [&](uint32_t idx) -> Label & { return high ? labels[idx].high : labels[idx].low; };
  bool ok = true;
  for (int32_t j = (int32_t)h + 2; j >= 0; j--) {
    const uint32_t r = wrap(j, h) * w, ra = wrap(j - 1, h) * w, rb = wrap(j + 1, h) * w;
    // ---- the maps of this thread's pixels (processing position k = tid * per + q, column w - 1 - k)
    uint8_t own[kBackMaxPer], others[kBackMaxPer];
    uint32_t f_chunk = kFnIdentity;
    for (uint32_t q = 0; q < per; q++) {
      const uint32_t i = w - 1 - (k0 + q), xr = wrap((int32_t)i + 1, w), xl = wrap((int32_t)i - 1, w);
      const uint32_t c = lab(r + i).distance;
      const uint32_t d0 = lab(ra + xr).distance, d2 = lab(rb + xr).distance, d3 = lab(rb + i).distance, d4 = lab(rb + xl).distance;
      uint32_t m = 5;
      if (d0 > 0 && d0 < m) m = d0;
      if (d2 > 0 && d2 < m) m = d2;
      if (d3 > 0 && d3 < m) m = d3;
      if (d4 > 0 && d4 < m) m = d4;
      own[q] = (uint8_t)c; others[q] = (uint8_t)m;
      uint32_t f = 0;
      bool upd;
#pragma unroll
      for (uint32_t x = 0; x < 5; x++) f |= back_distance(c, m, x, upd) << (3 * x);
      f_chunk = fn_after(f, f_chunk);
    }
    if (tid == 0) s_first_right = lab(r + 0);  // pre-state: column 0 is the row's LAST pixel
    // ---- exclusive scan of the maps over the threads
    uint32_t f_incl = f_chunk;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t up = __shfl_up_sync(0xffffffffu, f_incl, d);
      if (lane >= d) f_incl = fn_after(f_incl, up);
    }
    if (lane == 31) s_warp[wid] = f_incl;
    __syncthreads();
    if (wid == 0) {
      uint32_t v = lane < (nthreads >> 5) ? s_warp[lane] : kFnIdentity;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v = fn_after(v, up);
      }
      s_warp[lane] = v;  // inclusive over warps
    }
    __syncthreads();
    {
      uint32_t f_excl = __shfl_up_sync(0xffffffffu, f_incl, 1);
      if (lane == 0) f_excl = kFnIdentity;
      if (wid > 0) f_excl = fn_after(f_excl, s_warp[wid - 1]);
      s_prev[tid] = f_excl;
    }
    // the right neighbour of the row's first pixel is column 0 in its pre-state
    uint32_t x = fn_apply(s_prev[tid], s_first_right.distance);
    // ---- new distances of this thread's pixels, and which are rewritten (with what round)
    uint8_t right[kBackMaxPer], round[kBackMaxPer];
    for (uint32_t q = 0; q < per; q++) {
      bool upd;
      const uint32_t nd = back_distance(own[q], others[q], x, upd);
      right[q] = (uint8_t)x;
      round[q] = upd ? (uint8_t)nd : 0;
      x = nd;
    }
    // ---- the lists, in rounds of the new distance
    for (uint32_t rnd = 2; rnd <= 4; rnd++) {
      __syncthreads();
      for (uint32_t q = 0; q < per; q++) {
        if (round[q] != rnd) continue;
        const uint32_t k = k0 + q, i = w - 1 - k, xr = wrap((int32_t)i + 1, w), xl = wrap((int32_t)i - 1, w);
        Label &me = lab(r + i);
        const uint32_t md = rnd - 1;
        if (me.distance != rnd) me.nLabels = 0;
        const Label *nb[5] = {&lab(ra + xr), k == 0 ? &s_first_right : &lab(r + xr), &lab(rb + xr), &lab(rb + i), &lab(rb + xl)};
        for (int n = 0; n < 5; n++) {
          const uint32_t dn = n == 1 ? right[q] : nb[n]->distance;
          if (dn != md) continue;
          const uint32_t cnt = nb[n]->nLabels;
          for (uint32_t e = 0; e < cnt; e++) ok = add_idx(me, nb[n]->idxs[e]) && ok;  // Combine
        }
        me.distance = (uint8_t)rnd;
      }
    }
    __syncthreads();
  }
  if (!ok) atomicOr(overflow, 1u);
}

__global__ void pvr_low_high(const PixelLabels *__restrict__ labels, const float *__restrict__ intensity,
                             const uint32_t *__restrict__ img, uint32_t w, uint32_t h, uint32_t *__restrict__ fields) {
  const uint32_t bw = w >> 2, nb = bw * (h >> 2);
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  const size_t po = (size_t)blockIdx.y * w * h;
  fields[(size_t)blockIdx.y * nb + b] = low_high_block(labels + po, intensity + po, img + po, w, h, b % bw, b / bw);
}

__global__ void pvr_modulate(const uint32_t *__restrict__ fields, const uint32_t *__restrict__ img, uint32_t w, uint32_t h,
                             uint2 *__restrict__ out) {
  const uint32_t bw = w >> 2, nb = bw * (h >> 2);
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  const uint32_t bx = b % bw, by = b / bw;
  fields += (size_t)blockIdx.y * nb;
  img += (size_t)blockIdx.y * w * h;
  // the block's 64 bits: modulation in the low word, colour fields in the high word; blocks are
  // stored in the reference's interleaved (Morton) order
  out[(size_t)blockIdx.y * nb + block_index(bx, by)] = make_uint2(modulation_block(fields, img, w, h, bx, by), fields[b]);
}

// decoder: one thread per pixel of the block range
__global__ void pvr_decode(const uint2 *__restrict__ blocks, uint32_t w, uint32_t h, uint32_t first_block, uint32_t num_blocks,
                           uint32_t *__restrict__ out) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= num_blocks * 16) return;
  const uint32_t b = first_block + (t >> 4), bw = w >> 2;
  const uint32_t i = (b % bw) * 4 + (t & 3), j = (b / bw) * 4 + ((t >> 2) & 3);
  out[(size_t)j * w + i] = decode_pixel(blocks, w, h, i, j);
}

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

cudaError_t launch_pvrtc_decode(const void *cmp_dev, uint32_t width, uint32_t height, uint32_t first_block,
                                uint32_t num_blocks, void *rgba_dev, cudaStream_t stream) {
  if (num_blocks == 0) return cudaSuccess;
  pvr_decode<<<(num_blocks * 16 + 255) / 256, 256, 0, stream>>>(static_cast<const uint2 *>(cmp_dev), width, height, first_block,
                                                                num_blocks, static_cast<uint32_t *>(rgba_dev));
  return cudaGetLastError();
}

void pvrtc_free_workspace(PvrtcWorkspace &ws) {
  if (ws.base) cudaFree(ws.base);
  if (ws.host_flag) cudaFreeHost(ws.host_flag);
  ws.base = nullptr; ws.bytes = 0; ws.host_flag = nullptr;
}

size_t pvrtc_scratch_bytes(uint32_t width, uint32_t height) {
  const size_t n = (size_t)width * height, nb = n / 16;
  return n * (4 + 1 + 1 + sizeof(pvr::PixelLabels)) + nb * 4 + 4096;
}

// ntex textures of width x height, one after the other in rgba_dev; their blocks one after the other in
// out_dev.  The textures of a batch are independent: every kernel takes the texture index from its
// grid (the two labelling kernels: 2 x ntex CTAs), which is what lets PVRTC use the GPU at all.
cudaError_t launch_pvrtc(PvrtcWorkspace &ws, const void *rgba_dev, uint32_t width, uint32_t height, uint32_t ntex,
                         void *out_dev, cudaStream_t stream, uint32_t *launches) {
  if (ntex == 0) return cudaSuccess;
  const uint32_t w = width, h = height, nb = (w >> 2) * (h >> 2);
  const size_t n = (size_t)w * h * ntex;
  if (n > 0xFFFFFFFFull) return cudaErrorInvalidValue;
  const size_t o_int = 0, o_ib = o_int + align256(n * 4), o_cls = o_ib + align256(n), o_fields = o_cls + align256(n),
               o_flag = o_fields + align256((size_t)nb * ntex * 4), o_labels = o_flag + 256,
               total = o_labels + n * sizeof(pvr::PixelLabels);
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
  cudaError_t e = cudaMemsetAsync(base + o_flag, 0, 256 + n * sizeof(pvr::PixelLabels), stream);  // calloc'ed labels
  if (e != cudaSuccess) return e;
  const uint32_t grid = (uint32_t)std::min<size_t>((n + 255) / 256, 148 * 16);
  pvr_intensity<<<grid, 256, 0, stream>>>(img, (uint32_t)n, intensity, ibyte);
  pvr_classify<<<grid, 256, 0, stream>>>(ibyte, w, h, ntex, cls);
  pvr_forward_rows<<<dim3(2, ntex), std::min<uint32_t>(1024u, (h + 3 + 31) & ~31u), 0, stream>>>(labels, cls, w, h, flag);
  pvr_backward_rows<<<dim3(2, ntex), std::max<uint32_t>(32u, std::min<uint32_t>(w, kBackThreads)), 0, stream>>>(labels, w, h, flag);
  pvr_low_high<<<dim3((nb + 127) / 128, ntex), 128, 0, stream>>>(labels, intensity, img, w, h, fields);
  pvr_modulate<<<dim3((nb + 127) / 128, ntex), 128, 0, stream>>>(fields, img, w, h, static_cast<uint2 *>(out_dev));
  if (launches) *launches += 6;
  return cudaGetLastError();
}

}  // namespace fastc
