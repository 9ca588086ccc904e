// Block decoders (BC7, DXT1, DXT5, ETC1) and the PSNR reduction for sm_100a --
// the step right after the encode path (`tc` decodes every
// result to print its PSNR line and to write the -d image).
//
// Behavioural contract: bit-identical to the reference's decoders, including
// their departures from the format specifications:
//   BC7   reference/BPTCEncoder/src/Decompressor.cpp:32-190, 255-319 -- mode 4 with
//         idxMode == 1 keeps the mode's static 2-bit colour / 3-bit alpha interpolation
//         tables although the index arrays were swapped (:264-292);
//   DXT1  reference/DXTEncoder/src/Decompressor.cpp:28-67, 98-125 -- c0 <= c1 decodes index
//         3 as opaque black; alpha is always 255;
//   DXT5  reference/DXTEncoder/src/Decompressor.cpp:69-96, 127-156 -- colour half never
//         checks the endpoint order;
//   ETC1  reference/ETCEncoder/src/rg_etc1.cpp:945-1257 (unpack_etc1_block),
//         reference/ETCEncoder/src/Decompressor.cpp:27-47 -- out-of-range differential
//         bases are clamped.
// PSNR: reference/Base/src/Image.cpp:205-255 (alpha-premultiplied RGB, mse over W*H,
// peak 3*255^2).  The squared differences are exact integers
// ((a*c - a'*c')^2, scaled by 255^2), summed in uint64 -- order-independent, unlike
// the reference's running double.
//
// Mapping: one thread per block; a warp reads 256/512 contiguous bytes of blocks and
// writes four 512 B row segments (uint4 per thread per row).  Purely HBM-bound:
// 8 or 16 B in, 64 B out per block.
#include <cmath>

#include "bc7_tables.cuh"
#include "kernels.h"

namespace fastc {
namespace {

__constant__ uint16_t d_shape2[64];
__constant__ uint32_t d_shape3[64];
__constant__ uint8_t d_anchor2[64], d_anchor3a[64], d_anchor3b[64];
__constant__ uint8_t d_weight[64];  // [index_bits-1][16] weight of endpoint 2 (0..64)
__constant__ int d_inten[8][4] = {{-8, -2, 2, 8},     {-17, -5, 5, 17},   {-29, -9, 9, 29},    {-42, -13, 13, 42},
                                  {-60, -18, 18, 60}, {-80, -24, 24, 80}, {-106, -33, 33, 106}, {-183, -47, 47, 183}};
// mode attributes: partition bits, subsets, index bits, alpha index bits, colour bits, alpha bits,
// rotation, index mode, p-bit type (0 shared, 1 per endpoint, 2 none)
__constant__ uint8_t d_modes[8][9] = {
    {4, 3, 3, 0, 4, 0, 0, 0, 1}, {6, 2, 3, 0, 6, 0, 0, 0, 0}, {6, 3, 2, 0, 5, 0, 0, 0, 2}, {6, 2, 2, 0, 7, 0, 0, 0, 1},
    {0, 1, 2, 3, 5, 6, 1, 1, 2}, {0, 1, 2, 2, 7, 8, 1, 0, 2}, {0, 1, 4, 0, 7, 7, 0, 0, 1}, {6, 2, 2, 0, 5, 5, 0, 0, 1},
};

__device__ __forceinline__ int clamp255(int v) { return min(max(v, 0), 255); }

__device__ __forceinline__ void store_block(uint32_t *__restrict__ img, uint32_t width, uint32_t blocks_x, uint32_t bi,
                                            const uint32_t px[16]) {
  const uint32_t bx = bi % blocks_x, by = bi / blocks_x;
  uint4 *row = reinterpret_cast<uint4 *>(img + (size_t)by * 4 * width + (size_t)bx * 4);
  const uint32_t pitch4 = width >> 2;
#pragma unroll
  for (int j = 0; j < 4; j++) row[(size_t)j * pitch4] = make_uint4(px[4 * j], px[4 * j + 1], px[4 * j + 2], px[4 * j + 3]);
}

// ---------------------------------------------------------------- DXT
__device__ __forceinline__ void dxt_color(uint2 b, bool check_order, uint32_t alpha_keep_mask, uint32_t px[16]) {
  const uint32_t c0 = b.x & 0xFFFF, c1 = b.x >> 16;
  int col[4][3];
  col[0][0] = ((c0 >> 11) << 3) | (c0 >> 13); col[0][1] = (((c0 >> 5) & 63) << 2) | ((c0 >> 9) & 3);
  col[0][2] = ((c0 & 31) << 3) | ((c0 & 31) >> 2);
  col[1][0] = ((c1 >> 11) << 3) | (c1 >> 13); col[1][1] = (((c1 >> 5) & 63) << 2) | ((c1 >> 9) & 3);
  col[1][2] = ((c1 & 31) << 3) | ((c1 & 31) >> 2);
  const bool four = !check_order || c0 > c1;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    col[2][k] = four ? (col[0][k] * 2 + col[1][k]) / 3 : (col[0][k] + col[1][k]) / 2;
    col[3][k] = four ? (col[0][k] + col[1][k] * 2) / 3 : 0;
  }
  uint32_t pal[4];
#pragma unroll
  for (int s = 0; s < 4; s++) pal[s] = (uint32_t)col[s][0] | ((uint32_t)col[s][1] << 8) | ((uint32_t)col[s][2] << 16);
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const uint32_t s = (b.y >> (2 * i)) & 3;
    const uint32_t c = s == 0 ? pal[0] : (s == 1 ? pal[1] : (s == 2 ? pal[2] : pal[3]));
    px[i] = (px[i] & alpha_keep_mask) | c;
  }
}

__device__ __forceinline__ void dxt5_alpha(uint2 b, uint32_t px[16]) {
  const int a0 = b.x & 0xFF, a1 = (b.x >> 8) & 0xFF;
  int pal[8];
  pal[0] = a0; pal[1] = a1;
  if (a0 > a1) {
#pragma unroll
    for (int i = 2; i < 8; i++) pal[i] = ((8 - i) * a0 + (i - 1) * a1) / 7;
  } else {
#pragma unroll
    for (int i = 2; i < 6; i++) pal[i] = ((6 - i) * a0 + (i - 1) * a1) / 5;
    pal[6] = 0; pal[7] = 255;
  }
  const unsigned long long mod = ((unsigned long long)b.y << 16) | (b.x >> 16);
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const uint32_t s = (uint32_t)(mod >> (3 * i)) & 7;
    int a = pal[0];
#pragma unroll
    for (int q = 1; q < 8; q++) a = s == (uint32_t)q ? pal[q] : a;
    px[i] = (uint32_t)a << 24;
  }
}

// ---------------------------------------------------------------- ETC1
__device__ __forceinline__ void etc1_block(uint2 blk, uint32_t px[16]) {
  const uint32_t b3 = blk.x >> 24;
  const bool diff = b3 & 2, flip = b3 & 1;
  const int t0 = (b3 >> 5) & 7, t1 = (b3 >> 2) & 7;
  int base[2][3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const int v = (blk.x >> (8 * k)) & 0xFF;
    if (diff) {
      const int c5 = v >> 3;
      int d3 = v & 7;
      if (d3 >= 4) d3 -= 8;
      const int c2 = min(max(c5 + d3, 0), 31);
      base[0][k] = (c5 << 3) | (c5 >> 2);
      base[1][k] = (c2 << 3) | (c2 >> 2);
    } else {
      const int c0 = v >> 4, c1 = v & 15;
      base[0][k] = (c0 << 4) | c0;
      base[1][k] = (c1 << 4) | c1;
    }
  }
  // bytes 4,5 = MSB plane (big endian), bytes 6,7 = LSB plane
  const uint32_t msb = ((blk.y & 0xFF) << 8) | ((blk.y >> 8) & 0xFF);
  const uint32_t lsb = (((blk.y >> 16) & 0xFF) << 8) | (blk.y >> 24);
#pragma unroll
  for (int y = 0; y < 4; y++)
#pragma unroll
    for (int x = 0; x < 4; x++) {
      const int sub = flip ? (y >= 2) : (x >= 2);
      const int bit = x * 4 + y;
      const uint32_t code = ((lsb >> bit) & 1) | (((msb >> bit) & 1) << 1);
      const int sel = (0x1E >> (2 * code)) & 3;  // ETC1 code -> modifier index {2,3,1,0}
      const int m = d_inten[sub ? t1 : t0][sel];
      const int r = clamp255(base[sub][0] + m), g = clamp255(base[sub][1] + m), bl = clamp255(base[sub][2] + m);
      px[y * 4 + x] = (uint32_t)r | ((uint32_t)g << 8) | ((uint32_t)bl << 16) | 0xFF000000u;
    }
}

// ---------------------------------------------------------------- BC7
struct Bits128 {
  uint32_t w[4];
  int pos;
  __device__ __forceinline__ uint32_t word(int i) const {
    return i == 0 ? w[0] : (i == 1 ? w[1] : (i == 2 ? w[2] : (i == 3 ? w[3] : 0u)));
  }
  __device__ __forceinline__ uint32_t get(int n) {  // LSB first, n <= 8
    if (n == 0) return 0;
    const int wi = pos >> 5;
    const uint32_t v = __funnelshift_r(word(wi), word(wi + 1), pos & 31) & ((1u << n) - 1);
    pos += n;
    return v;
  }
};

__device__ __forceinline__ int d_subset_of(int idx, int shape, int nsub) {
  if (nsub == 2) return (d_shape2[shape] >> idx) & 1;
  if (nsub == 3) return (d_shape3[shape] >> (2 * idx)) & 3;
  return 0;
}
__device__ __forceinline__ int d_anchor_of(int subset, int shape, int nsub) {
  if (subset == 0) return 0;
  if (subset == 1) return nsub == 2 ? d_anchor2[shape] : d_anchor3a[shape];
  return d_anchor3b[shape];
}

__device__ void bc7_block(uint4 blk, uint32_t px[16]) {
  Bits128 s{{blk.x, blk.y, blk.z, blk.w}, 0};
  const int mode = blk.x & 0xFF ? __ffs(blk.x & 0xFF) - 1 : 8;
  if (mode >= 8) {
#pragma unroll
    for (int i = 0; i < 16; i++) px[i] = 0;
    return;
  }
  s.pos = mode + 1;
  const uint8_t *A = d_modes[mode];
  const int nsub = A[1], ibits = A[2], aibits = A[3];
  int shape = 0, rot = 0, idx_mode = 0;
  if (nsub > 1) shape = s.get(mode == 0 ? 4 : 6);
  else if (A[6]) {
    rot = s.get(2);
    if (A[7]) idx_mode = s.get(1);
  }
  int cp = A[4], ap = A[5];
  uint32_t eps[3][2];  // packed RGBA endpoints per subset
#pragma unroll
  for (int i = 0; i < 3; i++) eps[i][0] = eps[i][1] = 0;
  for (int ch = 0; ch < 3; ch++)
    for (int i = 0; i < nsub; i++)
      for (int e = 0; e < 2; e++) eps[i][e] |= ((s.get(cp) << (8 - cp)) & 0xFF) << (8 * ch);
  for (int i = 0; i < nsub; i++)
    for (int e = 0; e < 2; e++) eps[i][e] |= (ap == 0 ? 0xFFu : ((s.get(ap) << (8 - ap)) & 0xFF)) << 24;
  if (A[8] != 2) {
    cp += 1; ap += 1;
    for (int i = 0; i < nsub; i++) {
      const uint32_t p0 = s.get(1);
      const uint32_t p1 = A[8] == 0 ? p0 : s.get(1);
      for (int ch = 0; ch < 4; ch++) {
        const int sh = 8 - (ch == 3 ? ap : cp);
        eps[i][0] |= ((p0 << sh) & 0xFF) << (8 * ch);
        eps[i][1] |= ((p1 << sh) & 0xFF) << (8 * ch);
      }
    }
  }
  for (int i = 0; i < nsub; i++)
    for (int e = 0; e < 2; e++) {
      uint32_t v = 0;
      for (int ch = 0; ch < 4; ch++) {
        uint32_t c = (eps[i][e] >> (8 * ch)) & 0xFF;
        c |= c >> (ch == 3 ? ap : cp);
        v |= (c & 0xFF) << (8 * ch);
      }
      eps[i][e] = v;
    }
  unsigned long long cidx = 0, aidx = 0;
  for (int i = 0; i < 16; i++) {
    const int sub = d_subset_of(i, shape, nsub);
    cidx |= (unsigned long long)s.get(d_anchor_of(sub, shape, nsub) == i ? ibits - 1 : ibits) << (4 * i);
  }
  if (aibits == 0) {
    aidx = cidx;
  } else {
    for (int i = 0; i < 16; i++) aidx |= (unsigned long long)s.get(i == 0 ? aibits - 1 : aibits) << (4 * i);
    if (idx_mode) {  // indices swap, interpolation tables do not (reference quirk)
      const unsigned long long t = aidx; aidx = cidx; cidx = t;
    }
  }
  const uint8_t *wc = d_weight + 16 * (ibits - 1);
  const uint8_t *wa = d_weight + 16 * ((aibits ? aibits : ibits) - 1);
  for (int i = 0; i < 16; i++) {
    const int sub = d_subset_of(i, shape, nsub);
    const uint32_t e0 = sub == 0 ? eps[0][0] : (sub == 1 ? eps[1][0] : eps[2][0]);
    const uint32_t e1 = sub == 0 ? eps[0][1] : (sub == 1 ? eps[1][1] : eps[2][1]);
    const uint32_t w1c = wc[(cidx >> (4 * i)) & 15], w1a = wa[(aidx >> (4 * i)) & 15];
    uint32_t c[4];
#pragma unroll
    for (int ch = 0; ch < 4; ch++) {
      // 255 marks "no such index at this precision" (mode 4, idxMode 1: 3-bit indices looked up
      // in the 2-bit row): the reference's table holds weights {0, 0} there -> channel 0
      const uint32_t wr = (ch == 3 && aibits > 0) ? w1a : w1c;
      const uint32_t w1 = wr == 255 ? 0 : wr, w0 = wr == 255 ? 0 : 64 - wr;
      c[ch] = ((((e0 >> (8 * ch)) & 0xFF) * w0 + ((e1 >> (8 * ch)) & 0xFF) * w1 + 32) >> 6) & 0xFF;
    }
    if (rot) {
      const uint32_t t = c[3];
      if (rot == 1) { c[3] = c[0]; c[0] = t; }
      else if (rot == 2) { c[3] = c[1]; c[1] = t; }
      else { c[3] = c[2]; c[2] = t; }
    }
    px[i] = c[0] | (c[1] << 8) | (c[2] << 16) | (c[3] << 24);
  }
}

template <int FMT>
__global__ void __launch_bounds__(128)
decode_kernel(const void *__restrict__ cmp, uint32_t width, uint32_t blocks_x, uint32_t first_block, uint32_t num_blocks,
              uint32_t *__restrict__ out) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= num_blocks) return;
  const uint32_t bi = first_block + t;
  uint32_t px[16];
  if (FMT == 0) {  // DXT1
#pragma unroll
    for (int i = 0; i < 16; i++) px[i] = 0xFF000000u;
    dxt_color(__ldg(static_cast<const uint2 *>(cmp) + bi), true, 0xFF000000u, px);
  } else if (FMT == 1) {  // DXT5
    const uint4 b = __ldg(static_cast<const uint4 *>(cmp) + bi);
    dxt5_alpha(make_uint2(b.x, b.y), px);
    dxt_color(make_uint2(b.z, b.w), false, 0xFF000000u, px);
  } else if (FMT == 2) {
    etc1_block(__ldg(static_cast<const uint2 *>(cmp) + bi), px);
  } else {
    bc7_block(__ldg(static_cast<const uint4 *>(cmp) + bi), px);
  }
  store_block(out, width, blocks_x, bi, px);
}

// Sum over pixels and RGB channels of (a_alpha * a_c - b_alpha * b_c)^2 (exact, uint64).
__global__ void __launch_bounds__(256)
psnr_kernel(const uint32_t *__restrict__ a, const uint32_t *__restrict__ b, size_t n, unsigned long long *sum) {
  unsigned long long acc = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t p = __ldg(a + i), q = __ldg(b + i);
    const int pa = p >> 24, qa = q >> 24;
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const long long d = (long long)(pa * (int)((p >> (8 * c)) & 0xFF)) - (long long)(qa * (int)((q >> (8 * c)) & 0xFF));
      acc += (unsigned long long)(d * d);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(sum, acc);
}

}  // namespace

cudaError_t decode_upload_tables() {
  using namespace bc7tab;
  cudaError_t e;
#define UP(sym, src) if ((e = cudaMemcpyToSymbol(sym, src, sizeof(src))) != cudaSuccess) return e
  UP(d_shape2, kShape2); UP(d_shape3, kShape3); UP(d_anchor2, kAnchor2); UP(d_anchor3a, kAnchor3a);
  UP(d_anchor3b, kAnchor3b); UP(d_weight, kWeight);
#undef UP
  return cudaSuccess;
}

cudaError_t launch_decode(int format, const void *cmp_dev, uint32_t width, uint32_t first_block, uint32_t num_blocks,
                          void *rgba_dev, cudaStream_t stream) {
  if (num_blocks == 0) return cudaSuccess;
  const uint32_t grid = (num_blocks + 127) / 128, bx = width / 4;
  uint32_t *out = static_cast<uint32_t *>(rgba_dev);
  switch (format) {
    case 0: decode_kernel<0><<<grid, 128, 0, stream>>>(cmp_dev, width, bx, first_block, num_blocks, out); break;
    case 1: decode_kernel<1><<<grid, 128, 0, stream>>>(cmp_dev, width, bx, first_block, num_blocks, out); break;
    case 2: decode_kernel<2><<<grid, 128, 0, stream>>>(cmp_dev, width, bx, first_block, num_blocks, out); break;
    default: decode_kernel<3><<<grid, 128, 0, stream>>>(cmp_dev, width, bx, first_block, num_blocks, out); break;
  }
  return cudaGetLastError();
}

cudaError_t launch_psnr_sum(const void *a_dev, const void *b_dev, size_t num_pixels, unsigned long long *sum_dev,
                            cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(sum_dev, 0, sizeof(unsigned long long), stream);
  if (e != cudaSuccess) return e;
  if (num_pixels == 0) return cudaSuccess;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t want = (num_pixels + 255) / 256;
  const uint32_t grid = (uint32_t)(want < (size_t)sms * 8 ? want : (size_t)sms * 8);  // grid-stride, 8 CTAs per SM
  psnr_kernel<<<grid, 256, 0, stream>>>(static_cast<const uint32_t *>(a_dev), static_cast<const uint32_t *>(b_dev),
                                        num_pixels, sum_dev);
  return cudaGetLastError();
}

// Image.cpp:246-254: mse = sum / (W*H); PSNR = 10 log10(3 * 255^2 / mse).  `sum` carries a 255^2 scale.
double psnr_from_sum(unsigned long long sum, size_t num_pixels) {
  const double mse = ((double)sum / 65025.0) / (double)num_pixels;
  return 10.0 * log10((3.0 * 255.0 * 255.0) / mse);
}

}  // namespace fastc
