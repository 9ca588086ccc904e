// DXT1 / DXT5 block encoder for sm_100a.
//
// Behavioural contract: bit-identical to stb_dxt v1.06 driven the way the
// reference drives it (STB_DXT_DITHER, one refinement pass):
//   reference/DXTEncoder/src/stb_dxt.h:477-548 (colour block), :551-601 (alpha
//   block), reference/DXTEncoder/src/Compressor.cpp:47-95 (block loops).
//
// Mapping: one thread per 4x4 block, everything in registers (16 packed RGBA
// words + 16 dithered words), no shared memory: there is no data reuse between
// blocks, a warp's loads/stores are already contiguous (512 B per row load,
// 256/512 B per store), and the encoder is ~1.5k integer instructions per block,
// i.e. issue-bound long before HBM.  The few float ops use
// explicit round-to-nearest intrinsics so that no FMA contraction can change a
// result relative to the reference's SSE2 scalar build.
#include "common.cuh"
#include "kernels.h"

namespace fastc {
namespace {

// Single-colour optimal endpoints, stb__OMatch5/6 (stb_dxt.h:121-147, 619-620):
// [0..511] = OMatch5[256][2], [512..1023] = OMatch6[256][2].  Built on the host
// with the same scan order, uploaded once per device.
__device__ uint8_t g_omatch[1024];

__device__ __forceinline__ int mul8bit(int a, int b) {
  int t = a * b + 128;
  return (t + (t >> 8)) >> 8;
}
__device__ __forceinline__ int expand5(int v) { return (v << 3) | (v >> 2); }
__device__ __forceinline__ int expand6(int v) { return (v << 2) | (v >> 4); }
// stb__QuantRBTab/GTab[x + 8] with the table's clamp folded in (stb_dxt.h:612-617).
__device__ __forceinline__ int quant_rb(int x) { return expand5(mul8bit(min(max(x, 0), 255), 31)); }
__device__ __forceinline__ int quant_g(int x) { return expand6(mul8bit(min(max(x, 0), 255), 63)); }
__device__ __forceinline__ int lerp13(int a, int b) { return (2 * a + b) / 3; }

__device__ __forceinline__ uint32_t as16bit(uint32_t p) {
  return (mul8bit(p & 0xFF, 31) << 11) + (mul8bit((p >> 8) & 0xFF, 63) << 5) + mul8bit((p >> 16) & 0xFF, 31);
}

template <int CH>
__device__ __forceinline__ int chan(uint32_t p) {
  return (p >> (8 * CH)) & 0xFF;
}

// Floyd-Steinberg dither of one channel to the 565 grid (stb_dxt.h:159-183).
template <int CH>
__device__ __forceinline__ void dither_channel(const uint32_t (&px)[16], uint32_t (&d)[16]) {
  int e[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
#pragma unroll
  for (int y = 0; y < 4; y++) {
    int(&e1)[4] = e[y & 1];
    int(&e2)[4] = e[(y & 1) ^ 1];
    int b0 = chan<CH>(px[4 * y + 0]), b1 = chan<CH>(px[4 * y + 1]);
    int b2 = chan<CH>(px[4 * y + 2]), b3 = chan<CH>(px[4 * y + 3]);
    int q0, q1, q2, q3;
    if (CH == 1) {
      q0 = quant_g(b0 + ((3 * e2[1] + 5 * e2[0]) >> 4));
      e1[0] = b0 - q0;
      q1 = quant_g(b1 + ((7 * e1[0] + 3 * e2[2] + 5 * e2[1] + e2[0]) >> 4));
      e1[1] = b1 - q1;
      q2 = quant_g(b2 + ((7 * e1[1] + 3 * e2[3] + 5 * e2[2] + e2[1]) >> 4));
      e1[2] = b2 - q2;
      q3 = quant_g(b3 + ((7 * e1[2] + 5 * e2[3] + e2[2]) >> 4));
      e1[3] = b3 - q3;
    } else {
      q0 = quant_rb(b0 + ((3 * e2[1] + 5 * e2[0]) >> 4));
      e1[0] = b0 - q0;
      q1 = quant_rb(b1 + ((7 * e1[0] + 3 * e2[2] + 5 * e2[1] + e2[0]) >> 4));
      e1[1] = b1 - q1;
      q2 = quant_rb(b2 + ((7 * e1[1] + 3 * e2[3] + 5 * e2[2] + e2[1]) >> 4));
      e1[2] = b2 - q2;
      q3 = quant_rb(b3 + ((7 * e1[2] + 5 * e2[3] + e2[2]) >> 4));
      e1[3] = b3 - q3;
    }
    d[4 * y + 0] |= (uint32_t)q0 << (8 * CH);
    d[4 * y + 1] |= (uint32_t)q1 << (8 * CH);
    d[4 * y + 2] |= (uint32_t)q2 << (8 * CH);
    d[4 * y + 3] |= (uint32_t)q3 << (8 * CH);
  }
}

// stb__OptimizeColorsBlock (stb_dxt.h:283-385) on the dithered block.
__device__ __forceinline__ void optimize_colors(const uint32_t (&d)[16], uint32_t &max16, uint32_t &min16) {
  int mu[3], mn[3], mx[3];
#pragma unroll
  for (int ch = 0; ch < 3; ch++) {
    int s, lo, hi;
    s = lo = hi = (d[0] >> (8 * ch)) & 0xFF;
#pragma unroll
    for (int i = 1; i < 16; i++) {
      int v = (d[i] >> (8 * ch)) & 0xFF;
      s += v;
      lo = min(lo, v);
      hi = max(hi, v);
    }
    mu[ch] = (s + 8) >> 4;
    mn[ch] = lo;
    mx[ch] = hi;
  }
  int c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    int r = chan<0>(d[i]) - mu[0], g = chan<1>(d[i]) - mu[1], b = chan<2>(d[i]) - mu[2];
    c0 += r * r; c1 += r * g; c2 += r * b;
    c3 += g * g; c4 += g * b; c5 += b * b;
  }
  const float f0 = __fdiv_rn((float)c0, 255.0f), f1 = __fdiv_rn((float)c1, 255.0f);
  const float f2 = __fdiv_rn((float)c2, 255.0f), f3 = __fdiv_rn((float)c3, 255.0f);
  const float f4 = __fdiv_rn((float)c4, 255.0f), f5 = __fdiv_rn((float)c5, 255.0f);
  float vr = (float)(mx[0] - mn[0]), vg = (float)(mx[1] - mn[1]), vb = (float)(mx[2] - mn[2]);
#pragma unroll
  for (int it = 0; it < 4; it++) {
    float r = __fadd_rn(__fadd_rn(__fmul_rn(vr, f0), __fmul_rn(vg, f1)), __fmul_rn(vb, f2));
    float g = __fadd_rn(__fadd_rn(__fmul_rn(vr, f1), __fmul_rn(vg, f3)), __fmul_rn(vb, f4));
    float b = __fadd_rn(__fadd_rn(__fmul_rn(vr, f2), __fmul_rn(vg, f4)), __fmul_rn(vb, f5));
    vr = r; vg = g; vb = b;
  }
  float magf = fmaxf(fmaxf(fabsf(vr), fabsf(vg)), fabsf(vb));  // values are finite: plain max is exact
  int v_r, v_g, v_b;
  if (magf < 4.0f) {
    v_r = 299; v_g = 587; v_b = 114;
  } else {
    // reference does this part in double (stb_dxt.h:361-364)
    double magn = __ddiv_rn(512.0, (double)magf);
    v_r = (int)__dmul_rn((double)vr, magn);
    v_g = (int)__dmul_rn((double)vg, magn);
    v_b = (int)__dmul_rn((double)vb, magn);
  }
  int mind = 0x7fffffff, maxd = -0x7fffffff;
  uint32_t minp = d[0], maxp = d[0];
#pragma unroll
  for (int i = 0; i < 16; i++) {
    int dot = chan<0>(d[i]) * v_r + chan<1>(d[i]) * v_g + chan<2>(d[i]) * v_b;
    if (dot < mind) { mind = dot; minp = d[i]; }
    if (dot > maxd) { maxd = dot; maxp = d[i]; }
  }
  max16 = as16bit(maxp);
  min16 = as16bit(minp);
}

// Palette (stb__EvalColors, stb_dxt.h:149-155) + dithered index selection
// (stb__MatchColorsBlock dither branch, stb_dxt.h:186-280) on the ORIGINAL block.
__device__ __forceinline__ uint32_t match_colors(const uint32_t (&px)[16], uint32_t c0_16, uint32_t c1_16) {
  int col[4][3];
  col[0][0] = expand5((c0_16 >> 11) & 31); col[0][1] = expand6((c0_16 >> 5) & 63); col[0][2] = expand5(c0_16 & 31);
  col[1][0] = expand5((c1_16 >> 11) & 31); col[1][1] = expand6((c1_16 >> 5) & 63); col[1][2] = expand5(c1_16 & 31);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    col[2][k] = lerp13(col[0][k], col[1][k]);
    col[3][k] = lerp13(col[1][k], col[0][k]);
  }
  const int dr = col[0][0] - col[1][0], dg = col[0][1] - col[1][1], db = col[0][2] - col[1][2];
  int stops[4];
#pragma unroll
  for (int i = 0; i < 4; i++) stops[i] = col[i][0] * dr + col[i][1] * dg + col[i][2] * db;
  const int c0p = ((stops[1] + stops[3]) >> 1) << 4;
  const int halfp = ((stops[3] + stops[2]) >> 1) << 4;
  const int c3p = ((stops[2] + stops[0]) >> 1) << 4;

  uint32_t mask = 0;
  int e[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
#pragma unroll
  for (int y = 0; y < 4; y++) {
    int(&e1)[4] = e[y & 1];
    int(&e2)[4] = e[(y & 1) ^ 1];
    int dp[4];
#pragma unroll
    for (int x = 0; x < 4; x++) {
      uint32_t p = px[4 * y + x];
      dp[x] = chan<0>(p) * dr + chan<1>(p) * dg + chan<2>(p) * db;
    }
#pragma unroll
    for (int x = 0; x < 4; x++) {
      int acc;
      if (x == 0) acc = 3 * e2[1] + 5 * e2[0];
      else if (x == 1) acc = 7 * e1[0] + 3 * e2[2] + 5 * e2[1] + e2[0];
      else if (x == 2) acc = 7 * e1[1] + 3 * e2[3] + 5 * e2[2] + e2[1];
      else acc = 7 * e1[2] + 5 * e2[3] + e2[2];
      const int dot = (dp[x] << 4) + acc;
      const int step = dot < halfp ? (dot < c0p ? 1 : 3) : (dot < c3p ? 2 : 0);
      const int stop = step == 0 ? stops[0] : step == 1 ? stops[1] : step == 2 ? stops[2] : stops[3];
      e1[x] = dp[x] - stop;
      mask |= (uint32_t)step << (8 * y + 2 * x);
    }
  }
  return mask;
}

__device__ __forceinline__ int sclamp(float y, int hi) {
  int x = (int)y;  // cvt.rzi, same as x86 cvttss2si for these in-range values
  return x < 0 ? 0 : (x > hi ? hi : x);
}

// stb__RefineBlock (stb_dxt.h:398-474) on the dithered block.  Returns true if
// the endpoints changed.
__device__ __forceinline__ bool refine_block(const uint32_t (&d)[16], uint32_t &max16, uint32_t &min16,
                                             uint32_t mask) {
  const uint32_t old_min = min16, old_max = max16;
  uint32_t nmin, nmax;
  if ((mask ^ (mask << 2)) < 4u) {
    int r = 8, g = 8, b = 8;
#pragma unroll
    for (int i = 0; i < 16; i++) { r += chan<0>(d[i]); g += chan<1>(d[i]); b += chan<2>(d[i]); }
    r >>= 4; g >>= 4; b >>= 4;
    const uint8_t *o5 = g_omatch, *o6 = g_omatch + 512;
    nmax = ((uint32_t)o5[2 * r] << 11) | ((uint32_t)o6[2 * g] << 5) | o5[2 * b];
    nmin = ((uint32_t)o5[2 * r + 1] << 11) | ((uint32_t)o6[2 * g + 1] << 5) | o5[2 * b + 1];
  } else {
    int xx = 0, yy = 0, xy = 0;
    int a1r = 0, a1g = 0, a1b = 0, a2r = 0, a2g = 0, a2b = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const int step = (mask >> (2 * i)) & 3;
      const int w1 = (0x1203 >> (4 * step)) & 0xF;  // w1Tab = {3,0,2,1}
      const int w2 = 3 - w1;
      const int r = chan<0>(d[i]), g = chan<1>(d[i]), b = chan<2>(d[i]);
      xx += w1 * w1; yy += w2 * w2; xy += w1 * w2;  // == the packed `prods` accumulator
      a1r += w1 * r; a1g += w1 * g; a1b += w1 * b;
      a2r += r; a2g += g; a2b += b;
    }
    a2r = 3 * a2r - a1r; a2g = 3 * a2g - a1g; a2b = 3 * a2b - a1b;
    const float frb = __fdiv_rn(__fdiv_rn(93.0f, 255.0f), (float)(xx * yy - xy * xy));
    const float fg = __fdiv_rn(__fmul_rn(frb, 63.0f), 31.0f);
    nmax = (uint32_t)sclamp(__fadd_rn(__fmul_rn((float)(a1r * yy - a2r * xy), frb), 0.5f), 31) << 11;
    nmax |= (uint32_t)sclamp(__fadd_rn(__fmul_rn((float)(a1g * yy - a2g * xy), fg), 0.5f), 63) << 5;
    nmax |= (uint32_t)sclamp(__fadd_rn(__fmul_rn((float)(a1b * yy - a2b * xy), frb), 0.5f), 31);
    nmin = (uint32_t)sclamp(__fadd_rn(__fmul_rn((float)(a2r * xx - a1r * xy), frb), 0.5f), 31) << 11;
    nmin |= (uint32_t)sclamp(__fadd_rn(__fmul_rn((float)(a2g * xx - a1g * xy), fg), 0.5f), 63) << 5;
    nmin |= (uint32_t)sclamp(__fadd_rn(__fmul_rn((float)(a2b * xx - a1b * xy), frb), 0.5f), 31);
  }
  min16 = nmin;
  max16 = nmax;
  return old_min != nmin || old_max != nmax;
}

// stb__CompressColorBlock (stb_dxt.h:477-548), mode = STB_DXT_DITHER.
__device__ __forceinline__ uint2 compress_color_block(const uint32_t (&px)[16]) {
  uint32_t mask, max16, min16;
  bool constant = true;
#pragma unroll
  for (int i = 1; i < 16; i++) constant = constant && (px[i] == px[0]);  // full 32-bit test (alpha included)
  if (constant) {
    const int r = chan<0>(px[0]), g = chan<1>(px[0]), b = chan<2>(px[0]);
    const uint8_t *o5 = g_omatch, *o6 = g_omatch + 512;
    mask = 0xaaaaaaaau;
    max16 = ((uint32_t)o5[2 * r] << 11) | ((uint32_t)o6[2 * g] << 5) | o5[2 * b];
    min16 = ((uint32_t)o5[2 * r + 1] << 11) | ((uint32_t)o6[2 * g + 1] << 5) | o5[2 * b + 1];
  } else {
    uint32_t d[16];
#pragma unroll
    for (int i = 0; i < 16; i++) d[i] = 0;
    dither_channel<0>(px, d);
    dither_channel<1>(px, d);
    dither_channel<2>(px, d);
    optimize_colors(d, max16, min16);
    mask = (max16 != min16) ? match_colors(px, max16, min16) : 0u;
    if (refine_block(d, max16, min16, mask)) {
      mask = (max16 != min16) ? match_colors(px, max16, min16) : 0u;
    }
  }
  if (max16 < min16) {
    uint32_t t = min16; min16 = max16; max16 = t;
    mask ^= 0x55555555u;
  }
  return make_uint2(max16 | (min16 << 16), mask);
}

// stb__CompressAlphaBlock (stb_dxt.h:551-601).
__device__ __forceinline__ uint2 compress_alpha_block(const uint32_t (&px)[16]) {
  int mn, mx;
  mn = mx = chan<3>(px[0]);
#pragma unroll
  for (int i = 1; i < 16; i++) {
    int a = chan<3>(px[i]);
    mn = min(mn, a);
    mx = max(mx, a);
  }
  const int dist = mx - mn, dist4 = dist * 4, dist2 = dist * 2;
  const int bias = ((dist < 8) ? (dist - 1) : (dist / 2 + 2)) - mn * 7;
  unsigned long long bits = (unsigned long long)mx | ((unsigned long long)mn << 8);
#pragma unroll
  for (int i = 0; i < 16; i++) {
    int a = chan<3>(px[i]) * 7 + bias;
    int ind, t;
    t = (a >= dist4) ? -1 : 0; ind = t & 4; a -= dist4 & t;
    t = (a >= dist2) ? -1 : 0; ind += t & 2; a -= dist2 & t;
    ind += (a >= dist);
    ind = -ind & 7;
    ind ^= (2 > ind);
    bits |= (unsigned long long)ind << (16 + 3 * i);
  }
  return make_uint2((uint32_t)bits, (uint32_t)(bits >> 32));
}

template <bool DXT5>
__global__ void __launch_bounds__(128, 4)
dxt_encode_kernel(const uint32_t *__restrict__ img, uint32_t width, uint32_t blocks_x,
                  uint32_t first_block, uint32_t num_blocks, uint8_t *__restrict__ out) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= num_blocks) return;
  const uint32_t bi = first_block + t;
  uint32_t px[16];
  load_block(img, width, blocks_x, bi, px);
  const uint2 color = compress_color_block(px);
  if (DXT5) {
    const uint2 alpha = compress_alpha_block(px);
    reinterpret_cast<uint4 *>(out)[bi] = make_uint4(alpha.x, alpha.y, color.x, color.y);
  } else {
    reinterpret_cast<uint2 *>(out)[bi] = color;
  }
}

// Host-side construction of stb__OMatch5/6 with the reference's scan order
// (stb_dxt.h:121-147): first strict minimum over mn-major, mx-minor.
void build_omatch(uint8_t *table, int size, bool six) {
  for (int i = 0; i < 256; i++) {
    int best = 256;
    for (int mn = 0; mn < size; mn++)
      for (int mx = 0; mx < size; mx++) {
        int mine = six ? ((mn << 2) | (mn >> 4)) : ((mn << 3) | (mn >> 2));
        int maxe = six ? ((mx << 2) | (mx >> 4)) : ((mx << 3) | (mx >> 2));
        int err = abs((2 * maxe + mine) / 3 - i) + abs(maxe - mine) * 3 / 100;
        if (err < best) {
          table[2 * i] = (uint8_t)mx;
          table[2 * i + 1] = (uint8_t)mn;
          best = err;
        }
      }
  }
}

}  // namespace

cudaError_t dxt_upload_tables() {
  uint8_t host[1024];
  build_omatch(host, 32, false);
  build_omatch(host + 512, 64, true);
  return cudaMemcpyToSymbol(g_omatch, host, sizeof(host));
}

cudaError_t launch_dxt(bool dxt5, const void *rgba_dev, uint32_t width, uint32_t first_block,
                       uint32_t num_blocks, void *out_dev, cudaStream_t stream) {
  if (num_blocks == 0) return cudaSuccess;
  const uint32_t threads = 128, grid = (num_blocks + threads - 1) / threads;
  const uint32_t *img = static_cast<const uint32_t *>(rgba_dev);
  if (dxt5)
    dxt_encode_kernel<true><<<grid, threads, 0, stream>>>(img, width, width / 4, first_block, num_blocks,
                                                          static_cast<uint8_t *>(out_dev));
  else
    dxt_encode_kernel<false><<<grid, threads, 0, stream>>>(img, width, width / 4, first_block, num_blocks,
                                                           static_cast<uint8_t *>(out_dev));
  return cudaGetLastError();
}

}  // namespace fastc
