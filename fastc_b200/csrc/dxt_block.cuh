// The DXT1 / DXT5 block encoder of dxt.cu as host+device code, so that tests/native/dxt_host_check.cpp
// can run the very same arithmetic on the CPU against the oracle (the product only ever runs it on
// the device: nothing in libfastc_gpu.so calls these functions from host code).
//
// Behavioural contract: bit-identical to stb_dxt v1.06 driven the way the reference drives it
// (STB_DXT_DITHER, one refinement pass):
//   reference/DXTEncoder/src/stb_dxt.h:477-548 (colour block), :551-601 (alpha block).
//
// Formulation (not how stb_dxt is organised).  The encoder is ~integer-instruction bound, so the
// work is arranged for packed-byte instructions:
//   * the dithered block is kept PLANAR: one word per (channel, row) holding the row's four
//     quantised bytes.  Sums are dp4a(word, 0x01010101); the six covariance sums are
//     sum(a*b) - mu_b*sum(a) - mu_a*sum(b) + 16*mu_a*mu_b with sum(a*b) = four dp4a(A_row, B_row)
//     (exact integers, equal to the reference's sum((a-mu_a)*(b-mu_b))); the least-squares sums of
//     stb__RefineBlock are dp4a(A_row, W_row) with the row's four index weights spread to bytes;
//   * per-pixel projections (16-bit signed direction x unsigned bytes) are two dp2a;
//   * the extreme-projection pixels are found as min / max over keys dot * 16 + index (first index
//     wins ties, like the reference's strict compares) and fetched from a scratch copy of the
//     interleaved dithered block by index;
//   * the 5/6-bit quantisation tables are two multiply-shifts:
//     mul8bit(x, 31) == (x * 7967 + 32896) >> 16, expand5(q) == (q * 33) >> 2 (checked exhaustively
//     in tests/test_tables.py).
// The original pixels and the interleaved dithered pixels live in caller-provided row storage
// (shared memory columns on the device), not in registers: the kernel runs at 8 CTAs / SM.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FASTC_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define FASTC_HD inline
struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
#endif

namespace fastc {
namespace dxtb {

// ---- primitives (device: one instruction each; host: plain C for the CPU check)
FASTC_HD uint32_t dp4a_uu(uint32_t a, uint32_t b, uint32_t c) {
#ifdef __CUDA_ARCH__
  return __dp4a(a, b, c);
#else
  for (int k = 0; k < 4; k++) c += ((a >> (8 * k)) & 0xFF) * ((b >> (8 * k)) & 0xFF);
  return c;
#endif
}
// a: two signed 16-bit halves, b: four unsigned bytes.  lo: a.h0*b.b0 + a.h1*b.b1 + c, hi: bytes 2, 3.
FASTC_HD int dp2a_lo_su(uint32_t a, uint32_t b, int c) {
#ifdef __CUDA_ARCH__
  int d;
  asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
#else
  return c + (int)(int16_t)(a & 0xFFFF) * (int)(b & 0xFF) + (int)(int16_t)(a >> 16) * (int)((b >> 8) & 0xFF);
#endif
}
FASTC_HD int dp2a_hi_su(uint32_t a, uint32_t b, int c) {
#ifdef __CUDA_ARCH__
  int d;
  asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
#else
  return c + (int)(int16_t)(a & 0xFFFF) * (int)((b >> 16) & 0xFF) + (int)(int16_t)(a >> 16) * (int)(b >> 24);
#endif
}
FASTC_HD uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t s) {
#ifdef __CUDA_ARCH__
  return __byte_perm(a, b, s);
#else
  const uint64_t v = ((uint64_t)b << 32) | a;
  uint32_t r = 0;
  for (int k = 0; k < 4; k++) r |= (uint32_t)((v >> (8 * ((s >> (4 * k)) & 7))) & 0xFF) << (8 * k);
  return r;
#endif
}
FASTC_HD int imin(int a, int b) { return a < b ? a : b; }
FASTC_HD int imax(int a, int b) { return a > b ? a : b; }
// float arithmetic without contraction (the reference is an SSE2 scalar build)
FASTC_HD float fmul(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
FASTC_HD float fadd(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}
FASTC_HD float fdiv(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fdiv_rn(a, b);
#else
  return a / b;
#endif
}

// pack two signed 16-bit halves
FASTC_HD uint32_t pack16(int lo, int hi) { return ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16); }

// stb__Mul8Bit(x, 31) / (x, 63) for x in [0, 255] (stb_dxt.h:82-86)
FASTC_HD int q5(int x) { return (x * 7967 + 32896) >> 16; }
FASTC_HD int q6(int x) { return (x * 16191 + 32896) >> 16; }
FASTC_HD int expand5(int q) { return (q * 33) >> 2; }  // (q << 3) | (q >> 2)
FASTC_HD int expand6(int q) { return (q * 65) >> 4; }  // (q << 2) | (q >> 4)
// stb__QuantRBTab / stb__QuantGTab[x + 8] with the table's clamp folded in (stb_dxt.h:612-617)
template <int CH>
FASTC_HD int quant(int x) {
  x = imin(imax(x, 0), 255);
  return CH == 1 ? expand6(q6(x)) : expand5(q5(x));
}
FASTC_HD int lerp13(int a, int b) { return (2 * a + b) / 3; }
FASTC_HD uint32_t as16bit(uint32_t p) {  // stb__As16Bit (stb_dxt.h:88-91)
  return ((uint32_t)q5(p & 0xFF) << 11) + ((uint32_t)q6((p >> 8) & 0xFF) << 5) + (uint32_t)q5((p >> 16) & 0xFF);
}
template <int CH>
FASTC_HD int chan(uint32_t p) { return (p >> (8 * CH)) & 0xFF; }

// Row storage of one block: row y of the original pixels / of the interleaved dithered pixels is
// px[y * stride] / d[y * stride] (device: a shared-memory column per thread; host: stride 1).
struct Rows {
  uint4 *px;
  uint4 *d;
  int stride;
};

// Planar dithered block + what stb__OptimizeColorsBlock and stb__RefineBlock need from it.
struct Planar {
  uint32_t p[3][4];  // [channel][row]: bytes = the row's four dithered values
  int sum[3], lo[3], hi[3];
};

// stb__DitherBlock (stb_dxt.h:159-183), one channel: Floyd-Steinberg to the 565 grid, row-serial.
template <int CH>
FASTC_HD void dither_channel(const Rows &R, Planar &P) {
  int ea[4] = {0, 0, 0, 0}, eb[4] = {0, 0, 0, 0};
  int lo = 255, hi = 0, sum = 0;
#pragma unroll
  for (int y = 0; y < 4; y++) {
    const uint4 row = R.px[y * R.stride];
    int(&e1)[4] = (y & 1) ? eb : ea;
    int(&e2)[4] = (y & 1) ? ea : eb;
    const int b0 = chan<CH>(row.x), b1 = chan<CH>(row.y), b2 = chan<CH>(row.z), b3 = chan<CH>(row.w);
    const int q0 = quant<CH>(b0 + ((3 * e2[1] + 5 * e2[0]) >> 4));
    e1[0] = b0 - q0;
    const int q1 = quant<CH>(b1 + ((7 * e1[0] + 3 * e2[2] + 5 * e2[1] + e2[0]) >> 4));
    e1[1] = b1 - q1;
    const int q2 = quant<CH>(b2 + ((7 * e1[1] + 3 * e2[3] + 5 * e2[2] + e2[1]) >> 4));
    e1[2] = b2 - q2;
    const int q3 = quant<CH>(b3 + ((7 * e1[2] + 5 * e2[3] + e2[2]) >> 4));
    e1[3] = b3 - q3;
    lo = imin(lo, imin(imin(q0, q1), imin(q2, q3)));
    hi = imax(hi, imax(imax(q0, q1), imax(q2, q3)));
    sum += q0 + q1 + q2 + q3;
    P.p[CH][y] = (uint32_t)q0 | ((uint32_t)q1 << 8) | ((uint32_t)q2 << 16) | ((uint32_t)q3 << 24);
  }
  P.lo[CH] = lo;
  P.hi[CH] = hi;
  P.sum[CH] = sum;
}

// interleave the planar rows into R.d (alpha byte = a copy of blue: never used)
FASTC_HD void store_dithered(const Rows &R, const Planar &P) {
#pragma unroll
  for (int y = 0; y < 4; y++) {
    const uint32_t t = byte_perm(P.p[0][y], P.p[1][y], 0x5140), u = byte_perm(P.p[0][y], P.p[1][y], 0x7362);
    const uint32_t b = P.p[2][y];
    R.d[y * R.stride] = make_uint4(byte_perm(t, b, 0x4410), byte_perm(t, b, 0x5532), byte_perm(u, b, 0x6610),
                                   byte_perm(u, b, 0x7732));
  }
}

// stb__OptimizeColorsBlock (stb_dxt.h:283-385) on the dithered block.
FASTC_HD void optimize_colors(const Rows &R, const Planar &P, uint32_t &max16, uint32_t &min16) {
  int mu[3];
#pragma unroll
  for (int c = 0; c < 3; c++) mu[c] = (P.sum[c] + 8) >> 4;
  // cov[] in the reference's order: rr rg rb gg gb bb
  int cov[6];
  {
    int e = 0;
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
      for (int b = a; b < 3; b++, e++) {
        uint32_t s = 0;
#pragma unroll
        for (int y = 0; y < 4; y++) s = dp4a_uu(P.p[a][y], P.p[b][y], s);
        cov[e] = (int)s - mu[b] * P.sum[a] - mu[a] * P.sum[b] + 16 * mu[a] * mu[b];
      }
  }
  float f[6];
#pragma unroll
  for (int e = 0; e < 6; e++) f[e] = fdiv((float)cov[e], 255.0f);
  float vr = (float)(P.hi[0] - P.lo[0]), vg = (float)(P.hi[1] - P.lo[1]), vb = (float)(P.hi[2] - P.lo[2]);
#pragma unroll
  for (int it = 0; it < 4; it++) {
    const float r = fadd(fadd(fmul(vr, f[0]), fmul(vg, f[1])), fmul(vb, f[2]));
    const float g = fadd(fadd(fmul(vr, f[1]), fmul(vg, f[3])), fmul(vb, f[4]));
    const float b = fadd(fadd(fmul(vr, f[2]), fmul(vg, f[4])), fmul(vb, f[5]));
    vr = r; vg = g; vb = b;
  }
  const float magf = fmaxf(fmaxf(fabsf(vr), fabsf(vg)), fabsf(vb));  // finite values: plain max is exact
  int v_r, v_g, v_b;
  if (magf < 4.0f) {
    v_r = 299; v_g = 587; v_b = 114;
  } else {
    // the reference does this part in double (stb_dxt.h:361-364)
#ifdef __CUDA_ARCH__
    const double magn = __ddiv_rn(512.0, (double)magf);
    v_r = (int)__dmul_rn((double)vr, magn);
    v_g = (int)__dmul_rn((double)vg, magn);
    v_b = (int)__dmul_rn((double)vb, magn);
#else
    const double magn = 512.0 / (double)magf;
    v_r = (int)((double)vr * magn);
    v_g = (int)((double)vg * magn);
    v_b = (int)((double)vb * magn);
#endif
  }
  // extreme projections: keys dot * 16 + index; |dot| <= 3 * 255 * 587 < 2^19
  const uint32_t vlo = pack16(v_r, v_g), vhi = pack16(v_b, 0);
  int kmin = 0x7fffffff, kmax = -0x7fffffff - 1;
#pragma unroll
  for (int y = 0; y < 4; y++) {
    const uint4 row = R.d[y * R.stride];
    const uint32_t w[4] = {row.x, row.y, row.z, row.w};
#pragma unroll
    for (int x = 0; x < 4; x++) {
      const int dot = dp2a_hi_su(vhi, w[x], dp2a_lo_su(vlo, w[x], 0));
      kmin = imin(kmin, dot * 16 + (4 * y + x));
      kmax = imax(kmax, dot * 16 + (15 - (4 * y + x)));
    }
  }
  const int i_min = kmin & 15, i_max = 15 - (kmax & 15);
  const uint32_t *dw = reinterpret_cast<const uint32_t *>(R.d);
  const uint32_t minp = dw[(i_min >> 2) * R.stride * 4 + (i_min & 3)];
  const uint32_t maxp = dw[(i_max >> 2) * R.stride * 4 + (i_max & 3)];
  max16 = as16bit(maxp);
  min16 = as16bit(minp);
}

// Palette (stb__EvalColors, stb_dxt.h:149-155) + dithered index selection (stb__MatchColorsBlock's
// dither branch, stb_dxt.h:186-280) on the ORIGINAL block.
FASTC_HD uint32_t match_colors(const Rows &R, uint32_t c0_16, uint32_t c1_16) {
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
  const uint32_t dlo = pack16(dr, dg), dhi = pack16(db, 0);

  uint32_t mask = 0;
  int ea[4] = {0, 0, 0, 0}, eb[4] = {0, 0, 0, 0};
#pragma unroll
  for (int y = 0; y < 4; y++) {
    int(&e1)[4] = (y & 1) ? eb : ea;
    int(&e2)[4] = (y & 1) ? ea : eb;
    const uint4 row = R.px[y * R.stride];
    const uint32_t w[4] = {row.x, row.y, row.z, row.w};
#pragma unroll
    for (int x = 0; x < 4; x++) {
      const int dp = dp2a_hi_su(dhi, w[x], dp2a_lo_su(dlo, w[x], 0));
      int acc;
      if (x == 0) acc = 3 * e2[1] + 5 * e2[0];
      else if (x == 1) acc = 7 * e1[0] + 3 * e2[2] + 5 * e2[1] + e2[0];
      else if (x == 2) acc = 7 * e1[1] + 3 * e2[3] + 5 * e2[2] + e2[1];
      else acc = 7 * e1[2] + 5 * e2[3] + e2[2];
      const int dot = dp * 16 + acc;
      // (stop, step) picked together: key = stop * 4 + step
      const int k0 = stops[0] * 4, k1 = stops[1] * 4 + 1, k2 = stops[2] * 4 + 2, k3 = stops[3] * 4 + 3;
      const int key = dot < halfp ? (dot < c0p ? k1 : k3) : (dot < c3p ? k2 : k0);
      e1[x] = dp - (key >> 2);
      mask |= (uint32_t)(key & 3) << (8 * y + 2 * x);
    }
  }
  return mask;
}

FASTC_HD int sclamp(float y, int hi) {
  const int x = (int)y;  // cvt.rzi, same as x86 cvttss2si for these in-range values
  return x < 0 ? 0 : (x > hi ? hi : x);
}

// stb__RefineBlock (stb_dxt.h:398-474) on the dithered block.  omatch: stb__OMatch5 | stb__OMatch6.
// Returns true if the endpoints changed.
FASTC_HD bool refine_block(const Planar &P, const uint8_t *omatch, uint32_t &max16, uint32_t &min16, uint32_t mask) {
  const uint32_t old_min = min16, old_max = max16;
  uint32_t nmin, nmax;
  if ((mask ^ (mask << 2)) < 4u) {
    const int r = (P.sum[0] + 8) >> 4, g = (P.sum[1] + 8) >> 4, b = (P.sum[2] + 8) >> 4;
    const uint8_t *o5 = omatch, *o6 = omatch + 512;
    nmax = ((uint32_t)o5[2 * r] << 11) | ((uint32_t)o6[2 * g] << 5) | o5[2 * b];
    nmin = ((uint32_t)o5[2 * r + 1] << 11) | ((uint32_t)o6[2 * g + 1] << 5) | o5[2 * b + 1];
  } else {
    uint32_t xx = 0, yy = 0, xy = 0, a1[3] = {0, 0, 0};
#pragma unroll
    for (int y = 0; y < 4; y++) {
      // the row's four 2-bit indices spread to bytes, then index -> weight w1Tab = {3, 0, 2, 1}
      uint32_t s = (mask >> (8 * y)) & 0xFFu;
      s = (s | (s << 12)) & 0x000F000Fu;
      s = (s | (s << 6)) & 0x03030303u;
      const uint32_t s0 = s & 0x01010101u, s1 = (s >> 1) & 0x01010101u;
      const uint32_t w1 = ((s0 ^ 0x01010101u) << 1) | (s0 ^ s1 ^ 0x01010101u);
      const uint32_t w2 = 0x03030303u - w1;
      xx = dp4a_uu(w1, w1, xx);
      yy = dp4a_uu(w2, w2, yy);
      xy = dp4a_uu(w1, w2, xy);
#pragma unroll
      for (int c = 0; c < 3; c++) a1[c] = dp4a_uu(P.p[c][y], w1, a1[c]);
    }
    const int ixx = (int)xx, iyy = (int)yy, ixy = (int)xy;
    const int a1r = (int)a1[0], a1g = (int)a1[1], a1b = (int)a1[2];
    const int a2r = 3 * P.sum[0] - a1r, a2g = 3 * P.sum[1] - a1g, a2b = 3 * P.sum[2] - a1b;
    const float frb = fdiv(fdiv(93.0f, 255.0f), (float)(ixx * iyy - ixy * ixy));
    const float fg = fdiv(fmul(frb, 63.0f), 31.0f);
    nmax = (uint32_t)sclamp(fadd(fmul((float)(a1r * iyy - a2r * ixy), frb), 0.5f), 31) << 11;
    nmax |= (uint32_t)sclamp(fadd(fmul((float)(a1g * iyy - a2g * ixy), fg), 0.5f), 63) << 5;
    nmax |= (uint32_t)sclamp(fadd(fmul((float)(a1b * iyy - a2b * ixy), frb), 0.5f), 31);
    nmin = (uint32_t)sclamp(fadd(fmul((float)(a2r * ixx - a1r * ixy), frb), 0.5f), 31) << 11;
    nmin |= (uint32_t)sclamp(fadd(fmul((float)(a2g * ixx - a1g * ixy), fg), 0.5f), 63) << 5;
    nmin |= (uint32_t)sclamp(fadd(fmul((float)(a2b * ixx - a1b * ixy), frb), 0.5f), 31);
  }
  min16 = nmin;
  max16 = nmax;
  return old_min != nmin || old_max != nmax;
}

// stb__CompressColorBlock (stb_dxt.h:477-548), mode = STB_DXT_DITHER.  R.px holds the block;
// `constant`: all 16 pixels equal as 32-bit words (alpha included).
FASTC_HD uint2 compress_color_block(const Rows &R, bool constant, const uint8_t *omatch) {
  uint32_t mask, max16, min16;
  if (constant) {
    const uint32_t p = R.px[0].x;
    const int r = chan<0>(p), g = chan<1>(p), b = chan<2>(p);
    const uint8_t *o5 = omatch, *o6 = omatch + 512;
    mask = 0xaaaaaaaau;
    max16 = ((uint32_t)o5[2 * r] << 11) | ((uint32_t)o6[2 * g] << 5) | o5[2 * b];
    min16 = ((uint32_t)o5[2 * r + 1] << 11) | ((uint32_t)o6[2 * g + 1] << 5) | o5[2 * b + 1];
  } else {
    Planar P;
    dither_channel<0>(R, P);
    dither_channel<1>(R, P);
    dither_channel<2>(R, P);
    store_dithered(R, P);
    optimize_colors(R, P, max16, min16);
    mask = (max16 != min16) ? match_colors(R, max16, min16) : 0u;
    if (refine_block(P, omatch, max16, min16, mask)) {
      mask = (max16 != min16) ? match_colors(R, max16, min16) : 0u;
    }
  }
  if (max16 < min16) {
    const uint32_t t = min16; min16 = max16; max16 = t;
    mask ^= 0x55555555u;
  }
  return make_uint2(max16 | (min16 << 16), mask);
}

// stb__CompressAlphaBlock (stb_dxt.h:551-601).
FASTC_HD uint2 compress_alpha_block(const Rows &R) {
  uint32_t a[16];
#pragma unroll
  for (int y = 0; y < 4; y++) {
    const uint4 row = R.px[y * R.stride];
    a[4 * y + 0] = row.x >> 24; a[4 * y + 1] = row.y >> 24; a[4 * y + 2] = row.z >> 24; a[4 * y + 3] = row.w >> 24;
  }
  int mn = (int)a[0], mx = (int)a[0];
#pragma unroll
  for (int i = 1; i < 16; i++) {
    mn = imin(mn, (int)a[i]);
    mx = imax(mx, (int)a[i]);
  }
  const int dist = mx - mn, dist4 = dist * 4, dist2 = dist * 2;
  const int bias = ((dist < 8) ? (dist - 1) : (dist / 2 + 2)) - mn * 7;
  unsigned long long bits = (unsigned long long)mx | ((unsigned long long)mn << 8);
#pragma unroll
  for (int i = 0; i < 16; i++) {
    int v = (int)a[i] * 7 + bias;
    int ind, t;
    t = (v >= dist4) ? -1 : 0; ind = t & 4; v -= dist4 & t;
    t = (v >= dist2) ? -1 : 0; ind += t & 2; v -= dist2 & t;
    ind += (v >= dist);
    ind = -ind & 7;
    ind ^= (2 > ind);
    bits |= (unsigned long long)ind << (16 + 3 * i);
  }
  return make_uint2((uint32_t)bits, (uint32_t)(bits >> 32));
}

// Host-side construction of stb__OMatch5/6 with the reference's scan order
// (stb_dxt.h:121-147): first strict minimum over mn-major, mx-minor.
inline int iabs(int v) { return v < 0 ? -v : v; }
inline void build_omatch(uint8_t *table, int size, bool six) {
  for (int i = 0; i < 256; i++) {
    int best = 256;
    for (int mn = 0; mn < size; mn++)
      for (int mx = 0; mx < size; mx++) {
        int mine = six ? ((mn << 2) | (mn >> 4)) : ((mn << 3) | (mn >> 2));
        int maxe = six ? ((mx << 2) | (mx >> 4)) : ((mx << 3) | (mx >> 2));
        int err = iabs((2 * maxe + mine) / 3 - i) + abs(maxe - mine) * 3 / 100;
        if (err < best) {
          table[2 * i] = (uint8_t)mx;
          table[2 * i + 1] = (uint8_t)mn;
          best = err;
        }
      }
  }
}


}  // namespace dxtb
}  // namespace fastc
