// ETC1 block encoder for sm_100a.
//
// Behavioural contract: bit-identical to rg_etc1 v1.04, no dithering, at the quality the
// reference drives it with (cLowQuality, template argument Q = 0) and at the library's other two
// levels (Q = 1 cMediumQuality: 3^3 lattice scan; Q = 2 cHighQuality: 9^3 scan with the
// exhaustive per-pixel selector search of evaluate_solution, rg_etc1.cpp:1674-1765, plus the
// constrained solid-colour trial of one-colour sub-blocks, :2035-2147):
//   reference/ETCEncoder/src/Compressor.cpp:26-54       block loop
//   reference/ETCEncoder/src/rg_etc1.cpp:2192-2451      pack_etc1_block
//   reference/ETCEncoder/src/rg_etc1.cpp:1483-1672      etc1_optimizer::compute / init
//   reference/ETCEncoder/src/rg_etc1.cpp:1767-1885      evaluate_solution_fast
//   reference/ETCEncoder/src/rg_etc1.cpp:1951-2033      pack_etc1_block_solid_color
//
// B200 mapping (not how the reference is organised): the reference walks the
// four (flip, 444/555) candidates of a block one after the other with a
// running-best early-out; the candidates are independent, so here a QUAD of
// lanes owns a block and lane q evaluates candidate q (both sub-blocks, the
// second constrained to the first's base colour in 555 differential mode).  A
// quad shuffle-argmin with "lowest candidate index wins ties" reproduces the
// reference's strict-< scan order, and the winning lane packs and stores the
// 8 bytes.  Two observations remove the reference's data-dependent control flow:
//   * evaluate_solution_fast's sorted-luma walk assigns pixel p the selector
//     (2*luma_p >= mid0) + (2*luma_p >= mid1) + (2*luma_p >= mid2) because the
//     midpoints are non-decreasing -- no sort is needed, only min / max luma for
//     the two "skip this table" tests;
//   * squared RGB distances are VABSDIFF4 + DP4A on packed bytes (exact integers), and the four
//     colours of an intensity table come from 16x2 SIMD clamped adds (table_colors).
// A CTA of 128 threads stages its 32 blocks (2 KiB) through shared memory with
// 16 B coalesced row loads; stores are 8 B per block, contiguous per warp.
#include <initializer_list>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace fastc {
namespace {

// ETC1 specification: intensity modifier tables (rg_etc1.cpp:371-375).
__constant__ int c_inten[8][4] = {{-8, -2, 2, 8},     {-17, -5, 5, 17},   {-29, -9, 9, 29},    {-42, -13, 13, 42},
                                  {-60, -18, 18, 60}, {-80, -24, 24, 80}, {-106, -33, 33, 106}, {-183, -47, 47, 183}};

// Solid-colour tables (rg_etc1.cpp:385-507 and :1905-1936), derived on the host by
// the rule the reference's arrays follow (see build_solid_tables) and uploaded once.
constexpr int kMaxCfg = 4096;
__device__ uint16_t g_cfg_off[257];
__device__ uint16_t g_cfg[kMaxCfg];
__device__ uint16_t g_inverse[64 * 256];

__device__ __forceinline__ int clamp255(int v) { return min(max(v, 0), 255); }

struct Sol {
  uint32_t err;    // total squared error of the sub-block
  uint32_t color;  // unscaled base colour r | g << 8 | b << 16
  uint32_t sel;    // 8 x 2-bit selector indices, sub-block pixel order
  int inten;
};

__device__ __forceinline__ uint32_t scale_color(uint32_t c, bool color4) {
  // per byte: 4-bit -> c | c << 4, 5-bit -> c >> 2 | c << 3
  return color4 ? (c | (c << 4)) : (((c >> 2) & 0x07070707u) | (c << 3));
}

// The four colours of intensity table `it` around the scaled base colour, packed r | g << 8 | b << 16,
// and their luma sums.  base_rb = r | b << 16, base_g = g: the clamped adds are 16x2 SIMD
// instructions (add + min 255 / add + max 0 in one), one byte permute packs a colour.
__device__ __forceinline__ void table_colors(uint32_t base_rb, int base_g, int it, uint32_t (&bc)[4], uint32_t (&bi)[4]) {
#pragma unroll
  for (int s = 0; s < 4; s++) {
    const int yd = c_inten[it][s];  // {-L, -S, +S, +L}
    const uint32_t yd2 = ((uint32_t)yd & 0xFFFFu) * 0x10001u;
    const uint32_t rb = s < 2 ? __viaddmax_s16x2_relu(base_rb, yd2, 0u) : __viaddmin_s16x2(base_rb, yd2, 0x00FF00FFu);
    const uint32_t g = (uint32_t)clamp255(base_g + yd);
    bc[s] = __byte_perm(rb, g, 0x5240);
    bi[s] = __dp4a(bc[s], 0x00010101u, 0u);
  }
}

// evaluate_solution_fast (rg_etc1.cpp:1767-1885) for base colour `color` (unscaled).
// px: 8 packed pixels (alpha cleared), luma2: 2 * (r+g+b).  Squared distances are VABSDIFF4 + DP4A on
// packed bytes (exact integers).  The table loop only totals the error; the selectors are derived
// once, for the winning table (they are a pure function of the table and the lumas).
__device__ __forceinline__ void evaluate(const uint32_t (&px)[8], const uint32_t (&luma2)[8],
                                         uint32_t lmin, uint32_t lmax, uint32_t color, bool color4, Sol &trial) {
  const uint32_t base = scale_color(color, color4);
  const uint32_t base_rb = (base & 0xFFu) | ((base & 0xFF0000u));
  const int base_g = (base >> 8) & 0xFF;
  trial.err = 0xFFFFFFFFu;
  trial.color = color;
  trial.inten = 0;
  trial.sel = 0;
#pragma unroll 1
  for (int it = 7; it >= 0; --it) {
    uint32_t bc[4], bi[4];
    table_colors(base_rb, base_g, it, bc, bi);
    const uint32_t mid0 = bi[0] + bi[1], mid1 = bi[1] + bi[2], mid2 = bi[2] + bi[3];
    // the two "all pixels beyond one end" cases may skip the table (rg_etc1.cpp:1809-1836)
    if (lmax * 2 < mid0) {
      if (bi[0] > lmax && bi[0] - lmax >= trial.err) continue;
    } else if (lmin * 2 >= mid2) {
      if (lmin > bi[3] && lmin - bi[3] >= trial.err) continue;
    }
    uint32_t total = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const bool g0 = luma2[i] >= mid0, g1 = luma2[i] >= mid1, g2 = luma2[i] >= mid2;
      const uint32_t c = g2 ? bc[3] : (g1 ? bc[2] : (g0 ? bc[1] : bc[0]));
      const uint32_t d = __vabsdiffu4(px[i], c);
      total = __dp4a(d, d, total);
    }
    if (total < trial.err) {
      trial.err = total;
      trial.inten = it;
      if (!total) break;
    }
  }
  if (trial.err != 0xFFFFFFFFu) {
    uint32_t bc[4], bi[4];
    table_colors(base_rb, base_g, trial.inten, bc, bi);
    const uint32_t mid0 = bi[0] + bi[1], mid1 = bi[1] + bi[2], mid2 = bi[2] + bi[3];
    uint32_t sel = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const bool g0 = luma2[i] >= mid0, g1 = luma2[i] >= mid1, g2 = luma2[i] >= mid2;
      sel |= ((uint32_t)g0 + (uint32_t)g1 + (uint32_t)g2) << (2 * i);
    }
    trial.sel = sel;
  }
}

// etc1_optimizer::init + compute (rg_etc1.cpp:1627-1672, 1483-1625) for one 8-pixel
// sub-block at cLowQuality (scan delta {0}: one lattice point, <= 2 refinement trials).
// constrain: differential mode's second sub-block, base5 = first sub-block's colour.
__device__ __forceinline__ bool optimize(const uint32_t (&px)[8], bool color4, bool constrain, uint32_t base5, Sol &best) {
  const int limit = color4 ? 15 : 31;
  uint32_t luma2[8];
  uint32_t sr = 0, sg = 0, sb = 0, lmin = 0xFFFFFFFFu, lmax = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint32_t r = px[i] & 0xFF, g = (px[i] >> 8) & 0xFF, b = px[i] >> 16;
    sr += r; sg += g; sb += b;
    const uint32_t l = r + g + b;
    lmin = min(lmin, l);
    lmax = max(lmax, l);
    luma2[i] = 2 * l;
  }
  const float flimit = (float)limit;
  const float avg[3] = {__fmul_rn((float)sr, 0.125f), __fmul_rn((float)sg, 0.125f), __fmul_rn((float)sb, 0.125f)};
  int m[3];
#pragma unroll
  for (int k = 0; k < 3; k++)
    m[k] = min(max(__float2int_rz(__fadd_rn(__fdiv_rn(__fmul_rn(avg[k], flimit), 255.0f), 0.5f)), 0), limit);

  auto allowed = [&](int r, int g, int b) {
    if (!constrain) return true;
    const int dr = r - (int)(base5 & 0xFF), dg = g - (int)((base5 >> 8) & 0xFF), db = b - (int)(base5 >> 16);
    return min(dr, min(dg, db)) >= -4 && max(dr, max(dg, db)) <= 3;
  };

  best.err = 0xFFFFFFFFu;
  if (!allowed(m[0], m[1], m[2])) return false;
  evaluate(px, luma2, lmin, lmax, (uint32_t)m[0] | ((uint32_t)m[1] << 8) | ((uint32_t)m[2] << 16), color4, best);

#pragma unroll 1
  for (int trial = 0; trial < 2; trial++) {
    const uint32_t base = scale_color(best.color, color4);
    // sum of the clamped intensity deltas actually applied under the best selectors
    int cnt[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const uint32_t s = (best.sel >> (2 * i)) & 3;
#pragma unroll
      for (int q = 0; q < 4; q++) cnt[q] += (s == (uint32_t)q);
    }
    int ds[3] = {0, 0, 0};
#pragma unroll
    for (int s = 0; s < 4; s++) {
      const int yd = c_inten[best.inten][s];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const int bk = (base >> (8 * k)) & 0xFF;
        ds[k] += cnt[s] * (clamp255(bk + yd) - bk);
      }
    }
    if (!ds[0] && !ds[1] && !ds[2]) break;
    int n1[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const float ad = __fdiv_rn((float)ds[k], 8.0f);
      const float f = __fadd_rn(__fdiv_rn(__fmul_rn(__fsub_rn(avg[k], ad), flimit), 255.0f), 0.5f);
      n1[k] = min(max(__float2int_rz(f), 0), limit);  // x86 cvttss2si, then clamp<int>
    }
    if (n1[0] == m[0] && n1[1] == m[1] && n1[2] == m[2]) break;
    const uint32_t ncol = (uint32_t)n1[0] | ((uint32_t)n1[1] << 8) | ((uint32_t)n1[2] << 16);
    if (ncol == best.color) break;
    if (!allowed(n1[0], n1[1], n1[2])) break;
    Sol t;
    evaluate(px, luma2, lmin, lmax, ncol, color4, t);
    if (t.err < best.err) best = t;
    else break;
  }
  return true;
}

// evaluate_solution (rg_etc1.cpp:1674-1765): every selector of every intensity table for every
// pixel, first strict minimum in ascending selector / table order.  Per pixel the four squared
// distances (VABSDIFF4 + DP4A, < 2^18) become keys distance * 4 + selector, so one min over the
// keys picks the distance and the selector together, the lower selector on ties.
#ifndef FASTC_ETC1_PRUNE_MASK
#define FASTC_ETC1_PRUNE_MASK 0x7F  // the pixels after which the bound is tested (measured: every one, r02_ap_etc1.log)
#endif
__device__ __forceinline__ void evaluate_full(const uint32_t (&px)[8], uint32_t color, bool color4, uint32_t bound, Sol &trial) {
  const uint32_t base = scale_color(color, color4);
  const uint32_t base_rb = (base & 0xFFu) | ((base & 0xFF0000u));
  const int base_g = (base >> 8) & 0xFF;
  trial.err = 0xFFFFFFFFu;
  trial.color = color;
  trial.inten = 0;
  trial.sel = 0;
#pragma unroll 1
  for (int it = 0; it < 8; it++) {
    uint32_t bc[4], bi[4];
    table_colors(base_rb, base_g, it, bc, bi);
    uint32_t total = 0, sel = 0;
    bool pruned = false;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      uint32_t key = 0xFFFFFFFFu;
#pragma unroll
      for (uint32_t s2 = 0; s2 < 4; s2++) {
        const uint32_t d = __vabsdiffu4(px[i], bc[s2]);
        key = min(key, __dp4a(d, d, 0u) * 4u + s2);
      }
      total += key >> 2;
      sel |= (key & 3u) << (2 * i);
#ifndef FASTC_ETC1_NO_PRUNE
      // A table whose partial sum has reached the error of the best table of this point, or of the
      // best point so far (`bound`), cannot become either (the sums only grow; the caller accepts a
      // point on strict <): the reference's own "total >= trial error" break (rg_etc1.cpp:1739),
      // taken earlier.  The lanes of a warp scan the same lattice offset at the same time, so for an
      // offset that is far from the optimum every lane is past its bound after a few pixels and
      // the whole warp leaves the table; a lane whose neighbours go on just goes on with them.
      if (((FASTC_ETC1_PRUNE_MASK >> i) & 1) && __all_sync(__activemask(), total >= min(trial.err, bound))) { pruned = true; break; }
#endif
    }
    if (!pruned && total < trial.err) {
      trial.err = total;
      trial.inten = it;
      trial.sel = sel;
    }
  }
}

// etc1_optimizer at cMediumQuality / cHighQuality: init once, then compute() over one or two sets
// of scan deltas (rg_etc1.cpp:1483-1625, :2283-2340); the running best survives between the calls.
template <int Q>
struct Optimizer {
  uint32_t px[8], luma2[8];
  uint32_t lmin, lmax, base5;
  float avg[3];
  int m[3], limit;
  bool color4, constrain, valid;
  Sol best;

  __device__ __forceinline__ void init(const uint32_t (&sub)[8], bool c4, bool cons, uint32_t b5) {
    color4 = c4; constrain = cons; base5 = b5;
    limit = c4 ? 15 : 31;
    uint32_t sr = 0, sg = 0, sb = 0;
    lmin = 0xFFFFFFFFu; lmax = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      px[i] = sub[i];
      const uint32_t r = px[i] & 0xFF, g = (px[i] >> 8) & 0xFF, b = px[i] >> 16;
      sr += r; sg += g; sb += b;
      const uint32_t l = r + g + b;
      lmin = min(lmin, l);
      lmax = max(lmax, l);
      luma2[i] = 2 * l;
    }
    const float flimit = (float)limit;
    avg[0] = __fmul_rn((float)sr, 0.125f); avg[1] = __fmul_rn((float)sg, 0.125f); avg[2] = __fmul_rn((float)sb, 0.125f);
#pragma unroll
    for (int k = 0; k < 3; k++)
      m[k] = min(max(__float2int_rz(__fadd_rn(__fdiv_rn(__fmul_rn(avg[k], flimit), 255.0f), 0.5f)), 0), limit);
    best.err = 0xFFFFFFFFu; best.color = 0; best.sel = 0; best.inten = 0;
    valid = false;
  }

  __device__ __forceinline__ bool allowed(int r, int g, int b) const {
    if (!constrain) return true;
    const int dr = r - (int)(base5 & 0xFF), dg = g - (int)((base5 >> 8) & 0xFF), db = b - (int)(base5 >> 16);
    return min(dr, min(dg, db)) >= -4 && max(dr, max(dg, db)) <= 3;
  }

  // evaluate_solution(_fast) + "better than the running best?"
  __device__ __forceinline__ bool try_point(int r, int g, int b) {
    if (!allowed(r, g, b)) return false;
    const uint32_t col = (uint32_t)r | ((uint32_t)g << 8) | ((uint32_t)b << 16);
    Sol t;
    if (Q == 2) evaluate_full(px, col, color4, best.err, t);
    else evaluate(px, luma2, lmin, lmax, col, color4, t);
    if (t.err < best.err) { best = t; valid = true; return true; }
    return false;
  }

  // deltas: see pack_deltas
  __device__ __forceinline__ void compute(uint64_t deltas, int ndeltas) {
    const float flimit = (float)limit;
#pragma unroll 1
    for (int zi = 0; zi < ndeltas; zi++) {
      const int zd = (int)((deltas >> (5 * zi)) & 31) - 8, mbb = m[2] + zd;
      if (mbb < 0) continue;
      if (mbb > limit) break;
#pragma unroll 1
      for (int yi = 0; yi < ndeltas; yi++) {
        const int yd = (int)((deltas >> (5 * yi)) & 31) - 8, mbg = m[1] + yd;
        if (mbg < 0) continue;
        if (mbg > limit) break;
#pragma unroll 1
        for (int xi = 0; xi < ndeltas; xi++) {
          const int xd = (int)((deltas >> (5 * xi)) & 31) - 8, mbr = m[0] + xd;
          if (mbr < 0) continue;
          if (mbr > limit) break;
          if (!try_point(mbr, mbg, mbb)) continue;
          const int max_trials = ((xd | yd | zd) == 0) ? 4 : 2;
#pragma unroll 1
          for (int trial = 0; trial < max_trials; trial++) {
            const uint32_t base = scale_color(best.color, color4);
            int cnt[4] = {0, 0, 0, 0};
#pragma unroll
            for (int i = 0; i < 8; i++) {
              const uint32_t s = (best.sel >> (2 * i)) & 3;
#pragma unroll
              for (int q = 0; q < 4; q++) cnt[q] += (s == (uint32_t)q);
            }
            int ds[3] = {0, 0, 0};
#pragma unroll
            for (int s = 0; s < 4; s++) {
              const int ydl = c_inten[best.inten][s];
#pragma unroll
              for (int k = 0; k < 3; k++) {
                const int bk = (base >> (8 * k)) & 0xFF;
                ds[k] += cnt[s] * (clamp255(bk + ydl) - bk);
              }
            }
            if (!ds[0] && !ds[1] && !ds[2]) break;
            int n1[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
              const float ad = __fdiv_rn((float)ds[k], 8.0f);
              const float f = __fadd_rn(__fdiv_rn(__fmul_rn(__fsub_rn(avg[k], ad), flimit), 255.0f), 0.5f);
              n1[k] = min(max(__float2int_rz(f), 0), limit);  // x86 cvttss2si, then clamp<int>
            }
            if (n1[0] == mbr && n1[1] == mbg && n1[2] == mbb) break;
            const uint32_t ncol = (uint32_t)n1[0] | ((uint32_t)n1[1] << 8) | ((uint32_t)n1[2] << 16);
            if (ncol == best.color) break;
            if (n1[0] == m[0] && n1[1] == m[1] && n1[2] == m[2]) break;
            if (!try_point(n1[0], n1[1], n1[2])) break;
          }
        }
      }
    }
  }
};

// scan deltas packed five bits each (value + 8), in the reference's array order (rg_etc1.cpp:2283-2334)
constexpr uint64_t pack_deltas(std::initializer_list<int> d) {
  uint64_t v = 0;
  int k = 0;
  for (int x : d) v |= (uint64_t)(x + 8) << (5 * k++);
  return v;
}
constexpr uint64_t kScan1 = pack_deltas({-1, 0, 1});
constexpr uint64_t kScan4 = pack_deltas({-4, -3, -2, -1, 0, 1, 2, 3, 4});
constexpr uint64_t kScan23 = pack_deltas({-3, -2, 2, 3});
constexpr uint64_t kScan55 = pack_deltas({-5, 5});
constexpr uint64_t kScan58 = pack_deltas({-8, -7, -6, -5, 5, 6, 7, 8});

// pack_etc1_block_solid_color_constrained (rg_etc1.cpp:2035-2147) for an 8-pixel sub-block of one
// colour: best exact-table configuration of the given mode, optionally within the differential
// range of sub-block 0's base colour.  Returns false when no configuration is admissible.
__device__ __noinline__ bool solid_constrained(uint32_t pixel, bool use_diff, bool have_base, uint32_t base5, Sol &res) {
  const int col[3] = {(int)(pixel & 0xFF), (int)((pixel >> 8) & 0xFF), (int)((pixel >> 16) & 0xFF)};
  const int b5[3] = {(int)(base5 & 0xFF), (int)((base5 >> 8) & 0xFF), (int)(base5 >> 16)};
  const int next_comp[4] = {1, 2, 0, 1};
  uint32_t best_error = 0xFFFFFFFFu, best_x = 0, best_c1 = 0, best_c2 = 0;
  int best_i = 0;
  bool perfect = false;
  for (int i = 0; i < 3 && !perfect; i++) {
    const int c1 = col[next_comp[i]], c2 = col[next_comp[i + 1]];
    for (int delta = -1; delta <= 1 && !perfect; delta++) {
      const int cpd = clamp255(col[i] + delta);
      const int d0 = cpd - col[i];
      for (uint32_t k = g_cfg_off[cpd]; k < g_cfg_off[cpd + 1]; k++) {
        const uint32_t x = g_cfg[k];
        const bool diff = x & 1;
        if (diff != use_diff) continue;
        const bool lim = diff && have_base;
        if (lim) {
          const int d = (int)((x >> 8) & 255) - b5[i];
          if (d < -4 || d > 3) continue;
        }
        const uint32_t p1 = g_inverse[(x & 0xFF) * 256 + c1], p2 = g_inverse[(x & 0xFF) * 256 + c2];
        if (lim) {
          const int d1 = (int)(p1 & 0xFF) - b5[next_comp[i]], d2 = (int)(p2 & 0xFF) - b5[next_comp[i + 1]];
          if (d1 < -4 || d1 > 3 || d2 < -4 || d2 > 3) continue;
        }
        const uint32_t err = (uint32_t)(d0 * d0) + (p1 >> 8) * (p1 >> 8) + (p2 >> 8) * (p2 >> 8);
        if (err < best_error) {
          best_error = err; best_x = x; best_c1 = p1 & 0xFF; best_c2 = p2 & 0xFF; best_i = i;
          if (!err) { perfect = true; break; }
        }
      }
    }
  }
  if (best_error == 0xFFFFFFFFu) return false;
  res.err = best_error * 8u;
  res.inten = (int)((best_x >> 1) & 7);
  res.sel = ((best_x >> 4) & 3) * 0x5555u;
  uint32_t c[3] = {0, 0, 0};
  const uint32_t vals[3] = {(best_x >> 8) & 255, best_c1, best_c2};
  const int where[3] = {best_i, next_comp[best_i], next_comp[best_i + 1]};
#pragma unroll
  for (int k = 0; k < 3; k++)
#pragma unroll
    for (int w = 0; w < 3; w++)
      if (where[k] == w) c[w] = vals[k];
  res.color = c[0] | (c[1] << 8) | (c[2] << 16);
  return true;
}

// pack_etc1_block_solid_color (rg_etc1.cpp:1951-2033)
__device__ uint2 pack_solid(uint32_t pixel) {
  const int col[3] = {(int)(pixel & 0xFF), (int)((pixel >> 8) & 0xFF), (int)((pixel >> 16) & 0xFF)};
  const int next_comp[4] = {1, 2, 0, 1};
  uint32_t best_error = 0xFFFFFFFFu, best_x = 0, best_c1 = 0, best_c2 = 0;
  int best_i = 0;
  bool perfect = false;
  for (int i = 0; i < 3 && !perfect; i++) {
    const int c1 = col[next_comp[i]], c2 = col[next_comp[i + 1]];
    for (int delta = -1; delta <= 1 && !perfect; delta++) {
      const int cpd = clamp255(col[i] + delta);
      const int d0 = cpd - col[i];
      for (uint32_t k = g_cfg_off[cpd]; k < g_cfg_off[cpd + 1]; k++) {
        const uint32_t x = g_cfg[k];
        const uint32_t p1 = g_inverse[(x & 0xFF) * 256 + c1], p2 = g_inverse[(x & 0xFF) * 256 + c2];
        const uint32_t err = (uint32_t)(d0 * d0) + (p1 >> 8) * (p1 >> 8) + (p2 >> 8) * (p2 >> 8);
        if (err < best_error) {
          best_error = err; best_x = x; best_c1 = p1 & 0xFF; best_c2 = p2 & 0xFF; best_i = i;
          if (!err) { perfect = true; break; }
        }
      }
    }
  }
  const uint32_t diff = best_x & 1, inten = (best_x >> 1) & 7;
  const uint32_t e = (0x4B >> (2 * ((best_x >> 4) & 3))) & 3;  // selector index -> ETC1 code {3,2,0,1}
  uint32_t bytes[3];
  const uint32_t vals[3] = {(best_x >> 8) & 255, best_c1, best_c2};
  const int where[3] = {best_i, next_comp[best_i], next_comp[best_i + 1]};
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const uint32_t v = diff ? (vals[k] << 3) : (vals[k] | (vals[k] << 4));
#pragma unroll
    for (int w = 0; w < 3; w++)
      if (where[k] == w) bytes[w] = v & 0xFF;
  }
  const uint32_t b3 = ((inten | (inten << 3)) << 2) | (diff << 1);
  uint2 o;
  o.x = bytes[0] | (bytes[1] << 8) | (bytes[2] << 16) | (b3 << 24);
  o.y = ((e & 2) ? 0x0000FFFFu : 0u) | ((e & 1) ? 0xFFFF0000u : 0u);
  return o;
}

// The 8 bytes of a block from its winning candidate (rg_etc1.cpp:2369-2448).
__device__ __forceinline__ uint2 pack_block(bool flip, bool color4, const Sol &r0, const Sol &r1) {
  const uint32_t c0 = r0.color, c1 = r1.color;
  uint32_t bytes[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const int a = (c0 >> (8 * k)) & 0xFF, b = (c1 >> (8 * k)) & 0xFF;
    if (color4) bytes[k] = (uint32_t)(b | (a << 4));
    else bytes[k] = (uint32_t)((a << 3) | ((b - a) & 7));
  }
  const uint32_t b3 = ((uint32_t)r1.inten << 2) | ((uint32_t)r0.inten << 5) | ((color4 ? 0u : 1u) << 1) | (flip ? 1u : 0u);
  uint32_t lsb = 0, msb = 0;
#pragma unroll
  for (int y = 0; y < 4; y++)
#pragma unroll
    for (int x = 0; x < 4; x++) {
      const int sbf = y >> 1, kf = (y & 1) * 4 + x;  // flipped
      const int sbn = x >> 1, kn = (x & 1) * 4 + y;  // not flipped
      const uint32_t sf = ((sbf ? r1.sel : r0.sel) >> (2 * kf)) & 3, sn = ((sbn ? r1.sel : r0.sel) >> (2 * kn)) & 3;
      const uint32_t s = flip ? sf : sn;
      const uint32_t e = (0x4B >> (2 * s)) & 3;  // selector index -> ETC1 code {3,2,0,1}
      lsb |= (e & 1) << (x * 4 + y);
      msb |= (e >> 1) << (x * 4 + y);
    }
  uint2 o;
  o.x = bytes[0] | (bytes[1] << 8) | (bytes[2] << 16) | (b3 << 24);
  o.y = (msb >> 8) | ((msb & 0xFF) << 8) | ((lsb >> 8) << 16) | ((lsb & 0xFF) << 24);
  return o;
}

constexpr int kEtcThreads = 128;
constexpr int kEtcBlocksPerCta = kEtcThreads / 4;

template <int Q>
__global__ void __launch_bounds__(kEtcThreads, Q == 0 ? 6 : 4)
etc1_encode_kernel(const uint32_t *__restrict__ img, uint32_t width, uint32_t blocks_x, uint32_t first_block,
                   uint32_t num_blocks, uint2 *__restrict__ out) {
  __shared__ uint32_t s_px[kEtcBlocksPerCta][17];  // +1 word: quads of a warp hit distinct banks
  const int q = threadIdx.x >> 2, cand = threadIdx.x & 3;
  const uint32_t t = blockIdx.x * kEtcBlocksPerCta + q;
  const bool valid = t < num_blocks;
  const uint32_t bi = first_block + (valid ? t : 0);
  if (valid) {
    // lane `cand` of the quad loads row `cand` of the block: 16 B, contiguous across the warp's quads
    const uint32_t bx = bi % blocks_x, by = bi / blocks_x;
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(img + (size_t)(by * 4 + cand) * width + (size_t)bx * 4));
    s_px[q][4 * cand + 0] = v.x; s_px[q][4 * cand + 1] = v.y; s_px[q][4 * cand + 2] = v.z; s_px[q][4 * cand + 3] = v.w;
  }
  __syncwarp();
  if (!valid) return;  // whole quads leave together

  uint32_t blk[16];
  bool solid = true;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    blk[i] = s_px[q][i];
    solid = solid && (blk[i] == blk[0]);  // includes the alpha byte
  }
  if (solid) {
    if (cand == 0) out[bi] = pack_solid(blk[0]);
    return;
  }

  // ---- candidate `cand`: flip = cand >> 1, 444 mode = cand & 1 (reference loop order, rg_etc1.cpp:2251-2254)
  const bool flip = cand >> 1, color4 = cand & 1;
  Sol res[2] = {{0xFFFFFFFFu, 0, 0, 0}, {0xFFFFFFFFu, 0, 0, 0}};
  bool ok = true;
  uint32_t total = 0;
#pragma unroll
  for (int sb = 0; sb < 2; sb++) {
    uint32_t sub[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      // flipped: rows 2*sb, 2*sb+1 in raster order; else columns 2*sb, 2*sb+1, column-major
      const uint32_t a = blk[sb * 8 + i], b = blk[sb * 2 + (i >> 2) + 4 * (i & 3)];
      sub[i] = (flip ? a : b) & 0x00FFFFFFu;
    }
    if (ok) {
      if (Q == 0) {
        ok = optimize(sub, color4, !color4 && sb == 1, res[0].color, res[sb]);
      } else {
        // a one-colour sub-block also tries the exact solid-colour tables (rg_etc1.cpp:2259-2269)
        Sol solid_res;
        bool have_solid = false;
        if (sb == 1 || color4) {
          bool same = true;
#pragma unroll
          for (int i = 1; i < 8; i++) same = same && sub[i] == sub[0];
          if (same) have_solid = solid_constrained(sub[0], !color4, sb == 1 && !color4, res[0].color, solid_res);
        }
        Optimizer<Q> o;
        o.init(sub, color4, !color4 && sb == 1, res[0].color);
        o.compute(Q == 2 ? kScan4 : kScan1, Q == 2 ? 9 : 3);
        ok = o.valid;
        if (ok) {
          if (o.best.err > 3000u) {  // refinement_error_thresh0 / 1 (:2312-2340)
            if (Q == 1) o.compute(kScan23, 4);
            else if (o.best.err > 6000u) o.compute(kScan58, 8);
            else o.compute(kScan55, 2);
          }
          res[sb] = (have_solid && solid_res.err < o.best.err) ? solid_res : o.best;
        }
      }
      if (ok) total += res[sb].err;
    }
  }
  // strict-< scan over the candidates in index order == min over (error, index)
  uint32_t key = ok ? ((total << 2) | (uint32_t)cand) : 0xFFFFFFFFu;
  const uint32_t qmask = 0xFu << ((threadIdx.x & 31) & ~3);
  uint32_t best = key;
  best = min(best, __shfl_xor_sync(qmask, best, 1));
  best = min(best, __shfl_xor_sync(qmask, best, 2));
  if (best != key) return;

  out[bi] = pack_block(flip, color4, res[0], res[1]);
}

// ---- cLowQuality, dynamically scheduled.  In the quad kernel above a lane's candidate takes 2 to 6
// evaluations (one per sub-block plus up to two refinement trials each) and the warp waits for its
// slowest lane at every call: 19 of 32 lanes are busy on average.  Here a CTA owns a tile of 256
// blocks whose pixels it stages in shared memory, the (block, candidate) pairs are TASKS handed out
// through a shared cursor, and every lane runs the optimizer as a state machine -- fetch a task, set
// up a sub-block, evaluate, decide what comes next -- so that the one expensive step, evaluate(), is
// executed by the whole warp on every trip, each lane for whatever task / sub-block / trial it has
// reached.  Candidate results are parked in shared memory; a last pass picks each block's winner in
// the reference's order (strict <, lowest candidate index first) and packs it.
constexpr int kDynTile = 256;

__device__ __forceinline__ bool etc1_allowed(bool constrain, uint32_t base5, int r, int g, int b) {
  if (!constrain) return true;
  const int dr = r - (int)(base5 & 0xFF), dg = g - (int)((base5 >> 8) & 0xFF), db = b - (int)(base5 >> 16);
  return min(dr, min(dg, db)) >= -4 && max(dr, max(dg, db)) <= 3;
}

__global__ void __launch_bounds__(kEtcThreads, 6)
etc1_encode_dyn_kernel(const uint32_t *__restrict__ img, uint32_t width, uint32_t blocks_x, uint32_t first_block,
                       uint32_t num_blocks, uint2 *__restrict__ out) {
  __shared__ uint32_t s_blk[kDynTile][17];  // +1 word: lanes on different blocks hit different banks
  __shared__ uint4 s_res[4][kDynTile];      // per candidate: error, colour 0 | inten 0 << 24, colour 1 | inten 1 << 24, sel 0 | sel 1 << 16
  __shared__ uint32_t s_solid[kDynTile / 32];
  __shared__ uint32_t s_next;
  const int tid = threadIdx.x;
  const uint32_t tile0 = blockIdx.x * kDynTile;
  const int nbt = (int)min((uint32_t)kDynTile, num_blocks - tile0);
  if (tid < kDynTile / 32) s_solid[tid] = 0;
  if (tid == 0) s_next = 0;
  // stage the tile: consecutive threads take the same row of consecutive blocks (16 B each, contiguous
  // while the blocks share a block row)
  for (int item = tid; item < 4 * nbt; item += kEtcThreads) {
    const int row = item / nbt, b = item - row * nbt;
    const uint32_t bi = first_block + tile0 + b;
    const uint32_t bx = bi % blocks_x, by = bi / blocks_x;
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(img + (size_t)(by * 4 + row) * width + (size_t)bx * 4));
    s_blk[b][4 * row + 0] = v.x; s_blk[b][4 * row + 1] = v.y; s_blk[b][4 * row + 2] = v.z; s_blk[b][4 * row + 3] = v.w;
  }
  __syncthreads();
  for (int b = tid; b < nbt; b += kEtcThreads) {
    bool solid = true;
#pragma unroll
    for (int i = 1; i < 16; i++) solid = solid && (s_blk[b][i] == s_blk[b][0]);  // includes the alpha byte
    if (solid) atomicOr(&s_solid[b >> 5], 1u << (b & 31));
  }
  __syncthreads();

  // ---- the optimizer as a per-lane state machine (rg_etc1.cpp:2251-2367 with optimize() unrolled in time)
  enum { kNeedTask, kNeedSub, kReady, kDone };
  const int ntasks = 4 * nbt;
  int stage = kNeedTask, task = 0, blk = 0, sub = 0, ktrial = 0;
  bool flip = false, color4 = false, constrain = false;
  uint32_t px[8], luma2[8], lmin = 0, lmax = 0, mcol = 0, cur_color = 0;
  float avg[3] = {0.0f, 0.0f, 0.0f};
  Sol best = {0xFFFFFFFFu, 0, 0, 0}, res0 = {0xFFFFFFFFu, 0, 0, 0};
  for (;;) {
    // [setup] lanes between tasks / sub-blocks; a constrained second sub-block whose mean colour is
    // out of the differential range ends the candidate at once (rg_etc1.cpp:1511-1518), so try again
#pragma unroll 1
    for (int tries = 0; tries < 3 && (stage == kNeedTask || stage == kNeedSub); tries++) {
      if (stage == kNeedTask) {
        for (;;) {
          task = (int)atomicAdd(&s_next, 1u);
          if (task >= ntasks) { stage = kDone; break; }
          const int cand = task & 3;  // (block-major: measured 3 % faster than candidate-major)
          blk = task >> 2;
          if ((s_solid[blk >> 5] >> (blk & 31)) & 1u) continue;
          flip = cand >> 1; color4 = cand & 1;  // reference loop order, rg_etc1.cpp:2251-2254
          sub = 0;
          stage = kNeedSub;
          break;
        }
      }
      if (stage == kNeedSub) {
        uint32_t sr = 0, sg = 0, sb = 0;
        lmin = 0xFFFFFFFFu; lmax = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
          // flipped: rows 2*sub, 2*sub+1 in raster order; else columns 2*sub, 2*sub+1, column-major
          const int pos = flip ? sub * 8 + i : sub * 2 + (i >> 2) + 4 * (i & 3);
          px[i] = s_blk[blk][pos] & 0x00FFFFFFu;
          const uint32_t r = px[i] & 0xFF, g = (px[i] >> 8) & 0xFF, b = px[i] >> 16;
          sr += r; sg += g; sb += b;
          const uint32_t l = r + g + b;
          lmin = min(lmin, l);
          lmax = max(lmax, l);
          luma2[i] = 2 * l;
        }
        const int limit = color4 ? 15 : 31;
        const float flimit = (float)limit;
        avg[0] = __fmul_rn((float)sr, 0.125f); avg[1] = __fmul_rn((float)sg, 0.125f); avg[2] = __fmul_rn((float)sb, 0.125f);
        int m[3];
#pragma unroll
        for (int k = 0; k < 3; k++)
          m[k] = min(max(__float2int_rz(__fadd_rn(__fdiv_rn(__fmul_rn(avg[k], flimit), 255.0f), 0.5f)), 0), limit);
        mcol = (uint32_t)m[0] | ((uint32_t)m[1] << 8) | ((uint32_t)m[2] << 16);
        constrain = !color4 && sub == 1;
        if (!etc1_allowed(constrain, res0.color, m[0], m[1], m[2])) {
          s_res[task & 3][blk] = make_uint4(0xFFFFFFFFu, 0u, 0u, 0u);  // the candidate fails
          stage = kNeedTask;
        } else {
          cur_color = mcol;
          ktrial = 0;
          stage = kReady;
        }
      }
    }
    if (__all_sync(0xffffffffu, stage == kDone)) break;

    // [evaluate] the one expensive step, by every lane that has something to evaluate
    Sol t = {0xFFFFFFFFu, 0, 0, 0};
    if (stage == kReady) evaluate(px, luma2, lmin, lmax, cur_color, color4, t);

    // [decide] etc1_optimizer::compute's refinement loop (rg_etc1.cpp:1531-1612), one trip per evaluation
    if (stage == kReady) {
      bool finish = false;
      if (ktrial == 0) best = t;
      else if (t.err < best.err) best = t;
      else finish = true;
      if (!finish && ktrial == 2) finish = true;  // both refinement trials are used up
      if (!finish) {
        const int limit = color4 ? 15 : 31;
        const float flimit = (float)limit;
        const uint32_t base = scale_color(best.color, color4);
        // sum of the clamped intensity deltas actually applied under the best selectors
        int cnt[4] = {0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const uint32_t s = (best.sel >> (2 * i)) & 3;
#pragma unroll
          for (int q = 0; q < 4; q++) cnt[q] += (s == (uint32_t)q);
        }
        int ds[3] = {0, 0, 0};
#pragma unroll
        for (int s = 0; s < 4; s++) {
          const int yd = c_inten[best.inten][s];
#pragma unroll
          for (int k = 0; k < 3; k++) {
            const int bk = (base >> (8 * k)) & 0xFF;
            ds[k] += cnt[s] * (clamp255(bk + yd) - bk);
          }
        }
        if (!ds[0] && !ds[1] && !ds[2]) {
          finish = true;
        } else {
          int n1[3];
#pragma unroll
          for (int k = 0; k < 3; k++) {
            const float ad = __fdiv_rn((float)ds[k], 8.0f);
            const float f = __fadd_rn(__fdiv_rn(__fmul_rn(__fsub_rn(avg[k], ad), flimit), 255.0f), 0.5f);
            n1[k] = min(max(__float2int_rz(f), 0), limit);  // x86 cvttss2si, then clamp<int>
          }
          const uint32_t ncol = (uint32_t)n1[0] | ((uint32_t)n1[1] << 8) | ((uint32_t)n1[2] << 16);
          if (ncol == mcol || ncol == best.color || !etc1_allowed(constrain, res0.color, n1[0], n1[1], n1[2])) {
            finish = true;
          } else {
            cur_color = ncol;
            ktrial++;
          }
        }
      }
      if (finish) {
        if (sub == 0) {
          res0 = best;
          sub = 1;
          stage = kNeedSub;
        } else {
          s_res[task & 3][blk] = make_uint4(res0.err + best.err, res0.color | ((uint32_t)res0.inten << 24),
                                             best.color | ((uint32_t)best.inten << 24), res0.sel | (best.sel << 16));
          stage = kNeedTask;
        }
      }
    }
  }
  __syncthreads();

  // ---- winners: strict-< scan over the candidates in index order, then pack
  for (int b = tid; b < nbt; b += kEtcThreads) {
    const uint32_t bi = first_block + tile0 + b;
    if ((s_solid[b >> 5] >> (b & 31)) & 1u) {
      out[bi] = pack_solid(s_blk[b][0]);
      continue;
    }
    uint4 w = s_res[0][b];
    int wc = 0;
#pragma unroll
    for (int cand = 1; cand < 4; cand++) {
      const uint4 r = s_res[cand][b];
      if (r.x < w.x) { w = r; wc = cand; }
    }
    const Sol r0 = {w.x, w.y & 0x00FFFFFFu, w.w & 0xFFFFu, (int)(w.y >> 24)};
    const Sol r1 = {w.x, w.z & 0x00FFFFFFu, w.w >> 16, (int)(w.z >> 24)};
    out[bi] = pack_block(wc >> 1, wc & 1, r0, r1);
  }
}

// rg_etc1.cpp:1887-1901
int decode_value(int diff, int inten, int selector, int packed_c) {
  static const int kInten[8][4] = {{-8, -2, 2, 8},     {-17, -5, 5, 17},   {-29, -9, 9, 29},    {-42, -13, 13, 42},
                                   {-60, -18, 18, 60}, {-80, -24, 24, 80}, {-106, -33, 33, 106}, {-183, -47, 47, 183}};
  int c = diff ? ((packed_c >> 2) | (packed_c << 3)) : (packed_c | (packed_c << 4));
  c += kInten[inten][selector];
  return c < 0 ? 0 : (c > 255 ? 255 : c);
}

// Host construction of the solid-colour tables.
//   inverse[diff | inten << 1 | selector << 4][v] = best base | abs error << 8, first strict
//     minimum over ascending bases (pack_etc1_block_init, rg_etc1.cpp:1905-1936);
//   config list of value v = every (diff, inten, selector) that reproduces v exactly, with
//     the smallest such base, ordered by (diff, inten, base, selector) -- the order of the
//     reference's literal arrays (rg_etc1.cpp:385-507), which decides ties in the search.
void build_solid_tables(std::vector<uint16_t> &off, std::vector<uint16_t> &cfg, std::vector<uint16_t> &inverse) {
  inverse.assign(64 * 256, 0);
  for (int diff = 0; diff < 2; diff++)
    for (int inten = 0; inten < 8; inten++)
      for (int sel = 0; sel < 4; sel++)
        for (int v = 0; v < 256; v++) {
          uint32_t best = 0xFFFFFFFFu, best_c = 0;
          for (int pc = 0; pc < (diff ? 32 : 16); pc++) {
            const int d = decode_value(diff, inten, sel, pc) - v;
            const uint32_t err = (uint32_t)(d < 0 ? -d : d);
            if (err < best) { best = err; best_c = (uint32_t)pc; if (!best) break; }
          }
          inverse[(diff + (inten << 1) + (sel << 4)) * 256 + v] = (uint16_t)(best_c | (best << 8));
        }
  off.assign(257, 0);
  cfg.clear();
  for (int v = 0; v < 256; v++) {
    off[v] = (uint16_t)cfg.size();
    for (int diff = 0; diff < 2; diff++)
      for (int inten = 0; inten < 8; inten++) {
        // entries of this (diff, inten), by ascending base then selector
        for (int pc = 0; pc < (diff ? 32 : 16); pc++)
          for (int sel = 0; sel < 4; sel++) {
            if (decode_value(diff, inten, sel, pc) != v) continue;
            bool smaller = false;  // a smaller base already reproduces v with this selector
            for (int p2 = 0; p2 < pc; p2++) smaller = smaller || decode_value(diff, inten, sel, p2) == v;
            if (!smaller) cfg.push_back((uint16_t)(diff | (inten << 1) | (sel << 4) | (pc << 8)));
          }
      }
  }
  off[256] = (uint16_t)cfg.size();
}

}  // namespace

cudaError_t etc1_upload_tables() {
  static std::vector<uint16_t> off, cfg, inverse;
  if (off.empty()) build_solid_tables(off, cfg, inverse);
  if (cfg.size() > (size_t)kMaxCfg) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemcpyToSymbol(g_cfg_off, off.data(), off.size() * 2);
  if (e != cudaSuccess) return e;
  e = cudaMemcpyToSymbol(g_cfg, cfg.data(), cfg.size() * 2);
  if (e != cudaSuccess) return e;
  return cudaMemcpyToSymbol(g_inverse, inverse.data(), inverse.size() * 2);
}

cudaError_t launch_etc1(const void *rgba_dev, uint32_t width, uint32_t first_block, uint32_t num_blocks,
                        void *out_dev, int quality, cudaStream_t stream) {
  if (num_blocks == 0) return cudaSuccess;
  const uint32_t grid = (num_blocks + kEtcBlocksPerCta - 1) / kEtcBlocksPerCta;
  const uint32_t *img = static_cast<const uint32_t *>(rgba_dev);
  uint2 *out = static_cast<uint2 *>(out_dev);
  if (quality == 0)
    etc1_encode_dyn_kernel<<<(num_blocks + kDynTile - 1) / kDynTile, kEtcThreads, 0, stream>>>(img, width, width / 4, first_block,
                                                                                              num_blocks, out);
  else if (quality == 1) etc1_encode_kernel<1><<<grid, kEtcThreads, 0, stream>>>(img, width, width / 4, first_block, num_blocks, out);
  else if (quality == 2) etc1_encode_kernel<2><<<grid, kEtcThreads, 0, stream>>>(img, width, width / 4, first_block, num_blocks, out);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

}  // namespace fastc
