// TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the reference's ETC1
// block encoder.  Never linked into or called from the product path.
//
// Restates rg_etc1 v1.04 as FasTC drives it (cLowQuality, dithering off:
// /root/reference/ETCEncoder/src/Compressor.cpp:26-54 and
// /root/reference/ETCEncoder/src/rg_etc1.cpp).  Pinned bit-for-bit against the
// compiled reference (oracle/_ref) by tests/test_oracle_vs_ref.py and against
// tests/golden/*.npz.
//
// The solid-colour configuration lists (rg_etc1.cpp:385-507) are not stored
// here: they are re-derived by the rule they follow (every exact
// (diff, inten, selector, base) configuration of an 8-bit value, smallest base
// per selector, ordered by diff, inten table, base, selector);
// tests/test_tables.py checks the derivation against the reference's arrays.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "oracle.h"

namespace {

// ETC1 specification: intensity modifier tables (rg_etc1.cpp:371-375).
const int kInten[8][4] = {{-8, -2, 2, 8},     {-17, -5, 5, 17},   {-29, -9, 9, 29},    {-42, -13, 13, 42},
                          {-60, -18, 18, 60}, {-80, -24, 24, 80}, {-106, -33, 33, 106}, {-183, -47, 47, 183}};
// selector index (position in the modifier table) -> ETC1 2-bit code (rg_etc1.cpp:378)
const uint8_t kSelToEtc1[4] = {3, 2, 0, 1};

int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
int sq(int v) { return v * v; }

// rg_etc1.cpp:1887-1901
int decode_value(int diff, int inten, int selector, int packed_c) {
  int c = diff ? ((packed_c >> 2) | (packed_c << 3)) : (packed_c | (packed_c << 4));
  return clampi(c + kInten[inten][selector], 0, 255);
}

struct SolidTables {
  uint16_t inverse[64][256];              // rg_etc1.cpp:1905-1936
  std::vector<uint16_t> config[256];      // rg_etc1.cpp:385-507 (derived, see header)
};

const SolidTables &solid_tables() {
  static SolidTables t;
  static bool init = false;
  if (init) return t;
  for (int diff = 0; diff < 2; diff++) {
    const int limit = diff ? 32 : 16;
    for (int inten = 0; inten < 8; inten++)
      for (int sel = 0; sel < 4; sel++) {
        const int idx = diff + (inten << 1) + (sel << 4);
        for (int color = 0; color < 256; color++) {
          uint32_t best = 0xFFFFFFFFu, best_c = 0;
          for (int pc = 0; pc < limit; pc++) {
            const uint32_t err = (uint32_t)std::abs(decode_value(diff, inten, sel, pc) - color);
            if (err < best) {
              best = err;
              best_c = (uint32_t)pc;
              if (!best) break;
            }
          }
          t.inverse[idx][color] = (uint16_t)(best_c | (best << 8));
        }
      }
  }
  for (int color = 0; color < 256; color++) {
    struct E { int diff, inten, pc, sel; };
    std::vector<E> es;
    for (int diff = 0; diff < 2; diff++)
      for (int inten = 0; inten < 8; inten++)
        for (int sel = 0; sel < 4; sel++)
          for (int pc = 0; pc < (diff ? 32 : 16); pc++)
            if (decode_value(diff, inten, sel, pc) == color) {
              es.push_back({diff, inten, pc, sel});
              break;  // smallest base per (diff, inten, selector)
            }
    std::stable_sort(es.begin(), es.end(), [](const E &a, const E &b) {
      if (a.diff != b.diff) return a.diff < b.diff;
      if (a.inten != b.inten) return a.inten < b.inten;
      if (a.pc != b.pc) return a.pc < b.pc;
      return a.sel < b.sel;
    });
    for (const E &e : es) t.config[color].push_back((uint16_t)(e.diff | (e.inten << 1) | (e.sel << 4) | (e.pc << 8)));
  }
  init = true;
  return t;
}

// pack_etc1_block_solid_color (rg_etc1.cpp:1951-2033)
void pack_solid(uint8_t *out, const uint8_t *color) {
  const SolidTables &T = solid_tables();
  static const int next_comp[4] = {1, 2, 0, 1};
  uint32_t best_error = 0xFFFFFFFFu, best_i = 0;
  int best_x = 0, best_c1 = 0, best_c2 = 0;
  bool perfect = false;
  for (int i = 0; i < 3 && !perfect; i++) {
    const int c1 = color[next_comp[i]], c2 = color[next_comp[i + 1]];
    for (int delta = -1; delta <= 1 && !perfect; delta++) {
      const int cpd = clampi(color[i] + delta, 0, 255);
      for (uint16_t x : T.config[cpd]) {
        const uint16_t p1 = T.inverse[x & 0xFF][c1], p2 = T.inverse[x & 0xFF][c2];
        const uint32_t err = (uint32_t)(sq(cpd - color[i]) + sq(p1 >> 8) + sq(p2 >> 8));
        if (err < best_error) {
          best_error = err;
          best_x = x;
          best_c1 = p1 & 0xFF;
          best_c2 = p2 & 0xFF;
          best_i = (uint32_t)i;
          if (!best_error) { perfect = true; break; }
        }
      }
    }
  }
  const int diff = best_x & 1, inten = (best_x >> 1) & 7;
  out[3] = (uint8_t)(((inten | (inten << 3)) << 2) | (diff << 1));
  const int e = kSelToEtc1[(best_x >> 4) & 3];
  out[4] = out[5] = (e & 2) ? 0xFF : 0;
  out[6] = out[7] = (e & 1) ? 0xFF : 0;
  const int c0 = (best_x >> 8) & 255;
  const int vals[3] = {c0, best_c1, best_c2};
  const int where[3] = {(int)best_i, next_comp[best_i], next_comp[best_i + 1]};
  for (int k = 0; k < 3; k++)
    out[where[k]] = (uint8_t)(diff ? (vals[k] << 3) : (vals[k] | (vals[k] << 4)));
}

struct Solution {
  int r = 0, g = 0, b = 0;  // unscaled base colour (4 or 5 bit)
  int inten = 0;
  uint8_t sel[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  uint64_t error = ~0ull;
  bool valid = false;
};

// etc1_optimizer for one 8-pixel subblock (rg_etc1.cpp:1346-1885).  quality 0 (cLowQuality, what
// FasTC passes): one lattice point + at most two refinement trials, fast evaluation; 1 (cMedium):
// a 3^3 scan, fast evaluation; 2 (cHigh): a 9^3 scan, exhaustive per-pixel selector search.
struct Optimizer {
  const uint8_t (*px)[4];  // 8 source pixels
  int quality = 0;
  bool color4;
  bool constrain;
  int base5[3];
  int limit;
  float avg[3];
  int br, bg, bb;
  uint32_t sorted_luma[8];
  uint32_t sorted_idx[8];
  Solution best;

  // etc1_optimizer::init (rg_etc1.cpp:1627-1672)
  void init() {
    limit = color4 ? 15 : 31;
    float sum[3] = {0.0f, 0.0f, 0.0f};
    uint16_t luma[8];
    for (int i = 0; i < 8; i++) {
      for (int k = 0; k < 3; k++) sum[k] += (float)px[i][k];
      luma[i] = (uint16_t)(px[i][0] + px[i][1] + px[i][2]);
      sorted_idx[i] = (uint32_t)i;
    }
    for (int k = 0; k < 3; k++) avg[k] = sum[k] * (1.0f / 8.0f);
    br = clampi((int)(uint32_t)(avg[0] * limit / 255.0f + .5f), 0, limit);
    bg = clampi((int)(uint32_t)(avg[1] * limit / 255.0f + .5f), 0, limit);
    bb = clampi((int)(uint32_t)(avg[2] * limit / 255.0f + .5f), 0, limit);
    // indirect_radix_sort (rg_etc1.cpp:746-921): stable, ascending by luma
    std::stable_sort(sorted_idx, sorted_idx + 8, [&](uint32_t a, uint32_t b) { return luma[a] < luma[b]; });
    for (int i = 0; i < 8; i++) sorted_luma[i] = luma[sorted_idx[i]];
    best = Solution();
  }

  void scaled(int r, int g, int b, int out[3]) const {
    if (color4) { out[0] = r | (r << 4); out[1] = g | (g << 4); out[2] = b | (b << 4); }
    else { out[0] = (r >> 2) | (r << 3); out[1] = (g >> 2) | (g << 3); out[2] = (b >> 2) | (b << 3); }
  }

  bool violates_base5(int r, int g, int b) const {
    if (!constrain) return false;
    const int dr = r - base5[0], dg = g - base5[1], db = b - base5[2];
    return std::min(dr, std::min(dg, db)) < -4 || std::max(dr, std::max(dg, db)) > 3;
  }

  // evaluate_solution (rg_etc1.cpp:1674-1765): every selector of every intensity table per pixel
  bool evaluate_full(int r, int g, int b) {
    if (violates_base5(r, g, b)) return false;
    int base[3];
    scaled(r, g, b, base);
    Solution trial;
    uint8_t tmp[8];
    for (int it = 0; it < 8; it++) {
      int bc[4][3];
      for (int s = 0; s < 4; s++)
        for (int k = 0; k < 3; k++) bc[s][k] = clampi(base[k] + kInten[it][s], 0, 255);
      uint64_t total = 0;
      for (int c = 0; c < 8; c++) {
        uint32_t be = 0xFFFFFFFFu;
        int bs = 0;
        for (int sidx = 0; sidx < 4; sidx++) {
          const uint32_t e = (uint32_t)(sq(px[c][0] - bc[sidx][0]) + sq(px[c][1] - bc[sidx][1]) + sq(px[c][2] - bc[sidx][2]));
          if (e < be) { be = e; bs = sidx; }
        }
        tmp[c] = (uint8_t)bs;
        total += be;
        if (total >= trial.error) break;
      }
      if (total < trial.error) {
        trial.error = total;
        trial.inten = it;
        memcpy(trial.sel, tmp, 8);
        trial.valid = true;
      }
    }
    trial.r = r; trial.g = g; trial.b = b;
    if (trial.error < best.error) { best = trial; return true; }
    return false;
  }

  // evaluate_solution_fast (rg_etc1.cpp:1767-1885)
  bool evaluate_fast(int r, int g, int b) {
    if (violates_base5(r, g, b)) return false;
    int base[3];
    scaled(r, g, b, base);
    Solution trial;
    uint8_t tmp[8];
    for (int it = 7; it >= 0; --it) {
      int bc[4][3];
      uint32_t bi[4];
      for (int s = 0; s < 4; s++) {
        bi[s] = 0;
        for (int k = 0; k < 3; k++) { bc[s][k] = clampi(base[k] + kInten[it][s], 0, 255); bi[s] += (uint32_t)bc[s][k]; }
      }
      const uint32_t mid[3] = {bi[0] + bi[1], bi[1] + bi[2], bi[2] + bi[3]};
      auto dist = [&](int s, int p) { return (uint64_t)(sq(bc[s][0] - px[p][0]) + sq(bc[s][1] - px[p][1]) + sq(bc[s][2] - px[p][2])); };
      uint64_t total = 0;
      if (sorted_luma[7] * 2 < mid[0]) {
        if (bi[0] > sorted_luma[7] && (uint64_t)(bi[0] - sorted_luma[7]) >= trial.error) continue;
        memset(tmp, 0, 8);
        for (int c = 0; c < 8; c++) total += dist(0, c);
      } else if (sorted_luma[0] * 2 >= mid[2]) {
        if (sorted_luma[0] > bi[3] && (uint64_t)(sorted_luma[0] - bi[3]) >= trial.error) continue;
        memset(tmp, 3, 8);
        for (int c = 0; c < 8; c++) total += dist(3, c);
      } else {
        uint32_t cur = 0;
        int c = 0;
        bool done = false;
        for (; c < 8 && !done; c++) {
          const uint32_t y = sorted_luma[c];
          while (y * 2 >= mid[cur])
            if (++cur > 2) { done = true; break; }
          if (done) break;
          tmp[sorted_idx[c]] = (uint8_t)cur;
          total += dist((int)cur, (int)sorted_idx[c]);
        }
        for (; c < 8; c++) {
          tmp[sorted_idx[c]] = 3;
          total += dist(3, (int)sorted_idx[c]);
        }
      }
      if (total < trial.error) {
        trial.error = total;
        trial.inten = it;
        memcpy(trial.sel, tmp, 8);
        trial.valid = true;
        if (!total) break;
      }
    }
    trial.r = r; trial.g = g; trial.b = b;
    if (trial.error < best.error) { best = trial; return true; }
    return false;
  }

  bool evaluate(int r, int g, int b) { return quality == 2 ? evaluate_full(r, g, b) : evaluate_fast(r, g, b); }

  // etc1_optimizer::compute (rg_etc1.cpp:1483-1625) over the lattice points centre + deltas^3
  bool compute(const int *deltas, int ndeltas) {
    for (int zi = 0; zi < ndeltas; zi++) {
      const int zd = deltas[zi], mbb = bb + zd;
      if (mbb < 0) continue; else if (mbb > limit) break;
      for (int yi = 0; yi < ndeltas; yi++) {
        const int yd = deltas[yi], mbg = bg + yd;
        if (mbg < 0) continue; else if (mbg > limit) break;
        for (int xi = 0; xi < ndeltas; xi++) {
          const int xd = deltas[xi], mbr = br + xd;
          if (mbr < 0) continue; else if (mbr > limit) break;
          if (!evaluate(mbr, mbg, mbb)) continue;
          const int max_trials = quality == 0 ? 2 : (((xd | yd | zd) == 0) ? 4 : 2);
          for (int trial = 0; trial < max_trials; trial++) {
            int base[3];
            scaled(best.r, best.g, best.b, base);
            int ds[3] = {0, 0, 0};
            for (int i = 0; i < 8; i++) {
              const int ydl = kInten[best.inten][best.sel[i]];
              for (int k = 0; k < 3; k++) ds[k] += clampi(base[k] + ydl, 0, 255) - base[k];
            }
            if (!ds[0] && !ds[1] && !ds[2]) break;
            int n1[3];
            for (int k = 0; k < 3; k++) {
              const float ad = (float)ds[k] / 8.0f;
              const float f = (avg[k] - ad) * limit / 255.0f + .5f;
              // static_cast<uint>(float) on x86-64: cvttss2si (64-bit) then truncation to 32 bits,
              // reinterpreted as int by clamp<int>
              n1[k] = clampi((int)(uint32_t)(int64_t)f, 0, limit);
            }
            if (n1[0] == mbr && n1[1] == mbg && n1[2] == mbb) break;
            if (n1[0] == best.r && n1[1] == best.g && n1[2] == best.b) break;
            if (n1[0] == br && n1[1] == bg && n1[2] == bb) break;
            if (!evaluate(n1[0], n1[1], n1[2])) break;
          }
        }
      }
    }
    return best.valid;
  }
};

// pack_etc1_block_solid_color_constrained (rg_etc1.cpp:2035-2147) for an 8-pixel subblock of one
// colour: the best exact-table configuration in the given mode (diff or not), optionally within
// the differential range of subblock 0's base colour.  false: no admissible configuration.
bool solid_constrained(Solution &res, const uint8_t *color, bool use_diff, const int *base5) {
  const SolidTables &T = solid_tables();
  static const int next_comp[4] = {1, 2, 0, 1};
  uint32_t best_error = 0xFFFFFFFFu, best_i = 0;
  int best_x = 0, best_c1 = 0, best_c2 = 0;
  bool perfect = false;
  for (int i = 0; i < 3 && !perfect; i++) {
    const int c1 = color[next_comp[i]], c2 = color[next_comp[i + 1]];
    for (int delta = -1; delta <= 1 && !perfect; delta++) {
      const int cpd = clampi(color[i] + delta, 0, 255);
      for (uint16_t x : T.config[cpd]) {
        const int diff = x & 1;
        if ((int)use_diff != diff) continue;
        if (diff && base5) {
          const int d = ((x >> 8) & 255) - base5[i];
          if (d < -4 || d > 3) continue;
        }
        const uint16_t p1 = T.inverse[x & 0xFF][c1], p2 = T.inverse[x & 0xFF][c2];
        if (diff && base5) {
          const int d1 = (p1 & 0xFF) - base5[next_comp[i]], d2 = (p2 & 0xFF) - base5[next_comp[i + 1]];
          if (d1 < -4 || d1 > 3 || d2 < -4 || d2 > 3) continue;
        }
        const uint32_t err = (uint32_t)(sq(cpd - color[i]) + sq(p1 >> 8) + sq(p2 >> 8));
        if (err < best_error) {
          best_error = err;
          best_x = x;
          best_c1 = p1 & 0xFF;
          best_c2 = p2 & 0xFF;
          best_i = (uint32_t)i;
          if (!best_error) { perfect = true; break; }
        }
      }
    }
  }
  if (best_error == 0xFFFFFFFFu) return false;
  res = Solution();
  res.error = (uint64_t)(uint32_t)(best_error * 8u);
  res.inten = (best_x >> 1) & 7;
  memset(res.sel, (best_x >> 4) & 3, 8);
  int c[3];
  c[best_i] = (best_x >> 8) & 255;
  c[next_comp[best_i]] = best_c1;
  c[next_comp[best_i + 1]] = best_c2;
  res.r = c[0]; res.g = c[1]; res.b = c[2];
  res.valid = true;
  return true;
}

// pack_etc1_block (rg_etc1.cpp:2192-2451)
void pack_block(uint8_t *out, const uint8_t px[16][4], int quality) {
  uint32_t first;
  memcpy(&first, px[0], 4);
  bool solid = true;
  for (int i = 1; i < 16; i++) {
    uint32_t v;
    memcpy(&v, px[i], 4);
    solid = solid && v == first;
  }
  if (solid) { pack_solid(out, px[0]); return; }  // (same at every quality)

  uint64_t best_error = ~0ull;
  int best_flip = 0, best_c4 = 0;
  Solution best[2];
  for (int flip = 0; flip < 2; flip++)
    for (int c4 = 0; c4 < 2; c4++) {
      Solution res[2];
      uint64_t trial = 0;
      int sb;
      for (sb = 0; sb < 2; sb++) {
        uint8_t sub[8][4];
        if (flip) memcpy(sub, px[sb * 8], 32);
        else
          for (int i = 0; i < 8; i++) memcpy(sub[i], px[sb * 2 + (i >> 2) + 4 * (i & 3)], 4);
        // medium / high: a one-colour subblock also tries the exact solid-colour tables (:2259-2269)
        Solution solid_res;
        bool have_solid = false;
        if (quality >= 1 && (sb || c4)) {
          bool same = true;
          for (int i = 1; i < 8; i++) same = same && memcmp(sub[i], sub[0], 4) == 0;
          if (same) {
            const int b5[3] = {res[0].r, res[0].g, res[0].b};
            have_solid = solid_constrained(solid_res, sub[0], !c4, (sb && !c4) ? b5 : nullptr);
          }
        }
        Optimizer o;
        o.px = sub;
        o.quality = quality;
        o.color4 = c4 != 0;
        o.constrain = !c4 && sb;
        if (o.constrain) { o.base5[0] = res[0].r; o.base5[1] = res[0].g; o.base5[2] = res[0].b; }
        o.init();
        static const int d0[] = {0}, d1[] = {-1, 0, 1}, d4[] = {-4, -3, -2, -1, 0, 1, 2, 3, 4};
        static const int d23[] = {-3, -2, 2, 3}, d55[] = {-5, 5}, d58[] = {-8, -7, -6, -5, 5, 6, 7, 8};
        if (!(quality == 2 ? o.compute(d4, 9) : (quality == 1 ? o.compute(d1, 3) : o.compute(d0, 1)))) break;
        if (quality >= 1) {
          if (o.best.error > 3000) {  // refinement_error_thresh0 / 1 (:2312-2340)
            if (quality == 1) o.compute(d23, 4);
            else if (o.best.error > 6000) o.compute(d58, 8);
            else o.compute(d55, 2);
          }
          if (have_solid && solid_res.error < o.best.error) o.best = solid_res;
        }
        res[sb] = o.best;
        trial += res[sb].error;
        if (trial >= best_error) break;
      }
      if (sb < 2) continue;
      best_error = trial;
      best[0] = res[0];
      best[1] = res[1];
      best_flip = flip;
      best_c4 = c4;
    }

  if (best_c4) {
    out[0] = (uint8_t)(best[1].r | (best[0].r << 4));
    out[1] = (uint8_t)(best[1].g | (best[0].g << 4));
    out[2] = (uint8_t)(best[1].b | (best[0].b << 4));
  } else {
    int dr = best[1].r - best[0].r, dg = best[1].g - best[0].g, db = best[1].b - best[0].b;
    if (dr < 0) dr += 8;
    if (dg < 0) dg += 8;
    if (db < 0) db += 8;
    out[0] = (uint8_t)((best[0].r << 3) | dr);
    out[1] = (uint8_t)((best[0].g << 3) | dg);
    out[2] = (uint8_t)((best[0].b << 3) | db);
  }
  out[3] = (uint8_t)((best[1].inten << 2) | (best[0].inten << 5) | ((~best_c4 & 1) << 1) | best_flip);
  // selector bit planes: bit (x*4 + y) of each plane belongs to pixel (x, y)
  uint32_t lsb = 0, msb = 0;
  for (int y = 0; y < 4; y++)
    for (int x = 0; x < 4; x++) {
      int sb, k;
      if (best_flip) { sb = y >> 1; k = (y & 1) * 4 + x; }
      else { sb = x >> 1; k = (x & 1) * 4 + y; }
      const uint32_t e = kSelToEtc1[best[sb].sel[k]];
      lsb |= (e & 1) << (x * 4 + y);
      msb |= (e >> 1) << (x * 4 + y);
    }
  out[4] = (uint8_t)(msb >> 8); out[5] = (uint8_t)msb;
  out[6] = (uint8_t)(lsb >> 8); out[7] = (uint8_t)lsb;
}

}  // namespace

// ETCC::Compress_RG's block loop (ETCEncoder/src/Compressor.cpp:26-54)
static void run_etc1(const uint8_t *rgba, uint32_t width, uint32_t first_block, uint32_t num_blocks, uint8_t *out,
                     int quality) {
  const uint32_t bw = width / 4;
  for (uint32_t n = 0; n < num_blocks; n++) {
    const uint32_t bi = first_block + n, bx = bi % bw, by = bi / bw;
    uint8_t px[16][4];
    for (int j = 0; j < 4; j++) memcpy(px[4 * j], rgba + ((size_t)(by * 4 + j) * width + bx * 4) * 4, 16);
    pack_block(out + (size_t)bi * 8, px, quality);
  }
}

extern "C" void fastc_oracle_etc1(const uint8_t *rgba, uint32_t width, uint32_t height, uint32_t first_block,
                                  uint32_t num_blocks, uint8_t *out) {
  (void)height;
  run_etc1(rgba, width, first_block, num_blocks, out, 0);
}

// rg_etc1::pack_etc1_block with etc1_pack_params::m_quality = 0 cLowQuality / 1 cMediumQuality /
// 2 cHighQuality (dithering off).
extern "C" void fastc_oracle_etc1_quality(const uint8_t *rgba, uint32_t width, uint32_t height, uint32_t first_block,
                                          uint32_t num_blocks, uint8_t *out, int quality) {
  (void)height;
  run_etc1(rgba, width, first_block, num_blocks, out, quality);
}

// Derived solid-colour tables, for tests/test_tables.py: writes the config list of `color`
// (terminated by 0xFFFF) and returns its length; inverse[64*256] gets the inverse lookup.
extern "C" uint32_t fastc_oracle_etc1_tables(int color, uint16_t *config_out, uint16_t *inverse_out) {
  const SolidTables &T = solid_tables();
  uint32_t n = 0;
  for (uint16_t x : T.config[color & 255]) config_out[n++] = x;
  config_out[n] = 0xFFFF;
  if (inverse_out) memcpy(inverse_out, T.inverse, sizeof(T.inverse));
  return n;
}
