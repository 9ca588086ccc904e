// TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the reference's
// DXT1/DXT5 block encoder.  Never linked into or called from the product path.
//
// Restates stb_dxt v1.06 as FasTC drives it (mode = STB_DXT_DITHER, one refine
// pass): /root/reference/DXTEncoder/src/stb_dxt.h and
// /root/reference/DXTEncoder/src/Compressor.cpp:47-95.  Pinned bit-for-bit
// against the compiled reference (oracle/_ref) by tests/test_oracle_vs_ref.py and
// against tests/golden/*.npz.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "oracle.h"

namespace {

struct DxtTables {
  uint8_t expand5[32], expand6[64];
  uint8_t omatch5[256][2], omatch6[256][2];
  uint8_t quant_rb[256 + 16], quant_g[256 + 16];
};

// stb_dxt.h:74-78
int mul8bit(int a, int b) {
  int t = a * b + 128;
  return (t + (t >> 8)) >> 8;
}

// stb_dxt.h:98-108 (no rounding bias)
int lerp13(int a, int b) { return (2 * a + b) / 3; }

// stb_dxt.h:121-147
void prepare_opt_table(uint8_t (*table)[2], const uint8_t *expand, int size) {
  for (int i = 0; i < 256; i++) {
    int best = 256;
    for (int mn = 0; mn < size; mn++)
      for (int mx = 0; mx < size; mx++) {
        int mine = expand[mn], maxe = expand[mx];
        int err = std::abs(lerp13(maxe, mine) - i);
        err += std::abs(maxe - mine) * 3 / 100;
        if (err < best) {
          table[i][0] = (uint8_t)mx;
          table[i][1] = (uint8_t)mn;
          best = err;
        }
      }
  }
}

// stb_dxt.h:603-621
const DxtTables &tables() {
  static DxtTables t;
  static bool init = false;
  if (!init) {
    for (int i = 0; i < 32; i++) t.expand5[i] = (uint8_t)((i << 3) | (i >> 2));
    for (int i = 0; i < 64; i++) t.expand6[i] = (uint8_t)((i << 2) | (i >> 4));
    for (int i = 0; i < 256 + 16; i++) {
      int v = i - 8 < 0 ? 0 : i - 8 > 255 ? 255 : i - 8;
      t.quant_rb[i] = t.expand5[mul8bit(v, 31)];
      t.quant_g[i] = t.expand6[mul8bit(v, 63)];
    }
    prepare_opt_table(t.omatch5, t.expand5, 32);
    prepare_opt_table(t.omatch6, t.expand6, 64);
    init = true;
  }
  return t;
}

// stb_dxt.h:92-95
uint16_t as16bit(int r, int g, int b) {
  return (uint16_t)((mul8bit(r, 31) << 11) + (mul8bit(g, 63) << 5) + mul8bit(b, 31));
}

// stb_dxt.h:80-90,149-155: the 4 palette colours (RGB0 each)
void eval_colors(uint8_t color[16], uint16_t c0, uint16_t c1) {
  const DxtTables &t = tables();
  auto from16 = [&](uint8_t *o, uint16_t v) {
    o[0] = t.expand5[(v & 0xf800) >> 11];
    o[1] = t.expand6[(v & 0x07e0) >> 5];
    o[2] = t.expand5[v & 0x001f];
    o[3] = 0;
  };
  from16(color + 0, c0);
  from16(color + 4, c1);
  for (int k = 0; k < 3; k++) {
    color[8 + k] = (uint8_t)lerp13(color[0 + k], color[4 + k]);
    color[12 + k] = (uint8_t)lerp13(color[4 + k], color[0 + k]);
  }
}

// stb_dxt.h:159-183: Floyd-Steinberg dither of each channel to the 565 grid.
void dither_block(uint8_t dest[64], const uint8_t block[64]) {
  const DxtTables &t = tables();
  for (int ch = 0; ch < 3; ch++) {
    const uint8_t *quant = (ch == 1 ? t.quant_g : t.quant_rb) + 8;
    int err[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    for (int y = 0; y < 4; y++) {
      int *e1 = err[y & 1], *e2 = err[(y & 1) ^ 1];  // e1 = this row, e2 = previous
      const uint8_t *bp = block + 16 * y + ch;
      uint8_t *dp = dest + 16 * y + ch;
      dp[0] = quant[bp[0] + ((3 * e2[1] + 5 * e2[0]) >> 4)];
      e1[0] = bp[0] - dp[0];
      dp[4] = quant[bp[4] + ((7 * e1[0] + 3 * e2[2] + 5 * e2[1] + e2[0]) >> 4)];
      e1[1] = bp[4] - dp[4];
      dp[8] = quant[bp[8] + ((7 * e1[1] + 3 * e2[3] + 5 * e2[2] + e2[1]) >> 4)];
      e1[2] = bp[8] - dp[8];
      dp[12] = quant[bp[12] + ((7 * e1[2] + 5 * e2[3] + e2[2]) >> 4)];
      e1[3] = bp[12] - dp[12];
    }
  }
}

// stb_dxt.h:186-280, dither branch only (FasTC always passes STB_DXT_DITHER).
uint32_t match_colors_dither(const uint8_t block[64], const uint8_t color[16]) {
  int dirr = color[0] - color[4], dirg = color[1] - color[5], dirb = color[2] - color[6];
  int dots[16], stops[4];
  for (int i = 0; i < 16; i++)
    dots[i] = block[i * 4] * dirr + block[i * 4 + 1] * dirg + block[i * 4 + 2] * dirb;
  for (int i = 0; i < 4; i++)
    stops[i] = color[i * 4] * dirr + color[i * 4 + 1] * dirg + color[i * 4 + 2] * dirb;

  int c0 = ((stops[1] + stops[3]) >> 1) << 4;
  int half = ((stops[3] + stops[2]) >> 1) << 4;
  int c3 = ((stops[2] + stops[0]) >> 1) << 4;
  auto pick = [&](int dot) {
    return dot < half ? (dot < c0 ? 1 : 3) : (dot < c3 ? 2 : 0);
  };

  uint32_t mask = 0;
  int err[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
  for (int y = 0; y < 4; y++) {
    int *e1 = err[y & 1], *e2 = err[(y & 1) ^ 1];
    const int *dp = dots + 4 * y;
    int step, lmask;
    step = pick((dp[0] << 4) + (3 * e2[1] + 5 * e2[0]));
    e1[0] = dp[0] - stops[step];
    lmask = step;
    step = pick((dp[1] << 4) + (7 * e1[0] + 3 * e2[2] + 5 * e2[1] + e2[0]));
    e1[1] = dp[1] - stops[step];
    lmask |= step << 2;
    step = pick((dp[2] << 4) + (7 * e1[1] + 3 * e2[3] + 5 * e2[2] + e2[1]));
    e1[2] = dp[2] - stops[step];
    lmask |= step << 4;
    step = pick((dp[3] << 4) + (7 * e1[2] + 5 * e2[3] + e2[2]));
    e1[3] = dp[3] - stops[step];
    lmask |= step << 6;
    mask |= (uint32_t)lmask << (y * 8);
  }
  return mask;
}

// stb_dxt.h:283-385
void optimize_colors(const uint8_t block[64], uint16_t *pmax16, uint16_t *pmin16) {
  int mu[3], mn[3], mx[3];
  for (int ch = 0; ch < 3; ch++) {
    int muv, minv, maxv;
    muv = minv = maxv = block[ch];
    for (int i = 4; i < 64; i += 4) {
      int v = block[i + ch];
      muv += v;
      if (v < minv) minv = v;
      else if (v > maxv) maxv = v;
    }
    mu[ch] = (muv + 8) >> 4;
    mn[ch] = minv;
    mx[ch] = maxv;
  }
  int cov[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 16; i++) {
    int r = block[i * 4] - mu[0], g = block[i * 4 + 1] - mu[1], b = block[i * 4 + 2] - mu[2];
    cov[0] += r * r; cov[1] += r * g; cov[2] += r * b;
    cov[3] += g * g; cov[4] += g * b; cov[5] += b * b;
  }
  float covf[6];
  for (int i = 0; i < 6; i++) covf[i] = cov[i] / 255.0f;
  float vfr = (float)(mx[0] - mn[0]), vfg = (float)(mx[1] - mn[1]), vfb = (float)(mx[2] - mn[2]);
  for (int iter = 0; iter < 4; iter++) {
    float r = vfr * covf[0] + vfg * covf[1] + vfb * covf[2];
    float g = vfr * covf[1] + vfg * covf[3] + vfb * covf[4];
    float b = vfr * covf[2] + vfg * covf[4] + vfb * covf[5];
    vfr = r; vfg = g; vfb = b;
  }
  double magn = std::fabs((double)vfr);
  if (std::fabs((double)vfg) > magn) magn = std::fabs((double)vfg);
  if (std::fabs((double)vfb) > magn) magn = std::fabs((double)vfb);
  int v_r, v_g, v_b;
  if (magn < 4.0f) {
    v_r = 299; v_g = 587; v_b = 114;
  } else {
    magn = 512.0 / magn;
    v_r = (int)(vfr * magn); v_g = (int)(vfg * magn); v_b = (int)(vfb * magn);
  }
  int mind = 0x7fffffff, maxd = -0x7fffffff;
  const uint8_t *minp = block, *maxp = block;
  for (int i = 0; i < 16; i++) {
    int dot = block[i * 4] * v_r + block[i * 4 + 1] * v_g + block[i * 4 + 2] * v_b;
    if (dot < mind) { mind = dot; minp = block + i * 4; }
    if (dot > maxd) { maxd = dot; maxp = block + i * 4; }
  }
  *pmax16 = as16bit(maxp[0], maxp[1], maxp[2]);
  *pmin16 = as16bit(minp[0], minp[1], minp[2]);
}

int sclamp(float y, int p0, int p1) {
  int x = (int)y;
  return x < p0 ? p0 : x > p1 ? p1 : x;
}

// stb_dxt.h:398-474
bool refine_block(const uint8_t block[64], uint16_t *pmax16, uint16_t *pmin16, uint32_t mask) {
  static const int w1tab[4] = {3, 0, 2, 1};
  static const int prods[4] = {0x090000, 0x000900, 0x040102, 0x010402};
  const DxtTables &t = tables();
  uint16_t old_min = *pmin16, old_max = *pmax16, min16, max16;
  if ((mask ^ (mask << 2)) < 4) {
    int r = 8, g = 8, b = 8;
    for (int i = 0; i < 16; i++) { r += block[i * 4]; g += block[i * 4 + 1]; b += block[i * 4 + 2]; }
    r >>= 4; g >>= 4; b >>= 4;
    max16 = (uint16_t)((t.omatch5[r][0] << 11) | (t.omatch6[g][0] << 5) | t.omatch5[b][0]);
    min16 = (uint16_t)((t.omatch5[r][1] << 11) | (t.omatch6[g][1] << 5) | t.omatch5[b][1]);
  } else {
    int akku = 0, a1r = 0, a1g = 0, a1b = 0, a2r = 0, a2g = 0, a2b = 0;
    uint32_t cm = mask;
    for (int i = 0; i < 16; i++, cm >>= 2) {
      int step = cm & 3, w1 = w1tab[step];
      int r = block[i * 4], g = block[i * 4 + 1], b = block[i * 4 + 2];
      akku += prods[step];
      a1r += w1 * r; a1g += w1 * g; a1b += w1 * b;
      a2r += r; a2g += g; a2b += b;
    }
    a2r = 3 * a2r - a1r; a2g = 3 * a2g - a1g; a2b = 3 * a2b - a1b;
    int xx = akku >> 16, yy = (akku >> 8) & 0xff, xy = akku & 0xff;
    float frb = 3.0f * 31.0f / 255.0f / (xx * yy - xy * xy);
    float fg = frb * 63.0f / 31.0f;
    max16 = (uint16_t)(sclamp((a1r * yy - a2r * xy) * frb + 0.5f, 0, 31) << 11);
    max16 |= (uint16_t)(sclamp((a1g * yy - a2g * xy) * fg + 0.5f, 0, 63) << 5);
    max16 |= (uint16_t)(sclamp((a1b * yy - a2b * xy) * frb + 0.5f, 0, 31));
    min16 = (uint16_t)(sclamp((a2r * xx - a1r * xy) * frb + 0.5f, 0, 31) << 11);
    min16 |= (uint16_t)(sclamp((a2g * xx - a1g * xy) * fg + 0.5f, 0, 63) << 5);
    min16 |= (uint16_t)(sclamp((a2b * xx - a1b * xy) * frb + 0.5f, 0, 31));
  }
  *pmin16 = min16;
  *pmax16 = max16;
  return old_min != min16 || old_max != max16;
}

// stb_dxt.h:477-548 with mode = STB_DXT_DITHER (refinecount 1)
void compress_color_block(uint8_t dest[8], const uint8_t block[64]) {
  const DxtTables &t = tables();
  uint32_t mask;
  uint16_t max16, min16;
  uint32_t px[16];
  memcpy(px, block, 64);
  int i;
  for (i = 1; i < 16; i++)
    if (px[i] != px[0]) break;
  if (i == 16) {
    int r = block[0], g = block[1], b = block[2];
    mask = 0xaaaaaaaau;
    max16 = (uint16_t)((t.omatch5[r][0] << 11) | (t.omatch6[g][0] << 5) | t.omatch5[b][0]);
    min16 = (uint16_t)((t.omatch5[r][1] << 11) | (t.omatch6[g][1] << 5) | t.omatch5[b][1]);
  } else {
    uint8_t dblock[64], color[16];
    dither_block(dblock, block);
    optimize_colors(dblock, &max16, &min16);
    if (max16 != min16) {
      eval_colors(color, max16, min16);
      mask = match_colors_dither(block, color);
    } else {
      mask = 0;
    }
    if (refine_block(dblock, &max16, &min16, mask)) {
      if (max16 != min16) {
        eval_colors(color, max16, min16);
        mask = match_colors_dither(block, color);
      } else {
        mask = 0;
      }
    }
  }
  if (max16 < min16) {
    uint16_t tmp = min16; min16 = max16; max16 = tmp;
    mask ^= 0x55555555u;
  }
  dest[0] = (uint8_t)max16; dest[1] = (uint8_t)(max16 >> 8);
  dest[2] = (uint8_t)min16; dest[3] = (uint8_t)(min16 >> 8);
  dest[4] = (uint8_t)mask; dest[5] = (uint8_t)(mask >> 8);
  dest[6] = (uint8_t)(mask >> 16); dest[7] = (uint8_t)(mask >> 24);
}

// stb_dxt.h:551-601 (src = alpha bytes, stride 4)
void compress_alpha_block(uint8_t dest[8], const uint8_t *src) {
  int mn, mx;
  mn = mx = src[0];
  for (int i = 1; i < 16; i++) {
    if (src[i * 4] < mn) mn = src[i * 4];
    else if (src[i * 4] > mx) mx = src[i * 4];
  }
  dest[0] = (uint8_t)mx;
  dest[1] = (uint8_t)mn;
  dest += 2;
  int dist = mx - mn, dist4 = dist * 4, dist2 = dist * 2;
  int bias = (dist < 8) ? (dist - 1) : (dist / 2 + 2);
  bias -= mn * 7;
  int bits = 0, mask = 0;
  for (int i = 0; i < 16; i++) {
    int a = src[i * 4] * 7 + bias;
    int ind, t;
    t = (a >= dist4) ? -1 : 0; ind = t & 4; a -= dist4 & t;
    t = (a >= dist2) ? -1 : 0; ind += t & 2; a -= dist2 & t;
    ind += (a >= dist);
    ind = -ind & 7;
    ind ^= (2 > ind);
    mask |= ind << bits;
    if ((bits += 3) >= 8) {
      *dest++ = (uint8_t)mask;
      mask >>= 8;
      bits -= 8;
    }
  }
}

}  // namespace

// DXTEncoder/src/Compressor.cpp:47-95: raster block range, 4 rows x 16 B per block.
extern "C" void fastc_oracle_dxt(int dxt5, const uint8_t *rgba, uint32_t width, uint32_t height,
                                 uint32_t first_block, uint32_t num_blocks, uint8_t *out) {
  (void)height;
  const uint32_t bw = width / 4;
  const uint32_t bsz = dxt5 ? 16 : 8;
  for (uint32_t n = 0; n < num_blocks; n++) {
    uint32_t bi = first_block + n;
    uint32_t bx = bi % bw, by = bi / bw;
    uint8_t block[64];
    for (int j = 0; j < 4; j++)
      memcpy(block + 16 * j, rgba + ((size_t)(by * 4 + j) * width + bx * 4) * 4, 16);
    uint8_t *dst = out + (size_t)bi * bsz;
    if (dxt5) {
      compress_alpha_block(dst, block + 3);
      dst += 8;
    }
    compress_color_block(dst, block);
  }
}
