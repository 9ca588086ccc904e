// TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's DECODERS and of
// its PSNR definition, used to score parity.  Never part of the product path.
//
//   DXT1/DXT5  reference/DXTEncoder/src/Decompressor.cpp:28-156
//   ETC1       reference/ETCEncoder/src/rg_etc1.cpp:945-1257 (unpack_etc1_block),
//              reference/ETCEncoder/src/Decompressor.cpp:27-47
//   BC7        reference/BPTCEncoder/src/Decompressor.cpp:32-367
//   PSNR       reference/Base/src/Image.cpp:205-255
//
// Pinned against the compiled reference (oracle/_ref) by
// tests/test_oracle_vs_ref.py on encoder output and on random bit patterns.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>

#include "bc7_tables.h"
#include "oracle.h"

namespace {

// ---------------------------------------------------------------- DXT
inline int rep5(int v) { return (v << 3) | (v >> 2); }
inline int rep6(int v) { return (v << 2) | (v >> 4); }

// Decompressor.cpp:28-67.  check_order=false for the colour half of DXT5.
void dxt_color_block(const uint8_t *b, uint32_t out[16], bool check_order) {
  const uint32_t c0 = b[0] | (b[1] << 8), c1 = b[2] | (b[3] << 8);
  int col[4][3];
  col[0][0] = rep5((c0 >> 11) & 31); col[0][1] = rep6((c0 >> 5) & 63); col[0][2] = rep5(c0 & 31);
  col[1][0] = rep5((c1 >> 11) & 31); col[1][1] = rep6((c1 >> 5) & 63); col[1][2] = rep5(c1 & 31);
  for (int k = 0; k < 3; k++) {
    if (!check_order || c0 > c1) {
      col[2][k] = (col[0][k] * 2 + col[1][k]) / 3;
      col[3][k] = (col[0][k] + col[1][k] * 2) / 3;
    } else {
      col[2][k] = (col[0][k] + col[1][k]) / 2;
      col[3][k] = 0;  // "d already initialized to zero": opaque black, not transparent
    }
  }
  const uint32_t mod = b[4] | (b[5] << 8) | (b[6] << 16) | ((uint32_t)b[7] << 24);
  for (int i = 0; i < 16; i++) {
    const int *c = col[(mod >> (2 * i)) & 3];
    out[i] = (out[i] & 0xFF000000u) | (uint32_t)c[0] | ((uint32_t)c[1] << 8) | ((uint32_t)c[2] << 16);
  }
}

// Decompressor.cpp:69-96
void dxt5_alpha_block(const uint8_t *b, uint32_t out[16]) {
  int a0 = b[0], a1 = b[1], pal[8];
  pal[0] = a0; pal[1] = a1;
  if (a0 > a1) {
    for (int i = 2; i < 8; i++) pal[i] = ((8 - i) * a0 + (i - 1) * a1) / 7;
  } else {
    for (int i = 2; i < 6; i++) pal[i] = ((6 - i) * a0 + (i - 1) * a1) / 5;
    pal[6] = 0; pal[7] = 255;
  }
  uint64_t mod = 0;
  for (int i = 0; i < 6; i++) mod |= (uint64_t)b[2 + i] << (8 * i);
  for (int i = 0; i < 16; i++)
    out[i] = (out[i] & 0x00FFFFFFu) | ((uint32_t)pal[(mod >> (3 * i)) & 7] << 24);
}

// ---------------------------------------------------------------- ETC1
const int kInten[8][4] = {{-8, -2, 2, 8},     {-17, -5, 5, 17},   {-29, -9, 9, 29},    {-42, -13, 13, 42},
                          {-60, -18, 18, 60}, {-80, -24, 24, 80}, {-106, -33, 33, 106}, {-183, -47, 47, 183}};
const int kEtc1ToSelector[4] = {2, 3, 1, 0};

inline int clamp255(int v) { return v < 0 ? 0 : v > 255 ? 255 : v; }

void etc1_block(const uint8_t *b, uint32_t out[16]) {
  const bool diff = (b[3] & 2) != 0, flip = (b[3] & 1) != 0;
  const int t0 = (b[3] >> 5) & 7, t1 = (b[3] >> 2) & 7;
  int base[2][3];
  if (diff) {
    for (int k = 0; k < 3; k++) {
      int c5 = b[k] >> 3, d3 = b[k] & 7;
      if (d3 >= 4) d3 -= 8;
      int c2 = c5 + d3;
      c2 = c2 < 0 ? 0 : c2 > 31 ? 31 : c2;  // rg_etc1.cpp:972-981 (invalid blocks are clamped)
      base[0][k] = (c5 << 3) | (c5 >> 2);
      base[1][k] = (c2 << 3) | (c2 >> 2);
    }
  } else {
    for (int k = 0; k < 3; k++) {
      int c0 = b[k] >> 4, c1 = b[k] & 15;
      base[0][k] = (c0 << 4) | c0;
      base[1][k] = (c1 << 4) | c1;
    }
  }
  for (int y = 0; y < 4; y++)
    for (int x = 0; x < 4; x++) {
      const int sub = flip ? (y >= 2) : (x >= 2);
      const int bit = x * 4 + y;
      const int lsb = (b[7 - (bit >> 3)] >> (bit & 7)) & 1;
      const int msb = (b[5 - (bit >> 3)] >> (bit & 7)) & 1;
      const int sel = kEtc1ToSelector[lsb | (msb << 1)];
      const int m = kInten[sub ? t1 : t0][sel];
      const int r = clamp255(base[sub][0] + m), g = clamp255(base[sub][1] + m), bl = clamp255(base[sub][2] + m);
      out[y * 4 + x] = (uint32_t)r | ((uint32_t)g << 8) | ((uint32_t)bl << 16) | 0xFF000000u;
    }
}

// ---------------------------------------------------------------- BC7
struct BitReader {
  const uint8_t *p;
  int pos;
  uint32_t bit() {
    uint32_t v = (p[pos >> 3] >> (pos & 7)) & 1;
    pos++;
    return v;
  }
  uint32_t bits(int n) {  // LSB first (Base/include/FasTC/BitStream.h)
    uint32_t v = 0;
    for (int i = 0; i < n; i++) v |= bit() << i;
    return v;
  }
};

// Decompressor.cpp:32-190 + 255-319.  `quirks`: reproduce the reference
// decoder's handling of mode 4 with idxMode==1 -- it picks the interpolation
// table from the mode's static attributes (2-bit colour / 3-bit alpha) even
// though the index arrays were swapped (Decompressor.cpp:264-292), so such
// blocks decode differently from the BC7 specification.  quirks=false decodes
// per spec.
void bc7_block(const uint8_t *blk, uint32_t out[16], bool quirks) {
  using namespace bc7t;
  BitReader s{blk, 0};
  int mode = 0;
  while (mode < 8 && !s.bit()) mode++;
  if (mode >= 8) {
    for (int i = 0; i < 16; i++) out[i] = 0;
    return;
  }
  const ModeAttr &A = kModes[mode];
  int shape = 0, rot = 0, idx_mode = 0;
  if (A.subsets > 1) shape = s.bits(mode == 0 ? 4 : 6);
  else if (A.has_rotation) {
    rot = s.bits(2);
    if (A.has_idx_mode) idx_mode = s.bit();
  }
  int cp = A.color_bits, ap = A.alpha_bits;
  uint32_t eps[3][2][4];
  for (int ch = 0; ch < 3; ch++)
    for (int i = 0; i < A.subsets; i++)
      for (int e = 0; e < 2; e++) eps[i][e][ch] = s.bits(cp) << (8 - cp);
  for (int i = 0; i < A.subsets; i++)
    for (int e = 0; e < 2; e++) eps[i][e][3] = ap == 0 ? 0xFF : (s.bits(ap) << (8 - ap));
  if (A.pbit_type != kPbitNone) {
    cp += 1; ap += 1;
    for (int i = 0; i < A.subsets; i++) {
      uint32_t pb[2];
      pb[0] = s.bit();
      pb[1] = (A.pbit_type == kPbitShared) ? pb[0] : s.bit();  // shared: one bit for both endpoints
      for (int e = 0; e < 2; e++)
        for (int ch = 0; ch < 4; ch++) eps[i][e][ch] |= pb[e] << (8 - (ch == 3 ? ap : cp));
    }
  }
  for (int i = 0; i < A.subsets; i++)
    for (int e = 0; e < 2; e++)
      for (int ch = 0; ch < 4; ch++) {
        eps[i][e][ch] &= 0xFF;  // reference stores into uint8
        eps[i][e][ch] |= eps[i][e][ch] >> (ch == 3 ? ap : cp);
        eps[i][e][ch] &= 0xFF;
      }
  uint32_t cidx[16], aidx[16];
  for (int i = 0; i < 16; i++) {
    int sub = subset_of(i, shape, A.subsets);
    cidx[i] = s.bits(anchor_of(sub, shape, A.subsets) == i ? A.index_bits - 1 : A.index_bits);
  }
  int color_bits = A.index_bits, alpha_bits = A.alpha_index_bits;
  if (A.alpha_index_bits == 0) {
    memcpy(aidx, cidx, sizeof(aidx));
  } else {
    for (int i = 0; i < 16; i++) {
      int sub = subset_of(i, shape, A.subsets);
      aidx[i] = s.bits(anchor_of(sub, shape, A.subsets) == i ? A.alpha_index_bits - 1 : A.alpha_index_bits);
    }
    if (idx_mode) {
      for (int i = 0; i < 16; i++) std::swap(aidx[i], cidx[i]);
      if (!quirks) std::swap(color_bits, alpha_bits);
    }
  }
  for (int i = 0; i < 16; i++) {
    const int sub = subset_of(i, shape, A.subsets);
    uint8_t px[4];
    for (int ch = 0; ch < 4; ch++) {
      uint32_t w0, w1;
      if (ch == 3 && A.alpha_index_bits > 0) {
        w0 = kInterp[alpha_bits - 1][aidx[i] & 15][0];
        w1 = kInterp[alpha_bits - 1][aidx[i] & 15][1];
      } else {
        w0 = kInterp[color_bits - 1][cidx[i] & 15][0];
        w1 = kInterp[color_bits - 1][cidx[i] & 15][1];
      }
      px[ch] = (uint8_t)(((eps[sub][0][ch] * w0 + eps[sub][1][ch] * w1 + 32) >> 6) & 0xFF);
    }
    if (rot == 1) std::swap(px[0], px[3]);
    else if (rot == 2) std::swap(px[1], px[3]);
    else if (rot == 3) std::swap(px[2], px[3]);
    out[i] = px[0] | (px[1] << 8) | (px[2] << 16) | ((uint32_t)px[3] << 24);
  }
}

}  // namespace

extern "C" void fastc_oracle_decode2(int format, const uint8_t *cmp, uint32_t width, uint32_t height,
                                     uint8_t *rgba_out, int spec_correct_bc7) {
  const uint32_t bw = width / 4, bh = height / 4;
  uint32_t *outp = reinterpret_cast<uint32_t *>(rgba_out);
  uint32_t px[16];
  memset(px, 0xFF, sizeof(px));  // DXT decoders only overwrite part of each pixel
  for (uint32_t j = 0; j < bh; j++)
    for (uint32_t i = 0; i < bw; i++) {
      const uint32_t bi = j * bw + i;
      switch (format) {
        case FASTC_ORACLE_DXT1: dxt_color_block(cmp + (size_t)bi * 8, px, true); break;
        case FASTC_ORACLE_DXT5:
          dxt5_alpha_block(cmp + (size_t)bi * 16, px);
          dxt_color_block(cmp + (size_t)bi * 16 + 8, px, false);
          break;
        case FASTC_ORACLE_ETC1: etc1_block(cmp + (size_t)bi * 8, px); break;
        default: bc7_block(cmp + (size_t)bi * 16, px, !spec_correct_bc7); break;
      }
      for (int y = 0; y < 4; y++)
        memcpy(outp + (size_t)(j * 4 + y) * width + i * 4, px + 4 * y, 16);
    }
}

extern "C" void fastc_oracle_decode(int format, const uint8_t *cmp, uint32_t width, uint32_t height,
                                    uint8_t *rgba_out) {
  fastc_oracle_decode2(format, cmp, width, height, rgba_out, 0);
}

// Image.cpp:205-255: alpha-premultiplied RGB error, mse divided by W*H only,
// peak = 3*255^2.
extern "C" double fastc_oracle_psnr(const uint8_t *a, const uint8_t *b, uint32_t width, uint32_t height) {
  double mse = 0.0;
  const uint32_t n = width * height;
  for (uint32_t i = 0; i < n; i++) {
    const double ra = (double)a[4 * i + 3] / 255.0, ua = (double)b[4 * i + 3] / 255.0;
    for (int c = 0; c < 3; c++) {
      const double diff = ra * ((double)a[4 * i + c] * 1.0) - ua * ((double)b[4 * i + c] * 1.0);
      mse += diff * diff;
    }
  }
  mse /= (double)(width * height);
  const double maxi = (1.0 * 1.0 + 1.0 * 1.0 + 1.0 * 1.0) * (255.0 * 255.0);
  return 10 * log10(maxi / mse);
}
