// PVRTC 4bpp encoder of pvrtc.cu as host+device code (the per-pixel / per-block arithmetic), so that
// tests/native/pvrtc_host_check.cpp can run the same functions on the CPU, in the reference's raster
// order, against the compiled reference.  The product only runs them on the device.
//
// Behavioural contract: bit-identical to PVRTCC::Compress(job, eWrapMode_Wrap)
//   reference/PVRTCEncoder/src/Compressor.cpp:861-944   Compress
//   :174-262   intensity, ComputeLocalExtrema       :264-356  DilateLabelForward, LabelImageForward
//   :358-439   DilateLabelBackward, LabelImageBackward
//   :441-583   CollectLabel, GenerateLowHighImages   :585-753  BilerpPixels, GenerateModulationValues
//   reference/PVRTCEncoder/src/Block.cpp (colour fields, Pack), reference/Base/src/Pixel.cpp
//   (ChangeBitDepth, ToBits / FromBits), reference/Base/src/Color.cpp (Pack / Unpack),
//   reference/Base/include/FasTC/Bits.h (Replicate).
// Quirks kept: the labelling scans run three rows past the image (rows 0..2 are visited twice and
// their label lists accumulate), the backward pass reads the not-yet-revisited row above, colour B's
// transparency flag tests its BLUE channel (Compressor.cpp:571), label lists are not bounded by the
// reference's assert in a release build (here: capped at 16 entries, overflow is flagged).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PVR_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define PVR_HD inline
struct uint2 { uint32_t x, y; };
#endif

namespace fastc {
namespace pvr {

PVR_HD float pf_mul(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
PVR_HD float pf_add(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}
PVR_HD float pf_div(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fdiv_rn(a, b);
#else
  return a / b;
#endif
}

constexpr int kMaxIdx = 16;
struct Label {  // Compressor.cpp:152-165
  uint8_t distance, nLabels;
  uint8_t times[kMaxIdx];
  uint32_t idxs[kMaxIdx];
};
struct PixelLabels {
  Label high, low;
};

// Label::AddIdx (:159-172).  Returns false when the list is full (the reference would write past it).
PVR_HD bool add_idx(Label &l, uint32_t idx) {
  for (uint32_t i = 0; i < l.nLabels; i++)
    if (l.idxs[i] == idx) {
      l.times[i]++;
      return true;
    }
  if (l.nLabels >= kMaxIdx) return false;
  l.times[l.nLabels] = 1;
  l.idxs[l.nLabels] = idx;
  l.nLabels++;
  return true;
}

// power-of-two wrap (Indexer.h:40-48)
PVR_HD uint32_t wrap(int32_t v, uint32_t n) { return (uint32_t)(v + (int32_t)n) & (n - 1); }

// LookupIntensity (:174-190): premultiplied luminance, float
PVR_HD float intensity_of(uint32_t pixel) {
  const float a = pf_div((float)((pixel >> 24) & 0xFF), 255.0f);
  const float r = pf_div(pf_mul(a, (float)(pixel & 0xFF)), 255.0f);
  const float g = pf_div(pf_mul(a, (float)((pixel >> 8) & 0xFF)), 255.0f);
  const float b = pf_div(pf_mul(a, (float)((pixel >> 16) & 0xFF)), 255.0f);
  return pf_add(pf_add(pf_mul(r, 0.2126f), pf_mul(g, 0.7152f)), pf_mul(b, 0.0722f));
}
// LookupIntensityByte (:206-211)
PVR_HD uint32_t intensity_byte(float intensity) { return (uint32_t)(uint8_t)pf_add(pf_mul(255.0f, intensity), 0.5f); }

// ComputeLocalExtrema's classification (:213-262) from the 3x3 neighbourhood of intensity bytes:
// 0 neither, 1 local minimum, 2 local maximum.  ibyte: [h][w] bytes.
PVR_HD int classify_extremum(const uint8_t *ibyte, uint32_t w, uint32_t h, uint32_t x, uint32_t y) {
  const uint32_t i0 = ibyte[y * w + x];
  int ng = 0, nl = 0;
  for (int j = -1; j <= 1; j++)
    for (int i = -1; i <= 1; i++) {
      if (i == 0 && j == 0) continue;
      const uint32_t ix = ibyte[wrap((int32_t)y + j, h) * w + wrap((int32_t)x + i, w)];
      ng += ix >= i0;
      nl += ix <= i0;
    }
  if (ng == nl) return 0;
  if (ng >= 8) return 1;
  if (nl >= 8) return 2;
  return 0;
}

// DilateLabelForward (:264-320)
PVR_HD bool dilate_forward(Label &l, const Label &up, const Label &left) {
  if (l.distance == 1) return true;
  if (up.distance == 0 && left.distance == 0) return true;
  if (up.distance == 0) {
    if (left.distance < 4) {
      l.distance = left.distance + 1;
      return add_idx(l, left.idxs[0]);
    }
    return true;
  }
  if (left.distance == 0) {
    if (up.distance < 4) {
      l.distance = up.distance + 1;
      return add_idx(l, up.idxs[0]);
    }
    return true;
  }
  if (left.distance == up.distance) {
    if (left.idxs[0] == up.idxs[0]) {
      l.distance = left.distance;
      return add_idx(l, left.idxs[0]);
    }
    if (up.distance < 4) {
      l.distance = up.distance + 1;
      return add_idx(l, up.idxs[0]);
    }
    return true;
  }
  if (left.distance < up.distance) {
    l.distance = left.distance + 1;
    return add_idx(l, left.idxs[0]);
  }
  l.distance = up.distance + 1;
  return add_idx(l, up.idxs[0]);
}

// One pixel of LabelImageForward (:322-356).  cls: classify_extremum of the pixel; (x, y): y may run
// to h + 2 (wrapped).
PVR_HD bool forward_pixel(PixelLabels *labels, uint32_t w, uint32_t h, uint32_t x, uint32_t yy, int cls) {
  const uint32_t y = wrap((int32_t)yy, h);
  const uint32_t idx0 = y * w + x;
  PixelLabels &l = labels[idx0];
  bool ok = true;
  if (cls == 1) {
    l.low.distance = 1;
    ok = add_idx(l.low, idx0) && ok;
  } else if (cls == 2) {
    l.high.distance = 1;
    ok = add_idx(l.high, idx0) && ok;
  }
  const PixelLabels &up = labels[wrap((int32_t)yy - 1, h) * w + x];
  const PixelLabels &left = labels[y * w + wrap((int32_t)x - 1, w)];
  if (cls != 2) ok = dilate_forward(l.high, up.high, left.high) && ok;
  if (cls != 1) ok = dilate_forward(l.low, up.low, left.low) && ok;
  return ok;
}

// The same for one label kind (the high and the low labels never mix)
PVR_HD bool forward_pixel_kind(PixelLabels *labels, uint32_t w, uint32_t h, uint32_t x, uint32_t yy, int cls, bool high) {
  const uint32_t y = wrap((int32_t)yy, h);
  const uint32_t idx0 = y * w + x;
  Label &l = high ? labels[idx0].high : labels[idx0].low;
  const bool extremum = cls == (high ? 2 : 1);
  if (extremum) {
    l.distance = 1;
    return add_idx(l, idx0);
  }
  const PixelLabels &up = labels[wrap((int32_t)yy - 1, h) * w + x];
  const PixelLabels &left = labels[y * w + wrap((int32_t)x - 1, w)];
  return dilate_forward(l, high ? up.high : up.low, high ? left.high : left.low);
}

// DilateLabelBackward (:358-404)
PVR_HD bool dilate_backward(Label &l, const Label *const nbs[5]) {
  if (l.distance == 1) return true;
  uint32_t min_dist = 5;
  for (int i = 0; i < 5; i++)
    if (nbs[i]->distance > 0 && nbs[i]->distance < min_dist) min_dist = nbs[i]->distance;
  const uint32_t new_dist = min_dist + 1;
  if ((l.distance != 0 && l.distance < new_dist) || new_dist > 4) return true;
  if (l.distance != new_dist) l.nLabels = 0;
  bool ok = true;
  for (int i = 0; i < 5; i++)
    if (nbs[i]->distance == min_dist)
      for (uint32_t k = 0; k < nbs[i]->nLabels; k++) ok = add_idx(l, nbs[i]->idxs[k]) && ok;  // Combine
  l.distance = (uint8_t)new_dist;
  return ok;
}

// One pixel of LabelImageBackward (:406-439)
PVR_HD bool backward_pixel(PixelLabels *labels, uint32_t w, uint32_t h, uint32_t x, uint32_t yy) {
  const int32_t i = (int32_t)x, j = (int32_t)yy;
  PixelLabels &l = labels[wrap(j, h) * w + x];
  const PixelLabels *nb[5] = {
      &labels[wrap(j - 1, h) * w + wrap(i + 1, w)],  // top right (not revisited yet in this pass)
      &labels[wrap(j, h) * w + wrap(i + 1, w)],      // right
      &labels[wrap(j + 1, h) * w + wrap(i + 1, w)],  // bottom right
      &labels[wrap(j + 1, h) * w + x],               // bottom
      &labels[wrap(j + 1, h) * w + wrap(i - 1, w)],  // bottom left
  };
  const Label *hi[5], *lo[5];
  for (int k = 0; k < 5; k++) { hi[k] = &nb[k]->high; lo[k] = &nb[k]->low; }
  bool ok = dilate_backward(l.high, hi);
  ok = dilate_backward(l.low, lo) && ok;
  return ok;
}

// ---- colours.  FasTC::Color is (a, r, g, b) floats, FasTC::Pixel (a, r, g, b) int16 + bit depths.
struct Color4 { float v[4]; };  // a r g b
PVR_HD Color4 color_unpack(uint32_t rgba) {  // Color::Unpack
  Color4 c;
  c.v[1] = pf_div((float)(rgba & 0xFF), 255.0f);
  c.v[2] = pf_div((float)((rgba >> 8) & 0xFF), 255.0f);
  c.v[3] = pf_div((float)((rgba >> 16) & 0xFF), 255.0f);
  c.v[0] = pf_div((float)((rgba >> 24) & 0xFF), 255.0f);
  return c;
}
// Color::Pack followed by Pixel::Unpack at 8 bits: the (a, r, g, b) bytes
PVR_HD void color_to_bytes(const Color4 &c, int (&argb)[4]) {
  for (int k = 0; k < 4; k++) argb[k] = (int)((uint32_t)pf_add(pf_mul(c.v[k], 255.0f), 0.5f) & 0xFF);
}
// CollectLabel (:441-451)
PVR_HD Color4 collect_label(const uint32_t *pixels, const Label &label) {
  Color4 ret = {{0.0f, 0.0f, 0.0f, 0.0f}};
  uint32_t n = 0;
  for (uint32_t p = 0; p < label.nLabels; p++) {
    const Color4 c = color_unpack(pixels[label.idxs[p]]);
    const float t = (float)(int)label.times[p];
    for (int k = 0; k < 4; k++) ret.v[k] = pf_add(ret.v[k], pf_mul(c.v[k], t));
    n += label.times[p];
  }
  const float fn = (float)n;
  for (int k = 0; k < 4; k++) ret.v[k] = pf_div(ret.v[k], fn);
  return ret;
}

// Pixel::ChangeBitDepth for one channel (Pixel.cpp:103-131), Replicate (Bits.h:56-76)
PVR_HD int replicate(int val, uint32_t num_bits, uint32_t to_bit) {
  if (num_bits == 0 || to_bit == 0) return 0;
  const int v = val & ((1 << num_bits) - 1);
  int res = v;
  uint32_t reslen = num_bits;
  while (reslen < to_bit) {
    uint32_t comp = 0;
    if (num_bits > to_bit - reslen) {
      const uint32_t newshift = to_bit - reslen;
      comp = num_bits - newshift;
      num_bits = newshift;
    }
    res <<= num_bits;
    res |= v >> comp;
    reslen += num_bits;
  }
  return res;
}
PVR_HD int change_depth(int val, int old_depth, int new_depth) {
  if (old_depth == new_depth) return val;
  if (old_depth == 0 && new_depth != 0) return (1 << new_depth) - 1;
  if (new_depth > old_depth) return (int)(int16_t)replicate(val, (uint32_t)old_depth, (uint32_t)new_depth);
  if (new_depth == 0) return 0xFF;
  const int wasted = old_depth - new_depth;
  uint32_t v = (uint32_t)(uint16_t)val;
  v = ((v + (1u << (wasted - 1))) >> wasted) & 0xFFFFu;
  const uint32_t hi = (1u << new_depth) - 1;
  return (int)(v < hi ? v : hi);
}

// Block::SetColor (Block.cpp:59-78) + the colour's bits as Block::Pack lays them out (ToBits,
// Pixel.cpp:71-101).  which: 0 colour A (15 bits, byte 7:6), 1 colour B (14 bits at bit offset 1, byte 5:4).
// argb: 8-bit channels.  Returns the 16-bit field (opaque flag in bit 15, mode bit 0 for B clear).
PVR_HD uint32_t color_field(const int (&argb)[4], bool transparent, int which) {
  const int tbd[2][4] = {{3, 4, 4, 4}, {3, 4, 4, 3}}, obd[2][4] = {{0, 5, 5, 5}, {0, 5, 5, 4}};
  int ch[4], depth[4];
  bool opaque = !transparent;
  if (transparent) {
    for (int k = 0; k < 4; k++) { depth[k] = tbd[which][k]; ch[k] = change_depth(argb[k], 8, depth[k]); }
    if (ch[0] == 0x7) opaque = true;  // "effectively opaque": start over as opaque
  }
  if (opaque) {
    for (int k = 0; k < 4; k++) { depth[k] = obd[which][k]; ch[k] = change_depth(k == 0 ? 255 : argb[k], 8, depth[k]); }
  }
  // ToBits: channels B, G, R, A from the low bits up (A: bit offset 0, B: bit offset 1)
  uint8_t bits[2] = {0, 0};
  int byte_idx = 0, bit_idx = which;
  for (int i = 3; i >= 0; i--) {
    const int val = ch[i], d = depth[i];
    if (d + bit_idx > 8) {
      const int next = d - (8 - bit_idx);
      const uint32_t v = (uint32_t)(uint16_t)val;
      bits[byte_idx++] |= (uint8_t)((v << bit_idx) & 0xFF);
      bit_idx = next;
      bits[byte_idx] = (uint8_t)((v >> (d - bit_idx)) & 0xFF);
    } else {
      bits[byte_idx] |= (uint8_t)(((uint32_t)val << bit_idx) & 0xFF);
      bit_idx += d;
    }
    if (bit_idx == 8) { bit_idx = 0; byte_idx++; }
  }
  uint32_t field = (uint32_t)bits[0] | ((uint32_t)bits[1] << 8);
  if (which == 1) {  // Block::Pack (Block.cpp:196-211): opaque flag from the alpha VALUE, mode bit cleared
    field = ch[0] == 0xFF ? (field | 0x8000u) : (field & 0x7FFFu);
    field &= 0xFFFEu;
  }
  return field;
}

// Interleave (Compressor.cpp:43-64): GetBlockIndex(i, j) = Interleave(j, i) = bits of j even, of i odd
PVR_HD uint32_t spread16(uint32_t x) {
  x &= 0xFFFFu;
  x = (x | (x << 8)) & 0x00FF00FFu;
  x = (x | (x << 4)) & 0x0F0F0F0Fu;
  x = (x | (x << 2)) & 0x33333333u;
  x = (x | (x << 1)) & 0x55555555u;
  return x;
}
PVR_HD uint32_t block_index(uint32_t i, uint32_t j) { return spread16(j) | (spread16(i) << 1); }

// One block of GenerateLowHighImages (:457-583): returns colour fields A << 16 | B (the upper half
// of the 64-bit block; modulation bits are filled in afterwards).
// intensity: [h][w] floats.
PVR_HD uint32_t low_high_block(const PixelLabels *labels, const float *intensity, const uint32_t *pixels, uint32_t w,
                               uint32_t h, uint32_t bi, uint32_t bj) {
  float min_i = 1.1f, max_i = -0.1f;
  uint32_t min_idx = 0, max_idx = 0;
  Color4 high = {{0, 0, 0, 0}}, low = {{0, 0, 0, 0}};
  // first the extreme intensities of the 5x5 window (needed before the holes can be filled) ...
  for (uint32_t y = bj * 4; y <= (bj + 1) * 4; y++)
    for (uint32_t x = bi * 4; x <= (bi + 1) * 4; x++) {
      const uint32_t idx = wrap((int32_t)y, h) * w + wrap((int32_t)x, w);
      const float it = intensity[idx];
      if (it < min_i) { min_i = it; min_idx = idx; }
      if (it > max_i) { max_i = it; max_idx = idx; }
    }
  // ... then the averages over the block, in raster order
  for (uint32_t y = 0; y < 4; y++)
    for (uint32_t x = 0; x < 4; x++) {
      const uint32_t idx = (bj * 4 + y) * w + bi * 4 + x;
      const Color4 ch = labels[idx].high.distance > 0 ? collect_label(pixels, labels[idx].high) : color_unpack(pixels[max_idx]);
      const Color4 cl = labels[idx].low.distance > 0 ? collect_label(pixels, labels[idx].low) : color_unpack(pixels[min_idx]);
      for (int k = 0; k < 4; k++) {
        high.v[k] = pf_add(high.v[k], pf_mul(ch.v[k], 1.0f / 16.0f));
        low.v[k] = pf_add(low.v[k], pf_mul(cl.v[k], 1.0f / 16.0f));
      }
    }
  int pa[4], pb[4];
  color_to_bytes(high, pa);
  color_to_bytes(low, pb);
  const uint32_t fa = color_field(pa, pa[0] < 200, 0);
  const uint32_t fb = color_field(pb, pb[3] < 200, 1);  // (sic: the BLUE channel decides, :571)
  return (fa << 16) | fb;
}

// Block::GetColorA / GetColorB (Block.cpp:33-57, 92-116) followed by ChangePixelTo4555
// (Compressor.cpp:624-634): (a, r, g, b) at 4 / 5 / 5 / 5 bits.
PVR_HD void decode_4555(uint32_t field, int which, int (&argb)[4]) {
  const bool opaque = (field >> 15) & 1;
  const int od[2][4] = {{0, 5, 5, 5}, {0, 5, 5, 4}}, td[2][4] = {{3, 4, 4, 4}, {3, 4, 4, 3}};
  // FromBits: most significant bits first, after the flag bit
  int pos = 15;  // next bit to read (bit 15 is the flag)
  for (int k = 0; k < 4; k++) {
    const int d = opaque ? od[which][k] : td[which][k];
    int val;
    if (d == 0) {
      val = 0xFF;
    } else {
      pos -= d;
      val = (int)((field >> pos) & ((1u << d) - 1));
    }
    const int target = k == 0 ? 4 : 5;
    argb[k] = change_depth(val, d, target);
    if (k == 0 && d > 0) argb[0] &= 0xFE;
  }
}

// BilerpPixels (:585-622) for one channel set
PVR_HD void bilerp(uint32_t x, uint32_t y, const int (&tl)[4], const int (&tr)[4], const int (&bl)[4], const int (&br)[4],
                   int (&out)[4]) {
  const int wtl = (int)((4 - x) * (4 - y)), wtr = (int)(x * (4 - y)), wbl = (int)((4 - x) * y), wbr = (int)(x * y);
  for (int c = 0; c < 4; c++) {
    const int sum = (int)(int16_t)((int16_t)(tl[c] * wtl) + (int16_t)(tr[c] * wtr) + (int16_t)(bl[c] * wbl) + (int16_t)(br[c] * wbr));
    const int fp = sum & 15;
    int t = sum / 16;
    if (c == 0) {
      t = (int)(int16_t)((t << 4) | t);
      t += (fp * 17) >> 4;
    } else {
      t = (int)(int16_t)((t << 3) | (t >> 2));
      t += ((fp >> 1) * 33) >> 5;
    }
    out[c] = (int)(int16_t)t;
  }
}

// The modulation value of one pixel (GenerateModulationValues' inner loop, :702-726)
PVR_HD uint32_t best_modulation(const int (&ca)[4], const int (&cb)[4], uint32_t original) {
  const int o[4] = {(int)(original >> 24), (int)(original & 0xFF), (int)((original >> 8) & 0xFF), (int)((original >> 16) & 0xFF)};
  const int steps[4] = {8, 5, 3, 0};
  uint32_t best = 0, best_err = 0xFFFFFFFFu;
  for (uint32_t s = 0; s < 4; s++) {
    const int lv = steps[s];
    uint32_t err = 0;
    for (int c = 0; c < 4; c++) {
      const int r = (int)(int16_t)((int16_t)(ca[c] * (8 - lv)) + (int16_t)(cb[c] * lv)) / 8;
      const int d = r - o[c];
      err += (uint32_t)(d * d);
    }
    if (err < best_err) { best_err = err; best = s; }
  }
  return best;
}

// The 32 modulation bits of output block (bx, by): each of its texels belongs to the 4x4 window of
// one of four block corners (GenerateModulationValues, :636-753, rearranged by OUTPUT block: every
// texel is written exactly once there, so the order of the windows does not matter).
// fields: [blocks_h][blocks_w] colour fields (A << 16 | B) in raster order.
PVR_HD uint32_t modulation_block(const uint32_t *fields, const uint32_t *pixels, uint32_t w, uint32_t h, uint32_t bx,
                                 uint32_t by) {
  const uint32_t bw = w >> 2, bh = h >> 2;
  uint32_t bits = 0;
  for (uint32_t py = 0; py < 4; py++)
    for (uint32_t px = 0; px < 4; px++) {
      // the window (i, j) covers pixels [4i + 2, 4i + 6) x [4j + 2, 4j + 6)
      const uint32_t i = px >= 2 ? bx : wrap((int32_t)bx - 1, bw), x = px >= 2 ? px - 2 : px + 2;
      const uint32_t j = py >= 2 ? by : wrap((int32_t)by - 1, bh), y = py >= 2 ? py - 2 : py + 2;
      const uint32_t i1 = wrap((int32_t)i + 1, bw), j1 = wrap((int32_t)j + 1, bh);
      const uint32_t ftl = fields[j * bw + i], ftr = fields[j * bw + i1], fbl = fields[j1 * bw + i], fbr = fields[j1 * bw + i1];
      int tl[4], tr[4], bl[4], br[4], ca[4], cb[4];
      decode_4555(ftl >> 16, 0, tl); decode_4555(ftr >> 16, 0, tr); decode_4555(fbl >> 16, 0, bl); decode_4555(fbr >> 16, 0, br);
      bilerp(x, y, tl, tr, bl, br, ca);
      decode_4555(ftl & 0xFFFF, 1, tl); decode_4555(ftr & 0xFFFF, 1, tr); decode_4555(fbl & 0xFFFF, 1, bl); decode_4555(fbr & 0xFFFF, 1, br);
      bilerp(x, y, tl, tr, bl, br, cb);
      const uint32_t m = best_modulation(ca, cb, pixels[(by * 4 + py) * w + bx * 4 + px]);
      bits |= m << (2 * (py * 4 + px));
    }
  return bits;
}


// ---- decoder: PVRTCC::Decompress (Decompressor.cpp:245-352) for 4bpp, one pixel.
// blocks: the compressed texture (Morton order).  BilinearUpscale(2, 2) + ExpandTo8888
// (PVRTCImage.cpp:94-160, :437-462) are the encoder's BilerpPixels with x = (i + 2) % 4 between the
// blocks (i + 2) / 4 - 1 and (i + 2) / 4; then Decompress4BPP (:52-118) incl. the punch-through mode.
PVR_HD uint32_t decode_pixel(const uint2 *blocks, uint32_t w, uint32_t h, uint32_t i, uint32_t j) {
  const uint32_t bw = w >> 2, bh = h >> 2;
  const uint32_t hx = wrap((int32_t)((i + 2) >> 2), bw), lx = wrap((int32_t)((i + 2) >> 2) - 1, bw);
  const uint32_t hy = wrap((int32_t)((j + 2) >> 2), bh), ly = wrap((int32_t)((j + 2) >> 2) - 1, bh);
  const uint32_t x = (i + 2) & 3, y = (j + 2) & 3;
  const uint32_t ftl = blocks[block_index(lx, ly)].y, ftr = blocks[block_index(hx, ly)].y;
  const uint32_t fbl = blocks[block_index(lx, hy)].y, fbr = blocks[block_index(hx, hy)].y;
  int tl[4], tr[4], bl[4], br[4], ca[4], cb[4];
  decode_4555(ftl >> 16, 0, tl); decode_4555(ftr >> 16, 0, tr); decode_4555(fbl >> 16, 0, bl); decode_4555(fbr >> 16, 0, br);
  bilerp(x, y, tl, tr, bl, br, ca);
  decode_4555(ftl & 0xFFFF, 1, tl); decode_4555(ftr & 0xFFFF, 1, tr); decode_4555(fbl & 0xFFFF, 1, bl); decode_4555(fbr & 0xFFFF, 1, br);
  bilerp(x, y, tl, tr, bl, br, cb);
  const uint2 b = blocks[block_index(i >> 2, j >> 2)];
  uint32_t mod = (b.x >> (2 * ((j & 3) * 4 + (i & 3)))) & 3u;
  bool punch = false;
  int lerp;
  if (b.y & 1u) {  // mode bit: {8, 4, punch-through 4, 0}
    if (mod >= 2) { punch = mod == 2; mod -= 1; }
    lerp = mod == 0 ? 8 : (mod == 1 ? 4 : 0);
  } else {
    lerp = mod == 0 ? 8 : (mod == 1 ? 5 : (mod == 2 ? 3 : 0));
  }
  int res[4];
  for (int c = 0; c < 4; c++) res[c] = (int)(int16_t)((int16_t)(ca[c] * (8 - lerp)) + (int16_t)(cb[c] * lerp)) / 8;
  if (punch) res[0] = 0;
  // Pixel::Pack: a << 24 | b << 16 | g << 8 | r (the shifts carry any bits above 8 upwards)
  uint32_t r = (uint32_t)(uint16_t)res[0];
  r = (r << 8) | (uint32_t)(uint16_t)res[3];
  r = (r << 8) | (uint32_t)(uint16_t)res[2];
  r = (r << 8) | (uint32_t)(uint16_t)res[1];
  return r;
}
}  // namespace pvr
}  // namespace fastc
