// TEST INFRASTRUCTURE: runs the product's PVRTC arithmetic (fastc_b200/csrc/pvrtc_block.cuh, the code
// the kernels of pvrtc.cu execute) on the CPU in the reference's raster order and compares the blocks
// with the compiled reference (oracle/_ref/libfastc_ref.so: PVRTCC::Compress).  Built and run by
// tests/test_native_host.py:
//   g++ -O2 -ffp-contract=off -I fastc_b200/csrc pvrtc_host_check.cpp oracle/_ref/libfastc_ref.so
// Usage: pvrtc_host_check <width> <height> < rgba-bytes     -> prints "blocks N mismatches M"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "pvrtc_block.cuh"

extern "C" int fastc_ref_decompress(int format, const uint8_t *cmp, uint32_t width, uint32_t height, uint8_t *rgba_out);
extern "C" int fastc_ref_compress(int format, const uint8_t *rgba, uint32_t width, uint32_t height, uint8_t *out,
                                  uint32_t out_size, int quality, int threads, int job_size, double *ms);

using namespace fastc::pvr;

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  const uint32_t w = (uint32_t)atoi(argv[1]), h = (uint32_t)atoi(argv[2]);
  std::vector<uint8_t> img((size_t)w * h * 4);
  if (fread(img.data(), 1, img.size(), stdin) != img.size()) return 3;
  const uint32_t *pixels = reinterpret_cast<const uint32_t *>(img.data());
  const uint32_t bw = w / 4, bh = h / 4, nb = bw * bh;
  std::vector<uint64_t> want(nb), got(nb);
  if (fastc_ref_compress(4, img.data(), w, h, reinterpret_cast<uint8_t *>(want.data()), nb * 8, 0, 1, 0, nullptr)) return 4;

  std::vector<float> intensity((size_t)w * h);
  std::vector<uint8_t> ibyte((size_t)w * h), cls((size_t)w * h);
  for (size_t i = 0; i < (size_t)w * h; i++) {
    intensity[i] = intensity_of(pixels[i]);
    ibyte[i] = (uint8_t)intensity_byte(intensity[i]);
  }
  for (uint32_t y = 0; y < h; y++)
    for (uint32_t x = 0; x < w; x++) cls[y * w + x] = (uint8_t)classify_extremum(ibyte.data(), w, h, x, y);
  std::vector<PixelLabels> labels((size_t)w * h);
  memset(labels.data(), 0, labels.size() * sizeof(PixelLabels));
  bool ok = true;
  for (uint32_t j = 0; j < h + 3; j++)
    for (uint32_t i = 0; i < w; i++) ok = forward_pixel(labels.data(), w, h, i, j, cls[wrap((int32_t)j, h) * w + i]) && ok;
  for (int32_t j = (int32_t)h + 2; j >= 0; j--)
    for (int32_t i = (int32_t)w - 1; i >= 0; i--) ok = backward_pixel(labels.data(), w, h, (uint32_t)i, (uint32_t)j) && ok;
  std::vector<uint32_t> fields(nb);
  for (uint32_t bj = 0; bj < bh; bj++)
    for (uint32_t bi = 0; bi < bw; bi++)
      fields[bj * bw + bi] = low_high_block(labels.data(), intensity.data(), pixels, w, h, bi, bj);
  for (uint32_t bj = 0; bj < bh; bj++)
    for (uint32_t bi = 0; bi < bw; bi++)
      got[block_index(bi, bj)] =
          ((uint64_t)fields[bj * bw + bi] << 32) | modulation_block(fields.data(), pixels, w, h, bi, bj);
  uint32_t bad = 0, bad_col = 0;
  for (uint32_t b = 0; b < nb; b++) {
    bad += got[b] != want[b];
    bad_col += (got[b] >> 32) != (want[b] >> 32);
  }
  // the decoder, on the reference's blocks, against the reference's decoder
  std::vector<uint32_t> dec_want((size_t)w * h), dec_got((size_t)w * h);
  if (fastc_ref_decompress(4, reinterpret_cast<const uint8_t *>(want.data()), w, h, reinterpret_cast<uint8_t *>(dec_want.data()))) return 5;
  uint32_t bad_px = 0;
  for (uint32_t j = 0; j < h; j++)
    for (uint32_t i = 0; i < w; i++) {
      dec_got[j * w + i] = decode_pixel(reinterpret_cast<const uint2 *>(want.data()), w, h, i, j);
      bad_px += dec_got[j * w + i] != dec_want[j * w + i];
    }
  bad += bad_px;
  printf("blocks %u mismatches %u (colour fields %u) label lists %s decoded pixels differing %u\n", nb, bad - bad_px, bad_col,
         ok ? "ok" : "OVERFLOWED", bad_px);
  if (bad && argc > 3)
    for (uint32_t b = 0, shown = 0; b < nb && shown < 8; b++)
      if (got[b] != want[b]) { printf("  block %u: got %016llx want %016llx\n", b, (unsigned long long)got[b], (unsigned long long)want[b]); shown++; }
  return bad ? 1 : 0;
}
