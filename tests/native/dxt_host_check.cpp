// TEST INFRASTRUCTURE: runs the product's DXT block arithmetic (fastc_b200/csrc/dxt_block.cuh, the code
// dxt_encode_kernel executes per thread) on the CPU and compares every block with the oracle
// (oracle/dxt_oracle.cpp, pinned to stb_dxt).  Built and run by tests/test_native_host.py with
//   g++ -O2 -ffp-contract=off -I fastc_b200/csrc dxt_host_check.cpp oracle/libfastc_oracle.so
// Usage: dxt_host_check <dxt5> <width> <height> < rgba-bytes     -> prints "blocks N mismatches M"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "dxt_block.cuh"
#include "../../oracle/oracle.h"

int main(int argc, char **argv) {
  if (argc < 4) return 2;
  const int dxt5 = atoi(argv[1]);
  const uint32_t w = (uint32_t)atoi(argv[2]), h = (uint32_t)atoi(argv[3]);
  std::vector<uint8_t> img((size_t)w * h * 4);
  if (fread(img.data(), 1, img.size(), stdin) != img.size()) return 3;
  const uint32_t nb = (w / 4) * (h / 4), bsz = dxt5 ? 16 : 8;
  std::vector<uint8_t> want((size_t)nb * bsz), got((size_t)nb * bsz);
  fastc_oracle_dxt(dxt5, img.data(), w, h, 0, nb, want.data());
  uint8_t omatch[1024];
  fastc::dxtb::build_omatch(omatch, 32, false);
  fastc::dxtb::build_omatch(omatch + 512, 64, true);
  for (uint32_t bi = 0; bi < nb; bi++) {
    const uint32_t bx = bi % (w / 4), by = bi / (w / 4);
    uint4 px[4], d[4];
    bool constant = true;
    uint32_t first = 0;
    for (int j = 0; j < 4; j++) {
      memcpy(&px[j], &img[((size_t)(by * 4 + j) * w + bx * 4) * 4], 16);
      if (j == 0) first = px[0].x;
      constant = constant && px[j].x == first && px[j].y == first && px[j].z == first && px[j].w == first;
    }
    const fastc::dxtb::Rows R = {px, d, 1};
    const uint2 color = fastc::dxtb::compress_color_block(R, constant, omatch);
    uint32_t o[4];
    if (dxt5) {
      const uint2 alpha = fastc::dxtb::compress_alpha_block(R);
      o[0] = alpha.x; o[1] = alpha.y; o[2] = color.x; o[3] = color.y;
    } else {
      o[0] = color.x; o[1] = color.y;
    }
    memcpy(&got[(size_t)bi * bsz], o, bsz);
  }
  uint32_t bad = 0;
  for (uint32_t bi = 0; bi < nb; bi++) bad += memcmp(&got[(size_t)bi * bsz], &want[(size_t)bi * bsz], bsz) != 0;
  printf("blocks %u mismatches %u\n", nb, bad);
  return bad ? 1 : 0;
}
