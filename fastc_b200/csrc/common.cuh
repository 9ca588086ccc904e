// Shared device/host helpers for the sm_100a block encoders.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fastc {

// One 4x4 RGBA8 block = 4 rows x 16 B.  Thread t of a warp loads block bi+t, so a
// warp row-load covers 32 x 16 B = 512 contiguous bytes (fully coalesced 128-bit
// loads; rows are 16 B aligned because width % 4 == 0).  Mirrors the reference's
// GetBlock / ExtractBlock gathers (BPTCEncoder/src/Compressor.cpp:1460-1466,
// DXTEncoder/src/Compressor.cpp:33-40).
__device__ __forceinline__ void load_block(const uint32_t *__restrict__ img, uint32_t width,
                                           uint32_t blocks_x, uint32_t bi, uint32_t px[16]) {
  const uint32_t bx = bi % blocks_x, by = bi / blocks_x;
  const uint4 *row = reinterpret_cast<const uint4 *>(img + (size_t)by * 4 * width + (size_t)bx * 4);
  const uint32_t pitch4 = width >> 2;  // row pitch in uint4 units
#pragma unroll
  for (int j = 0; j < 4; j++) {
    uint4 v = __ldg(row + (size_t)j * pitch4);
    px[4 * j + 0] = v.x;
    px[4 * j + 1] = v.y;
    px[4 * j + 2] = v.z;
    px[4 * j + 3] = v.w;
  }
}

}  // namespace fastc
