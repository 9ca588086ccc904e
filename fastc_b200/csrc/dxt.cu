// DXT1 / DXT5 block encoder for sm_100a.
//
// Behavioural contract: bit-identical to stb_dxt v1.06 driven the way the
// reference drives it (STB_DXT_DITHER, one refinement pass):
//   reference/DXTEncoder/src/stb_dxt.h:477-548 (colour block), :551-601 (alpha
//   block), reference/DXTEncoder/src/Compressor.cpp:47-95 (block loops).
//
// Mapping: one thread per 4x4 block.  There is no data reuse between blocks and a
// warp's loads / stores are already contiguous (512 B per row load, 256 / 512 B
// per store); the encoder is integer-instruction bound long before HBM (DESIGN.md
// section 3), so the arithmetic is arranged for packed-byte instructions (dp4a on
// planar rows, dp2a projections: dxt_block.cuh) and the two 64-byte per-block
// arrays (original pixels, interleaved dithered pixels) live in the thread's
// shared-memory column instead of registers, which lets 8 CTAs of 128 threads
// share an SM.  The few float ops use explicit round-to-nearest intrinsics so that
// no FMA contraction can change a result relative to the reference's SSE2 build.
#include "common.cuh"
#include "dxt_block.cuh"
#include "kernels.h"

namespace fastc {
namespace {

// Single-colour optimal endpoints, stb__OMatch5/6 (stb_dxt.h:121-147, 619-620):
// [0..511] = OMatch5[256][2], [512..1023] = OMatch6[256][2].  Built on the host
// with the same scan order, uploaded once per device.
__device__ uint8_t g_omatch[1024];

constexpr int kDxtThreads = 128;

template <bool DXT5>
__global__ void __launch_bounds__(kDxtThreads, 8)
dxt_encode_kernel(const uint32_t *__restrict__ img, uint32_t width, uint32_t blocks_x,
                  uint32_t first_block, uint32_t num_blocks, uint8_t *__restrict__ out) {
  // [row][thread]: a thread's 16-byte accesses to its own column are conflict-free
  // (each quarter warp covers 128 contiguous bytes)
  __shared__ uint4 s_px[4][kDxtThreads];
  __shared__ uint4 s_d[4][kDxtThreads];
  const uint32_t t = blockIdx.x * kDxtThreads + threadIdx.x;
  if (t >= num_blocks) return;
  const uint32_t bi = first_block + t;
  const uint32_t bx = bi % blocks_x, by = bi / blocks_x;
  const uint4 *row = reinterpret_cast<const uint4 *>(img + (size_t)by * 4 * width + (size_t)bx * 4);
  const uint32_t pitch4 = width >> 2;  // row pitch in uint4 units
  bool constant = true;
  uint32_t first = 0;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const uint4 v = __ldg(row + (size_t)j * pitch4);
    if (j == 0) first = v.x;
    constant = constant && v.x == first && v.y == first && v.z == first && v.w == first;  // alpha included (T13)
    s_px[j][threadIdx.x] = v;
  }
  const dxtb::Rows R = {&s_px[0][threadIdx.x], &s_d[0][threadIdx.x], kDxtThreads};
  const uint2 color = dxtb::compress_color_block(R, constant, g_omatch);
  if (DXT5) {
    const uint2 alpha = dxtb::compress_alpha_block(R);
    reinterpret_cast<uint4 *>(out)[bi] = make_uint4(alpha.x, alpha.y, color.x, color.y);
  } else {
    reinterpret_cast<uint2 *>(out)[bi] = color;
  }
}

}  // namespace

cudaError_t dxt_upload_tables() {
  uint8_t host[1024];
  dxtb::build_omatch(host, 32, false);
  dxtb::build_omatch(host + 512, 64, true);
  return cudaMemcpyToSymbol(g_omatch, host, sizeof(host));
}

cudaError_t launch_dxt(bool dxt5, const void *rgba_dev, uint32_t width, uint32_t first_block,
                       uint32_t num_blocks, void *out_dev, cudaStream_t stream) {
  if (num_blocks == 0) return cudaSuccess;
  const uint32_t threads = kDxtThreads, grid = (num_blocks + threads - 1) / threads;
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
