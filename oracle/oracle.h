// TEST INFRASTRUCTURE ONLY -- C interface of the CPU oracle (libfastc_oracle.so).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load it.  The product (fastc_b200/, include/) never does.
#ifndef FASTC_ORACLE_H_
#define FASTC_ORACLE_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

// Format numbering shared with include/fastc_gpu.h.
enum { FASTC_ORACLE_DXT1 = 0, FASTC_ORACLE_DXT5 = 1, FASTC_ORACLE_ETC1 = 2, FASTC_ORACLE_BPTC = 3 };

// All encoders take a row-major RGBA8 image (pitch = width*4) and encode the
// raster-order block range [first_block, first_block+num_blocks); block i is
// written at out + i*block_size (same addressing as the reference's
// CompressionJob loops).
void fastc_oracle_dxt(int dxt5, const uint8_t *rgba, uint32_t width, uint32_t height,
                      uint32_t first_block, uint32_t num_blocks, uint8_t *out);
void fastc_oracle_etc1(const uint8_t *rgba, uint32_t width, uint32_t height,
                       uint32_t first_block, uint32_t num_blocks, uint8_t *out);

// quality: rg_etc1::etc1_quality (0 low -- what FasTC passes --, 1 medium, 2 high)
void fastc_oracle_etc1_quality(const uint8_t *rgba, uint32_t width, uint32_t height,
                               uint32_t first_block, uint32_t num_blocks, uint8_t *out, int quality);

// rng_mode 0: the reference's single global LCG, *lcg_state is read and updated
//             (pins the restatement against oracle/_ref at q>0, -t 1);
// rng_mode 1: per-chain keyed streams derived from (seed, block index, chain id)
//             -- the scheme the CUDA path uses (bit-comparable with the GPU).
// wm_base = number of solid-colour blocks that precede first_block (watermark T1).
void fastc_oracle_bc7(const uint8_t *rgba, uint32_t width, uint32_t height,
                      uint32_t first_block, uint32_t num_blocks, uint8_t *out,
                      int quality, int rng_mode, uint32_t *lcg_state, uint64_t seed,
                      uint32_t wm_base);

void fastc_oracle_bc7_keyed(const uint8_t *rgba, uint32_t width, uint32_t height,
                            uint32_t first_block, uint32_t num_blocks, uint8_t *out,
                            int quality, uint64_t seed, uint32_t wm_base, uint32_t block_index_base);

// The same with the reference's per-call BPTCC::CompressionSettings: m_BlockModes (bit m = mode m
// allowed) and m_ErrorMetric (0 uniform, 1 non-uniform).
void fastc_oracle_bc7_settings(const uint8_t *rgba, uint32_t width, uint32_t height,
                               uint32_t first_block, uint32_t num_blocks, uint8_t *out,
                               int quality, int rng_mode, uint32_t *lcg_state, uint64_t seed,
                               uint32_t wm_base, uint32_t block_index_base, uint32_t block_modes,
                               int error_metric);

// Decoders + the reference's PSNR definition (Base/src/Image.cpp:205-255).
void fastc_oracle_decode(int format, const uint8_t *cmp, uint32_t width, uint32_t height,
                         uint8_t *rgba_out);
double fastc_oracle_psnr(const uint8_t *a, const uint8_t *b, uint32_t width, uint32_t height);

#ifdef __cplusplus
}
#endif
#endif
