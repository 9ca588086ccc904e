// Internal launcher interface between the C ABI (capi.cu) and the per-format
// kernels.  Not part of the public boundary (that is include/fastc_gpu.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fastc {

// Per-device one-time constant uploads.
cudaError_t dxt_upload_tables();
cudaError_t etc1_upload_tables();
cudaError_t bc7_upload_tables();
cudaError_t decode_upload_tables();

cudaError_t launch_dxt(bool dxt5, const void *rgba_dev, uint32_t width, uint32_t first_block,
                       uint32_t num_blocks, void *out_dev, cudaStream_t stream);

cudaError_t launch_etc1(const void *rgba_dev, uint32_t width, uint32_t first_block, uint32_t num_blocks,
                        void *out_dev, int quality, cudaStream_t stream);

// PVRTC 4bpp (pvrtc.cu): image-level encoder, the whole (square, power-of-two) texture per call.
// Scratch (intensities, extremum classes, per-pixel label lists: 168 B per pixel) is owned by the
// per-device context and grown on demand.
struct PvrtcWorkspace {
  void *base = nullptr;
  size_t bytes = 0;
  uint32_t *host_flag = nullptr;
};
void pvrtc_free_workspace(PvrtcWorkspace &ws);
// ntex textures of width x height back to back in rgba_dev, their compressed blocks back to back in out_dev
cudaError_t launch_pvrtc(PvrtcWorkspace &ws, const void *rgba_dev, uint32_t width, uint32_t height, uint32_t ntex,
                         void *out_dev, cudaStream_t stream, uint32_t *launches);
size_t pvrtc_scratch_bytes(uint32_t width, uint32_t height);  // per texture
cudaError_t launch_pvrtc_decode(const void *cmp_dev, uint32_t width, uint32_t height, uint32_t first_block,
                                uint32_t num_blocks, void *rgba_dev, cudaStream_t stream);

// Decoders + PSNR (decode.cu).  format: include/fastc_gpu.h numbering.
cudaError_t launch_decode(int format, const void *cmp_dev, uint32_t width, uint32_t first_block, uint32_t num_blocks,
                          void *rgba_dev, cudaStream_t stream);
cudaError_t launch_psnr_sum(const void *a_dev, const void *b_dev, size_t num_pixels, unsigned long long *sum_dev,
                            cudaStream_t stream);
double psnr_from_sum(unsigned long long sum, size_t num_pixels);

// Scratch for the multi-kernel BC7 pipeline (chain records, per-block
// selections, prefix counters).  Owned by the per-device context, grown on
// demand, reused across calls.
struct Bc7Workspace {
  void *base = nullptr;
  size_t bytes = 0;
  uint32_t *host_count = nullptr;  // pinned, 1 word (count_solid result)
  uint32_t *wm_running = nullptr;  // device: watermark base of the chunk being packed
  unsigned long long *counters = nullptr;  // device: [0] QuantizedError calls, [1] pixel-bucket evaluations
  // optional per-stage timing (CUDA events on the launching stream)
  bool timing = false;
  int timed_chunks = 0;
  static constexpr int kMaxTimedChunks = 64;
  cudaEvent_t ev[kMaxTimedChunks][5] = {};
  cudaEvent_t ev_mid[kMaxTimedChunks] = {};  // between setup(+sort) and anneal
  cudaEvent_t *cur_ev = nullptr;             // the timed chunk bc7_back closes
  bool nu = false;                           // the submission in flight uses the non-uniform metric
};
void bc7_free_workspace(Bc7Workspace &ws);

// block_index_base: raster index (in the full texture) of the buffer's block 0;
// keys the per-chain RNG streams so sharded / chunked runs are bit-identical to
// a single submission.  wm_base: solid blocks preceding first_block.
// Encoder settings as the kernels see them: BPTCC::CompressionSettings (reference
// BPTCEncoder/include/FasTC/BPTCCompressor.h:123-158) and rg_etc1's quality (ETCEncoder/src/rg_etc1.h:24-29)
struct EncodeParams {
  int quality = 50;            // m_NumSimulatedAnnealingSteps
  uint64_t seed = 0;           // keys the per-chain RNG streams
  uint32_t block_modes = 0xFF; // m_BlockModes
  int error_metric = 0;        // m_ErrorMetric: 0 uniform, 1 non-uniform
  int etc1_quality = 0;        // rg_etc1::etc1_quality: 0 cLowQuality (what FasTC uses), 1 medium, 2 high
};
cudaError_t launch_bc7(Bc7Workspace &ws, const void *rgba_dev, uint32_t width, uint32_t height,
                       uint32_t first_block, uint32_t num_blocks, void *out_dev, const EncodeParams &prm,
                       uint32_t wm_base, uint32_t block_index_base, cudaStream_t stream, uint32_t *launches);
// The same in two halves for one submission of <= bc7_max_submission() blocks (the host path chains
// watermark bases across streams and GPUs between the two): front = everything but the pack, with
// the submission's solid-block count copied to ws.host_count and `count_ready` recorded right after
// the classification; back = set the watermark base (unless it already sits in ws.wm_running) and pack.
uint32_t bc7_max_submission();
cudaError_t bc7_front(Bc7Workspace &ws, const void *rgba_dev, uint32_t width, uint32_t first_block,
                      uint32_t num_blocks, const EncodeParams &prm, uint32_t block_index_base, cudaStream_t stream,
                      cudaEvent_t count_ready, uint32_t *launches);
// stats_dev: optional device array of bc7_stat_doubles() doubles per block of the submission
// (mode, path, error of every mode tried: the records of the reference's CompressWithStats)
cudaError_t bc7_back(Bc7Workspace &ws, const void *rgba_dev, uint32_t width, uint32_t first_block,
                     uint32_t num_blocks, void *out_dev, uint32_t wm_base, bool base_on_device, cudaStream_t stream,
                     uint32_t *launches, double *stats_dev = nullptr);
uint32_t bc7_stat_doubles();

cudaError_t bc7_count_solid(Bc7Workspace &ws, const void *rgba_dev, uint32_t width, uint32_t first_block,
                            uint32_t num_blocks, cudaStream_t stream, uint32_t *count_out);

// Debug: copies sel[nblocks] then results[nblocks][16][8] of the last launch_bc7 (<= one chunk).
cudaError_t bc7_debug_dump(Bc7Workspace &ws, uint32_t nblocks, uint32_t *sel_out, uint32_t *results_out);

cudaError_t bc7_stage_timing(Bc7Workspace &ws, int enable, double *ms6);

cudaError_t bc7_read_counters(Bc7Workspace &ws, uint64_t *qe_calls, uint64_t *pbe);

}  // namespace fastc
