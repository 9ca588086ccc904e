/* fastc_gpu.h -- C ABI of the B200 (sm_100a) block-compression library
 * (libfastc_gpu.so).  Plain pointers and sizes only: no C++/torch types.
 *
 * This is the boundary the reference's host code binds to.  Each entry point
 * names the reference interface it replaces (paths relative to the reference
 * tree, GammaUNC/FasTC @ 0f8cef65):
 *
 *   fastc_gpu_compress         <- bool CompressImageData(data, w, h, cmpData, cmpDataSz, settings)
 *                                 Core/include/FasTC/TexComp.h:82-89, Core/src/TexComp.cpp:427-525
 *                                 and, via first_block/num_blocks, one CompressionFunc call
 *                                 `void(*)(const FasTC::CompressionJob&)` Core/src/CompressionFuncs.h:29
 *                                 (BPTCC::Compress BPTCEncoder/src/Compressor.cpp:1473,
 *                                  DXTC::CompressImageDXT1/5 DXTEncoder/src/Compressor.cpp:47/74,
 *                                  ETCC::Compress_RG ETCEncoder/src/Compressor.cpp:26)
 *   fastc_gpu_compress_batch   <- FasTC::CompressionJobList (Base/include/FasTC/CompressionJob.h:172-209)
 *                                 / BPTCC::CompressAtomic (BPTCEncoder/src/Compressor.cpp:1542-1574)
 *   fastc_gpu_compress_device  <- the same CompressionFunc, with device-resident
 *   fastc_gpu_count_solid_device  buffers (kernel-only timing; multi-process sharding)
 *   fastc_gpu_decompress(_device) <- CompressedImage::DecompressImage (Core/src/CompressedImage.cpp:86-119):
 *                                 BPTCC::Decompress BPTCEncoder/src/Decompressor.cpp:345,
 *                                 DXTC::DecompressDXT1/5 DXTEncoder/src/Decompressor.cpp:98/127,
 *                                 ETCC::Decompress ETCEncoder/src/Decompressor.cpp:27
 *   fastc_gpu_psnr(_device)    <- FasTC::Image<Pixel>::ComputePSNR (Base/src/Image.cpp:205-255)
 *   fastc_gpu_compressed_size  <- CompressedImage::GetCompressedSize (Core/src/CompressedImage.cpp:136-149)
 *   fastc_gpu_last_error       <- ReportError()'s "TexComp -- %s" message (Core/src/TexComp.cpp:157-159)
 *
 * Conventions: every function returns 0 on success and non-zero on failure
 * (message via fastc_gpu_last_error(), thread-local).  Nothing is allocated
 * that the caller must free.  There is NO CPU fallback: if no CUDA device is
 * usable the call fails.
 *
 * Threads: every entry point may be called from any thread.  Host-pointer
 * submissions to one device are serialised inside the library (the reference's
 * ThreadGroup / WorkerQueue call a CompressionFunc from up to 256 threads at
 * once on disjoint block ranges, Core/src/ThreadGroup.cpp:146-188).  The
 * device-pointer BPTC entry points of one device share one scratch area: calls
 * on different streams are ordered one after the other on the device (each
 * waits, on its stream, for the previous call's kernels), never corrupted.
 *
 * Host buffers may be pageable (new[] / malloc: staged through pinned memory
 * owned by the library) or pinned (cudaHostAlloc / cudaHostRegister: used
 * directly by the copy engines, ~2.5x faster end to end for DXT / ETC1).
 *
 * Pixel layout: row-major RGBA8, R in the lowest byte, pitch = width*4
 * (RGBAEndpoints.h:64-72, Pixel.cpp:165-179).  Block i (raster order over
 * width/4 x height/4 blocks) is written at out + i*block_bytes, exactly like
 * the reference's encoder loops.
 */
#ifndef FASTC_GPU_H_
#define FASTC_GPU_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum fastc_gpu_format {
  FASTC_GPU_DXT1 = 0, /* FasTC::eCompressionFormat_DXT1, 8 B/block  */
  FASTC_GPU_DXT5 = 1, /* FasTC::eCompressionFormat_DXT5, 16 B/block */
  FASTC_GPU_ETC1 = 2, /* FasTC::eCompressionFormat_ETC1, 8 B/block (rg_etc1 cLowQuality) */
  FASTC_GPU_BPTC = 3, /* FasTC::eCompressionFormat_BPTC, 16 B/block (BC7) */
  /* FasTC::eCompressionFormat_PVRTC4, 8 B/block, PVRTCC::Compress(job, eWrapMode_Wrap)
   * (PVRTCEncoder/src/Compressor.cpp:861-944).  Image-level: the texture must be square with a
   * power-of-two side >= 8 (Core/src/TexComp.cpp:477-482), a call encodes all of it on one GPU
   * (first_block 0, every block), and blocks are stored in the reference's interleaved (Morton)
   * order.  Its labelling scan is one serial chain per texture, so it scales over the textures of
   * a batch, not inside one. */
  FASTC_GPU_PVRTC4 = 4
};

/* One texture of a batch submission (CompressionJob: Base/include/FasTC/CompressionJob.h:40-140). */
typedef struct fastc_gpu_job {
  const uint8_t *rgba_host; /* width*height*4 bytes */
  uint8_t *out_host;        /* >= fastc_gpu_compressed_size() bytes */
  uint32_t width, height;   /* multiples of 4 */
} fastc_gpu_job;

/* Timing breakdown returned by the host-pointer entry points (milliseconds). */
typedef struct fastc_gpu_timing {
  double kernel_ms; /* max over GPUs of the summed kernel time (CUDA events)   */
  double total_ms;  /* wall time of the whole call: H2D + kernels + D2H        */
  uint64_t h2d_bytes, d2h_bytes;
  uint32_t kernel_launches; /* number of our kernels launched by the call */
} fastc_gpu_timing;

/* Encoder settings beyond (format, quality): what the reference passes per call as
 * BPTCC::CompressionSettings (BPTCEncoder/include/FasTC/BPTCCompressor.h:123-158; applied at
 * BPTCEncoder/src/Compressor.cpp:1848-1857) and rg_etc1's quality level
 * (ETCEncoder/src/rg_etc1.h:24-29; FasTC itself always passes cLowQuality,
 * ETCEncoder/src/Compressor.cpp:36-37).  Taken by the *_opt variants of the compress entry
 * points; a NULL pointer means the defaults below, which is what the plain entry points use. */
/* Per-block record of the BPTC statistics (the lines BPTCC::CompressWithStats logs per block,
 * BPTCEncoder/src/Compressor.cpp:1577-1624, :1954-2183): the mode packed, the path taken
 * (0 solid colour, 1 transparent, 2 a partition shape estimated to ~zero error, 3 full search) and
 * the best error of every mode tried (-1: not tried).  The reference's stats variant also logs
 * per-mode shape ESTIMATES; the GPU selection kernel does not keep them. */
typedef struct fastc_gpu_bptc_block_stat {
  double mode;
  double path;
  double mode_error[8];
} fastc_gpu_bptc_block_stat;

typedef struct fastc_gpu_options {
  uint32_t struct_size;       /* = sizeof(fastc_gpu_options) */
  uint32_t bptc_block_modes;  /* m_BlockModes: bit m set = BC7 mode m may be used (default 0xFF)     */
  int32_t bptc_error_metric;  /* m_ErrorMetric: 0 = eErrorMetric_Uniform (default), 1 = _Nonuniform  */
  int32_t etc1_quality;       /* 0 = cLowQuality (default), 1 = cMediumQuality, 2 = cHighQuality     */
  /* host-pointer BPTC submissions only (fastc_gpu_compress_opt): NULL, or one record per block of the
   * IMAGE (index = raster block index); the records of the encoded block range are filled */
  fastc_gpu_bptc_block_stat *bptc_block_stats;
} fastc_gpu_options;
#define FASTC_GPU_OPTIONS_INIT { (uint32_t)sizeof(fastc_gpu_options), 0xFFu, 0, 0, 0 }

/* Number of visible CUDA devices (<0 on error). Creates nothing. */
int fastc_gpu_device_count(void);

/* Optional: pre-create per-device contexts/streams/workspaces for devices
 * [0, num_gpus).  num_gpus <= 0 means "all visible".  Called lazily otherwise. */
int fastc_gpu_init(int num_gpus);
void fastc_gpu_shutdown(void);

uint32_t fastc_gpu_block_bytes(int format);
uint64_t fastc_gpu_compressed_size(int format, uint32_t width, uint32_t height);

/* Host -> host.  Encodes blocks [first_block, first_block+num_blocks) of the
 * image (num_blocks == 0 means "to the end").  quality = SA steps (BPTC only;
 * SCompressionSettings::iQuality), seed keys the per-block RNG streams,
 * chunk_blocks = blocks per pipeline chunk (SCompressionSettings::iJobSize;
 * 0 = auto), num_gpus = devices to shard block rows over (<= 0: all visible
 * devices).  timing may be NULL. */
int fastc_gpu_compress(int format, const uint8_t *rgba_host, uint32_t width, uint32_t height,
                       uint32_t first_block, uint32_t num_blocks, uint8_t *out_host,
                       int quality, uint64_t seed, uint32_t chunk_blocks, int num_gpus,
                       fastc_gpu_timing *timing);

/* N independent textures in one submission, whole textures dealt round-robin
 * to the GPUs. */
int fastc_gpu_compress_batch(int format, const fastc_gpu_job *jobs, uint32_t num_jobs,
                             int quality, uint64_t seed, int num_gpus, fastc_gpu_timing *timing);

/* The same three with explicit encoder settings (BPTCC::Compress(job, settings),
 * BPTCEncoder/src/Compressor.cpp:1473; rg_etc1::pack_etc1_block(..., pack_params),
 * ETCEncoder/src/rg_etc1.cpp:2192). */
int fastc_gpu_compress_opt(int format, const uint8_t *rgba_host, uint32_t width, uint32_t height,
                           uint32_t first_block, uint32_t num_blocks, uint8_t *out_host,
                           int quality, uint64_t seed, uint32_t chunk_blocks, int num_gpus,
                           fastc_gpu_timing *timing, const fastc_gpu_options *options);
int fastc_gpu_compress_batch_opt(int format, const fastc_gpu_job *jobs, uint32_t num_jobs,
                                 int quality, uint64_t seed, int num_gpus, fastc_gpu_timing *timing,
                                 const fastc_gpu_options *options);

/* Device -> device on the CURRENT device, asynchronous on `cuda_stream`
 * (a cudaStream_t passed as void*; NULL = legacy default stream).
 * rgba_dev/out_dev address the WHOLE image / whole output; only the block
 * range is touched.  wm_base = number of solid-colour blocks that precede
 * first_block in raster order (BC7 watermark sequence, Compressor.cpp:1457);
 * ignored by the other formats.  block_index_base = raster index, in the full
 * texture, of this buffer's block 0 when the buffer is a slab of a larger
 * texture (0 otherwise): it keys the per-block RNG streams, so a sharded run is
 * bit-identical to an unsharded one.  launches_out (may be NULL) receives the
 * number of kernels enqueued. */
int fastc_gpu_compress_device(int format, const void *rgba_dev, uint32_t width, uint32_t height,
                              uint32_t first_block, uint32_t num_blocks, void *out_dev,
                              int quality, uint64_t seed, uint32_t wm_base, uint32_t block_index_base,
                              void *cuda_stream, uint32_t *launches_out);

int fastc_gpu_compress_device_opt(int format, const void *rgba_dev, uint32_t width, uint32_t height,
                                  uint32_t first_block, uint32_t num_blocks, void *out_dev,
                                  int quality, uint64_t seed, uint32_t wm_base, uint32_t block_index_base,
                                  void *cuda_stream, uint32_t *launches_out,
                                  const fastc_gpu_options *options);

/* Counts solid-colour blocks in the range (needed to chain wm_base across
 * shards).  Synchronous with respect to `cuda_stream`. */
int fastc_gpu_count_solid_device(const void *rgba_dev, uint32_t width, uint32_t height,
                                 uint32_t first_block, uint32_t num_blocks, void *cuda_stream,
                                 uint32_t *count_out);

/* Decoders (bit-identical to the reference's, including its departures from the format
 * specifications -- see fastc_b200/csrc/decode.cu).  Host -> host on the current device;
 * the device variant is asynchronous on `cuda_stream`.  rgba_out: width*height*4 bytes. */
int fastc_gpu_decompress(int format, const uint8_t *cmp_host, uint32_t width, uint32_t height,
                         uint8_t *rgba_out_host, fastc_gpu_timing *timing);
int fastc_gpu_decompress_device(int format, const void *cmp_dev, uint32_t width, uint32_t height,
                                void *rgba_out_dev, void *cuda_stream);

/* The reference's PSNR between two RGBA8 images of the same size (alpha-premultiplied RGB error,
 * peak 3*255^2).  Identical images give +inf.  Both variants synchronise before returning. */
int fastc_gpu_psnr(const uint8_t *a_host, const uint8_t *b_host, uint32_t width, uint32_t height,
                   double *psnr_out);
int fastc_gpu_psnr_device(const void *a_dev, const void *b_dev, uint32_t width, uint32_t height,
                          void *cuda_stream, double *psnr_out);

/* BC7 work counters of the last BPTC call on this thread's device context
 * (QuantizedError calls and pixel-bucket evaluations), only
 * maintained when the library is built with -DFASTC_GPU_COUNTERS. */
int fastc_gpu_bc7_counters(uint64_t *qe_calls, uint64_t *pixel_bucket_evals);

/* Per-stage device time of the BC7 pipeline on this device's device-API context:
 * enable != 0 arms CUDA-event recording around each stage of subsequent
 * fastc_gpu_compress_device(BPTC) calls on their stream; ms6 (may be NULL) receives
 * the summed milliseconds of the LAST call's stages {classify+scan, select,
 * setup(+sort), anneal, pack, total}, synchronising on its events. */
int fastc_gpu_bc7_stage_ms(int enable, double *ms6);

/* Diagnostics: after a BPTC fastc_gpu_compress_device call of nblocks (<= 2^22)
 * blocks on this device, copies the per-block selection word and the per-chain
 * fit results (nblocks x 16 slots x 8 words) out of the scratch. */
int fastc_gpu_debug_bc7_dump(uint32_t nblocks, uint32_t *sel_out, uint32_t *results_out);

const char *fastc_gpu_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* FASTC_GPU_H_ */
