// C ABI (include/fastc_gpu.h) over the sm_100a kernels: per-device contexts,
// workspaces, the host->host sharded path and the device->device path.
//
// Replaces, on the reference side: CompressImageData's dispatch + the Serial /
// ThreadGroup / WorkerQueue schedulers (reference/Core/src/TexComp.cpp:161-365,
// 427-525; Core/src/ThreadGroup.cpp:133-192; Core/src/WorkerQueue.cpp:196-241).
// Where the reference splits the raster block range over <=256 pthreads, this
// splits it over GPUs (contiguous block-row slabs, one host thread + streams
// per GPU) and, inside a GPU, over pipeline chunks of `chunk_blocks` blocks so
// that H2D, kernels and D2H of neighbouring chunks overlap.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/fastc_gpu.h"
#include "kernels.h"

namespace fastc {
namespace {

thread_local char tl_error[512] = "";

int fail(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tl_error, sizeof(tl_error), fmt, ap);
  va_end(ap);
  return 1;
}

#define CU_TRY(expr)                                                                         \
  do {                                                                                       \
    cudaError_t e_ = (expr);                                                                 \
    if (e_ != cudaSuccess) return fail("%s failed: %s", #expr, cudaGetErrorString(e_));      \
  } while (0)

constexpr int kMaxDevices = 16;
constexpr int kPipeDepth = 3;  // staging slots per device for the host path
constexpr int kStagePieces = 4;               // pinned upload ring (pageable inputs)
constexpr size_t kStagePieceBytes = 16u << 20;

struct DeviceCtx {
  bool tables_ready = false;
  bool ready = false;
  cudaStream_t streams[kPipeDepth] = {};
  cudaEvent_t ev_start[kPipeDepth] = {}, ev_stop[kPipeDepth] = {};
  void *in_buf[kPipeDepth] = {};
  void *out_buf[kPipeDepth] = {};
  size_t in_cap[kPipeDepth] = {}, out_cap[kPipeDepth] = {};
  // slot i for pipeline slot i of the host path, slot kPipeDepth for the device API;
  // grown on demand, reused across calls
  Bc7Workspace bc7ws[kPipeDepth + 1];
  // host-path pipeline state: the slot rotation carries over from one submission to the next of a
  // batch, so texture j+1 uploads while texture j encodes and texture j-1 downloads
  uint64_t next_chunk = 0;
  bool slot_busy[kPipeDepth] = {};
  // one host-path submission (or batch) at a time per device: FasTC's ThreadGroup / WorkerQueue
  // call a CompressionFunc from many threads at once (Core/src/ThreadGroup.cpp:146-188)
  std::mutex host_mu;
  // Pageable caller memory (what FasTC hands over: new[] / malloc) would make every
  // cudaMemcpyAsync synchronous and serialise the pipeline, so it is staged through pinned
  // memory owned by the library: uploads through a ring of kStagePieces pieces (filled by a few
  // host threads while the previous piece is on the wire), downloads into a pinned buffer per
  // slot that is copied out when the slot's chunk has finished.
  void *pin_in[kStagePieces] = {};
  cudaEvent_t pin_in_ev[kStagePieces] = {};
  bool pin_in_used[kStagePieces] = {};
  uint64_t pin_in_next = 0;
  void *pin_out[kPipeDepth] = {};
  size_t pin_out_cap[kPipeDepth] = {};
  uint8_t *pend_dst[kPipeDepth] = {};  // copy-out the slot still owes the caller
  size_t pend_bytes[kPipeDepth] = {};
  unsigned long long *psnr_sum = nullptr, *psnr_host = nullptr;  // device / pinned accumulators of fastc_gpu_psnr*
  std::mutex mu;
};

DeviceCtx g_ctx[kMaxDevices];
std::mutex g_init_mu;
int g_num_init = 0;

int device_count() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return -1;
  return std::min(n, kMaxDevices);
}

// Uploads constant tables for the current device once (any entry point).
int ensure_tables(int dev) {
  DeviceCtx &c = g_ctx[dev];
  std::lock_guard<std::mutex> lk(c.mu);
  if (c.tables_ready) return 0;
  CU_TRY(dxt_upload_tables());
  CU_TRY(etc1_upload_tables());
  CU_TRY(bc7_upload_tables());
  CU_TRY(decode_upload_tables());
  c.tables_ready = true;
  return 0;
}

int ensure_ctx(int dev) {
  CU_TRY(cudaSetDevice(dev));
  if (ensure_tables(dev)) return 1;
  DeviceCtx &c = g_ctx[dev];
  std::lock_guard<std::mutex> lk(c.mu);
  if (c.ready) return 0;
  for (int i = 0; i < kPipeDepth; i++) {
    CU_TRY(cudaStreamCreateWithFlags(&c.streams[i], cudaStreamNonBlocking));
    CU_TRY(cudaEventCreate(&c.ev_start[i]));
    CU_TRY(cudaEventCreate(&c.ev_stop[i]));
  }
  c.ready = true;
  return 0;
}

int grow(void **buf, size_t *cap, size_t need) {
  if (*cap >= need) return 0;
  if (*buf) CU_TRY(cudaFree(*buf));
  *buf = nullptr;
  *cap = 0;
  CU_TRY(cudaMalloc(buf, need));
  *cap = need;
  return 0;
}

bool valid_format(int f) { return f >= FASTC_GPU_DXT1 && f <= FASTC_GPU_BPTC; }

int check_dims(int format, uint32_t width, uint32_t height) {
  if (!valid_format(format)) return fail("unknown compression format %d", format);
  // reference: "Image dimensions must be multiples of the block size" (TexComp.cpp:472-476)
  if (width == 0 || height == 0 || (width & 3) || (height & 3))
    return fail("image dimensions %ux%u are not non-zero multiples of the 4x4 block", width, height);
  return 0;
}

// Enqueue every kernel needed to encode [first_block, first_block+num_blocks)
// on `stream` of the current device.
int enqueue(int dev, int ws_slot, int format, const void *rgba_dev, uint32_t width, uint32_t height,
            uint32_t first_block, uint32_t num_blocks, void *out_dev, int quality, uint64_t seed,
            uint32_t wm_base, uint32_t block_index_base, cudaStream_t stream, uint32_t *launches) {
  uint32_t n = 0;
  switch (format) {
    case FASTC_GPU_DXT1:
    case FASTC_GPU_DXT5:
      CU_TRY(launch_dxt(format == FASTC_GPU_DXT5, rgba_dev, width, first_block, num_blocks, out_dev, stream));
      n = num_blocks ? 1 : 0;
      break;
    case FASTC_GPU_ETC1:
      CU_TRY(launch_etc1(rgba_dev, width, first_block, num_blocks, out_dev, stream));
      n = num_blocks ? 1 : 0;
      break;
    case FASTC_GPU_BPTC: {
      DeviceCtx &c = g_ctx[dev];
      std::lock_guard<std::mutex> lk(c.mu);
      CU_TRY(launch_bc7(c.bc7ws[ws_slot], rgba_dev, width, height, first_block, num_blocks, out_dev, quality, seed,
                        wm_base, block_index_base, stream, &n));
      break;
    }
  }
  if (launches) *launches += n;
  return 0;
}

// Solid-colour blocks in the raster range [lo, hi) of a host image (64 B compare per
// block, split over a few host threads).  Solid-ness depends on the input only, so
// the BC7 watermark order (Compressor.cpp:135-140,1457) can be fixed before encoding.
uint32_t host_count_solid(const uint8_t *rgba_host, uint32_t width, uint32_t lo, uint32_t hi) {
  if (hi <= lo) return 0;
  const uint32_t bx = width / 4;
  const uint32_t *img = reinterpret_cast<const uint32_t *>(rgba_host);
  auto count = [&](uint32_t a, uint32_t b) {
    uint32_t cnt = 0;
    for (uint32_t bi = a; bi < b; bi++) {
      const uint32_t *p = img + (size_t)(bi / bx) * 4 * width + (size_t)(bi % bx) * 4;
      const uint32_t v = p[0];
      bool same = true;
      for (int j = 0; j < 4 && same; j++)
        for (int i = 0; i < 4; i++)
          if (p[(size_t)j * width + i] != v) { same = false; break; }
      cnt += same;
    }
    return cnt;
  };
  const uint32_t n = hi - lo;
  const int nt = (int)std::min<uint32_t>(8, std::max<uint32_t>(1, n >> 16));
  if (nt == 1) return count(lo, hi);
  std::vector<uint32_t> part(nt, 0);
  std::vector<std::thread> th;
  for (int k = 0; k < nt; k++)
    th.emplace_back([&, k] {
      part[k] = count(lo + (uint32_t)((uint64_t)n * k / nt), lo + (uint32_t)((uint64_t)n * (k + 1) / nt));
    });
  for (auto &t : th) t.join();
  uint32_t total = 0;
  for (uint32_t c : part) total += c;
  return total;
}

struct Shard {
  int dev;
  uint32_t first_block, num_blocks;  // within the image
  uint32_t wm_base = 0;
  double kernel_ms = 0;
  uint32_t launches = 0;
  uint64_t h2d = 0, d2h = 0;
  int rc = 0;
  char err[512] = "";
};

// Host->host for one GPU's slab.  The slab is cut into chunks of whole block
// rows; chunk k uses staging slot k % kPipeDepth, so its H2D overlaps the
// kernels of chunk k-1 and the D2H of chunk k-2 (three copy/compute engines).
// true: plain pageable host memory (not cudaHostAlloc'ed / cudaHostRegister'ed / managed)
bool is_pageable(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();  // older drivers report unregistered memory as an error: clear it
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}

// memcpy split over a few host threads (one thread moves ~10 GB/s; PCIe 5 x16 wants ~50)
void par_memcpy(void *dst, const void *src, size_t n) {
  const int nt = (int)std::min<size_t>(8, n >> 20);  // >= 1 MiB per thread
  if (nt <= 1) {
    memcpy(dst, src, n);
    return;
  }
  std::vector<std::thread> th;
  for (int k = 1; k < nt; k++) {
    const size_t a = n * k / nt, b = n * (k + 1) / nt;
    th.emplace_back([=] { memcpy((uint8_t *)dst + a, (const uint8_t *)src + a, b - a); });
  }
  memcpy(dst, src, n / nt);
  for (auto &t : th) t.join();
}

// Host -> device on `st`.  Pinned sources go straight to the copy engine; pageable ones through
// the pinned ring, piece by piece.
int upload(DeviceCtx &c, void *dst_dev, const uint8_t *src, size_t bytes, cudaStream_t st) {
  if (!is_pageable(src)) {
    CU_TRY(cudaMemcpyAsync(dst_dev, src, bytes, cudaMemcpyHostToDevice, st));
    return 0;
  }
  for (size_t off = 0; off < bytes; off += kStagePieceBytes) {
    const int i = (int)(c.pin_in_next++ % kStagePieces);
    const size_t n = std::min(kStagePieceBytes, bytes - off);
    if (!c.pin_in[i]) {
      CU_TRY(cudaHostAlloc(&c.pin_in[i], kStagePieceBytes, cudaHostAllocDefault));
      CU_TRY(cudaEventCreateWithFlags(&c.pin_in_ev[i], cudaEventDisableTiming));
    }
    if (c.pin_in_used[i]) CU_TRY(cudaEventSynchronize(c.pin_in_ev[i]));  // its previous upload has left the buffer
    par_memcpy(c.pin_in[i], src + off, n);
    CU_TRY(cudaMemcpyAsync((uint8_t *)dst_dev + off, c.pin_in[i], n, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaEventRecord(c.pin_in_ev[i], st));
    c.pin_in_used[i] = true;
  }
  return 0;
}

// Device -> host on `st` for the chunk in `slot`.  Pageable destinations receive their bytes
// from the slot's pinned buffer once the chunk has finished (finish_slot).
int download(DeviceCtx &c, int slot, uint8_t *dst, const void *src_dev, size_t bytes, cudaStream_t st) {
  if (!is_pageable(dst)) {
    CU_TRY(cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToHost, st));
    return 0;
  }
  if (c.pin_out_cap[slot] < bytes) {
    if (c.pin_out[slot]) CU_TRY(cudaFreeHost(c.pin_out[slot]));
    c.pin_out[slot] = nullptr;
    c.pin_out_cap[slot] = 0;
    CU_TRY(cudaHostAlloc(&c.pin_out[slot], bytes, cudaHostAllocDefault));
    c.pin_out_cap[slot] = bytes;
  }
  CU_TRY(cudaMemcpyAsync(c.pin_out[slot], src_dev, bytes, cudaMemcpyDeviceToHost, st));
  c.pend_dst[slot] = dst;
  c.pend_bytes[slot] = bytes;
  return 0;
}

// Waits for the chunk in `slot`, books its kernel time and hands over a staged download.
int finish_slot(DeviceCtx &c, int slot, double *kernel_ms) {
  if (!c.slot_busy[slot]) return 0;
  CU_TRY(cudaStreamSynchronize(c.streams[slot]));
  float ms = 0;
  CU_TRY(cudaEventElapsedTime(&ms, c.ev_start[slot], c.ev_stop[slot]));
  if (kernel_ms) *kernel_ms += ms;
  if (c.pend_dst[slot]) {
    par_memcpy(c.pend_dst[slot], c.pin_out[slot], c.pend_bytes[slot]);
    c.pend_dst[slot] = nullptr;
    c.pend_bytes[slot] = 0;
  }
  c.slot_busy[slot] = false;
  return 0;
}

// Waits for every chunk still in flight on the device's staging slots.
int drain_slots(DeviceCtx &c, double *kernel_ms) {
  for (int slot = 0; slot < kPipeDepth; slot++)
    if (finish_slot(c, slot, kernel_ms)) return 1;
  return 0;
}

// drain = false leaves the last chunks in flight (batch submissions: the caller drains once at
// the end, so consecutive textures overlap).
int run_shard(Shard &s, int format, const uint8_t *rgba_host, uint32_t width, uint32_t height,
              uint8_t *out_host, int quality, uint64_t seed, uint32_t chunk_blocks, bool drain = true) {
  if (ensure_ctx(s.dev)) return 1;
  DeviceCtx &c = g_ctx[s.dev];
  const uint32_t bx = width / 4;
  const uint32_t bsz = fastc_gpu_block_bytes(format);
  // Work in whole block rows: rows [row0, row1) cover the shard's block range.
  const uint32_t row0 = s.first_block / bx;
  const uint32_t row1 = (s.first_block + s.num_blocks + bx - 1) / bx;
  uint32_t rows_per_chunk;
  if (chunk_blocks == 0 && format == FASTC_GPU_BPTC) {
    // BC7 is compute-bound (copies are ~1% of the time) and its watermark chain is kept
    // on the device inside one submission: no host-side chunking unless asked for.
    rows_per_chunk = row1 - row0;
  } else if (chunk_blocks == 0) {
    // auto: ~4 Mi pixels per chunk keeps the copy engines busy without making the
    // pipeline too coarse.
    rows_per_chunk = std::max<uint32_t>(1, (1u << 18) / bx);
  } else {
    rows_per_chunk = std::max<uint32_t>(1, chunk_blocks / bx);
  }
  const uint32_t total_rows = row1 - row0;
  rows_per_chunk = std::max<uint32_t>(1, rows_per_chunk);
  // chunk k covers block rows [bounds[k], bounds[k + 1])
  std::vector<uint32_t> bounds;
  for (uint32_t r = row0; r < row1; r += rows_per_chunk) bounds.push_back(r);
  bounds.push_back(row1);
  if (chunk_blocks == 0 && format == FASTC_GPU_BPTC && total_rows >= 64 &&
      (size_t)total_rows * 4 * width * 4 >= ((size_t)96 << 20)) {
    // BC7 auto, big uploads (>= 96 MiB, ~2 ms on the wire): two halves on two streams.  The second
    // half's upload and shape selection / fits run under the first half's annealing, the first
    // half's download under the second's kernels, and each half is still large enough for the
    // persistent annealing kernel's ~1.5 ms tail not to matter.  Measured at 8192^2
    // (tools/time_e2e_chunks.py): one chunk 238.9 ms, two halves 236.6 ms, 1/16 + 15/16 250.4 ms.
    bounds.assign({row0, row0 + total_rows / 2, row1});
  }
  const uint32_t nchunks = (uint32_t)bounds.size() - 1;
  // BC7's watermark chain needs the solid-block count of every earlier chunk;
  // bc7 tracks that itself when the whole shard is submitted as one range, so
  // BPTC shards are uploaded chunk-wise but encoded per chunk with a running base.
  uint32_t wm_base = s.wm_base;

  // nothing in flight (every submission except the later textures of a batch): restart the slot
  // rotation, so that the same chunk lands in the same slot call after call and its staging
  // buffers / BC7 workspace are grown once
  bool idle = true;
  for (int i = 0; i < kPipeDepth; i++) idle = idle && !c.slot_busy[i];
  if (idle) c.next_chunk = 0;
  for (uint32_t k = 0; k < nchunks; k++) {
    const int slot = (int)(c.next_chunk++ % kPipeDepth);
    cudaStream_t st = c.streams[slot];
    const uint32_t r0 = bounds[k], r1 = bounds[k + 1];
    const size_t in_bytes = (size_t)(r1 - r0) * 4 * width * 4;
    const size_t out_bytes = (size_t)(r1 - r0) * bx * bsz;
    // slot reuse: the slot's previous chunk must have finished (its timing is booked and a staged
    // download handed over) before its buffers are overwritten or regrown
    if (finish_slot(c, slot, &s.kernel_ms)) return 1;
    if (c.in_cap[slot] < in_bytes || c.out_cap[slot] < out_bytes) {
      CU_TRY(cudaStreamSynchronize(st));
      if (grow(&c.in_buf[slot], &c.in_cap[slot], in_bytes)) return 1;
      if (grow(&c.out_buf[slot], &c.out_cap[slot], out_bytes)) return 1;
    }
    if (upload(c, c.in_buf[slot], rgba_host + (size_t)r0 * 4 * width * 4, in_bytes, st)) return 1;
    s.h2d += in_bytes;
    // block range of this chunk in chunk-local coordinates (the staged slab is an
    // image of (r1-r0)*4 rows)
    uint32_t lo = (k == 0) ? s.first_block - row0 * bx : 0;
    uint32_t hi = (r1 - r0) * bx;
    if (r1 == row1) hi = s.first_block + s.num_blocks - r0 * bx;
    uint32_t solid = 0;
    if (format == FASTC_GPU_BPTC && nchunks > 1) {
      // need this chunk's solid count before the next chunk can be packed
      if (fastc_gpu_count_solid_device(c.in_buf[slot], width, (r1 - r0) * 4, lo, hi - lo, st, &solid)) return 1;
      s.launches += 1;
    }
    CU_TRY(cudaEventRecord(c.ev_start[slot], st));
    if (enqueue(s.dev, slot, format, c.in_buf[slot], width, (r1 - r0) * 4, lo, hi - lo, c.out_buf[slot], quality,
                seed, wm_base, r0 * bx, st, &s.launches))
      return 1;
    CU_TRY(cudaEventRecord(c.ev_stop[slot], st));
    c.slot_busy[slot] = true;
    wm_base += solid;
    if (download(c, slot, out_host + ((size_t)r0 * bx + lo) * bsz, (uint8_t *)c.out_buf[slot] + (size_t)lo * bsz,
                 (size_t)(hi - lo) * bsz, st))
      return 1;
    s.d2h += (size_t)(hi - lo) * bsz;
  }
  if (drain && drain_slots(c, &s.kernel_ms)) return 1;
  return 0;
}

int run_shard_locked(Shard &s, int format, const uint8_t *rgba_host, uint32_t width, uint32_t height,
                     uint8_t *out_host, int quality, uint64_t seed, uint32_t chunk_blocks) {
  if (s.dev < 0 || s.dev >= kMaxDevices) return fail("bad device %d", s.dev);
  std::lock_guard<std::mutex> lk(g_ctx[s.dev].host_mu);
  return run_shard(s, format, rgba_host, width, height, out_host, quality, seed, chunk_blocks);
}

}  // namespace
}  // namespace fastc

using namespace fastc;

extern "C" {

int fastc_gpu_device_count(void) { return device_count(); }

int fastc_gpu_init(int num_gpus) {
  std::lock_guard<std::mutex> lk(g_init_mu);
  int n = device_count();
  if (n <= 0) return fail("no CUDA device available (there is no CPU fallback)");
  if (num_gpus <= 0 || num_gpus > n) num_gpus = n;
  int prev = 0;
  cudaGetDevice(&prev);
  for (int d = 0; d < num_gpus; d++)
    if (ensure_ctx(d)) return 1;
  cudaSetDevice(prev);
  g_num_init = std::max(g_num_init, num_gpus);
  return 0;
}

void fastc_gpu_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_init_mu);
  int n = device_count();
  for (int d = 0; d < n && d < kMaxDevices; d++) {
    DeviceCtx &c = g_ctx[d];
    if (!c.ready && !c.tables_ready) continue;
    cudaSetDevice(d);
    for (int i = 0; i < kPipeDepth; i++) {
      if (c.streams[i]) cudaStreamDestroy(c.streams[i]);
      if (c.ev_start[i]) cudaEventDestroy(c.ev_start[i]);
      if (c.ev_stop[i]) cudaEventDestroy(c.ev_stop[i]);
      if (c.in_buf[i]) cudaFree(c.in_buf[i]);
      if (c.out_buf[i]) cudaFree(c.out_buf[i]);
      c.streams[i] = nullptr; c.ev_start[i] = c.ev_stop[i] = nullptr;
      c.in_buf[i] = c.out_buf[i] = nullptr; c.in_cap[i] = c.out_cap[i] = 0;
    }
    for (int i = 0; i <= kPipeDepth; i++) bc7_free_workspace(c.bc7ws[i]);
    for (int i = 0; i < kStagePieces; i++) {
      if (c.pin_in[i]) cudaFreeHost(c.pin_in[i]);
      if (c.pin_in_ev[i]) cudaEventDestroy(c.pin_in_ev[i]);
      c.pin_in[i] = nullptr; c.pin_in_ev[i] = nullptr; c.pin_in_used[i] = false;
    }
    for (int i = 0; i < kPipeDepth; i++) {
      if (c.pin_out[i]) cudaFreeHost(c.pin_out[i]);
      c.pin_out[i] = nullptr; c.pin_out_cap[i] = 0; c.pend_dst[i] = nullptr; c.pend_bytes[i] = 0;
      c.slot_busy[i] = false;
    }
    c.next_chunk = 0; c.pin_in_next = 0;
    if (c.psnr_sum) cudaFree(c.psnr_sum);
    if (c.psnr_host) cudaFreeHost(c.psnr_host);
    c.psnr_sum = c.psnr_host = nullptr;
    c.ready = false;
  }
  g_num_init = 0;
}

uint32_t fastc_gpu_block_bytes(int format) {
  return (format == FASTC_GPU_DXT1 || format == FASTC_GPU_ETC1) ? 8u : 16u;
}

uint64_t fastc_gpu_compressed_size(int format, uint32_t width, uint32_t height) {
  return (uint64_t)((width + 3) / 4) * ((height + 3) / 4) * fastc_gpu_block_bytes(format);
}

int fastc_gpu_compress_device(int format, const void *rgba_dev, uint32_t width, uint32_t height,
                              uint32_t first_block, uint32_t num_blocks, void *out_dev, int quality,
                              uint64_t seed, uint32_t wm_base, uint32_t block_index_base, void *cuda_stream,
                              uint32_t *launches_out) {
  if (check_dims(format, width, height)) return 1;
  const uint32_t total = (width / 4) * (height / 4);
  if (first_block > total) return fail("first_block %u beyond the image's %u blocks", first_block, total);
  if (num_blocks == 0) num_blocks = total - first_block;
  if (first_block + num_blocks > total) return fail("block range exceeds the image");
  if (!rgba_dev || !out_dev) return fail("null device pointer");
  if (quality < 0) return fail("quality must be >= 0");
  int dev = 0;
  CU_TRY(cudaGetDevice(&dev));
  if (dev >= kMaxDevices) return fail("device index %d not supported", dev);
  if (ensure_tables(dev)) return 1;
  uint32_t n = 0;
  if (enqueue(dev, kPipeDepth, format, rgba_dev, width, height, first_block, num_blocks, out_dev, quality, seed, wm_base,
              block_index_base, static_cast<cudaStream_t>(cuda_stream), &n))
    return 1;
  if (launches_out) *launches_out = n;
  return 0;
}

int fastc_gpu_count_solid_device(const void *rgba_dev, uint32_t width, uint32_t height, uint32_t first_block,
                                 uint32_t num_blocks, void *cuda_stream, uint32_t *count_out) {
  if (check_dims(FASTC_GPU_BPTC, width, height)) return 1;
  int dev = 0;
  CU_TRY(cudaGetDevice(&dev));
  if (ensure_tables(dev)) return 1;
  DeviceCtx &c = g_ctx[dev];
  std::lock_guard<std::mutex> lk(c.mu);
  CU_TRY(bc7_count_solid(c.bc7ws[kPipeDepth], rgba_dev, width, first_block, num_blocks, static_cast<cudaStream_t>(cuda_stream),
                         count_out));
  return 0;
}

int fastc_gpu_compress(int format, const uint8_t *rgba_host, uint32_t width, uint32_t height,
                       uint32_t first_block, uint32_t num_blocks, uint8_t *out_host, int quality, uint64_t seed,
                       uint32_t chunk_blocks, int num_gpus, fastc_gpu_timing *timing) {
  auto t0 = std::chrono::steady_clock::now();
  if (check_dims(format, width, height)) return 1;
  if (!rgba_host || !out_host) return fail("null host pointer");
  if (quality < 0) return fail("quality must be >= 0");
  const uint32_t bx = width / 4, total = bx * (height / 4);
  if (first_block > total) return fail("first_block %u beyond the image's %u blocks", first_block, total);
  if (num_blocks == 0) num_blocks = total - first_block;
  if (first_block + num_blocks > total) return fail("block range exceeds the image");
  int ndev = device_count();
  if (ndev <= 0) return fail("no CUDA device available (there is no CPU fallback)");
  if (num_gpus <= 0) num_gpus = g_num_init > 0 ? g_num_init : 1;
  num_gpus = std::min(num_gpus, ndev);
  int prev = 0;
  cudaGetDevice(&prev);

  // Contiguous block-row slabs, one per GPU.
  const uint32_t row0 = first_block / bx, row1 = (first_block + num_blocks + bx - 1) / bx;
  const uint32_t rows = row1 - row0;
  num_gpus = std::max(1, std::min<int>(num_gpus, rows));
  std::vector<Shard> shards(num_gpus);
  for (int g = 0; g < num_gpus; g++) {
    uint32_t a = row0 + (uint32_t)((uint64_t)rows * g / num_gpus);
    uint32_t b = row0 + (uint32_t)((uint64_t)rows * (g + 1) / num_gpus);
    uint32_t lo = std::max(first_block, a * bx), hi = std::min(first_block + num_blocks, b * bx);
    shards[g].dev = g;
    shards[g].first_block = lo;
    shards[g].num_blocks = hi > lo ? hi - lo : 0;
  }
  // BC7 watermark order: the word index of a solid block is the number of solid blocks
  // before it in raster order over the WHOLE image (the reference's single-threaded
  // process-global counter, Compressor.cpp:135-140,1457).  Counting on the host keeps a
  // sub-range / sharded submission bit-identical to the same bytes of a full submission.
  if (format == FASTC_GPU_BPTC) {
    // the blocks before the submission and every shard but the last, counted concurrently (the
    // scan is memory-bound: one host thread per slab instead of one slab after the other)
    std::vector<uint32_t> counts(num_gpus, 0);
    uint32_t before = 0;
    {
      std::vector<std::thread> th;
      for (int g = 0; g + 1 < num_gpus; g++)
        th.emplace_back([&, g] {
          counts[g] = host_count_solid(rgba_host, width, shards[g].first_block, shards[g].first_block + shards[g].num_blocks);
        });
      before = host_count_solid(rgba_host, width, 0, first_block);
      for (auto &t : th) t.join();
    }
    uint32_t run = before;
    for (int g = 0; g < num_gpus; g++) {
      shards[g].wm_base = run;
      run += counts[g];
    }
  }
  if (num_gpus == 1) cudaGetDevice(&shards[0].dev);  // one GPU: the caller's current device

  if (num_gpus == 1) {
    shards[0].rc = run_shard_locked(shards[0], format, rgba_host, width, height, out_host, quality, seed, chunk_blocks);
    if (shards[0].rc) snprintf(shards[0].err, sizeof(shards[0].err), "%s", tl_error);
  } else {
    std::vector<std::thread> th;
    for (int g = 0; g < num_gpus; g++)
      th.emplace_back([&, g] {
        if (shards[g].num_blocks == 0) return;
        shards[g].rc = run_shard_locked(shards[g], format, rgba_host, width, height, out_host, quality, seed, chunk_blocks);
        if (shards[g].rc) snprintf(shards[g].err, sizeof(shards[g].err), "%s", tl_error);
      });
    for (auto &t : th) t.join();
  }
  cudaSetDevice(prev);
  fastc_gpu_timing tm = {};
  for (auto &s : shards) {
    if (s.rc) return fail("GPU %d: %s", s.dev, s.err);
    tm.kernel_ms = std::max(tm.kernel_ms, s.kernel_ms);
    tm.kernel_launches += s.launches;
    tm.h2d_bytes += s.h2d;
    tm.d2h_bytes += s.d2h;
  }
  tm.total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  if (timing) *timing = tm;
  return 0;
}

int fastc_gpu_compress_batch(int format, const fastc_gpu_job *jobs, uint32_t num_jobs, int quality, uint64_t seed,
                             int num_gpus, fastc_gpu_timing *timing) {
  auto t0 = std::chrono::steady_clock::now();
  if (!jobs && num_jobs) return fail("null job list");
  int ndev = device_count();
  if (ndev <= 0) return fail("no CUDA device available (there is no CPU fallback)");
  if (num_gpus <= 0) num_gpus = g_num_init > 0 ? g_num_init : 1;
  num_gpus = std::max(1, std::min(num_gpus, ndev));
  for (uint32_t j = 0; j < num_jobs; j++) {
    if (check_dims(format, jobs[j].width, jobs[j].height)) return 1;
    if (!jobs[j].rgba_host || !jobs[j].out_host) return fail("job %u has a null pointer", j);
  }
  int prev = 0;
  cudaGetDevice(&prev);
  // Whole textures are dealt round-robin to the GPUs; each job
  // is an independent compression (its own watermark sequence), like one
  // CompressImageData call per texture in the reference.
  std::vector<fastc_gpu_timing> per(num_gpus);
  std::vector<int> rcs(num_gpus, 0);
  std::vector<std::string> errs(num_gpus);
  auto worker = [&](int g) {
    int last_dev = -1, dev = g;
    if (num_gpus == 1) cudaGetDevice(&dev);
    if (dev < 0 || dev >= kMaxDevices) { rcs[g] = 1; errs[g] = "bad device"; return; }
    std::lock_guard<std::mutex> lk(g_ctx[dev].host_mu);
    for (uint32_t j = g; j < num_jobs; j += num_gpus) {
      Shard s;
      s.dev = dev;
      s.first_block = 0;
      s.num_blocks = (jobs[j].width / 4) * (jobs[j].height / 4);
      // no drain between textures: the staging slots keep rotating across jobs
      if (run_shard(s, format, jobs[j].rgba_host, jobs[j].width, jobs[j].height, jobs[j].out_host, quality,
                    seed + ((uint64_t)j << 40), 0, /*drain=*/false)) {
        drain_slots(g_ctx[s.dev], nullptr);
        rcs[g] = 1;
        errs[g] = tl_error;
        return;
      }
      per[g].kernel_ms += s.kernel_ms;
      per[g].kernel_launches += s.launches;
      per[g].h2d_bytes += s.h2d;
      per[g].d2h_bytes += s.d2h;
      last_dev = s.dev;
    }
    if (last_dev >= 0) {
      double ms = 0;
      if (drain_slots(g_ctx[last_dev], &ms)) {
        rcs[g] = 1;
        errs[g] = tl_error;
        return;
      }
      per[g].kernel_ms += ms;
    }
  };
  if (num_gpus == 1) {
    worker(0);
  } else {
    std::vector<std::thread> th;
    for (int g = 0; g < num_gpus; g++) th.emplace_back(worker, g);
    for (auto &t : th) t.join();
  }
  cudaSetDevice(prev);
  fastc_gpu_timing tm = {};
  for (int g = 0; g < num_gpus; g++) {
    if (rcs[g]) return fail("GPU %d: %s", g, errs[g].c_str());
    tm.kernel_ms = std::max(tm.kernel_ms, per[g].kernel_ms);
    tm.kernel_launches += per[g].kernel_launches;
    tm.h2d_bytes += per[g].h2d_bytes;
    tm.d2h_bytes += per[g].d2h_bytes;
  }
  tm.total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  if (timing) *timing = tm;
  return 0;
}

int fastc_gpu_decompress_device(int format, const void *cmp_dev, uint32_t width, uint32_t height, void *rgba_out_dev,
                                void *cuda_stream) {
  if (check_dims(format, width, height)) return 1;
  if (!cmp_dev || !rgba_out_dev) return fail("null device pointer");
  int dev = 0;
  CU_TRY(cudaGetDevice(&dev));
  if (dev >= kMaxDevices) return fail("device index %d not supported", dev);
  if (ensure_tables(dev)) return 1;
  CU_TRY(launch_decode(format, cmp_dev, width, 0, (width / 4) * (height / 4), rgba_out_dev,
                       static_cast<cudaStream_t>(cuda_stream)));
  return 0;
}

int fastc_gpu_decompress(int format, const uint8_t *cmp_host, uint32_t width, uint32_t height, uint8_t *rgba_out_host,
                         fastc_gpu_timing *timing) {
  auto t0 = std::chrono::steady_clock::now();
  if (check_dims(format, width, height)) return 1;
  if (!cmp_host || !rgba_out_host) return fail("null host pointer");
  int dev = 0;
  CU_TRY(cudaGetDevice(&dev));
  if (dev >= kMaxDevices) return fail("device index %d not supported", dev);
  if (ensure_ctx(dev)) return 1;
  DeviceCtx &c = g_ctx[dev];
  const size_t cmp_bytes = fastc_gpu_compressed_size(format, width, height);
  const size_t out_bytes = (size_t)width * height * 4;
  std::lock_guard<std::mutex> hl(c.host_mu);  // shares the staging slots with the compress path
  cudaStream_t st = c.streams[0];
  // staging slot 0: the image buffer holds the decoded pixels, the output buffer the blocks
  CU_TRY(cudaStreamSynchronize(st));
  if (grow(&c.in_buf[0], &c.in_cap[0], out_bytes)) return 1;
  if (grow(&c.out_buf[0], &c.out_cap[0], cmp_bytes)) return 1;
  CU_TRY(cudaMemcpyAsync(c.out_buf[0], cmp_host, cmp_bytes, cudaMemcpyHostToDevice, st));
  CU_TRY(cudaEventRecord(c.ev_start[0], st));
  CU_TRY(launch_decode(format, c.out_buf[0], width, 0, (width / 4) * (height / 4), c.in_buf[0], st));
  CU_TRY(cudaEventRecord(c.ev_stop[0], st));
  CU_TRY(cudaMemcpyAsync(rgba_out_host, c.in_buf[0], out_bytes, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  if (timing) {
    float ms = 0;
    CU_TRY(cudaEventElapsedTime(&ms, c.ev_start[0], c.ev_stop[0]));
    *timing = fastc_gpu_timing{};
    timing->kernel_ms = ms;
    timing->h2d_bytes = cmp_bytes;
    timing->d2h_bytes = out_bytes;
    timing->kernel_launches = 1;
    timing->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  }
  return 0;
}

int fastc_gpu_psnr_device(const void *a_dev, const void *b_dev, uint32_t width, uint32_t height, void *cuda_stream,
                          double *psnr_out) {
  if (!a_dev || !b_dev || !psnr_out) return fail("null pointer");
  if (width == 0 || height == 0) return fail("empty image");
  int dev = 0;
  CU_TRY(cudaGetDevice(&dev));
  if (dev >= kMaxDevices) return fail("device index %d not supported", dev);
  DeviceCtx &c = g_ctx[dev];
  std::lock_guard<std::mutex> lk(c.mu);
  if (!c.psnr_sum) {
    CU_TRY(cudaMalloc(reinterpret_cast<void **>(&c.psnr_sum), 64));
    CU_TRY(cudaMallocHost(reinterpret_cast<void **>(&c.psnr_host), 64));
  }
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const size_t n = (size_t)width * height;
  CU_TRY(launch_psnr_sum(a_dev, b_dev, n, c.psnr_sum, st));
  CU_TRY(cudaMemcpyAsync(c.psnr_host, c.psnr_sum, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  *psnr_out = psnr_from_sum(*c.psnr_host, n);
  return 0;
}

int fastc_gpu_psnr(const uint8_t *a_host, const uint8_t *b_host, uint32_t width, uint32_t height, double *psnr_out) {
  if (!a_host || !b_host || !psnr_out) return fail("null pointer");
  if (width == 0 || height == 0) return fail("empty image");
  int dev = 0;
  CU_TRY(cudaGetDevice(&dev));
  if (dev >= kMaxDevices) return fail("device index %d not supported", dev);
  if (ensure_ctx(dev)) return 1;
  DeviceCtx &c = g_ctx[dev];
  const size_t bytes = (size_t)width * height * 4;
  std::lock_guard<std::mutex> hl(c.host_mu);  // shares the staging slots with the compress path
  cudaStream_t st = c.streams[0];
  CU_TRY(cudaStreamSynchronize(st));
  if (grow(&c.in_buf[0], &c.in_cap[0], bytes)) return 1;
  if (grow(&c.in_buf[1], &c.in_cap[1], bytes)) return 1;
  CU_TRY(cudaStreamSynchronize(c.streams[1]));
  CU_TRY(cudaMemcpyAsync(c.in_buf[0], a_host, bytes, cudaMemcpyHostToDevice, st));
  CU_TRY(cudaMemcpyAsync(c.in_buf[1], b_host, bytes, cudaMemcpyHostToDevice, st));
  return fastc_gpu_psnr_device(c.in_buf[0], c.in_buf[1], width, height, st, psnr_out);
}

int fastc_gpu_bc7_counters(uint64_t *qe_calls, uint64_t *pixel_bucket_evals) {
  int dev = 0;
  CU_TRY(cudaGetDevice(&dev));
  DeviceCtx &c = g_ctx[dev];
  std::lock_guard<std::mutex> lk(c.mu);
  CU_TRY(bc7_read_counters(c.bc7ws[kPipeDepth], qe_calls, pixel_bucket_evals));
  return 0;
}

int fastc_gpu_bc7_stage_ms(int enable, double *ms6) {
  int dev = 0;
  CU_TRY(cudaGetDevice(&dev));
  DeviceCtx &c = g_ctx[dev];
  std::lock_guard<std::mutex> lk(c.mu);
  CU_TRY(bc7_stage_timing(c.bc7ws[kPipeDepth], enable, ms6));
  return 0;
}

int fastc_gpu_debug_bc7_dump(uint32_t nblocks, uint32_t *sel_out, uint32_t *results_out) {
  int dev = 0;
  CU_TRY(cudaGetDevice(&dev));
  DeviceCtx &c = g_ctx[dev];
  std::lock_guard<std::mutex> lk(c.mu);
  CU_TRY(bc7_debug_dump(c.bc7ws[kPipeDepth], nblocks, sel_out, results_out));
  return 0;
}

const char *fastc_gpu_last_error(void) { return tl_error; }

}  // extern "C"
