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
#include <atomic>
#include <deque>
#include <chrono>
#include <condition_variable>
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
constexpr int kStagePieces = 16;              // pinned upload ring (pageable inputs)
constexpr size_t kStagePieceBytes = 4u << 20;
constexpr int kStageWindow = 8;               // pieces being filled at once, one host thread each

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
  std::atomic<int> out_left[kPipeDepth] = {};  // parts of a slot's copy-out still running on the pool
  // batch submissions from pageable memory: whole textures are staged by threads running ahead of the
  // submitting thread (compress_batch_impl)
  static constexpr int kBatchRing = 16;
  void *batch_pin[kBatchRing] = {};
  size_t batch_pin_cap[kBatchRing] = {};
  unsigned long long *psnr_sum = nullptr, *psnr_host = nullptr;  // device / pinned accumulators of fastc_gpu_psnr*
  void *stats_buf[kPipeDepth] = {};  // BPTC per-block statistics of the chunk in the slot (only when asked for)
  size_t stats_cap[kPipeDepth] = {};
  // BC7 host path: the chunk in a slot is packed once the watermark base is known (complete_piece)
  cudaEvent_t ev_count[kPipeDepth] = {};
  struct Pending {
    bool active = false;
    int piece = 0;                 // index in the submission's WmChain
    uint32_t lo = 0, nblk = 0, height = 0;
    uint8_t *dst = nullptr;        // caller memory of the chunk's first encoded block
    fastc_gpu_bptc_block_stat *stats = nullptr;  // caller's record of the chunk's first encoded block (or NULL)
  } pending[kPipeDepth];
  // The device API's BC7 scratch (bc7ws[kPipeDepth]) is shared by every caller stream of the device:
  // each use waits for the previous one (api_done) before it touches the scratch
  cudaEvent_t api_done = nullptr;
  bool api_used = false;
  // PVRTC scratch (one per device, uses ordered by pvr_done like the BC7 device-API scratch)
  PvrtcWorkspace pvrws;
  cudaEvent_t pvr_done = nullptr;
  bool pvr_used = false;
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
    CU_TRY(cudaEventCreateWithFlags(&c.ev_count[i], cudaEventDisableTiming));
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

bool valid_format(int f) { return f >= FASTC_GPU_DXT1 && f <= FASTC_GPU_PVRTC4; }

int check_dims(int format, uint32_t width, uint32_t height) {
  if (!valid_format(format)) return fail("unknown compression format %d", format);
  // reference: "Image dimensions must be multiples of the block size" (TexComp.cpp:472-476)
  if (width == 0 || height == 0 || (width & 3) || (height & 3))
    return fail("image dimensions %ux%u are not non-zero multiples of the 4x4 block", width, height);
  // reference: "PVRTC4 images must be square and power-of-two" (TexComp.cpp:477-482)
  if (format == FASTC_GPU_PVRTC4 && (width != height || (width & (width - 1)) || width < 8 || width > 16384))
    return fail("PVRTC4 images must be square with a power-of-two side in [8, 16384] (got %ux%u)", width, height);
  return 0;
}

// Device API: enqueue every kernel needed to encode [first_block, first_block+num_blocks) on
// `stream` of the current device.
int enqueue(int dev, int format, const void *rgba_dev, uint32_t width, uint32_t height, uint32_t first_block,
            uint32_t num_blocks, void *out_dev, const EncodeParams &prm, uint32_t wm_base, uint32_t block_index_base,
            cudaStream_t stream, uint32_t *launches) {
  uint32_t n = 0;
  switch (format) {
    case FASTC_GPU_DXT1:
    case FASTC_GPU_DXT5:
      CU_TRY(launch_dxt(format == FASTC_GPU_DXT5, rgba_dev, width, first_block, num_blocks, out_dev, stream));
      n = num_blocks ? 1 : 0;
      break;
    case FASTC_GPU_ETC1:
      CU_TRY(launch_etc1(rgba_dev, width, first_block, num_blocks, out_dev, prm.etc1_quality, stream));
      n = num_blocks ? 1 : 0;
      break;
    case FASTC_GPU_PVRTC4: {
      if (first_block != 0 || num_blocks != (width / 4) * (height / 4))
        return fail("PVRTC4 encodes whole textures: block ranges are not supported");
      DeviceCtx &c = g_ctx[dev];
      std::lock_guard<std::mutex> lk(c.mu);
      // one scratch per device: order this use after the previous one, whatever its stream
      if (!c.pvr_done) CU_TRY(cudaEventCreateWithFlags(&c.pvr_done, cudaEventDisableTiming));
      if (c.pvr_used) CU_TRY(cudaStreamWaitEvent(stream, c.pvr_done, 0));
      CU_TRY(launch_pvrtc(c.pvrws, rgba_dev, width, height, 1, out_dev, stream, &n));
      CU_TRY(cudaEventRecord(c.pvr_done, stream));
      c.pvr_used = true;
      break;
    }
    case FASTC_GPU_BPTC: {
      DeviceCtx &c = g_ctx[dev];
      std::lock_guard<std::mutex> lk(c.mu);
      // one scratch for every caller stream of the device: order this use after the previous one
      if (!c.api_done) CU_TRY(cudaEventCreateWithFlags(&c.api_done, cudaEventDisableTiming));
      if (c.api_used) CU_TRY(cudaStreamWaitEvent(stream, c.api_done, 0));
      CU_TRY(launch_bc7(c.bc7ws[kPipeDepth], rgba_dev, width, height, first_block, num_blocks, out_dev, prm, wm_base,
                        block_index_base, stream, &n));
      CU_TRY(cudaEventRecord(c.api_done, stream));
      c.api_used = true;
      break;
    }
  }
  if (launches) *launches += n;
  return 0;
}

// Solid-colour blocks of block row `row`, columns [c0, c1) of a host image (64 B compare per block).
uint32_t host_count_solid_row(const uint32_t *img, uint32_t width, uint32_t row, uint32_t c0, uint32_t c1) {
  uint32_t cnt = 0;
  const uint32_t *base = img + (size_t)row * 4 * width;
  for (uint32_t b = c0; b < c1; b++) {
    const uint32_t *p = base + (size_t)b * 4;
    const uint32_t v = p[0];
    bool same = true;
    for (int j = 0; j < 4 && same; j++)
      for (int i = 0; i < 4; i++)
        if (p[(size_t)j * width + i] != v) { same = false; break; }
    cnt += same;
  }
  return cnt;
}

// BC7 watermark order (Compressor.cpp:135-140,1457): the word a solid block carries is the number
// of solid blocks before it in raster order over the WHOLE image, so a sub-range submission
// (first_block > 0: one CompressionFunc call of a ThreadGroup-style split) needs the solid count of
// the blocks before it, which never reach the GPU.  Those are counted on the host, once: the
// per-block-row cumulative counts of the last image are kept, so T ranged calls on one image scan
// it once in total instead of O(T^2 / 2) times.  The entry is keyed by (pointer, size) and checked
// against a fingerprint of sampled pixels of the scanned rows; the watermark is a signature in
// otherwise unused index bits and never changes a decoded pixel.
struct PrefixCache {
  std::mutex mu;
  const uint8_t *ptr = nullptr;
  uint32_t width = 0, height = 0;
  std::vector<uint32_t> cum;  // cum[r] = solid blocks in block rows [0, r); rows scanned = cum.size() - 1
  uint64_t fingerprint = 0;
} g_prefix;

uint64_t sample_fingerprint(const uint32_t *img, uint32_t width, uint32_t rows_scanned) {
  const uint64_t npix = (uint64_t)rows_scanned * 4 * width;
  uint64_t h = 0x9E3779B97F4A7C15ull ^ npix, x = 0x2545F4914F6CDD1Dull;
  for (int k = 0; k < 2048 && npix; k++) {
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;  // xorshift64
    h = (h ^ img[x % npix]) * 0x100000001B3ull;
  }
  return h;
}

uint32_t host_prefix_solid(const uint8_t *rgba_host, uint32_t width, uint32_t height, uint32_t first_block) {
  if (first_block == 0) return 0;
  const uint32_t bx = width / 4, row = first_block / bx, rem = first_block % bx;
  const uint32_t *img = reinterpret_cast<const uint32_t *>(rgba_host);
  std::lock_guard<std::mutex> lk(g_prefix.mu);
  PrefixCache &pc = g_prefix;
  const bool hit = pc.ptr == rgba_host && pc.width == width && pc.height == height && !pc.cum.empty() &&
                   sample_fingerprint(img, width, (uint32_t)pc.cum.size() - 1) == pc.fingerprint;
  if (!hit) {
    pc.ptr = rgba_host; pc.width = width; pc.height = height;
    pc.cum.assign(1, 0u);
  }
  const uint32_t done = (uint32_t)pc.cum.size() - 1;
  if (row > done) {  // extend the scan to the rows before `row` (memory-bound: a few host threads)
    const uint32_t n = row - done;
    std::vector<uint32_t> cnt(n);
    const int nt = (int)std::min<uint32_t>(8, std::max<uint32_t>(1, (uint32_t)(((uint64_t)n * bx) >> 16)));
    auto work = [&](int k) {
      for (uint32_t r = (uint32_t)((uint64_t)n * k / nt); r < (uint32_t)((uint64_t)n * (k + 1) / nt); r++)
        cnt[r] = host_count_solid_row(img, width, done + r, 0, bx);
    };
    std::vector<std::thread> th;
    for (int k = 1; k < nt; k++) th.emplace_back(work, k);
    work(0);
    for (auto &t : th) t.join();
    for (uint32_t r = 0; r < n; r++) pc.cum.push_back(pc.cum.back() + cnt[r]);
  }
  if (!hit || row > done) pc.fingerprint = sample_fingerprint(img, width, (uint32_t)pc.cum.size() - 1);
  return pc.cum[row] + (rem ? host_count_solid_row(img, width, row, 0, rem) : 0u);
}

// Watermark bases of one host submission.  Its pieces (the pipeline chunks of every GPU's slab,
// numbered in raster order) are encoded concurrently; piece p packs with
//   base(p) = blocks before the submission + solid blocks of pieces 0 .. p-1,
// each count coming from the piece's own classification pass on its GPU (no host scan of the
// pixels, no GPU waits for another: only the pack, the last and cheapest kernel, needs the base).
struct WmChain {
  std::mutex mu;
  std::condition_variable cv;
  uint32_t before = 0;
  std::vector<uint32_t> counts;
  std::vector<char> known;
  bool failed = false;
  void resize(size_t n) { counts.assign(n, 0); known.assign(n, 0); }
  size_t size() const { return counts.size(); }
  void publish(int p, uint32_t n) {
    { std::lock_guard<std::mutex> lk(mu); counts[p] = n; known[p] = 1; }
    cv.notify_all();
  }
  void fail_all() {
    { std::lock_guard<std::mutex> lk(mu); failed = true; }
    cv.notify_all();
  }
  bool base_of(int p, uint32_t *base) {  // false: a piece it depends on failed
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] {
      if (failed) return true;
      for (int i = 0; i < p; i++) if (!known[i]) return false;
      return true;
    });
    if (failed) return false;
    uint32_t b = before;
    for (int i = 0; i < p; i++) b += counts[i];
    *base = b;
    return true;
  }
};

struct Shard {
  int dev;
  uint32_t first_block, num_blocks;  // within the image
  int piece0 = 0;                    // first piece of the slab in the submission's WmChain
  std::vector<uint32_t> bounds;      // chunk k covers block rows [bounds[k], bounds[k + 1])
  double kernel_ms = 0;
  uint32_t launches = 0;
  uint64_t h2d = 0, d2h = 0;
  int rc = 0;
  char err[512] = "";
};

// The pipeline chunks of a slab, in whole block rows.
void plan_chunks(Shard &s, int format, uint32_t width, uint32_t chunk_blocks) {
  s.bounds.clear();
  if (s.num_blocks == 0) return;
  const uint32_t bx = width / 4;
  const uint32_t row0 = s.first_block / bx;
  const uint32_t row1 = (s.first_block + s.num_blocks + bx - 1) / bx;
  const uint32_t total_rows = row1 - row0;
  uint32_t rows_per_chunk;
  if (format == FASTC_GPU_PVRTC4) {
    rows_per_chunk = total_rows;  // image-level encoder: the whole texture is one submission
  } else if (chunk_blocks == 0 && format == FASTC_GPU_BPTC) {
    // BC7 is compute-bound (copies are ~1% of the time): no chunking unless asked for
    rows_per_chunk = total_rows;
  } else if (chunk_blocks == 0) {
    // auto: ~4 Mi pixels per chunk keeps the copy engines busy without making the pipeline too coarse
    rows_per_chunk = (1u << 18) / bx;
  } else {
    rows_per_chunk = chunk_blocks / bx;
  }
  if (format == FASTC_GPU_BPTC) rows_per_chunk = std::min(rows_per_chunk, bc7_max_submission() / bx);  // scratch bound
  rows_per_chunk = std::max<uint32_t>(1, rows_per_chunk);
  for (uint32_t r = row0; r < row1; r += rows_per_chunk) s.bounds.push_back(r);
  s.bounds.push_back(row1);
  if (chunk_blocks == 0 && format == FASTC_GPU_BPTC && s.bounds.size() == 2 && total_rows >= 64 &&
      (size_t)total_rows * 4 * width * 4 >= ((size_t)192 << 20)) {
    // BC7 auto, very big uploads (>= 192 MiB, ~4 ms on the wire): two halves on two streams.  The
    // second half's upload and shape selection / fits run under the first half's annealing, the
    // first half's download under the second's kernels.  Each half pays the persistent annealing
    // kernel's tail once more, so this only wins when the copies are long: measured
    // (tools/time_e2e_chunks.py, profiles/r02_e2e_chunks.log) 8192^2 (256 MiB): one chunk 211.9 ms,
    // two halves 210.0 ms; 8192 x 4096 (128 MiB, the slab of one of two GPUs): 106.6 vs 115.6 ms.
    s.bounds.assign({row0, row0 + total_rows / 2, row1});
  }
}

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

// Host copies on a few threads (one thread moves ~8 GB/s; PCIe 5 x16 wants ~50).  The threads are a
// process-wide pool created on first use (spawning threads per piece costs as much as the copy
// itself), shared by the per-GPU host threads of a submission.
class CopyPool {
 public:
  static CopyPool &get() {
    static CopyPool *pool = new CopyPool();  // leaked on purpose: its threads may outlive static destructors
    return *pool;
  }
  // One whole copy as one task for one worker (`left` drops by one when it is done).  Copying whole
  // pieces on several threads at once moves ~50 GB/s on the GPU box's host, splitting each piece over
  // threads spawned for it ~23 GB/s (tools/host_copy_bw.cu).
  void submit(void *dst, const void *src, size_t n, std::atomic<int> *left) {
    if (workers_ == 0) {
      run(Task{(uint8_t *)dst, (const uint8_t *)src, n, left});
      return;
    }
    {
      std::lock_guard<std::mutex> lk(mu_);
      tasks_.push_back(Task{(uint8_t *)dst, (const uint8_t *)src, n, left});
    }
    cv_.notify_one();
  }

 private:
  struct Task {
    uint8_t *dst;
    const uint8_t *src;
    size_t n;
    std::atomic<int> *left;
  };
  CopyPool() {
    const unsigned hc = std::thread::hardware_concurrency();
    workers_ = (int)std::min<unsigned>(12, hc > 2 ? hc - 2 : 0);
    for (int k = 0; k < workers_; k++) std::thread([this] { loop(); }).detach();
  }
  static void run(const Task &t) {
    memcpy(t.dst, t.src, t.n);
    t.left->fetch_sub(1, std::memory_order_acq_rel);
  }
  void loop() {
    for (;;) {
      Task t;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [this] { return !tasks_.empty(); });
        t = tasks_.front();
        tasks_.pop_front();
      }
      run(t);
    }
  }
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<Task> tasks_;
  int workers_ = 0;
};

// Host -> device on `st`.  Pinned sources go straight to the copy engine; pageable ones through
// the pinned ring, piece by piece.
int upload(DeviceCtx &c, void *dst_dev, const uint8_t *src, size_t bytes, cudaStream_t st) {
  if (!is_pageable(src)) {
    CU_TRY(cudaMemcpyAsync(dst_dev, src, bytes, cudaMemcpyHostToDevice, st));
    return 0;
  }
  // Pieces of <= 4 MiB go through a ring of pinned buffers: up to kStageWindow pieces are being filled at
  // once, each by one pool thread, and every piece is sent as soon as it and the pieces before it are
  // full -- the host copies run alongside the wire instead of in front of it.  Small uploads (one
  // texture of a batch) are cut finer, so that they too are filled by a window of threads.
  const size_t piece = std::min(kStagePieceBytes, std::max<size_t>((size_t)256 << 10, ((bytes / kStageWindow + 65535) >> 16) << 16));
  const size_t np = (bytes + piece - 1) / piece;
  std::vector<std::atomic<int>> done(np);
  std::vector<int> slot_of(np);
  CopyPool &pool = CopyPool::get();
  auto start_piece = [&](size_t k) -> int {
    const int i = (int)(c.pin_in_next++ % kStagePieces);
    slot_of[k] = i;
    if (!c.pin_in[i]) {
      CU_TRY(cudaHostAlloc(&c.pin_in[i], kStagePieceBytes, cudaHostAllocDefault));
      CU_TRY(cudaEventCreateWithFlags(&c.pin_in_ev[i], cudaEventDisableTiming));
    }
    if (c.pin_in_used[i]) CU_TRY(cudaEventSynchronize(c.pin_in_ev[i]));  // its previous upload has left the buffer
    const size_t off = k * piece;
    done[k].store(1, std::memory_order_relaxed);
    pool.submit(c.pin_in[i], src + off, std::min(piece, bytes - off), &done[k]);
    return 0;
  };
  int rc = 0;
  size_t started = 0;
  while (started < np && started < (size_t)kStageWindow && !rc) {
    rc = start_piece(started);
    if (!rc) started++;
  }
  for (size_t k = 0; k < started; k++) {  // (on an error: only wait for what was started)
    // (without helping: this thread's job is to send every piece the moment it is full)
    while (done[k].load(std::memory_order_acquire) > 0) std::this_thread::yield();
    if (rc) continue;
    const size_t off = k * piece, n = std::min(piece, bytes - off);
    const int i = slot_of[k];
    if (cudaMemcpyAsync((uint8_t *)dst_dev + off, c.pin_in[i], n, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaEventRecord(c.pin_in_ev[i], st) != cudaSuccess) {
      rc = fail("CUDA: %s", cudaGetErrorString(cudaGetLastError()));
      continue;
    }
    c.pin_in_used[i] = true;
    if (started < np) { rc = start_piece(started); started += rc ? 0 : 1; }
  }
  return rc;
}

// The copy-out of a staged download runs on the pool while the submission goes on; it has to be
// over before its pinned source is reused and before the call returns.
void wait_copy_out(DeviceCtx &c, int slot) {
  while (c.out_left[slot].load(std::memory_order_acquire) > 0) std::this_thread::yield();
}

// Device -> host on `st` for the chunk in `slot`.  Pageable destinations receive their bytes
// from the slot's pinned buffer once the chunk has finished (finish_slot).
int download(DeviceCtx &c, int slot, uint8_t *dst, const void *src_dev, size_t bytes, cudaStream_t st) {
  if (!is_pageable(dst)) {
    CU_TRY(cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToHost, st));
    return 0;
  }
  wait_copy_out(c, slot);  // the pinned buffer may still be the source of the previous copy-out
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

// BC7: packs the chunk waiting in `slot` -- publishes its solid count if a later piece needs it,
// waits for the counts of the pieces before it, then enqueues the pack and the download.
int complete_piece(DeviceCtx &c, int slot, uint32_t width, WmChain &chain, Shard &s) {
  DeviceCtx::Pending &p = c.pending[slot];
  if (!p.active) return 0;
  p.active = false;
  cudaStream_t st = c.streams[slot];
  if ((size_t)p.piece + 1 < chain.size()) {
    CU_TRY(cudaEventSynchronize(c.ev_count[slot]));  // recorded right after the chunk's classification
    chain.publish(p.piece, *c.bc7ws[slot].host_count);
  }
  uint32_t base = 0;
  if (!chain.base_of(p.piece, &base)) return fail("an earlier part of the submission failed");
  double *stats_dev = nullptr;
  if (p.stats) {
    static_assert(sizeof(fastc_gpu_bptc_block_stat) == 10 * sizeof(double), "record layout");
    if (grow(&c.stats_buf[slot], &c.stats_cap[slot], (size_t)p.nblk * sizeof(fastc_gpu_bptc_block_stat))) return 1;
    stats_dev = static_cast<double *>(c.stats_buf[slot]);
  }
  CU_TRY(bc7_back(c.bc7ws[slot], c.in_buf[slot], width, p.lo, p.nblk, c.out_buf[slot], base, false, st, &s.launches,
                  stats_dev));
  CU_TRY(cudaEventRecord(c.ev_stop[slot], st));
  if (p.stats)  // (an analysis path: a plain copy into the caller's memory, pinned or not)
    CU_TRY(cudaMemcpyAsync(p.stats, stats_dev, (size_t)p.nblk * sizeof(fastc_gpu_bptc_block_stat), cudaMemcpyDeviceToHost, st));
  const size_t bytes = (size_t)p.nblk * 16;
  if (download(c, slot, p.dst, (uint8_t *)c.out_buf[slot] + (size_t)p.lo * 16, bytes, st)) return 1;
  s.d2h += bytes;
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
    // pinned -> caller memory on the pool, ~1 MiB per task; waited for in download() / drain_slots()
    const size_t n = c.pend_bytes[slot], parts = std::max<size_t>(1, std::min<size_t>(8, n >> 20));
    c.out_left[slot].store((int)parts, std::memory_order_release);
    for (size_t k = 0; k < parts; k++) {
      const size_t a = n * k / parts, b = n * (k + 1) / parts;
      CopyPool::get().submit(c.pend_dst[slot] + a, (const uint8_t *)c.pin_out[slot] + a, b - a, &c.out_left[slot]);
    }
    c.pend_dst[slot] = nullptr;
    c.pend_bytes[slot] = 0;
    static const bool sync_out = getenv("FASTC_GPU_SYNC_COPYOUT") != nullptr;  // (measurement switch)
    if (sync_out) wait_copy_out(c, slot);
  }
  c.slot_busy[slot] = false;
  return 0;
}

// Waits for every chunk still in flight on the device's staging slots.
int drain_slots(DeviceCtx &c, double *kernel_ms) {
  int rc = 0;
  for (int slot = 0; slot < kPipeDepth; slot++)
    if (finish_slot(c, slot, kernel_ms)) { rc = 1; break; }
  for (int slot = 0; slot < kPipeDepth; slot++) wait_copy_out(c, slot);  // the caller's buffers are complete on return
  return rc;
}

// Error path: nothing of a failed submission may linger in the slots (a later submission would
// otherwise copy stale bytes into this caller's, possibly freed, buffer).
void abort_slots(DeviceCtx &c) {
  for (int slot = 0; slot < kPipeDepth; slot++) {
    if (c.streams[slot]) cudaStreamSynchronize(c.streams[slot]);
    while (c.out_left[slot].load(std::memory_order_acquire) > 0) std::this_thread::yield();  // copy-outs under way finish
    c.slot_busy[slot] = false;
    c.pending[slot].active = false;
    c.pend_dst[slot] = nullptr;
    c.pend_bytes[slot] = 0;
  }
  cudaGetLastError();
}

// drain = false leaves the last chunks in flight (batch submissions: the caller drains once at
// the end, so consecutive textures overlap).
int run_shard_impl(Shard &s, int format, const uint8_t *rgba_host, uint32_t width, uint32_t height,
                   uint8_t *out_host, const EncodeParams &prm, WmChain &chain, bool drain,
                   fastc_gpu_bptc_block_stat *stats) {
  (void)height;
  DeviceCtx &c = g_ctx[s.dev];
  const uint32_t bx = width / 4;
  const uint32_t bsz = fastc_gpu_block_bytes(format);
  const uint32_t nchunks = s.bounds.empty() ? 0 : (uint32_t)s.bounds.size() - 1;
  if (nchunks == 0) return 0;
  const uint32_t row0 = s.bounds.front(), row1 = s.bounds.back();
  const bool bc7 = format == FASTC_GPU_BPTC;

  // nothing in flight (every submission except the later textures of a batch): restart the slot
  // rotation, so that the same chunk lands in the same slot call after call and its staging
  // buffers / BC7 workspace are grown once
  bool idle = true;
  for (int i = 0; i < kPipeDepth; i++) idle = idle && !c.slot_busy[i];
  if (idle) c.next_chunk = 0;
  int slots[kPipeDepth];  // slots of the chunks whose pack is still owed, oldest first
  int nowed = 0;
  for (uint32_t k = 0; k < nchunks; k++) {
    const int slot = (int)(c.next_chunk++ % kPipeDepth);
    cudaStream_t st = c.streams[slot];
    const uint32_t r0 = s.bounds[k], r1 = s.bounds[k + 1];
    const size_t in_bytes = (size_t)(r1 - r0) * 4 * width * 4;
    const size_t out_bytes = (size_t)(r1 - r0) * bx * bsz;
    // slot reuse: the slot's previous chunk must be packed and finished (its timing booked, a
    // staged download handed over) before its buffers are overwritten or regrown
    if (nowed == kPipeDepth) {  // the oldest owed pack sits in this very slot
      if (complete_piece(c, slots[0], width, chain, s)) return 1;
      for (int i = 1; i < nowed; i++) slots[i - 1] = slots[i];
      nowed--;
    }
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
    uint8_t *dst = out_host + ((size_t)r0 * bx + lo) * bsz;
    CU_TRY(cudaEventRecord(c.ev_start[slot], st));
    c.slot_busy[slot] = true;
    if (bc7) {
      // everything but the pack; the chunk's solid count travels to the host right after its
      // classification when a later piece of the submission needs it
      const int piece = s.piece0 + (int)k;
      const bool later = (size_t)piece + 1 < chain.size();
      CU_TRY(bc7_front(c.bc7ws[slot], c.in_buf[slot], width, lo, hi - lo, prm, r0 * bx, st,
                       later ? c.ev_count[slot] : nullptr, &s.launches));
      DeviceCtx::Pending &p = c.pending[slot];
      p.active = true; p.piece = piece; p.lo = lo; p.nblk = hi - lo; p.height = (r1 - r0) * 4; p.dst = dst;
      p.stats = stats ? stats + ((size_t)r0 * bx + lo) : nullptr;
      slots[nowed++] = slot;
      CU_TRY(cudaEventRecord(c.ev_stop[slot], st));  // (re-recorded after the pack)
    } else {
      EncodeParams none = prm;
      if (enqueue(s.dev, format, c.in_buf[slot], width, (r1 - r0) * 4, lo, hi - lo, c.out_buf[slot], none, 0, 0, st,
                  &s.launches))
        return 1;
      CU_TRY(cudaEventRecord(c.ev_stop[slot], st));
      if (download(c, slot, dst, (uint8_t *)c.out_buf[slot] + (size_t)lo * bsz, (size_t)(hi - lo) * bsz, st)) return 1;
      s.d2h += (size_t)(hi - lo) * bsz;
    }
  }
  for (int i = 0; i < nowed; i++)
    if (complete_piece(c, slots[i], width, chain, s)) return 1;
  if (drain && drain_slots(c, &s.kernel_ms)) return 1;
  return 0;
}

int run_shard(Shard &s, int format, const uint8_t *rgba_host, uint32_t width, uint32_t height, uint8_t *out_host,
              const EncodeParams &prm, WmChain &chain, bool drain = true, fastc_gpu_bptc_block_stat *stats = nullptr) {
  if (ensure_ctx(s.dev)) {
    chain.fail_all();
    return 1;
  }
  const int rc = run_shard_impl(s, format, rgba_host, width, height, out_host, prm, chain, drain, stats);
  if (rc) {
    chain.fail_all();  // pieces of other GPUs that wait for this slab's counts give up
    abort_slots(g_ctx[s.dev]);
  }
  return rc;
}

int run_shard_locked(Shard &s, int format, const uint8_t *rgba_host, uint32_t width, uint32_t height,
                     uint8_t *out_host, const EncodeParams &prm, WmChain &chain, fastc_gpu_bptc_block_stat *stats) {
  if (s.dev < 0 || s.dev >= kMaxDevices) {
    chain.fail_all();
    return fail("bad device %d", s.dev);
  }
  std::lock_guard<std::mutex> lk(g_ctx[s.dev].host_mu);
  return run_shard(s, format, rgba_host, width, height, out_host, prm, chain, true, stats);
}

// Block range checks shared by the entry points (no 32-bit wrap).
int check_range(uint32_t width, uint32_t height, uint32_t first_block, uint32_t *num_blocks) {
  const uint32_t total = (width / 4) * (height / 4);
  if (first_block > total) return fail("first_block %u beyond the image's %u blocks", first_block, total);
  if (*num_blocks == 0) *num_blocks = total - first_block;
  if (*num_blocks > total - first_block) return fail("block range exceeds the image");
  return 0;
}

int current_device(int *dev) {
  CU_TRY(cudaGetDevice(dev));
  if (*dev < 0 || *dev >= kMaxDevices) return fail("device index %d not supported (at most %d devices)", *dev, kMaxDevices);
  return 0;
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
      if (c.stats_buf[i]) cudaFree(c.stats_buf[i]);
      c.stats_buf[i] = nullptr; c.stats_cap[i] = 0;
      c.streams[i] = nullptr; c.ev_start[i] = c.ev_stop[i] = nullptr;
      c.in_buf[i] = c.out_buf[i] = nullptr; c.in_cap[i] = c.out_cap[i] = 0;
    }
    for (int i = 0; i <= kPipeDepth; i++) bc7_free_workspace(c.bc7ws[i]);
    pvrtc_free_workspace(c.pvrws);
    for (int i = 0; i < DeviceCtx::kBatchRing; i++) {
      if (c.batch_pin[i]) cudaFreeHost(c.batch_pin[i]);
      c.batch_pin[i] = nullptr; c.batch_pin_cap[i] = 0;
    }
    if (c.pvr_done) cudaEventDestroy(c.pvr_done);
    c.pvr_done = nullptr; c.pvr_used = false;
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
  return (format == FASTC_GPU_DXT1 || format == FASTC_GPU_ETC1 || format == FASTC_GPU_PVRTC4) ? 8u : 16u;
}

uint64_t fastc_gpu_compressed_size(int format, uint32_t width, uint32_t height) {
  return (uint64_t)((width + 3) / 4) * ((height + 3) / 4) * fastc_gpu_block_bytes(format);
}

namespace {

// The settings one submission runs with: the per-call arguments plus the thread's BPTC options.
EncodeParams make_params(int quality, uint64_t seed, const fastc_gpu_options *opt) {
  EncodeParams prm;
  prm.quality = quality;
  prm.seed = seed;
  if (opt) {
    prm.block_modes = opt->bptc_block_modes & 0xFFu;
    prm.error_metric = opt->bptc_error_metric;
    prm.etc1_quality = opt->etc1_quality;
  }
  return prm;
}

int check_options(const fastc_gpu_options *opt) {
  if (!opt) return 0;
  if (opt->struct_size != sizeof(fastc_gpu_options)) return fail("fastc_gpu_options::struct_size does not match this library");
  if (opt->bptc_error_metric != 0 && opt->bptc_error_metric != 1) return fail("unknown BPTC error metric %d", opt->bptc_error_metric);
  if (opt->etc1_quality < 0 || opt->etc1_quality > 2) return fail("unknown ETC1 quality %d", opt->etc1_quality);
  if ((opt->bptc_block_modes & 0xFFu) == 0) return fail("BPTC block-mode mask selects no mode");
  return 0;
}

int compress_device_impl(int format, const void *rgba_dev, uint32_t width, uint32_t height, uint32_t first_block,
                         uint32_t num_blocks, void *out_dev, int quality, uint64_t seed, uint32_t wm_base,
                         uint32_t block_index_base, void *cuda_stream, uint32_t *launches_out,
                         const fastc_gpu_options *opt) {
  if (check_dims(format, width, height)) return 1;
  if (check_range(width, height, first_block, &num_blocks)) return 1;
  if (!rgba_dev || !out_dev) return fail("null device pointer");
  if (quality < 0) return fail("quality must be >= 0");
  if (check_options(opt)) return 1;
  int dev = 0;
  if (current_device(&dev)) return 1;
  if (ensure_tables(dev)) return 1;
  uint32_t n = 0;
  if (enqueue(dev, format, rgba_dev, width, height, first_block, num_blocks, out_dev, make_params(quality, seed, opt),
              wm_base, block_index_base, static_cast<cudaStream_t>(cuda_stream), &n))
    return 1;
  if (launches_out) *launches_out = n;
  return 0;
}

int compress_impl(int format, const uint8_t *rgba_host, uint32_t width, uint32_t height, uint32_t first_block,
                  uint32_t num_blocks, uint8_t *out_host, int quality, uint64_t seed, uint32_t chunk_blocks,
                  int num_gpus, fastc_gpu_timing *timing, const fastc_gpu_options *opt) {
  auto t0 = std::chrono::steady_clock::now();
  if (check_dims(format, width, height)) return 1;
  if (!rgba_host || !out_host) return fail("null host pointer");
  if (quality < 0) return fail("quality must be >= 0");
  if (check_options(opt)) return 1;
  const uint32_t bx = width / 4;
  if (check_range(width, height, first_block, &num_blocks)) return 1;
  int ndev = device_count();
  if (ndev <= 0) return fail("no CUDA device available (there is no CPU fallback)");
  if (num_gpus <= 0) num_gpus = ndev;  // "all visible" (SCompressionSettings::iNumGPUs == 0, tc -g 0)
  num_gpus = std::min(num_gpus, ndev);
  int prev = 0;
  cudaGetDevice(&prev);
  const EncodeParams prm = make_params(quality, seed, opt);
  fastc_gpu_bptc_block_stat *stats = (opt && format == FASTC_GPU_BPTC) ? opt->bptc_block_stats : nullptr;

  // Contiguous block-row slabs, one per GPU.
  const uint32_t row0 = first_block / bx, row1 = (first_block + num_blocks + bx - 1) / bx;
  const uint32_t rows = row1 - row0;
  num_gpus = std::max(1, std::min<int>(num_gpus, rows));
  if (format == FASTC_GPU_PVRTC4) {  // neighbour-coupled over the whole texture: one GPU, all of it
    if (first_block != 0 || num_blocks != bx * (height / 4)) return fail("PVRTC4 encodes whole textures: block ranges are not supported");
    num_gpus = 1;
  }
  std::vector<Shard> shards(num_gpus);
  WmChain chain;
  int npieces = 0;
  for (int g = 0; g < num_gpus; g++) {
    uint32_t a = row0 + (uint32_t)((uint64_t)rows * g / num_gpus);
    uint32_t b = row0 + (uint32_t)((uint64_t)rows * (g + 1) / num_gpus);
    uint32_t lo = std::max(first_block, a * bx), hi = std::min(first_block + num_blocks, b * bx);
    shards[g].dev = g;
    shards[g].first_block = lo;
    shards[g].num_blocks = hi > lo ? hi - lo : 0;
    plan_chunks(shards[g], format, width, chunk_blocks);
    shards[g].piece0 = npieces;
    npieces += shards[g].bounds.empty() ? 0 : (int)shards[g].bounds.size() - 1;
  }
  chain.resize((size_t)npieces);
  // BC7 watermark order: the word index of a solid block is the number of solid blocks before it
  // in raster order over the WHOLE image (the reference's single-threaded process-global counter,
  // Compressor.cpp:135-140,1457).  The counts of the submission's own pieces come from the GPUs
  // (WmChain); only the blocks before a sub-range submission are counted on the host (cached).
  if (format == FASTC_GPU_BPTC) chain.before = host_prefix_solid(rgba_host, width, height, first_block);
  if (num_gpus == 1) {
    cudaGetDevice(&shards[0].dev);  // one GPU: the caller's current device
    shards[0].rc = run_shard_locked(shards[0], format, rgba_host, width, height, out_host, prm, chain, stats);
    if (shards[0].rc) snprintf(shards[0].err, sizeof(shards[0].err), "%s", tl_error);
  } else {
    std::vector<std::thread> th;
    for (int g = 0; g < num_gpus; g++)
      th.emplace_back([&, g] {
        if (shards[g].num_blocks == 0) return;
        shards[g].rc = run_shard_locked(shards[g], format, rgba_host, width, height, out_host, prm, chain, stats);
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

int compress_batch_impl(int format, const fastc_gpu_job *jobs, uint32_t num_jobs, int quality, uint64_t seed,
                        int num_gpus, fastc_gpu_timing *timing, const fastc_gpu_options *opt) {
  auto t0 = std::chrono::steady_clock::now();
  if (!jobs && num_jobs) return fail("null job list");
  if (quality < 0) return fail("quality must be >= 0");
  if (check_options(opt)) return 1;
  int ndev = device_count();
  if (ndev <= 0) return fail("no CUDA device available (there is no CPU fallback)");
  if (num_gpus <= 0) num_gpus = ndev;
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
    int dev = g;
    if (num_gpus == 1) cudaGetDevice(&dev);
    if (dev < 0 || dev >= kMaxDevices) { rcs[g] = 1; errs[g] = "bad device"; return; }
    std::lock_guard<std::mutex> lk(g_ctx[dev].host_mu);
    if (format == FASTC_GPU_PVRTC4) {
      // PVRTC: a texture is (nearly) a serial job for two CTAs, so the textures of a batch are encoded
      // SIDE BY SIDE: runs of equally sized jobs go through the kernels together (grid.y = texture),
      // as many at a time as fit ~8 GiB of scratch.
      if (ensure_ctx(dev)) { rcs[g] = 1; errs[g] = tl_error; return; }
      DeviceCtx &c = g_ctx[dev];
      cudaStream_t st = c.streams[0];
      std::vector<uint32_t> mine;
      for (uint32_t j = g; j < num_jobs; j += num_gpus) mine.push_back(j);
      auto failed = [&]() { rcs[g] = 1; errs[g] = tl_error; abort_slots(c); };
      for (size_t a = 0; a < mine.size();) {
        const uint32_t w = jobs[mine[a]].width, h = jobs[mine[a]].height;
        const size_t in_b = (size_t)w * h * 4, out_b = (size_t)(w / 4) * (h / 4) * 8;
        const size_t cap = std::max<size_t>(1, std::min<size_t>(((size_t)8 << 30) / pvrtc_scratch_bytes(w, h), 0xFFFFFFFFull / ((size_t)w * h)));
        size_t b = a;
        while (b < mine.size() && b - a < cap && jobs[mine[b]].width == w && jobs[mine[b]].height == h) b++;
        const uint32_t n = (uint32_t)(b - a);
        if (cudaStreamSynchronize(st) != cudaSuccess || grow(&c.in_buf[0], &c.in_cap[0], in_b * n) ||
            grow(&c.out_buf[0], &c.out_cap[0], out_b * n)) { failed(); return; }
        for (uint32_t k = 0; k < n; k++)
          if (upload(c, (uint8_t *)c.in_buf[0] + in_b * k, jobs[mine[a + k]].rgba_host, in_b, st)) { failed(); return; }
        uint32_t nl = 0;
        bool ok = cudaEventRecord(c.ev_start[0], st) == cudaSuccess;
        {
          std::lock_guard<std::mutex> dl(c.mu);
          if (!c.pvr_done) ok = ok && cudaEventCreateWithFlags(&c.pvr_done, cudaEventDisableTiming) == cudaSuccess;
          if (ok && c.pvr_used) ok = cudaStreamWaitEvent(st, c.pvr_done, 0) == cudaSuccess;
          ok = ok && launch_pvrtc(c.pvrws, c.in_buf[0], w, h, n, c.out_buf[0], st, &nl) == cudaSuccess;
          ok = ok && cudaEventRecord(c.pvr_done, st) == cudaSuccess;
          c.pvr_used = true;
        }
        ok = ok && cudaEventRecord(c.ev_stop[0], st) == cudaSuccess;
        for (uint32_t k = 0; ok && k < n; k++)
          ok = cudaMemcpyAsync(jobs[mine[a + k]].out_host, (uint8_t *)c.out_buf[0] + out_b * k, out_b, cudaMemcpyDeviceToHost, st) == cudaSuccess;
        ok = ok && cudaStreamSynchronize(st) == cudaSuccess;
        float ms = 0;
        ok = ok && cudaEventElapsedTime(&ms, c.ev_start[0], c.ev_stop[0]) == cudaSuccess;
        if (!ok) { fail("PVRTC batch: %s", cudaGetErrorString(cudaGetLastError())); failed(); return; }
        per[g].kernel_ms += ms;
        per[g].kernel_launches += nl;
        per[g].h2d_bytes += in_b * n;
        per[g].d2h_bytes += out_b * n;
        a = b;
      }
      return;
    }
    // Pageable inputs of small textures: staging texture by texture inside upload() leaves the host copy
    // (one thread moves ~8 GB/s) in front of every transfer.  Instead a few threads stage WHOLE textures
    // into a ring of pinned buffers, running ahead of this thread, which then submits from pinned
    // memory: the copies proceed at the rate of several threads (~50 GB/s on the GPU box's host).
    constexpr int kStagers = 8;
    constexpr size_t kStageMaxJob = (size_t)16 << 20;
    std::vector<uint32_t> mine;
    for (uint32_t j = g; j < num_jobs; j += num_gpus) mine.push_back(j);
    std::vector<const uint8_t *> src_of(mine.size());
    std::vector<std::atomic<int>> staged(mine.size());
    std::atomic<long> submitted(-1);
    std::atomic<bool> stop(false);
    std::vector<std::thread> stagers;
    {
      bool want = false;
      if (ensure_ctx(dev)) { rcs[g] = 1; errs[g] = tl_error; return; }
      DeviceCtx &c = g_ctx[dev];
      for (size_t k = 0; k < mine.size(); k++) {
        const fastc_gpu_job &jb = jobs[mine[k]];
        const size_t bytes = (size_t)jb.width * jb.height * 4;
        src_of[k] = jb.rgba_host;
        const bool stage = mine.size() >= 4 && bytes <= kStageMaxJob && is_pageable(jb.rgba_host);
        staged[k].store(stage ? 0 : 1, std::memory_order_relaxed);
        if (!stage) continue;
        want = true;
        void *&buf = c.batch_pin[k % DeviceCtx::kBatchRing];
        size_t &cap = c.batch_pin_cap[k % DeviceCtx::kBatchRing];
        if (cap < bytes) {
          if (buf) cudaFreeHost(buf);
          buf = nullptr; cap = 0;
          if (cudaHostAlloc(&buf, kStageMaxJob, cudaHostAllocDefault) != cudaSuccess) {
            fail("cudaHostAlloc failed: %s", cudaGetErrorString(cudaGetLastError()));
            rcs[g] = 1; errs[g] = tl_error;
            return;
          }
          cap = kStageMaxJob;
        }
        src_of[k] = static_cast<const uint8_t *>(buf);
      }
      if (want)
        for (int t = 0; t < kStagers; t++)
          stagers.emplace_back([&, t] {
            for (size_t k = (size_t)t; k < mine.size() && !stop.load(std::memory_order_acquire); k += kStagers) {
              if (staged[k].load(std::memory_order_acquire)) continue;
              // the ring buffer's previous texture (k - ring) has left it once kPipeDepth more textures were
              // submitted: their chunks have cycled through every staging slot (finish_slot synchronises)
              while (submitted.load(std::memory_order_acquire) < (long)k - DeviceCtx::kBatchRing + kPipeDepth &&
                     !stop.load(std::memory_order_acquire))
                std::this_thread::yield();
              if (stop.load(std::memory_order_acquire)) break;
              const fastc_gpu_job &jb = jobs[mine[k]];
              memcpy(const_cast<uint8_t *>(src_of[k]), jb.rgba_host, (size_t)jb.width * jb.height * 4);
              staged[k].store(1, std::memory_order_release);
            }
          });
    }
    auto join_stagers = [&] {
      stop.store(true, std::memory_order_release);
      for (auto &t : stagers) t.join();
      stagers.clear();
    };
    bool any = false;
    for (size_t k = 0; k < mine.size(); k++) {
      const uint32_t j = mine[k];
      while (!staged[k].load(std::memory_order_acquire)) std::this_thread::yield();
      Shard s;
      s.dev = dev;
      s.first_block = 0;
      s.num_blocks = (jobs[j].width / 4) * (jobs[j].height / 4);
      plan_chunks(s, format, jobs[j].width, 0);
      WmChain chain;
      chain.resize(s.bounds.size() - 1);
      EncodeParams prm = make_params(quality, seed + ((uint64_t)j << 40), opt);
      // no drain between textures: the staging slots keep rotating across jobs
      if (run_shard(s, format, src_of[k], jobs[j].width, jobs[j].height, jobs[j].out_host, prm, chain,
                    /*drain=*/false)) {
        rcs[g] = 1;
        errs[g] = tl_error;
        join_stagers();
        return;
      }
      submitted.store((long)k, std::memory_order_release);
      per[g].kernel_ms += s.kernel_ms;
      per[g].kernel_launches += s.launches;
      per[g].h2d_bytes += s.h2d;
      per[g].d2h_bytes += s.d2h;
      any = true;
    }
    join_stagers();
    if (any) {
      double ms = 0;
      if (drain_slots(g_ctx[dev], &ms)) {
        abort_slots(g_ctx[dev]);
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

}  // namespace

int fastc_gpu_compress_device(int format, const void *rgba_dev, uint32_t width, uint32_t height,
                              uint32_t first_block, uint32_t num_blocks, void *out_dev, int quality,
                              uint64_t seed, uint32_t wm_base, uint32_t block_index_base, void *cuda_stream,
                              uint32_t *launches_out) {
  return compress_device_impl(format, rgba_dev, width, height, first_block, num_blocks, out_dev, quality, seed, wm_base,
                              block_index_base, cuda_stream, launches_out, nullptr);
}

int fastc_gpu_compress_device_opt(int format, const void *rgba_dev, uint32_t width, uint32_t height,
                                  uint32_t first_block, uint32_t num_blocks, void *out_dev, int quality,
                                  uint64_t seed, uint32_t wm_base, uint32_t block_index_base, void *cuda_stream,
                                  uint32_t *launches_out, const fastc_gpu_options *options) {
  return compress_device_impl(format, rgba_dev, width, height, first_block, num_blocks, out_dev, quality, seed, wm_base,
                              block_index_base, cuda_stream, launches_out, options);
}

int fastc_gpu_count_solid_device(const void *rgba_dev, uint32_t width, uint32_t height, uint32_t first_block,
                                 uint32_t num_blocks, void *cuda_stream, uint32_t *count_out) {
  if (check_dims(FASTC_GPU_BPTC, width, height)) return 1;
  if (!rgba_dev || !count_out) return fail("null pointer");
  if (check_range(width, height, first_block, &num_blocks)) return 1;
  int dev = 0;
  if (current_device(&dev)) return 1;
  if (ensure_tables(dev)) return 1;
  DeviceCtx &c = g_ctx[dev];
  std::lock_guard<std::mutex> lk(c.mu);
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  // shares the device API's scratch with fastc_gpu_compress_device (see enqueue)
  if (!c.api_done) CU_TRY(cudaEventCreateWithFlags(&c.api_done, cudaEventDisableTiming));
  if (c.api_used) CU_TRY(cudaStreamWaitEvent(st, c.api_done, 0));
  CU_TRY(bc7_count_solid(c.bc7ws[kPipeDepth], rgba_dev, width, first_block, num_blocks, st, count_out));
  CU_TRY(cudaEventRecord(c.api_done, st));
  c.api_used = true;
  return 0;
}

int fastc_gpu_compress(int format, const uint8_t *rgba_host, uint32_t width, uint32_t height,
                       uint32_t first_block, uint32_t num_blocks, uint8_t *out_host, int quality, uint64_t seed,
                       uint32_t chunk_blocks, int num_gpus, fastc_gpu_timing *timing) {
  return compress_impl(format, rgba_host, width, height, first_block, num_blocks, out_host, quality, seed, chunk_blocks,
                       num_gpus, timing, nullptr);
}

int fastc_gpu_compress_opt(int format, const uint8_t *rgba_host, uint32_t width, uint32_t height,
                           uint32_t first_block, uint32_t num_blocks, uint8_t *out_host, int quality, uint64_t seed,
                           uint32_t chunk_blocks, int num_gpus, fastc_gpu_timing *timing,
                           const fastc_gpu_options *options) {
  return compress_impl(format, rgba_host, width, height, first_block, num_blocks, out_host, quality, seed, chunk_blocks,
                       num_gpus, timing, options);
}

int fastc_gpu_compress_batch(int format, const fastc_gpu_job *jobs, uint32_t num_jobs, int quality, uint64_t seed,
                             int num_gpus, fastc_gpu_timing *timing) {
  return compress_batch_impl(format, jobs, num_jobs, quality, seed, num_gpus, timing, nullptr);
}

int fastc_gpu_compress_batch_opt(int format, const fastc_gpu_job *jobs, uint32_t num_jobs, int quality, uint64_t seed,
                                 int num_gpus, fastc_gpu_timing *timing, const fastc_gpu_options *options) {
  return compress_batch_impl(format, jobs, num_jobs, quality, seed, num_gpus, timing, options);
}

int fastc_gpu_decompress_device(int format, const void *cmp_dev, uint32_t width, uint32_t height, void *rgba_out_dev,
                                void *cuda_stream) {
  if (check_dims(format, width, height)) return 1;
  if (!cmp_dev || !rgba_out_dev) return fail("null device pointer");
  int dev = 0;
  if (current_device(&dev)) return 1;
  if (ensure_tables(dev)) return 1;
  if (format == FASTC_GPU_PVRTC4)
    CU_TRY(launch_pvrtc_decode(cmp_dev, width, height, 0, (width / 4) * (height / 4), rgba_out_dev,
                               static_cast<cudaStream_t>(cuda_stream)));
  else
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
  if (current_device(&dev)) return 1;
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
  if (format == FASTC_GPU_PVRTC4)
    CU_TRY(launch_pvrtc_decode(c.out_buf[0], width, height, 0, (width / 4) * (height / 4), c.in_buf[0], st));
  else
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
  if (current_device(&dev)) return 1;
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
  if (current_device(&dev)) return 1;
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
  if (current_device(&dev)) return 1;
  DeviceCtx &c = g_ctx[dev];
  std::lock_guard<std::mutex> lk(c.mu);
  CU_TRY(bc7_read_counters(c.bc7ws[kPipeDepth], qe_calls, pixel_bucket_evals));
  return 0;
}

int fastc_gpu_bc7_stage_ms(int enable, double *ms6) {
  int dev = 0;
  if (current_device(&dev)) return 1;
  DeviceCtx &c = g_ctx[dev];
  std::lock_guard<std::mutex> lk(c.mu);
  CU_TRY(bc7_stage_timing(c.bc7ws[kPipeDepth], enable, ms6));
  return 0;
}

int fastc_gpu_debug_bc7_dump(uint32_t nblocks, uint32_t *sel_out, uint32_t *results_out) {
  int dev = 0;
  if (current_device(&dev)) return 1;
  DeviceCtx &c = g_ctx[dev];
  std::lock_guard<std::mutex> lk(c.mu);
  CU_TRY(bc7_debug_dump(c.bc7ws[kPipeDepth], nblocks, sel_out, results_out));
  return 0;
}

const char *fastc_gpu_last_error(void) { return tl_error; }

}  // extern "C"
