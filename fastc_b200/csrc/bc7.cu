// BC7 (BPTC) block encoder for sm_100a.
//
// Behavioural contract: the per-block search of the reference's BPTCEncoder
//   reference/BPTCEncoder/src/Compressor.cpp   (CompressBC7Block :1819, BoxSelection :1670,
//       CompressClusters :1752, CompressionMode::Compress :1300, CompressCluster :921 / :632,
//       OptimizeEndpointsForCluster :538, PickBestNeighboringEndpoints :426, Pack :1096)
//   reference/BPTCEncoder/src/RGBAEndpoints.cpp (QuantizedError :190, GetPrincipalAxis :327,
//       QuantizeChannel :126)
// replayed with the same float semantics (no FMA: this TU is built with
// -fmad=false; IEEE div/sqrt), so that quality 0 is bit-identical to the
// reference and quality > 0 runs the reference's annealing schedule on keyed
// per-chain RNG streams (bit-identical to oracle/bc7_oracle.cpp rng_mode 1).
//
// B200 mapping (this is NOT how the reference is organised):
//   bc7_classify   1 thread / block : solid / transparent flags, per-tile solid counts
//   bc7_wm_scan    1 CTA            : exclusive scan of tile counts (watermark order, T1)
//   bc7_select     1 warp / block   : lanes = the 64 partition shapes; bbox + 8/4-bucket
//                                     error estimate per shape, warp argmin (first index wins)
//   bc7_setup      1 thread / endpoint-fit "chain": PCA + k-means + least squares + grid clamp +
//                                     first evaluation; chains that share a cluster and an index
//                                     precision share the fit (twin_slot); writes either a final
//                                     result or the start state of the chain's annealing.  Five
//                                     instantiations <index precision, rotation fit or not>, each
//                                     with only the code its chains run (the kernel is bound by
//                                     instruction fetch); float work in packed f32x2 pairs
//   bc7_bin_offsets / bc7_scatter   : counting sort of the start states by (precision, cluster size)
//   bc7_anneal     persistent lanes : 1 chain / lane, refilled from the sorted list; all-integer
//                                     evaluation (paired palette rows, VABSDIFF4 + DP4A)
//   bc7_pack       1 thread / block : best mode by total error in the reference's mode
//                                     order, anchor fix-ups, 128-bit pack
// The unit of parallelism for the expensive part is the *chain* (one subset of
// one candidate mode/shape/rotation): ~14 independent serial chains per block.
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <type_traits>

#include "bc7_tables.cuh"
#include "common.cuh"
#include "kernels.h"

namespace fastc {
namespace {

// ------------------------------------------------------------------ tables
__constant__ uint16_t c_shape2[64];
__constant__ uint32_t c_shape3[64];
__constant__ uint8_t c_anchor2[64], c_anchor3a[64], c_anchor3b[64];
__constant__ uint8_t c_weight[64];  // [index_bits-1][16] weight of endpoint 2 (0..64)
__constant__ uint8_t c_opt7[512];   // [v][2]
__constant__ uint8_t c_opt6[1536];  // [v][2][3]
__constant__ uint32_t c_wm[9];

// kErrorMetrics[eErrorMetric_Nonuniform] (Compressor.cpp:205-208): sqrtf(0.3f), sqrtf(0.56f), sqrtf(0.11f), 1
// (computed on the host like the reference's static initialiser, uploaded by bc7_upload_tables)
__constant__ float c_nu_weights[4];

// Per-mode attributes (BC7 spec; reference kModeAttributes Compressor.cpp:170-203).
struct ModeAttr {
  uint8_t partition_bits, subsets, index_bits, alpha_index_bits, color_bits, alpha_bits, rotation, idx_mode, pbit;
};
enum { kPbitShared = 0, kPbitPerEndpoint = 1, kPbitNone = 2 };
__constant__ ModeAttr c_modes[8] = {
    {4, 3, 3, 0, 4, 0, 0, 0, 1}, {6, 2, 3, 0, 6, 0, 0, 0, 0}, {6, 3, 2, 0, 5, 0, 0, 0, 2}, {6, 2, 2, 0, 7, 0, 0, 0, 1},
    {0, 1, 2, 3, 5, 6, 1, 1, 2}, {0, 1, 2, 2, 7, 8, 1, 0, 2}, {0, 1, 4, 0, 7, 7, 0, 0, 1}, {6, 2, 2, 0, 5, 5, 0, 0, 1},
};

// CompressSingleColor (Compressor.cpp:252-353) is a per-channel exhaustive search
// that only depends on (mode, index mode, p-bit combo, channel class, byte value):
// precomputed on the host with the reference's scan order (first strict minimum,
// i-major / j-minor).  Entry = v1 | v2 << 8 | dist << 16.
// index: ((((mode * 2 + idx_mode) * 4 + pbi) * 2 + is_alpha) * 256 + val)
__device__ uint32_t g_single[8 * 2 * 4 * 2 * 256];

// ------------------------------------------------------------------ layout of the scratch
// sel word per block
//  [0:5] best 2-subset shape  [6:11] best 3-subset shape  [12:19] mode mask
//  [20:21] number of shapes   [22] layout B (alpha path)   [23] a shape estimate was ~0 (early-out)   [24:25] type
enum { kTypeNormal = 0, kTypeSolid = 1, kTypeTransparent = 2 };
constexpr int kSlots = 16;       // result slots per block
constexpr int kResWords = 8;     // 32 B per chain result
// result words: 0 err, 1 p1, 2 p2, 3 p-bit combo, 4-5 colour indices (4 bit each,
// cluster-local order), 6-7 alpha indices (modes 4/5)
// Start state of one annealing chain (8 words), written by bc7_setup, read by bc7_anneal.
//  w0: subset pixel mask [0:15] | mode [16:18] | rot [19:20] | idx_mode [21] | combo [22:23] | n [24:28] | valid [31]
//  w1/w2: start endpoints (bytes on the grid)   w3: start error   w4: RNG state
//  w5: alpha error (modes 4/5)   w6: rounded alpha endpoint bytes a1 | a2 << 8 (modes 4/5)
constexpr int kStateWords = 8;
constexpr int kStatDoubles = 10; // per-block statistics record: mode, path, error of modes 0..7 (-1: not tried)
constexpr int kTile = 256;       // blocks per watermark tile (= classify CTA)

struct Ws {
  uint32_t *sel;        // [nblocks]
  uint32_t *tile_count; // [ntiles] solid blocks per tile, then exclusive-scanned in place
  uint32_t *total_solid;
  uint32_t *results;    // [nblocks][kSlots][kResWords]
  uint32_t *states;     // [nblocks][kSlots][kStateWords] annealing start states
  uint32_t *sa_mask;    // [nblocks] bit s: slot s holds a start state (zeroed per submission, set by bc7_setup)
  uint4 *sorted;        // [nblocks*kSlots][2] the live chains' start states, sorted by descending
                        // (index precision, cluster size); word 7 = the chain's id (block * kSlots + slot)
  uint32_t *bins;       // histogram / offsets / cursors (see bc7_bin_offsets)
  uint32_t *tail_list;  // [kTailCap] chains handed from bc7_anneal to bc7_anneal_tail
  double *stats;        // [nblocks][kStatDoubles] per-block statistics (mode, path, error of every mode tried), or NULL
  double *err64;        // [nblocks][kSlots] non-uniform metric only: the chains' errors as doubles (else NULL)
  const uint32_t *wm_running;    // watermark base of this chunk (device side, chunks chain without a host sync)
  unsigned long long *counters;  // qe calls, pbe
};

// ------------------------------------------------------------------ small helpers
__device__ __forceinline__ int chan(uint32_t p, int k) { return (p >> (8 * k)) & 0xFF; }

__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
  h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
  return h;
}
// Keyed per-chain RNG stream; identical to oracle/bc7_oracle.cpp chain_seed().
__device__ __forceinline__ uint32_t chain_seed(uint64_t seed, uint32_t block, uint32_t chain) {
  const uint32_t h = (uint32_t)seed ^ ((uint32_t)(seed >> 32) * 0x9E3779B9u);
  return fmix32(h + fmix32(block * 64u + chain));
}
// reference fastrand() (Compressor.cpp:358-362): 16-bit draws on glibc (RAND_MAX = 2^31-1).
__device__ __forceinline__ uint32_t lcg_next(uint32_t &s) {
  s = 214013u * s + 2531011u;
  return s >> 16;
}

__device__ __forceinline__ int subset_of(int idx, int shape, int nsub) {
  if (nsub == 2) return (c_shape2[shape] >> idx) & 1;
  if (nsub == 3) return (c_shape3[shape] >> (2 * idx)) & 3;
  return 0;
}
__device__ __forceinline__ int anchor_of(int subset, int shape, int nsub) {
  if (subset == 0) return 0;
  if (subset == 1) return nsub == 2 ? c_anchor2[shape] : c_anchor3a[shape];
  return c_anchor3b[shape];
}

// GetQuantizationMask (CompressionMode.h:212-232): per-channel mask of the kept top bits.
__device__ __forceinline__ uint32_t quant_mask(const ModeAttr &A) {
  const uint32_t cm = (0xFF00u >> A.color_bits) & 0xFF;
  const uint32_t am = A.alpha_bits ? ((0xFF00u >> A.alpha_bits) & 0xFF) : 0u;
  return cm | (cm << 8) | (cm << 16) | (am << 24);
}

// QuantizeChannel (RGBAEndpoints.cpp:126-165).  prec = number of mask bits.
__device__ __noinline__ uint32_t quantize_channel(uint32_t val, uint32_t mask, int pbit) {
  if (mask == 0xFF) return val;
  if (mask == 0) return 0xFF;
  int prec = __popc(mask);
  const uint32_t step = 1u << (8 - prec);
  uint32_t lval = val & mask, hval = lval + step;
  if (pbit >= 0) {
    prec++;
    lval |= (uint32_t)(pbit != 0) << (8 - prec);
    hval |= (uint32_t)(pbit != 0) << (8 - prec);
  }
  if (lval > val) { lval -= step; hval -= step; }
  lval |= lval >> prec;
  hval |= hval >> prec;
  const uint32_t l8 = lval & 0xFF, h8 = hval & 0xFF;  // sad<uint8>
  const uint32_t dl = val > l8 ? val - l8 : l8 - val;
  const uint32_t dh = val > h8 ? val - h8 : h8 - val;
  return (dl < dh ? lval : hval) & 0xFF;
}
// QuantizeChannel for every (kept bits, p-bit, value), filled by bc7_build_quant_table at table upload:
// row = class * 3 + (pbit + 1), class 0..3 = 4..7 kept bits, 4 = all eight (identity), 5 = none (0xFF).
__device__ uint8_t g_quant[18 * 256];
__global__ void bc7_build_quant_table() {
  for (int e = threadIdx.x; e < 18 * 256; e += blockDim.x) {
    const int row = e >> 8, cls = row / 3, pbit = row % 3 - 1;
    const uint32_t mask = cls == 5 ? 0u : (cls == 4 ? 0xFFu : ((0xFF00u >> (cls + 4)) & 0xFFu));
    g_quant[e] = (uint8_t)quantize_channel((uint32_t)(e & 255), mask, pbit);
  }
}
__device__ __forceinline__ int quant_row(uint32_t mask8, int pbit) {
  const int bits = __popc(mask8);
  return (bits == 0 ? 5 : bits - 4) * 3 + pbit + 1;
}
// ToPixel for endpoint bytes that are already integers (RGBAEndpoints.cpp:167-177): four lookups
// (bc7_setup spent 12 % of its instructions in the arithmetic version).
__device__ __noinline__ uint32_t to_pixel_b(uint32_t p, uint32_t mask, int pbit) {
  const uint8_t *tc = g_quant + 256 * quant_row(mask & 0xFF, pbit), *ta = g_quant + 256 * quant_row(mask >> 24, pbit);
  return (uint32_t)__ldg(tc + (p & 0xFF)) | ((uint32_t)__ldg(tc + ((p >> 8) & 0xFF)) << 8) |
         ((uint32_t)__ldg(tc + ((p >> 16) & 0xFF)) << 16) | ((uint32_t)__ldg(ta + (p >> 24)) << 24);
}
// uint32(x + 0.5) & 0xFF, x in [0, 255.5): exact without fp64 (x - floor(x) is exact).
// floor(x + 0.5) = (floor(2 x) + 1) >> 1, and 2 x is exact; the conversion maps NaN to 0 like x86's
// cvttsd2si does in the low byte.
__device__ __forceinline__ uint32_t round_byte(float x) {
  return (uint32_t)((__float2int_rd(__fadd_rn(x, x)) + 1) >> 1) & 0xFF;
}
__device__ __forceinline__ uint32_t pack_round(const float p[4]) {
  return round_byte(p[0]) | (round_byte(p[1]) << 8) | (round_byte(p[2]) << 16) | (round_byte(p[3]) << 24);
}

// p-bit pair of combo `idx` (CompressionMode.h:244-251): returns pb[0] | pb[1] << 1, or -1/-1 via has=false
__device__ __forceinline__ void pbit_combo(int pbit_type, int idx, int &pb0, int &pb1) {
  if (pbit_type == kPbitShared) { pb0 = pb1 = (idx ? 1 : 0); }
  else if (pbit_type == kPbitPerEndpoint) { pb0 = (idx >> 1) & 1; pb1 = idx & 1; }
  else { pb0 = pb1 = -1; }
}

// ------------------------------------------------------------------ QuantizedError, integer form
// RGBACluster::QuantizedError (RGBAEndpoints.cpp:190-310) for the uniform metric.
// With metric (1,1,1,1) every per-pixel error is an integer (sum of squared byte
// differences <= 4*255^2) and every partial sum stays below 2^24, so the
// reference's float accumulation is exact and an int32 sum reproduces it
// bit-for-bit.  The projection uses the reference's float ops (one correctly
// rounded division per pixel; the dot products are exact integers).
//
// pts: points used for the projection (== pix except in the rotated mode-4/5 fit),
// pix: original pixel bytes the error is measured against (T16).
struct QeEndpoints {
  int e1[4], d[4];  // quantised endpoint 1 and (endpoint 2 - endpoint 1), per channel
  int den;          // |e2 - e1|^2
};
__device__ __forceinline__ void qe_prepare(QeEndpoints &q, uint32_t q1, uint32_t q2) {
  q.den = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    q.e1[k] = chan(q1, k);
    q.d[k] = chan(q2, k) - q.e1[k];
    q.den += q.d[k] * q.d[k];
  }
}
// error of pixel bytes `pb` against bucket with weight w (0..64)
__device__ __forceinline__ int qe_bucket_error(const QeEndpoints &q, const int pb[4], int w) {
  int err = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int ip = q.e1[k] + ((q.d[k] * w + 32) >> 6);  // == ((64-w)*e1 + w*e2 + 32) >> 6
    const int df = pb[k] - ip;
    err += df * df;
  }
  return err;
}
// One pixel: returns min error; *best = chosen bucket.
__device__ __forceinline__ int qe_pixel(const QeEndpoints &q, uint32_t pt, uint32_t px, int nbm1,
                                        const uint8_t *__restrict__ wtab, int *best) {
  int pb[4];
#pragma unroll
  for (int k = 0; k < 4; k++) pb[k] = chan(px, k);
  if (q.den == 0) {
    *best = 0;
    return qe_bucket_error(q, pb, 0);
  }
  int num = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) num += (chan(pt, k) - q.e1[k]) * q.d[k];
  const float pct = __fdiv_rn((float)num, (float)q.den);
  const float t = __fmul_rn(pct, (float)nbm1);
  int j1 = (int)floorf(t), j2 = (int)ceilf(t);
  j1 = max(0, j1);
  j1 = min(j1, nbm1);
  j2 = min(j2, nbm1);
  int e = qe_bucket_error(q, pb, wtab[j1]);
  int b = j1;
  if (j1 + 1 <= j2) {  // at most two candidates: floor and ceil of the projection
    const int e2 = qe_bucket_error(q, pb, wtab[j1 + 1]);
    if (e2 < e) { e = e2; b = j1 + 1; }
  }
  *best = b;
  return e;
}

// The same under a non-uniform metric (m_ErrorMetric = eErrorMetric_Nonuniform): the error of a
// pixel against one interpolated colour is the float  sum_k (float(|p_k - c_k|) * w_k)^2, summed
// in channel order exactly like the reference (RGBAEndpoints.cpp:291-296, VectorBase::Dot); the
// totals are float sums in pixel order.  Kernels take the metric as a template flag (NU); the
// uniform instantiations keep the all-integer fast paths.
__device__ __forceinline__ float nu_error(uint32_t colour, uint32_t px, const float w[4]) {
  const uint32_t d = __vabsdiffu4(colour, px);
  const float e0 = __fmul_rn((float)(d & 0xFF), w[0]), e1 = __fmul_rn((float)((d >> 8) & 0xFF), w[1]);
  const float e2 = __fmul_rn((float)((d >> 16) & 0xFF), w[2]), e3 = __fmul_rn((float)(d >> 24), w[3]);
  float s = __fmul_rn(e0, e0);
  s = __fadd_rn(s, __fmul_rn(e1, e1));
  s = __fadd_rn(s, __fmul_rn(e2, e2));
  s = __fadd_rn(s, __fmul_rn(e3, e3));
  return s;
}
// metric of a chain: the weights follow the rotation (CompressionMode::GetErrorMetric, CompressionMode.h:196-205)
__device__ __forceinline__ void nu_metric(int rot, float w[4]) {
#pragma unroll
  for (int k = 0; k < 4; k++) w[k] = c_nu_weights[k];
  if (rot == 1) { w[0] = c_nu_weights[3]; w[3] = c_nu_weights[0]; }
  else if (rot == 2) { w[1] = c_nu_weights[3]; w[3] = c_nu_weights[1]; }
  else if (rot == 3) { w[2] = c_nu_weights[3]; w[3] = c_nu_weights[2]; }
}

// ------------------------------------------------------------------ classify + watermark scan
__device__ __forceinline__ uint32_t classify_block(const uint32_t px[16]) {
  bool solid = true, transparent = true;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    solid = solid && (px[i] == px[0]);
    transparent = transparent && ((px[i] >> 24) == 0);
  }
  return solid ? kTypeSolid : (transparent ? kTypeTransparent : kTypeNormal);
}

__global__ void __launch_bounds__(kTile)
bc7_classify(const uint32_t *__restrict__ img, uint32_t width, uint32_t blocks_x, uint32_t first_block,
             uint32_t num_blocks, uint32_t *__restrict__ sel, uint32_t *__restrict__ tile_count) {
  const uint32_t t = blockIdx.x * kTile + threadIdx.x;
  uint32_t type = kTypeNormal;
  if (t < num_blocks) {
    uint32_t px[16];
    load_block(img, width, blocks_x, first_block + t, px);
    type = classify_block(px);
    if (sel) sel[t] = type << 24;
  }
  const int cnt = __syncthreads_count(type == kTypeSolid);
  if (threadIdx.x == 0) tile_count[blockIdx.x] = cnt;
}

// Exclusive scan of the per-tile solid counts (single CTA; ntiles = nblocks/256).
__global__ void __launch_bounds__(1024) bc7_wm_scan(uint32_t *tile_count, uint32_t ntiles, uint32_t *total) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < ntiles; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < ntiles ? tile_count[i] : 0;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      uint32_t w = warp_sums[threadIdx.x];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
        if (threadIdx.x >= o) w += y;
      }
      warp_sums[threadIdx.x] = w;
    }
    __syncthreads();
    const uint32_t warp_off = (threadIdx.x >> 5) ? warp_sums[(threadIdx.x >> 5) - 1] : 0;
    const uint32_t incl = carry + warp_off + x;
    if (i < ntiles) tile_count[i] = incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

__global__ void bc7_set_u32(uint32_t *p, uint32_t v) { *p = v; }
__global__ void bc7_add_u32(uint32_t *p, const uint32_t *v) { *p += *v; }

// ------------------------------------------------------------------ shape selection
// BoxSelection (Compressor.cpp:1670-1750).  One warp per block; lane l evaluates
// shapes l and l+32.  Estimate for one shape = sum over its subsets of
//   0 if the subset's bounding box is a point, else 0.0001 + QuantizedError(bbox min, bbox max,
//   8 (two subsets) or 4 (three subsets) buckets, no quantisation)
// accumulated in double exactly like the reference (the tie-break between shapes
// with equal integer error depends on those roundings).
//
// Lanes of a warp hold different shapes, so every per-pixel decision is a select, never a
// branch: the pixel's subset picks (bbox min, extent, |extent|^2, min . extent, reciprocal) from
// registers.  The endpoints are the bbox corners, hence extent >= 0 per byte and
// 0 <= (p - min) . extent <= |extent|^2: the projection needs no clamping.  Bucket colours are
// interpolated two channels per 32-bit multiply (a byte times a weight <= 64 fits 16 bits), the
// squared error is |p|^2 + |c|^2 - 2 dp4a(p, c), and the bucket pair comes from one multiply by
// a per-subset reciprocal; whenever that product lands within 2^-16 of an integer the
// reference's own divide / multiply sequence decides (RGBAEndpoints.cpp:262-289).
constexpr int kSelWarps = 4;


// interpolation weights of the 3-bit / 2-bit index precisions (bc7tab::kWeight rows 2 / 1; checked
// against the table in bc7_upload_tables): compile-time constants for the unrolled palette build
__device__ constexpr uint32_t kSelWeights3[8] = {0, 9, 18, 27, 37, 46, 55, 64};
__device__ constexpr uint32_t kSelWeights2[4] = {0, 21, 43, 64};
__device__ constexpr uint32_t kWeights4[16] = {0, 4, 9, 13, 17, 21, 26, 30, 34, 38, 43, 47, 51, 55, 60, 64};
constexpr int kSelPalRows = 22;  // 16 palette rows (2 subsets x 8 buckets, or 3 x 4) + 2 rows of box scalars per subset
constexpr int kSelBoxRow = 16;   // row 16 + 2 s: (|extent|^2, min . extent), row 17 + 2 s: (reciprocal bits, extent bytes)

// `pal`: the lane's column of shared memory.  Row (subset * NB + j) = (colour j, colour j + 1) of that
// subset's bounding-box endpoints (the last row of a subset: colour NB-1 twice); rows kSelBoxRow.. hold
// the subset's scalars.  The kernel is bound by the ALU pipe: fetching the subset's operands with two
// 64-bit shared loads instead of selecting them from registers takes 5-10 ALU instructions off every pixel.
template <int NB, bool NU = false>
__device__ __forceinline__ typename std::conditional<NU, float, uint32_t>::type box_pixel_error(const uint2 *pal, int s, uint32_t p) {
  constexpr int NBM1 = NB - 1, STRIDE = kSelWarps * 32;
  const uint2 b0 = pal[(kSelBoxRow + 2 * s) * STRIDE], b1 = pal[(kSelBoxRow + 2 * s + 1) * STRIDE];
  const uint32_t den = b0.x, base = b0.y, d = b1.y;
  const float inv16 = __uint_as_float(b1.x);
  const uint32_t num = __dp4a(p, d, 0u) - base;  // (p - min) . extent, exact
  const float fnum = (float)num;
  const int v = __float2int_rd(__fmul_rn(fnum, inv16));
  int ja = min(v >> 16, NBM1);
  bool two = ja < NBM1;
  if (num == 0 || num == den) {
    two = false;  // pct is exactly 0 or 1: floor == ceil
    ja = num ? NBM1 : 0;
  } else if ((((uint32_t)v + 1u) & 0xFFFFu) <= 1u) {
    const float t = __fmul_rn(__fdiv_rn(fnum, (float)den), (float)NBM1);
    const int x1 = min(max(0, (int)floorf(t)), NBM1), x2 = min((int)ceilf(t), NBM1);
    ja = x1;
    two = x1 + 1 <= x2;
  }
  const uint2 c = pal[(s * NB + ja) * STRIDE];
  if constexpr (NU) {
    float w[4];
    nu_metric(0, w);
    const float ea = nu_error(c.x, p, w), eb = nu_error(c.y, p, w);
    return (two && eb < ea) ? eb : ea;
  } else {
    const uint32_t da = __vabsdiffu4(c.x, p), db = __vabsdiffu4(c.y, p);
    const uint32_t ea = __dp4a(da, da, 0u), eb = __dp4a(db, db, 0u);
    return two ? min(ea, eb) : ea;
  }
}

template <int NSUB, bool NU = false>
__device__ __forceinline__ double estimate_shape(const uint32_t *__restrict__ px, const uint32_t *__restrict__ plo,
                                                 const uint32_t *__restrict__ phi, int shape, uint2 *pal) {
  constexpr int NB = NSUB == 2 ? 8 : 4, NBM1 = NB - 1;
  // bounding boxes on pixels spread over 16-bit halves (bytes 0,2 / 1,3): unsigned 16x2 min / max
  // is one instruction on sm_100a, byte-wise min / max is not
  uint32_t mnl[NSUB], mnh[NSUB], mxl[NSUB], mxh[NSUB];
#pragma unroll
  for (int s = 0; s < NSUB; s++) { mnl[s] = mnh[s] = 0xFFFFFFFFu; mxl[s] = mxh[s] = 0; }
  const uint32_t m2 = NSUB == 2 ? c_shape2[shape] : 0;
  const uint32_t m3 = NSUB == 3 ? c_shape3[shape] : 0;
#pragma unroll 1
  for (int i = 0; i < 16; i += 2) {  // two pixels per three-input 16x2 min / max
    const int s0 = NSUB == 2 ? ((m2 >> i) & 1) : ((m3 >> (2 * i)) & 3);
    const int s1 = NSUB == 2 ? ((m2 >> (i + 1)) & 1) : ((m3 >> (2 * i + 2)) & 3);
    const uint32_t l0 = plo[i], h0 = phi[i], l1 = plo[i + 1], h1 = phi[i + 1];
#pragma unroll
    for (int q = 0; q < NSUB; q++) {
      mnl[q] = __vimin3_u16x2(mnl[q], s0 == q ? l0 : 0xFFFFFFFFu, s1 == q ? l1 : 0xFFFFFFFFu);
      mnh[q] = __vimin3_u16x2(mnh[q], s0 == q ? h0 : 0xFFFFFFFFu, s1 == q ? h1 : 0xFFFFFFFFu);
      mxl[q] = __vimax3_u16x2(mxl[q], s0 == q ? l0 : 0u, s1 == q ? l1 : 0u);
      mxh[q] = __vimax3_u16x2(mxh[q], s0 == q ? h0 : 0u, s1 == q ? h1 : 0u);
    }
  }
  uint32_t dens[NSUB];
#pragma unroll
  for (int s = 0; s < NSUB; s++) {
    // per byte mx >= mn: no borrow (every BC7 partition uses all its subsets)
    const uint32_t dlo = mxl[s] - mnl[s], dhi = mxh[s] - mnh[s];
    const uint32_t d = dlo | (dhi << 8), mn = mnl[s] | (mnh[s] << 8);
    const uint32_t den = __dp4a(d, d, 0u);
    dens[s] = den;
    const float inv16 = den ? __fdiv_rn(65536.0f * (float)NBM1, (float)den) : 0.0f;
    pal[(kSelBoxRow + 2 * s) * (kSelWarps * 32)] = make_uint2(den, __dp4a(mn, d, 0u));
    pal[(kSelBoxRow + 2 * s + 1) * (kSelWarps * 32)] = make_uint2(__float_as_uint(inv16), d);
    // the subset's palette, once per shape instead of twice per pixel: colour j = min + ((extent *
    // w_j + 32) >> 6) per channel, two channels per 32-bit multiply (a byte times w <= 64 fits 16 bits)
    uint32_t cur = mn;  // w_0 = 0
#pragma unroll
    for (int j = 1; j <= NBM1; j++) {
      const uint32_t w = NSUB == 2 ? kSelWeights3[j] : kSelWeights2[j];
      const uint32_t nxt = mn + ((((dlo * w + 0x00200020u) >> 6) & 0x00FF00FFu) |
                                 ((((dhi * w + 0x00200020u) >> 6) & 0x00FF00FFu) << 8));
      pal[(s * NB + j - 1) * (kSelWarps * 32)] = make_uint2(cur, nxt);
      cur = nxt;
    }
    pal[(s * NB + NBM1) * (kSelWarps * 32)] = make_uint2(cur, cur);
  }
  typename std::conditional<NU, float, uint32_t>::type tot[NSUB];
#pragma unroll
  for (int s = 0; s < NSUB; s++) tot[s] = 0;
#pragma unroll 2
  for (int i = 0; i < 16; i++) {
    const int s = NSUB == 2 ? ((m2 >> i) & 1) : ((m3 >> (2 * i)) & 3);
    // a point-sized box contributes nothing; its pixels evaluate to 0 anyway (p == min, extent 0)
    const auto e = box_pixel_error<NB, NU>(pal, s, px[i]);
    // (NU: float sums in the subset's pixel order; adding 0 to the other subsets' sums is exact)
#pragma unroll
    for (int q = 0; q < NSUB; q++) {
      if constexpr (NU) tot[q] = __fadd_rn(tot[q], (s == q) ? e : 0.0f);
      else tot[q] += (s == q) ? e : 0u;
    }
  }
  double err = 0.0;
#pragma unroll
  for (int s = 0; s < NSUB; s++) {
    const double e = dens[s] == 0 ? 0.0 : __dadd_rn(0.0001, (double)tot[s]);
    err = __dadd_rn(err, e);
  }
  return err;
}

// warp argmin with "first index wins" (strict < in scan order, T10)
__device__ __forceinline__ void warp_argmin(double &err, int &idx) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double e2 = __shfl_xor_sync(0xffffffffu, err, o);
    const int i2 = __shfl_xor_sync(0xffffffffu, idx, o);
    if (e2 < err || (e2 == err && i2 < idx)) { err = e2; idx = i2; }
  }
}

template <bool NU>
__global__ void __launch_bounds__(kSelWarps * 32)
bc7_select(const uint32_t *__restrict__ img, uint32_t width, uint32_t blocks_x, uint32_t first_block,
           uint32_t num_blocks, uint32_t *__restrict__ sel, uint32_t block_modes) {
  // block_modes: BPTCC::CompressionSettings::m_BlockModes, ANDed into the selection's mode set
  // (Compressor.cpp:1857); solid / transparent blocks never reach it (:1822-1846)
  const uint32_t mode_keep = ~(0xFFu << 12) | ((block_modes & 0xFFu) << 12);
  __shared__ uint32_t s_px[kSelWarps][16], s_plo[kSelWarps][16], s_phi[kSelWarps][16];
  __shared__ uint2 s_pal[kSelPalRows][kSelWarps * 32];  // one palette column per lane (estimate_shape)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint2 *pal = &s_pal[0][threadIdx.x];
  const uint32_t t = blockIdx.x * kSelWarps + warp;
  const bool valid = t < num_blocks;
  uint32_t type = kTypeNormal;
  if (valid) {
    type = sel[t] >> 24;
    if (lane < 16) {
      const uint32_t bi = first_block + t, bx = bi % blocks_x, by = bi / blocks_x;
      const uint32_t p = __ldg(img + (size_t)(by * 4 + (lane >> 2)) * width + bx * 4 + (lane & 3));
      s_px[warp][lane] = p;
      s_plo[warp][lane] = p & 0x00FF00FFu;
      s_phi[warp][lane] = (p >> 8) & 0x00FF00FFu;
    }
  }
  __syncthreads();
  if (!valid || type != kTypeNormal) return;
  const uint32_t *px = s_px[warp], *plo = s_plo[warp], *phi = s_phi[warp];

  bool opaque = true;
#pragma unroll
  for (int i = 0; i < 16; i++) opaque = opaque && ((px[i] >> 24) >= 250);

  // ---- two-subset shapes
  // the lane's two shapes go through ONE copy of the estimate code (rolled loop): the kernel is
  // instruction-fetch bound otherwise
  double e0 = 0.0, e1 = 0.0;
#pragma unroll 1
  for (int h = 0; h < 2; h++) {
    const double e = estimate_shape<2, NU>(px, plo, phi, lane + 32 * h, pal);  // 8 buckets: 3-bit weights
    if (h == 0) e0 = e; else e1 = e;
  }
  // early-out: first shape (scan order) with estimate < 1e-9 (Compressor.cpp:1706-1710)
  const uint32_t z0 = __ballot_sync(0xffffffffu, e0 < 1e-9), z1 = __ballot_sync(0xffffffffu, e1 < 1e-9);
  uint32_t word;
  if (z0 | z1) {
    const int s = z0 ? (__ffs(z0) - 1) : (32 + __ffs(z1) - 1);
    word = (uint32_t)s | (0x8Au << 12) | (1u << 20) | (1u << 23);  // modes {1,3,7}, one shape; bit 23: early-out
    if (lane == 0) sel[t] = word & mode_keep;
    return;
  }
  double be = e0;
  int bi2 = lane;
  if (e1 < be) { be = e1; bi2 = lane + 32; }
  warp_argmin(be, bi2);
  if (!opaque) {
    word = (uint32_t)bi2 | (0xF0u << 12) | (1u << 20) | (1u << 22);  // modes {4,5,6,7}, layout B
    if (lane == 0) sel[t] = word & mode_keep;
    return;
  }
  // ---- three-subset shapes (opaque blocks only)
#pragma unroll 1
  for (int h = 0; h < 2; h++) {
    const double e = estimate_shape<3, NU>(px, plo, phi, lane + 32 * h, pal);  // 4 buckets: 2-bit weights
    if (h == 0) e0 = e; else e1 = e;
  }
  const uint32_t y0 = __ballot_sync(0xffffffffu, e0 < 1e-9), y1 = __ballot_sync(0xffffffffu, e1 < 1e-9);
  if (y0 | y1) {
    const int s = y0 ? (__ffs(y0) - 1) : (32 + __ffs(y1) - 1);
    word = (uint32_t)bi2 | ((uint32_t)s << 6) | (0x05u << 12) | (2u << 20) | (1u << 23);  // modes {0,2}
    if (lane == 0) sel[t] = word & mode_keep;
    return;
  }
  double be3 = e0;
  int bi3 = lane;
  if (e1 < be3) { be3 = e1; bi3 = lane + 32; }
  warp_argmin(be3, bi3);
  word = (uint32_t)bi2 | ((uint32_t)bi3 << 6) | (0xCFu << 12) | (2u << 20);  // all but modes 4,5
  if (lane == 0) sel[t] = word & mode_keep;
}

// ------------------------------------------------------------------ chains
// Decoded description of the chain a (block, slot) pair stands for.
struct Chain {
  int mode, shape, nsub, subset, rot, idx_mode, chain_id;
  bool active;
};

__device__ __noinline__ Chain decode_chain(uint32_t selw, int slot) {
  Chain c;
  c.active = false;
  c.rot = 0; c.idx_mode = 0; c.subset = 0; c.shape = 0; c.nsub = 1; c.mode = 0; c.chain_id = 0;
  if ((selw >> 24) != kTypeNormal) return c;
  const int shape2 = selw & 63, shape3 = (selw >> 6) & 63;
  const uint32_t modes = (selw >> 12) & 0xFF;
  const int nshapes = (selw >> 20) & 3;
  const bool layout_b = (selw >> 22) & 1;
  int mode = -1, subset = 0, si = 0;
  if (!layout_b) {
    if (slot < 3) { mode = 0; subset = slot; si = 1; }
    else if (slot < 6) { mode = 2; subset = slot - 3; si = 1; }
    else if (slot < 8) { mode = 1; subset = slot - 6; }
    else if (slot < 10) { mode = 3; subset = slot - 8; }
    else if (slot < 12) { mode = 7; subset = slot - 10; }
    else if (slot < 14) { mode = 6; si = slot - 12; }
    else return c;
  } else {
    if (slot < 8) { mode = 4; c.rot = slot >> 1; c.idx_mode = slot & 1; }
    else if (slot < 12) { mode = 5; c.rot = slot - 8; }
    else if (slot == 12) { mode = 6; }
    else if (slot < 15) { mode = 7; subset = slot - 13; }
    else return c;
  }
  if (!((modes >> mode) & 1)) return c;
  const int nsub = c_modes[mode].subsets;
  if (si >= nshapes) return c;                       // needs the three-subset shape slot
  if (nsub == 3 && mode == 0 && shape3 >= 16) return c;  // mode 0 has 4 partition bits (Compressor.cpp:1790)
  c.mode = mode;
  c.nsub = nsub;
  c.subset = subset;
  c.shape = nsub == 3 ? shape3 : (nsub == 2 ? shape2 : (si ? shape3 : shape2));
  c.chain_id = (mode == 4) ? (32 + c.rot * 2 + c.idx_mode) : (mode == 5 ? 40 + c.rot : mode * 8 + si * 4 + subset);
  c.active = true;
  return c;
}

// a / c for an integer a in [0, 4080] and a count c in [1, 16], correctly rounded like the
// reference's float division, in three instructions: with rc = RN(1 / c), q = RN(a * rc),
// r = a - q * c (exact in an FMA), RN(q + r * rc) is the correctly rounded quotient (Markstein);
// tests/test_tables.py::test_small_division_exact checks every (a, c) of that domain.
__device__ __forceinline__ float div_small(float a, float c, float rc) {
  const float q = __fmul_rn(a, rc);
  const float r = __fmaf_rn(-q, c, a);
  return __fmaf_rn(r, rc, q);
}

// x / 3, correctly rounded, in three instructions (Markstein's sequence with RN(1/3)): equal to the IEEE
// quotient for EVERY normal float -- checked exhaustively by tests/native/div3_check.c.
__device__ __forceinline__ float div3(float x) {
  const float y = 0.3333333432674407958984375f;  // RN(1 / 3)
  const float q = __fmul_rn(x, y);
  return __fmaf_rn(__fmaf_rn(-q, 3.0f, x), y, q);
}

// IEEE division for the once-per-chain code of bc7_setup, out of line: the inline expansion is ~10
// instructions per site, and that kernel is bound by instruction fetch (see setup_chain)
__device__ __noinline__ float div_cold(float a, float b) { return __fdiv_rn(a, b); }

// Packed f32x2 arithmetic (sm_100 FFMA2 / FMUL2: one instruction, two independent IEEE RN results).
// ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into a fused FFMA2 even with --fmad=false (it never does
// that to the scalar forms), which would break the reference's separately rounded products and sums
// (T4).  So a packed sum is written as RN(a * one + b) with `one` = 1.0f LOADED AT RUN TIME: the
// multiplication by one is exact, the result is the correctly rounded sum, and there is nothing left to
// contract (a product can feed an FFMA2 only as an operand).  A difference p - c is RN(c * -1 + p).
__device__ __forceinline__ float2 sub2(float2 p, float c) { return __ffma2_rn(make_float2(c, c), make_float2(-1.0f, -1.0f), p); }
__device__ __forceinline__ float2 sub2(float2 p, float2 c) { return __ffma2_rn(c, make_float2(-1.0f, -1.0f), p); }
__device__ __forceinline__ float2 add2(float2 a, float2 b, float one) { return __ffma2_rn(a, make_float2(one, one), b); }
__device__ __forceinline__ float2 mul2(float2 a, float b) { return __fmul2_rn(a, make_float2(b, b)); }

struct F4 { float v[4]; };
__device__ __forceinline__ float dot4(const float a[4], const float b[4]) {
  float s = __fmul_rn(a[0], b[0]);  // 0 + x == x
  s = __fadd_rn(s, __fmul_rn(a[1], b[1]));
  s = __fadd_rn(s, __fmul_rn(a[2], b[2]));
  s = __fadd_rn(s, __fmul_rn(a[3], b[3]));
  return s;
}
__device__ __forceinline__ float length4(const float a[4]) { return __fsqrt_rn(dot4(a, a)); }

// single colour lookup (see g_single)
__device__ __forceinline__ uint32_t single_color(int mode, int idx_mode, int npbit, uint32_t pixel, uint32_t &p1,
                                                 uint32_t &p2, int &best_combo) {
  uint32_t best_err = 0xFFFFFFFFu;
  for (int pbi = 0; pbi < npbit; pbi++) {
    uint32_t v1 = 0, v2 = 0, err = 0;
#pragma unroll
    for (int ci = 0; ci < 4; ci++) {
      const uint32_t e = g_single[((((mode * 2 + idx_mode) * 4 + pbi) * 2 + (ci == 3)) << 8) + chan(pixel, ci)];
      v1 |= (e & 0xFF) << (8 * ci);
      v2 |= ((e >> 8) & 0xFF) << (8 * ci);
      const uint32_t d = e >> 16;
      err += d * d;
    }
    if (err < best_err) { best_err = err; best_combo = pbi; p1 = v1; p2 = v2; }
  }
  return best_err;
}

// The same under the non-uniform metric (Compressor.cpp:334-338): the combo's error is the float
// sum_c (float(dist_c) * w_c)^2 with the UN-rotated weights; returns its bit pattern.
__device__ __forceinline__ uint32_t single_color_nu(int mode, int idx_mode, int npbit, uint32_t pixel, uint32_t &p1,
                                                    uint32_t &p2, int &best_combo) {
  float best_err = FLT_MAX;
  for (int pbi = 0; pbi < npbit; pbi++) {
    uint32_t v1 = 0, v2 = 0;
    float err = 0.0f;
#pragma unroll
    for (int ci = 0; ci < 4; ci++) {
      const uint32_t e = g_single[((((mode * 2 + idx_mode) * 4 + pbi) * 2 + (ci == 3)) << 8) + chan(pixel, ci)];
      v1 |= (e & 0xFF) << (8 * ci);
      v2 |= ((e >> 8) & 0xFF) << (8 * ci);
      const float d = __fmul_rn((float)(e >> 16), c_nu_weights[ci]);
      err = __fadd_rn(err, __fmul_rn(d, d));
    }
    if (err < best_err) { best_err = err; best_combo = pbi; p1 = v1; p2 = v2; }
  }
  return __float_as_uint(best_err);
}

#ifdef FASTC_GPU_COUNTERS
#define COUNT_QE(ws, ncalls, npbe) do { atomicAdd(&(ws).counters[0], (unsigned long long)(ncalls)); atomicAdd(&(ws).counters[1], (unsigned long long)(npbe)); } while (0)
#else
#define COUNT_QE(ws, ncalls, npbe) do { } while (0)
#endif

// A thread's column of a [16][kChainThreads] shared-memory plane: element i of the thread's private
// 16-entry array.  bc7_setup keeps its per-chain arrays (the block, the cluster's points, the unique
// points, the k-means accumulators) in such columns: rolled loops index them dynamically, which in
// per-thread arrays means local memory -- and that kernel's L1 could not hold 640 threads' arrays.
constexpr int kChainThreads = 128;
#ifndef FASTC_ALPHA_UNROLL
#define FASTC_ALPHA_UNROLL 4
#endif
constexpr int kAlphaUnroll = FASTC_ALPHA_UNROLL;
struct Col {
  uint32_t *p;
  __device__ __forceinline__ uint32_t operator[](int i) const { return p[i * kChainThreads]; }
  __device__ __forceinline__ void set(int i, uint32_t v) const { p[i * kChainThreads] = v; }
};

// Full QuantizedError over a cluster (returns the integer total; optionally the indices).
__device__ __forceinline__ uint32_t qe_cluster(const Col pts, const Col pix, int n, uint32_t q1, uint32_t q2,
                                               int nbm1, const uint8_t *__restrict__ wtab, unsigned long long *indices) {
  QeEndpoints q;
  qe_prepare(q, q1, q2);
  uint32_t total = 0;
  unsigned long long idx = 0;
  for (int i = 0; i < n; i++) {
    int b;
    total += qe_pixel(q, pts[i], pix[i], nbm1, wtab, &b);
    idx |= (unsigned long long)b << (4 * i);
  }
  if (indices) *indices = idx;
  return total;
}

// QuantizedError under the non-uniform metric `w` (already rotated for the chain): float error per
// candidate bucket, first strict minimum of the (at most two) candidates, float total in pixel
// order (RGBAEndpoints.cpp:254-306).  Returns the total's bit pattern (non-negative floats order
// like their bit patterns, so the callers' comparisons stay integer compares).
__device__ __forceinline__ float qe_bucket_error_nu(const QeEndpoints &q, const int pb[4], int wgt, const float w[4]) {
  float err = 0.0f;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int ip = q.e1[k] + ((q.d[k] * wgt + 32) >> 6);
    const float e = __fmul_rn((float)abs(pb[k] - ip), w[k]);
    err = __fadd_rn(err, __fmul_rn(e, e));
  }
  return err;
}
__device__ __forceinline__ uint32_t qe_cluster_nu(const Col pts, const Col pix, int n, uint32_t q1, uint32_t q2,
                                                  int nbm1, const uint8_t *__restrict__ wtab, const float w[4],
                                                  unsigned long long *indices) {
  QeEndpoints q;
  qe_prepare(q, q1, q2);
  float total = 0.0f;
  unsigned long long idx = 0;
  for (int i = 0; i < n; i++) {
    int pb[4];
#pragma unroll
    for (int k = 0; k < 4; k++) pb[k] = chan(pix[i], k);
    int b = 0;
    float e;
    if (q.den == 0) {
      e = qe_bucket_error_nu(q, pb, 0, w);
    } else {
      int num = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) num += (chan(pts[i], k) - q.e1[k]) * q.d[k];
      const float t = __fmul_rn(__fdiv_rn((float)num, (float)q.den), (float)nbm1);
      int j1 = (int)floorf(t), j2 = (int)ceilf(t);
      j1 = min(max(0, j1), nbm1);
      j2 = min(j2, nbm1);
      e = qe_bucket_error_nu(q, pb, wtab[j1], w);
      b = j1;
      if (j1 + 1 <= j2) {
        const float e2 = qe_bucket_error_nu(q, pb, wtab[j1 + 1], w);
        if (e2 < e) { e = e2; b = j1 + 1; }
      }
    }
    total = __fadd_rn(total, e);
    idx |= (unsigned long long)b << (4 * i);
  }
  if (indices) *indices = idx;
  return __float_as_uint(total);
}

// The fit of one chain: CompressCluster (Compressor.cpp:921-1094) followed by
// OptimizeEndpointsForCluster (:538-630).  pts/pix are this thread's private
// copies of the cluster (n points).  avg/mn/mx are the cluster statistics the
// reference would see (for the rotated mode-4/5 fit these are the STALE ones of
// the un-rotated block, T16).  Returns the total error.
struct FitResult {
  double err64;  // non-uniform metric: the chain's error as the reference's double
  uint32_t err, p1, p2;
  int combo;
  unsigned long long indices;
  bool need_sa;  // true: (p1, p2, combo, err) is the START state of the annealing chain
};

// The fit is split in two so that chains which share a cluster AND an index precision share the
// expensive part (see twin_slot below):
//   fit_core    mode-independent: PCA axis, k-means over the 2^ibits interpolation points,
//               least-squares endpoints (float)                      (Compressor.cpp:936-1077)
//   fit_finish  per mode: single-colour shortcuts, ClampEndpointsToGrid, first error evaluation
struct FitCore {
  int kind;        // 0: all points equal, 1: k-means left one bucket (colour in `single`), 2: p1/p2 valid
  uint32_t single;
  float p1[4], p2[4];
};

// GetPrincipalAxis and the two extreme projections along it: the float endpoints the k-means starts from.
// `upts` is a scratch column for the unique points.
// `one`: see add2.
__device__ __forceinline__ void fit_pca(const Col pts, const Col upts, int n, const float avg[4], float one, float p1[4], float p2[4]) {
  // ---- GetPrincipalAxis (RGBAEndpoints.cpp:327-428)
  float axis[4];
  {
    // unique points; entries past the unique count stay (-1,-1,-1,-1) (T7)
    int nu = 0;
#pragma unroll 1
    for (int i = 0; i < n; i++) {
      const uint32_t p = pts[i];
      bool has = false;
#pragma unroll 1
      for (int j = 0; j < nu; j++) has = has || (upts[j] == p);
      if (!has) upts.set(nu++, p);
    }
    if (nu == 1) {
      axis[0] = axis[1] = axis[2] = axis[3] = 0.0f;
    } else {
      float dir[4];
#pragma unroll
      for (int k = 0; k < 4; k++) dir[k] = (float)(chan(upts[1], k) - chan(upts[0], k));
      {
        const float len = length4(dir);
#pragma unroll
        for (int k = 0; k < 4; k++) dir[k] = div_cold(dir[k], len);
      }
      bool collinear = true;
#pragma unroll 1
      for (int i = 2; i < n; i++) {
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; k++)
          v[k] = i < nu ? (float)(chan(upts[i], k) - chan(upts[0], k)) : __fsub_rn(-1.0f, (float)chan(upts[0], k));
        const double a = fabs((double)dot4(v, dir));
        const double b = (double)length4(v);
        if (fabs(a - b) > 1e-7) { collinear = false; break; }
      }
      if (collinear) {
#pragma unroll
        for (int k = 0; k < 4; k++) axis[k] = dir[k];
      } else {
        // covariance of (pt - avg), divided by 3 (T8): ten running sums (lower triangle), each
        // accumulated over the points in order exactly like the reference's per-entry loops
        // (packed: the products of one row of the triangle share a factor, so a row is one or two FMUL2
        // with that factor broadcast; two of the twelve lanes hold products nobody reads)
        float cs[10];
        {
          const float2 avg01 = make_float2(avg[0], avg[1]), avg23 = make_float2(avg[2], avg[3]);
          float2 c0 = make_float2(0.0f, 0.0f), c1 = c0, c2 = c0, c3 = c0, c4 = c0, c5 = c0;
#pragma unroll 1
          for (int k = 0; k < n; k++) {
            const uint32_t q = pts[k];
            const float2 a01 = sub2(make_float2((float)chan(q, 0), (float)chan(q, 1)), avg01);
            const float2 a23 = sub2(make_float2((float)chan(q, 2), (float)chan(q, 3)), avg23);
            c0 = add2(mul2(a01, a01.x), c0, one);  // (0,0) (1,0)
            c1 = add2(mul2(a01, a01.y), c1, one);  //   -   (1,1)
            c2 = add2(mul2(a23, a01.x), c2, one);  // (2,0) (3,0)
            c3 = add2(mul2(a23, a01.y), c3, one);  // (2,1) (3,1)
            c4 = add2(mul2(a23, a23.x), c4, one);  // (2,2) (3,2)
            c5 = add2(mul2(a23, a23.y), c5, one);  //   -   (3,3)
          }
          cs[0] = c0.x; cs[1] = c0.y; cs[2] = c1.y; cs[3] = c2.x; cs[4] = c3.x; cs[5] = c4.x;
          cs[6] = c2.y; cs[7] = c3.y; cs[8] = c4.y; cs[9] = c5.y;
        }
        float cov[4][4];
        {
          int e = 0;
#pragma unroll
          for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j <= i; j++, e++) {
              cov[i][j] = div3(cs[e]);
              cov[j][i] = cov[i][j];
            }
        }
        // MatrixSquare::PowerMethod (MatrixSquare.h:44-105): <= 4 iterations from (.5,.5,.5,.5)
        float b[4] = {0.5f, 0.5f, 0.5f, 0.5f};
        bool bad = false, fixed = false;
        int it = 0;
#pragma unroll 1
        while (!fixed && ++it < 5) {
          float nbv[4];
#pragma unroll
          for (int j = 0; j < 4; j++) {
            float r = __fmul_rn(cov[j][0], b[0]);
            r = __fadd_rn(r, __fmul_rn(cov[j][1], b[1]));
            r = __fadd_rn(r, __fmul_rn(cov[j][2], b[2]));
            r = __fadd_rn(r, __fmul_rn(cov[j][3], b[3]));
            nbv[j] = r;
          }
          const float len = length4(nbv);
          if ((double)len < 1e-10) {
            if (bad) break;
            b[0] = 1.0f; b[1] = 1.0f;
            const float l2 = length4(b);
#pragma unroll
            for (int k = 0; k < 4; k++) b[k] = div_cold(b[k], l2);
            bad = true;
            continue;
          }
#pragma unroll
          for (int k = 0; k < 4; k++) nbv[k] = div_cold(nbv[k], len);
          if (fabs((double)__fsub_rn(1.0f, dot4(b, nbv))) < 1e-8) fixed = true;
#pragma unroll
          for (int k = 0; k < 4; k++) b[k] = nbv[k];
        }
#pragma unroll
        for (int k = 0; k < 4; k++) axis[k] = b[k];
      }
    }
  }

  // ---- endpoints along the axis (Compressor.cpp:946-959)
  {
    float mindp = FLT_MAX, maxdp = -FLT_MAX;
#pragma unroll 1
    for (int i = 0; i < n; i += 2) {  // two points per trip (an odd cluster's last point twice: min / max do not mind)
      const uint32_t qa = pts[i], qb = pts[min(i + 1, n - 1)];
      float2 m[4];
#pragma unroll
      for (int k = 0; k < 4; k++) m[k] = mul2(sub2(make_float2((float)chan(qa, k), (float)chan(qb, k)), avg[k]), axis[k]);
      const float2 dp = add2(m[3], add2(m[2], add2(m[1], m[0], one), one), one);  // dot4's order
      if (dp.x < mindp) mindp = dp.x;
      if (dp.x > maxdp) maxdp = dp.x;
      if (dp.y < mindp) mindp = dp.y;
      if (dp.y > maxdp) maxdp = dp.y;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      p1[k] = __fadd_rn(avg[k], __fmul_rn(axis[k], mindp));
      p2[k] = __fadd_rn(avg[k], __fmul_rn(axis[k], maxdp));
      p1[k] = (p1[k] < 0.0f) ? 0.0f : ((p1[k] > 255.0f) ? 255.0f : p1[k]);
      p2[k] = (p2[k] < 0.0f) ? 0.0f : ((p2[k] > 255.0f) ? 255.0f : p2[k]);
    }
  }

}

// fit_core: one instantiation per bucket count NB = 2^(index bits).  The centroids live in registers
// (every loop over the buckets is unrolled, so their indices are static), the interpolation fractions
// i / (NB - 1) are compile-time constants, the bucket of each point is a nibble of a register pair, the
// points and the per-bucket sums are columns of shared memory: the k-means, the part of the fit that
// dominates, runs without a single local-memory access.
__host__ __device__ constexpr float ratio_c(int a, int b) { return (float)a / (float)b; }  // IEEE RN division, folded

template <int NB>
__device__ __forceinline__ void fit_core(const Col pts, int n, const float avg[4], bool all_same,
                                      uint32_t (*s_acc)[16][kChainThreads], const float *__restrict__ s_rcp, int tid, FitCore &C) {
  constexpr int nbm1 = NB - 1;
  if (all_same) {  // AllSamePoint -> CompressSingleColor on point 0 (fit_finish)
    C.kind = 0;
    C.single = pts[0];
    return;
  }
  float p1[4], p2[4];
  const float one = s_rcp[1];  // 1.0f the compiler cannot see (add2)
  fit_pca(pts, Col{&s_acc[1][0][tid]}, n, avg, one, p1, p2);  // (the accumulator planes are free until the k-means starts)

  // ---- k-means over the NB interpolation points until a fixed point (:961-1026, T15)
  float cen[NB][4];
#pragma unroll
  for (int i = 0; i < NB; i++) {
    const float s = ratio_c(i, nbm1);
    const float oms = 1.0f - s;  // one rounding, like the reference's runtime subtraction; folded
#pragma unroll
    for (int k = 0; k < 4; k += 2) {
      const float2 c = add2(mul2(make_float2(p1[k], p1[k + 1]), oms), mul2(make_float2(p2[k], p2[k + 1]), s), one);
      cen[i][k] = c.x; cen[i][k + 1] = c.y;
    }
  }
  {
    bool fixed = false;
    int guard = 0;
    while (!fixed && guard++ < 4096) {
      // two points per pass over the centroids: two independent dependency chains (an odd
      // cluster's last point is paired with itself); buckets: nibble i & 7 of blo (i < 8) / bhi
      uint32_t blo = 0, bhi = 0;
#pragma unroll 1
      for (int i = 0; i < n; i += 2) {
        const uint32_t qa = pts[i], qb = pts[min(i + 1, n - 1)];
        float2 pab[4];  // (point a, point b) per channel: the two points go through the packed pipe together
#pragma unroll
        for (int k = 0; k < 4; k++) pab[k] = make_float2((float)chan(qa, k), (float)chan(qb, k));
        int mba = 0, mbb = 0;
        float mda = FLT_MAX, mdb = FLT_MAX;
#pragma unroll
        for (int j = 0; j < NB; j++) {
          float2 m[4];
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const float2 v = sub2(pab[k], cen[j][k]);
            m[k] = __fmul2_rn(v, v);
          }
          const float2 d = add2(m[3], add2(m[2], add2(m[1], m[0], one), one), one);  // dot4's order
          if (d.x < mda) { mda = d.x; mba = j; }
          if (d.y < mdb) { mdb = d.y; mbb = j; }
        }
        // (i is even: the pair shares a word; a self-paired last point ORs the same nibble value into
        // the next, unused, nibble position -- never read, since reads stop at n)
        const uint32_t pair = ((uint32_t)mba | ((uint32_t)mbb << 4)) << (4 * (i & 7));
        if (i < 8) blo |= pair; else bhi |= pair;
      }
      // centroids: bucket sums are exact small integers (<= 16 * 255), so they are accumulated as
      // packed 16-bit pairs per bucket in shared memory -- O(n) instead of the reference's
      // O(n * buckets) scan -- and only the division is done in float, as the reference does
#pragma unroll
      for (int j = 0; j < NB; j++) { s_acc[0][j][tid] = 0; s_acc[1][j][tid] = 0; s_acc[2][j][tid] = 0; }
#pragma unroll 1
      for (int i = 0; i < n; i++) {
        const int b = (int)(((i < 8 ? blo : bhi) >> (4 * (i & 7))) & 15u);
        const uint32_t q = pts[i];
        s_acc[0][b][tid] += q & 0x00FF00FFu;
        s_acc[1][b][tid] += (q >> 8) & 0x00FF00FFu;
        s_acc[2][b][tid] += 1u;
      }
      fixed = true;
#pragma unroll
      for (int j = 0; j < NB; j++) {
        const uint32_t rb = s_acc[0][j][tid], ga = s_acc[1][j][tid];
        const int c = (int)s_acc[2][j][tid];
        float2 s01 = make_float2((float)(rb & 0xFFFFu), (float)(ga & 0xFFFFu)), s23 = make_float2((float)(rb >> 16), (float)(ga >> 16));
        if (c != 0) {  // div_small on two channels at a time
          const float nfc = -(float)c, rc = s_rcp[c];
          const float2 q01 = mul2(s01, rc), q23 = mul2(s23, rc);
          const float2 r01 = __ffma2_rn(q01, make_float2(nfc, nfc), s01), r23 = __ffma2_rn(q23, make_float2(nfc, nfc), s23);
          s01 = __ffma2_rn(r01, make_float2(rc, rc), q01);
          s23 = __ffma2_rn(r23, make_float2(rc, rc), q23);
        }
        const float sum[4] = {s01.x, s01.y, s23.x, s23.y};
#pragma unroll
        for (int k = 0; k < 4; k++) {
          if (!(cen[j][k] == sum[k])) fixed = false;
          cen[j][k] = sum[k];
        }
      }
    }
  }
  int filled = 0;
  float lastc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
  for (int j = 0; j < NB; j++)
    if (s_acc[2][j][tid] > 0) {  // the last pass's counts
      filled++;
#pragma unroll
      for (int k = 0; k < 4; k++) lastc[k] = cen[j][k];
    }
  if (filled == 1) {  // one bucket -> CompressSingleColor on its centroid (:1038-1047, fit_finish)
    C.kind = 1;
    C.single = pack_round(lastc);
    return;
  }

  // ---- least squares endpoints (:1053-1077)
  {
    float ab = 0.0f;
    float2 sq = make_float2(0.0f, 0.0f), abx[4] = {sq, sq, sq, sq};
#pragma unroll
    for (int i = 0; i < NB; i++) {
      const float fn = (float)s_acc[2][i][tid];
      const float2 wab = make_float2(ratio_c(nbm1 - i, nbm1), ratio_c(i, nbm1));  // (a, b)
      const float2 fab = __fmul2_rn(make_float2(fn, fn), wab);                      // (fn * a, fn * b)
      sq = add2(__fmul2_rn(fab, wab), sq, one);                                     // (asq, bsq)
      ab = __fadd_rn(ab, __fmul_rn(fab.x, wab.y));
#pragma unroll
      for (int k = 0; k < 4; k++) abx[k] = add2(mul2(mul2(wab, cen[i][k]), fn), abx[k], one);  // (ax[k], bx[k])
    }
    const float asq = sq.x, bsq = sq.y;
    float ax[4], bx[4];
#pragma unroll
    for (int k = 0; k < 4; k++) { ax[k] = abx[k].x; bx[k] = abx[k].y; }
    const float f = div_cold(1.0f, __fsub_rn(__fmul_rn(asq, bsq), __fmul_rn(ab, ab)));
#pragma unroll
    for (int k = 0; k < 4; k++) {
      C.p1[k] = __fmul_rn(__fsub_rn(__fmul_rn(ax[k], bsq), __fmul_rn(bx[k], ab)), f);
      C.p2[k] = __fmul_rn(__fsub_rn(__fmul_rn(bx[k], asq), __fmul_rn(ax[k], ab)), f);
    }
  }
  C.kind = 2;
}

template <bool NU>
__device__ __forceinline__ void fit_finish(const Ws &ws, const ModeAttr &A, int mode, int idx_mode, int rot, const FitCore &C,
                                        const Col pts, const Col pix, int n, int sa_steps,
                                        const uint8_t *__restrict__ s_w, FitResult &R) {
  R.need_sa = false;
  R.err64 = 0.0;
  const int ibits = idx_mode == 0 ? A.index_bits : A.alpha_index_bits;
  const int nbm1 = (1 << ibits) - 1;
  const uint8_t *wtab = s_w + 16 * (ibits - 1);
  const int npbit = A.pbit == kPbitShared ? 2 : (A.pbit == kPbitPerEndpoint ? 4 : 1);
  const uint32_t qm = quant_mask(A);
  if (C.kind != 2) {  // CompressSingleColor (:252-353) on point 0 / on the only bucket's centroid
    int combo = 0;
    uint32_t q1 = 0, q2 = 0;
    uint32_t e;
    if constexpr (NU) {
      e = single_color_nu(mode, idx_mode, npbit, C.single, q1, q2, combo);
      R.err64 = (double)n * (double)__uint_as_float(e);  // cluster.GetNumPoints() * CompressSingleColor(...)
    } else {
      e = single_color(mode, idx_mode, npbit, C.single, q1, q2, combo);
    }
    R.err = (uint32_t)n * e;
    R.p1 = q1; R.p2 = q2; R.combo = combo;
    R.indices = 0x1111111111111111ull;
    return;
  }
  float p1[4], p2[4];
#pragma unroll
  for (int k = 0; k < 4; k++) { p1[k] = C.p1[k]; p2[k] = C.p2[k]; }

  // ---- ClampEndpointsToGrid (:212-250)
  uint32_t c1, c2;  // current endpoints, integer bytes from here on
  int combo = 0;
  {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      p1[k] = (p1[k] < 0.0f) ? 0.0f : ((p1[k] > 255.0f) ? 255.0f : p1[k]);
      p2[k] = (p2[k] < 0.0f) ? 0.0f : ((p2[k] > 255.0f) ? 255.0f : p2[k]);
    }
    const uint32_t r1 = pack_round(p1), r2 = pack_round(p2);
    float md = FLT_MAX;
    c1 = c2 = 0;
#pragma unroll 1
    for (int i = 0; i < npbit; i++) {
      int pb0, pb1;
      pbit_combo(A.pbit, i, pb0, pb1);
      const uint32_t q1 = to_pixel_b(r1, qm, pb0), q2 = to_pixel_b(r2, qm, pb1);
      float d1[4], d2[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        d1[k] = __fsub_rn((float)chan(q1, k), p1[k]);
        d2[k] = __fsub_rn((float)chan(q2, k), p2[k]);
      }
      const float dist = __fadd_rn(dot4(d1, d1), dot4(d2, d2));
      if (dist < md) { md = dist; c1 = q1; c2 = q2; combo = i; }
    }
  }

  // ---- OptimizeEndpointsForCluster (:538-630), integer state.
  // Endpoints are bytes on the mode's grid; a step moves every channel by one
  // grid step in a random direction (p-bit modes: the p-bit always flips).
  // Reference quirk: QuantizedError always receives a non-NULL p-bit pair -- for the modes
  // WITHOUT p-bits GetPBitCombo() returns {0,0} (CompressionMode.h:244-251), so inside the error
  // evaluation (and only there) their endpoints are quantised as if a p-bit of 0 followed the
  // colour bits.  The endpoint state itself and Pack quantise without a p-bit.
  const bool has_pbit = A.pbit != kPbitNone;
  const uint32_t cur1 = c1, cur2 = c2;  // on the mode's grid already (quantisation is idempotent)
  // what QuantizedError sees: the same, except that modes without p-bits are quantised with a p-bit of 0
  const uint32_t qe1 = has_pbit ? c1 : to_pixel_b(c1, qm, 0), qe2 = has_pbit ? c2 : to_pixel_b(c2, qm, 0);
  // one evaluation serves both outcomes: its error starts the annealing chain, its indices are
  // the result when there is no annealing (same quirky quantisation either way)
  unsigned long long indices;
  uint32_t cur_err;
  if constexpr (NU) {
    float w[4];
    nu_metric(A.rotation ? rot : 0, w);
    cur_err = qe_cluster_nu(pts, pix, n, qe1, qe2, nbm1, wtab, w, &indices);
    R.err64 = (double)__uint_as_float(cur_err);
  } else {
    cur_err = qe_cluster(pts, pix, n, qe1, qe2, nbm1, wtab, &indices);
  }
  COUNT_QE(ws, 1, 0);
  if (sa_steps > 0 && cur_err > 0) {  // hand over to bc7_anneal
    R.need_sa = true;
    R.err = cur_err;
    R.p1 = cur1; R.p2 = cur2; R.combo = combo;
    R.indices = indices;  // of the start state: final if annealing never improves on it
    return;
  }
  R.err = cur_err;
  R.p1 = cur1; R.p2 = cur2; R.combo = combo;
  R.indices = indices;
}

// Sort key of an annealing chain: ((index bits - 2) * 17 + cluster size) * 4 + expected-length level.
// A chain runs until 50 (-q) consecutive steps bring no new best, so its length is unknown in
// advance, but it correlates strongly with the start error per pixel (measured on the SURVEY 8d
// image at -q 50: < 128 per pixel: 57 steps on average, at most ~270; 128..4096: ~170; above --
// the mode 4/5 fits, whose error carries the stale-alpha term of T16 and which make up more than
// half of all annealing steps -- 245 to 586 on average, the longest > 2500 steps: 10+ ms for one
// lane, a sizeable part of the whole kernel on a 1/8 shard).  Chains are queued longest-expected
// first within their (precision, size) bin, in eight levels, so that the later a chain is
// dequeued the shorter its worst case: the kernel's tail shrinks.
constexpr int kLenLevels = 8;
constexpr int kSortKeys = 51 * kLenLevels;
__device__ __forceinline__ int sort_key(int ibits, int n, uint32_t err) {
  // start error per pixel: < 128 | < 4096 | < 8192 | < 16384 | < 32768 | < 49152 | < 65536 | above
  // (an ordering heuristic, not a result: the quotient may be off by one at a level's edge, but bc7_setup's
  // histogram and bc7_scatter evaluate this same function, which is all that has to agree)
  const uint32_t e = __float2uint_rz(__fmul_rn((float)err, __frcp_rn((float)max(n, 1))));
  const int lvl = (e >= 128u) + (e >= 4096u) + (e >= 8192u) + (e >= 16384u) + (e >= 32768u) + (e >= 49152u) + (e >= 65536u);
#ifdef FASTC_SORT_SIZE_MAJOR
  return ((ibits - 2) * 17 + n) * kLenLevels + lvl;
#else
  // level-major: EVERY expected-long chain of a precision class is dequeued before any expected-short
  // one (size-major order left the long chains of the small clusters for the end of the queue: the
  // persistent kernel's tail); within a level the clusters still come sorted by size, so the lanes
  // of a warp keep fitting clusters of (nearly) one size
  return ((ibits - 2) * kLenLevels + lvl) * 17 + n;
#endif
}
__device__ __forceinline__ void sort_key_parts(int k, int &cls, int &lvl, int &n) {
#ifdef FASTC_SORT_SIZE_MAJOR
  const int base = k / kLenLevels;
  lvl = k % kLenLevels; cls = base / 17; n = base % 17;
#else
  const int base = k / 17;
  n = k % 17; cls = base / kLenLevels; lvl = base % kLenLevels;
#endif
}
// mean steps of a chain of each level (same measurement), for the work estimate of bc7_bin_offsets
__constant__ float c_level_steps[kLenLevels] = {57.0f, 170.0f, 245.0f, 297.0f, 372.0f, 437.0f, 500.0f, 586.0f};
// bins layout (uint32 words)
constexpr int kBinCount = 0;      // [kSortKeys] chains per key (histogram, bc7_setup)
constexpr int kBinOffset = 512;   // [kSortKeys] start of each key's range in the sorted order
constexpr int kBinCursor = 1024;  // [kSortKeys] scatter cursors
constexpr int kBinTotal = 1536;
constexpr int kBinFetch = 1537;   // [3] fetch cursor of precision class c (= index bits - 2)
constexpr int kBinEnd = 1540;     // [3] end of class c's region
constexpr int kBinHome = 1544;    // [3] first CTA whose home class is c or lower
constexpr int kBinTailCount = 1600;  // chains handed to bc7_anneal_tail
constexpr int kBinTailFetch = 1601;  // its fetch cursor
constexpr int kBinWords = 2048;
constexpr int kTailChains = 16;      // a warp of bc7_anneal hands its chains over once it is down to this many
constexpr uint32_t kTailCap = 1u << 18;  // capacity of the hand-over list
constexpr uint32_t kTailMaxBlocks = 1000000u;  // submissions below this size use the hand-over (see launch_bc7)

// (under the non-uniform metric R.err is a float's bit pattern: the length predictor takes its value)
__device__ __forceinline__ uint32_t sort_error(const Ws &ws, uint32_t err) {
  return ws.err64 ? (uint32_t)__uint_as_float(err) : err;
}
__device__ __forceinline__ void write_state(const Ws &ws, uint32_t gid, uint32_t mask, const Chain &c, int n,
                                            const FitResult &R, uint32_t rng, uint32_t alpha_err, uint32_t abytes) {
  uint32_t *st = ws.states + (size_t)gid * kStateWords;
  st[1] = R.p1; st[2] = R.p2; st[3] = R.err; st[4] = rng; st[5] = alpha_err; st[6] = abytes;
  atomicOr(&ws.sa_mask[gid / kSlots], 1u << (gid % kSlots));
  st[0] = mask | ((uint32_t)c.mode << 16) | ((uint32_t)c.rot << 19) | ((uint32_t)c.idx_mode << 21) |
          ((uint32_t)R.combo << 22) | ((uint32_t)n << 24) | (1u << 31);
  // histogram by (index precision, cluster size): bc7_anneal runs chains sorted by that key so the
  // lanes of a warp build palettes of the same length and walk the same number of pixels
  const int ibits = c.idx_mode == 0 ? c_modes[c.mode].index_bits : c_modes[c.mode].alpha_index_bits;
  {
    // one atomic per distinct key among the lanes that arrive here together
    const int key = sort_key(ibits, n, sort_error(ws, R.err));
    const unsigned act = __activemask();
    const unsigned peers = __match_any_sync(act, key);
    if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&ws.bins[kBinCount + key], (uint32_t)__popc(peers));
  }
}

// Chains that fit the SAME cluster with the SAME index precision share everything up to the
// least-squares endpoints (fit_core); only the grid clamp and the first evaluation depend on the
// mode.  The first chain of such a pair does the shared part once and finishes both:
//   opaque layout : mode 3 subset s -> mode 7 subset s (slots 8,9 -> 10,11; same two-subset shape,
//                   2-bit indices), mode 6 on shape slot 0 -> mode 6 on shape slot 1 (12 -> 13: the
//                   reference fits mode 6 once per candidate shape, T12; only the RNG stream differs)
//   alpha layout  : mode 4 rotation r, index mode 0 -> mode 5 rotation r (slots 0,2,4,6 -> 8..11:
//                   same rotated points, 2-bit colour indices; the scalar alpha fits differ and
//                   run per chain)
__device__ __forceinline__ int twin_slot(int layout_b, int slot) {
  if (!layout_b) return (slot == 8 || slot == 9) ? slot + 2 : (slot == 12 ? 13 : -1);
  return (slot < 8 && !(slot & 1)) ? 8 + (slot >> 1) : -1;
}
// Is the chain in `slot` a twin whose primary chain is live (and therefore fits it)?  A primary is live
// exactly when its mode is enabled: modes 3, 6 and 4 have no other condition in decode_chain (their
// shape slot always exists).
__device__ __forceinline__ bool fitted_by_primary(uint32_t selw, int slot) {
  const uint32_t modes = (selw >> 12) & 0xFF;
  if (!((selw >> 22) & 1)) return ((slot == 10 || slot == 11) && ((modes >> 3) & 1)) || (slot == 13 && ((modes >> 6) & 1));
  return slot >= 8 && slot < 12 && ((modes >> 4) & 1);
}

// Per-mode tail of one chain: ClampEndpointsToGrid + first evaluation (fit_finish), the scalar alpha fit
// of modes 4/5, and the result / annealing start state.
template <bool NU, bool ROT>
__device__ __forceinline__ void setup_variant(const Ws &ws, const Chain &c, const ModeAttr &A, const FitCore &core,
                                              const Col pts, const Col pix, int n, uint32_t mask,
                                              int sa_steps, const uint8_t *__restrict__ s_w, const float *__restrict__ s_rcp,
                                              uint32_t (*s_acc)[16][kChainThreads], int tid, uint32_t gid, uint32_t rng,
                                              uint32_t *res, float amin, float amax) {
  FitResult R;
  fit_finish<NU>(ws, A, c.mode, c.idx_mode, c.rot, core, pts, pix, n, sa_steps, s_w, R);  // the one call site
  res[4] = (uint32_t)R.indices; res[5] = (uint32_t)(R.indices >> 32);
  if constexpr (!ROT) {
    if (R.need_sa) {
      write_state(ws, gid, mask, c, n, R, rng, 0, 0);
      return;
    }
    res[0] = R.err; res[1] = R.p1; res[2] = R.p2; res[3] = (uint32_t)R.combo;
    res[4] = (uint32_t)R.indices; res[5] = (uint32_t)(R.indices >> 32);
    if constexpr (NU) ws.err64[gid] = R.err64;
    return;
  } else {

  const int abits = c.idx_mode == 0 ? A.alpha_index_bits : A.index_bits;
  const int nba = 1 << abits;
  const uint8_t *wa = s_w + 16 * (abits - 1);
  float a1 = amin, a2 = amax;
  uint32_t alpha_err = 0;
  // non-uniform metric: the alpha error is the double sum of the floats (weight * |difference|)^2,
  // weight = the rotated metric's alpha entry (Compressor.cpp:712, :756-760, :905-915)
  double alpha_err64 = 0.0;
  float wgt_a = 1.0f;
  if constexpr (NU) {
    float w[4];
    nu_metric(c.rot, w);
    wgt_a = w[3];
  }
  unsigned long long aidx = 0;
  if (a1 == a2) {
    const int a1be = (int)a1;
    if (c.mode == 5) {
      aidx = 0;
      alpha_err = 0;
    } else {
      const uint8_t *t1 = c_opt6 + 6 * a1be;
      if (t1[0]) {
        a1 = (float)((t1[4] << 2) | (t1[1] >> 4));
        a2 = (float)((t1[5] << 2) | (t1[1] >> 4));
      } else {
        a1 = (float)((t1[1] << 2) | (t1[1] >> 4));
        a2 = (float)((t1[2] << 2) | (t1[1] >> 4));
      }
      const int ai = c.idx_mode == 1 ? 1 : 2;
      aidx = ai * 0x1111111111111111ull;
      const int w1 = wa[ai], w0 = 64 - w1;
      const int ip = (((int)a1 * w0 + (int)a2 * w1 + 32) >> 6) & 0xFF;
      const int d = a1be > ip ? a1be - ip : ip - a1be;
      alpha_err = 16u * (uint32_t)(d * d);
      if constexpr (NU) {
        const float px = __fmul_rn(wgt_a, (float)d);
        alpha_err64 = (double)__fmul_rn(16.0f, __fmul_rn(px, px));
      }
    }
  } else {
    // scalar k-means over the alpha interpolation points (:770-842).  The alpha values are re-read
    // from the block (`pix`), the interpolation points live in the lane's column of accumulator
    // plane 1 (float bits), the buckets in the nibbles of a register pair.
    const int ash = c.rot == 0 ? 24 : 8 * (c.rot - 1);  // the channel the rotation put into alpha
    const Col vals{&s_acc[1][0][tid]};
    uint32_t blo = 0, bhi = 0;
#pragma unroll 1
    for (int i = 0; i < nba; i++)
      vals.set(i, __float_as_uint(__fadd_rn(amin, __fmul_rn(div_cold((float)i, (float)(nba - 1)), __fsub_rn(amax, amin)))));
    // (one copy of the assignment pass: the reference's initial assignment is the loop's with every
    // previous bucket 0; the pass after the update that found the fixed point still runs, as there)
    bool fixed = false;
    int guard = 0;
    for (;;) {
      uint32_t nlo = 0, nhi = 0;
#pragma unroll 1
      for (int i = 0; i < 16; i += 2) {  // two pixels per trip through the packed pipe
        const float2 av = make_float2((float)((pix[i] >> ash) & 0xFFu), (float)((pix[i + 1] >> ash) & 0xFFu));
        float mda = 255.0f, mdb = 255.0f;
        const uint32_t old = (i < 8 ? blo : bhi) >> (4 * (i & 7));
        uint32_t ba = old & 15u, bb = (old >> 4) & 15u;  // reference keeps the previous bucket when nothing is closer than 255
#pragma unroll kAlphaUnroll
        for (int j = 0; j < nba; j++) {  // (nba is 4 or 8)
          const float2 d = sub2(av, __uint_as_float(vals[j]));
          if (fabsf(d.x) < mda) { mda = fabsf(d.x); ba = (uint32_t)j; }
          if (fabsf(d.y) < mdb) { mdb = fabsf(d.y); bb = (uint32_t)j; }
        }
        const uint32_t pair = (ba | (bb << 4)) << (4 * (i & 7));
        if (i < 8) nlo |= pair; else nhi |= pair;
      }
      blo = nlo; bhi = nhi;
      if (fixed || guard++ >= 4096) break;
      fixed = true;
      // bucket sums / counts are small integers (exact in any order): one pass over the pixels into
      // the lane's shared accumulators instead of the reference's bucket x pixel scan (:790-806)
#pragma unroll 1
      for (int i = 0; i < nba; i++) { s_acc[0][i][tid] = 0; s_acc[2][i][tid] = 0; }
#pragma unroll 1
      for (int j = 0; j < 16; j++) {
        const int b = (int)(((j < 8 ? blo : bhi) >> (4 * (j & 7))) & 15u);
        s_acc[0][b][tid] += (pix[j] >> ash) & 0xFFu;
        s_acc[2][b][tid] += 1u;
      }
#pragma unroll 1
      for (int i = 0; i < nba; i++) {
        const int cnt = (int)s_acc[2][i][tid];
        float s = (float)s_acc[0][i][tid];
        if (cnt > 0) s = div_small(s, (float)cnt, s_rcp[cnt]);
        fixed = fixed && (s == __uint_as_float(vals[i]));
        vals.set(i, __float_as_uint(s));
      }
    }
    float asq = 0.0f, bsq = 0.0f, ab = 0.0f, ax = 0.0f, bx = 0.0f;
    const float fb = (float)(nba - 1);
#pragma unroll 1
    for (int i = 0; i < nba; i++) {
      const float a = div_cold((float)(nba - 1 - i), fb), b = div_cold((float)i, fb);
      const float nn = (float)s_acc[2][i][tid], x = __uint_as_float(vals[i]);  // the last pass's counts
      asq = __fadd_rn(asq, __fmul_rn(__fmul_rn(nn, a), a));
      bsq = __fadd_rn(bsq, __fmul_rn(__fmul_rn(nn, b), b));
      ab = __fadd_rn(ab, __fmul_rn(__fmul_rn(nn, a), b));
      ax = __fadd_rn(ax, __fmul_rn(__fmul_rn(x, a), nn));
      bx = __fadd_rn(bx, __fmul_rn(__fmul_rn(x, b), nn));
    }
    const float f = div_cold(1.0f, __fsub_rn(__fmul_rn(asq, bsq), __fmul_rn(ab, ab)));
    a1 = __fmul_rn(f, __fsub_rn(__fmul_rn(ax, bsq), __fmul_rn(bx, ab)));
    a2 = __fmul_rn(f, __fsub_rn(__fmul_rn(bx, asq), __fmul_rn(ax, ab)));
    // std::min(255.0f, std::max(0.0f, a)) -- NaN maps to 0 through std::max's argument order
    a1 = (a1 > 0.0f) ? a1 : 0.0f; a1 = (a1 < 255.0f) ? a1 : 255.0f;
    a2 = (a2 > 0.0f) ? a2 : 0.0f; a2 = (a2 < 255.0f) ? a2 : 255.0f;
    const uint32_t qmask8 = (0xFF00u >> A.alpha_bits) & 0xFF;
    const int a1b = (int)quantize_channel((uint32_t)(int)a1, qmask8, -1);  // uint8(a1): truncation
    const int a2b = (int)quantize_channel((uint32_t)(int)a2, qmask8, -1);
#pragma unroll 1
    for (int i = 0; i < 16; i++) {
      const int val = (int)((pix[i] >> ash) & 0xFFu);
      int me = 0x7fffffff, bb = 0;
#pragma unroll 1
      for (int j = 0; j < nba; j++) {
        const int w1 = wa[j], w0 = 64 - w1;
        const int ip = ((a1b * w0 + a2b * w1 + 32) >> 6) & 0xFF;
        const int d = val - ip;
        const int e = d * d;
        if (e < me) { me = e; bb = j; }
      }
      alpha_err += (uint32_t)me;
      if constexpr (NU) {
        const float px = __fmul_rn(wgt_a, __fsqrt_rn((float)me));  // |difference| (exact: a perfect square < 2^16)
        alpha_err64 = __dadd_rn(alpha_err64, (double)__fmul_rn(px, px));
      }
      aidx |= (unsigned long long)bb << (4 * i);
    }
  }
  // endpoints handed to Pack: rgb from the fit, alpha = the float a1/a2 (Pack rounds them)
  res[6] = (uint32_t)aidx; res[7] = (uint32_t)(aidx >> 32);
  if (R.need_sa) {
    if constexpr (NU) ws.err64[gid] = alpha_err64;  // bc7_anneal adds the colour fit's error
    write_state(ws, gid, 0xFFFFu, c, 16, R, rng, alpha_err, round_byte(a1) | (round_byte(a2) << 8));
    return;
  }
  const uint32_t e1 = (R.p1 & 0x00FFFFFFu) | (round_byte(a1) << 24);
  const uint32_t e2 = (R.p2 & 0x00FFFFFFu) | (round_byte(a2) << 24);
  res[0] = R.err + alpha_err; res[1] = e1; res[2] = e2; res[3] = 0;
  res[4] = (uint32_t)R.indices; res[5] = (uint32_t)(R.indices >> 32);
  if constexpr (NU) ws.err64[gid] = __dadd_rn(R.err64, alpha_err64);
  }
}

// One endpoint-fit chain: cluster statistics, CompressCluster's start (fit_cluster), and for
// modes 4/5 the scalar alpha fit.  Writes either the finished result or the start state of the
// annealing chain.
template <bool NU, int IB, bool ROT>
__device__ __forceinline__ void setup_chain(const uint32_t *__restrict__ img, uint32_t width, uint32_t blocks_x,
                                            uint32_t first_block, const Ws &ws, int sa_steps, uint64_t seed,
                                            uint32_t block_index_base, uint32_t t, int slot,
                                            const uint8_t *__restrict__ s_w, const float *__restrict__ s_rcp,
                                            uint32_t (*s_acc)[16][kChainThreads], uint32_t (*s_blk)[kChainThreads],
                                            uint32_t (*s_pts)[kChainThreads], int tid) {
  const uint32_t selw = ws.sel[t];
  const Chain c = decode_chain(selw, slot), &c0 = c;
  if (!c.active) return;  // (the caller's list only holds live chains)
  int twin = twin_slot((selw >> 22) & 1, slot);
  Chain ctwin = c;
  if (twin >= 0) {
    ctwin = decode_chain(selw, twin);
    if (!ctwin.active) twin = -1;
  }

  // The kernel's code is large (the resident warps are in different phases: "no instruction" is one of
  // its top stalls), so the loops over the points stay rolled.  Rolled loops index their arrays
  // dynamically; so that this does not mean local memory, the block and the cluster's points are
  // columns of shared memory (Col).  The error pixels (`pix`) are the points themselves, except for
  // modes 4/5 (n == 16), where they are the block.
  {
    uint32_t blk[16];
    load_block(img, width, blocks_x, first_block + t, blk);
#pragma unroll
    for (int i = 0; i < 16; i++) s_blk[i][tid] = blk[i];
  }
  const Col blk{&s_blk[0][tid]}, pts{&s_pts[0][tid]};

  int n = 0;
  uint32_t mask = 0xFFFFu;  // pixels of the subset
  float sum[4] = {0, 0, 0, 0};
  uint32_t mn = 0xFFFFFFFFu, mx = 0;
  float amin = FLT_MAX, amax = -FLT_MAX;
  if constexpr (!ROT) {
    // Cluster of this chain: points in raster order of the subset (m_PointMap).
    if (c.nsub == 2) {
      mask = c.subset ? (uint32_t)c_shape2[c.shape] : (~(uint32_t)c_shape2[c.shape] & 0xFFFFu);
    } else if (c.nsub == 3) {
      const uint32_t m3 = c_shape3[c.shape], lo = m3 & 0x55555555u, hi = (m3 >> 1) & 0x55555555u;
      const uint32_t sel = c.subset == 0 ? ~(lo | hi) & 0x55555555u : (c.subset == 1 ? lo & ~hi : hi & ~lo);
      mask = 0;  // one bit per 2-bit field
#pragma unroll 1
      for (int i = 0; i < 16; i++) mask |= ((sel >> (2 * i)) & 1u) << i;
    }
#pragma unroll 1
    for (int i = 0; i < 16; i++) {
      if ((mask >> i) & 1u) {
        const uint32_t p = blk[i];
        pts.set(n++, p);
#pragma unroll
        for (int k = 0; k < 4; k++) sum[k] = __fadd_rn(sum[k], (float)chan(p, k));  // exact integers
        mn = __vminu4(mn, p);
        mx = __vmaxu4(mx, p);
      }
    }
  } else {
    // ---- modes 4/5: CompressCluster alpha variant (Compressor.cpp:632-919), the whole block.
    // Points are rotated and their alpha forced to 255, but avg / bounds / error
    // pixels stay those of the original block (T16).
    n = 16;
#pragma unroll 1
    for (int i = 0; i < 16; i++) {
      const uint32_t p = blk[i];
#pragma unroll
      for (int k = 0; k < 4; k++) sum[k] = __fadd_rn(sum[k], (float)chan(p, k));  // exact integers
      mn = __vminu4(mn, p);
      mx = __vmaxu4(mx, p);
      const uint32_t a = c.rot == 0 ? (p >> 24) : chan(p, c.rot - 1);
      uint32_t q = p;
      if (c.rot) q = (p & ~(0xFFu << (8 * (c.rot - 1)))) | ((p >> 24) << (8 * (c.rot - 1)));  // channel <- old alpha
      pts.set(i, q | 0xFF000000u);
      amin = fminf(amin, (float)a);
      amax = fmaxf(amax, (float)a);
    }
  }
  float avg[4];  // channel sums <= 16 * 255 over a count <= 16: div_small's domain
#pragma unroll
  for (int k = 0; k < 4; k++) avg[k] = div_small(sum[k], (float)n, s_rcp[n]);
  const bool all_same = mn == mx;
  const uint32_t gblock = block_index_base + first_block + t;
  const Col pix = ROT ? blk : pts;
  // the expensive, mode-independent part runs once for the chain and its twin
  FitCore core;
  fit_core<(1 << IB)>(pts, n, avg, all_same, s_acc, s_rcp, tid, core);
  const int nvariants = twin >= 0 ? 2 : 1;
#pragma unroll 1
  for (int variant = 0; variant < nvariants; variant++) {
    const int vslot = variant == 0 ? slot : twin;
    const Chain &c = variant == 0 ? c0 : ctwin;
    const ModeAttr A = c_modes[c.mode];
    const uint32_t gid = t * kSlots + vslot;
    const uint32_t rng = chain_seed(seed, gblock, (uint32_t)c.chain_id);
    uint32_t *res = ws.results + (size_t)gid * kResWords;
    setup_variant<NU, ROT>(ws, c, A, core, pts, pix, n, mask, sa_steps, s_w, s_rcp, s_acc, tid, gid, rng, res, amin, amax);
  }
}

// Number of pixels in the subset a chain fits (its cluster size), from the partition tables.
__device__ __forceinline__ int chain_pixels(const Chain &c) {
  if (c.nsub == 1) return 16;
  if (c.nsub == 2) {
    const int ones = __popc((uint32_t)c_shape2[c.shape]);
    return c.subset ? ones : 16 - ones;
  }
  const uint32_t m = c_shape3[c.shape], lo = m & 0x55555555u, hi = (m >> 1) & 0x55555555u;
  return c.subset == 0 ? 16 - __popc(lo | hi) : (c.subset == 1 ? __popc(lo & ~hi) : __popc(hi & ~lo));
}

// One kernel per index precision IB (2, 3, 4 bits = 4, 8, 16 buckets), so that the bucket count is a
// compile-time constant of the fit (fit_core<NB>: centroids in registers) while each kernel still holds
// a single copy of the code.  A CTA owns 128 consecutive blocks.  It first compacts the live chains OF
// ITS PRECISION (from all 16 slots of those blocks) into a shared list sorted by cluster size, then its
// lanes walk the list: no lane idles on a dead slot (mode 0 is only tried for a quarter of the shapes,
// solid / transparent blocks have no chains), and the lanes of a warp fit clusters of (nearly) the same
// size, so the loops over the points run with uniform trip counts.
// Index precision of the chain in a slot, per layout: 2 bits per slot, value = IB - 2, 3 = none
//   opaque: 0-2 mode 0 (3), 3-5 mode 2 (2), 6-7 mode 1 (3), 8-9 mode 3 (2), 10-11 mode 7 (2), 12-13 mode 6 (4)
//   alpha : 0-7 mode 4 rotation r, index mode 0 / 1 (2 / 3), 8-11 mode 5 (2), 12 mode 6 (4), 13-14 mode 7 (2)
__device__ __forceinline__ int slot_class(int layout_b, int slot) {
  const uint32_t opaque = 0xFA005015u;  // slots 15..0, two bits each
  const uint32_t alpha = 0xC2004444u;
  return (int)(((layout_b ? alpha : opaque) >> (2 * slot)) & 3u);
}
constexpr int kClassList = 6 * kChainThreads;  // at most six primary chains of one precision per block
#ifndef FASTC_SETUP_CTAS
#define FASTC_SETUP_CTAS 5
#endif
#ifndef FASTC_SETUP_CTAS16
#define FASTC_SETUP_CTAS16 4
#endif
template <bool NU, int IB, bool ROT>
__global__ void __launch_bounds__(kChainThreads, IB == 4 ? FASTC_SETUP_CTAS16 : FASTC_SETUP_CTAS)
bc7_setup(const uint32_t *__restrict__ img, uint32_t width, uint32_t blocks_x, uint32_t first_block,
          uint32_t num_blocks, Ws ws, int sa_steps, uint64_t seed, uint32_t block_index_base) {
  __shared__ uint8_t s_w[64];
  __shared__ uint32_t s_acc[3][16][kChainThreads];  // k-means accumulators (before that: the unique points)
  __shared__ uint32_t s_blk[16][kChainThreads];     // the chain's block
  __shared__ uint32_t s_pts[16][kChainThreads];     // the chain's cluster
  __shared__ uint16_t s_list[kClassList];
  __shared__ uint32_t s_hist[17], s_cur[17];
  __shared__ float s_rcp[17];  // RN(1 / c) for the bucket counts 1..16 (div_small)
  const int tid = threadIdx.x;
  if (tid < 64) s_w[tid] = c_weight[tid];
  if (tid < 17) { s_hist[tid] = 0; s_cur[tid] = 0; s_rcp[tid] = tid ? __frcp_rn((float)tid) : 0.0f; }
  __syncthreads();
  const uint32_t tile = blockIdx.x;
  const uint32_t t = tile * kChainThreads + tid;
  const uint32_t selw = t < num_blocks ? ws.sel[t] : (uint32_t)kTypeSolid << 24;
  const int layout_b = (selw >> 22) & 1;
  // pass 1: histogram of the live chains of this precision by cluster size; `live` = their slots
  uint32_t live = 0, sizes = 0;
  int nlive = 0;
  if ((selw >> 24) == kTypeNormal) {
#pragma unroll 1
    for (int k = 0; k < 15; k++) {
      // a twin is fitted by its primary chain's lane (see twin_slot)
      if (slot_class(layout_b, k) != IB - 2 || fitted_by_primary(selw, k)) continue;
      if (ROT != (layout_b && k < 12)) continue;  // the rotation fits (modes 4/5) have kernels of their own
      const Chain c = decode_chain(selw, k);
      if (c.active) {
        const int sz = chain_pixels(c);
        live |= 1u << k;
        sizes = (sizes << 5) | (uint32_t)sz;  // at most six live chains: 30 bits, first chain in the top field
        nlive++;
        atomicAdd(&s_hist[sz], 1u);
      }
    }
  }
  __syncthreads();
  if (tid == 0) {  // start offsets, largest clusters first
    uint32_t off = 0;
#pragma unroll 1
    for (int n = 16; n >= 0; n--) { const uint32_t c = s_hist[n]; s_hist[n] = off; off += c; }
    s_cur[0] = off;  // total
  }
  __syncthreads();
  while (live) {
    const int k = __ffs(live) - 1;
    live &= live - 1;
    const int sz = (int)((sizes >> (5 * --nlive)) & 31u);
    s_list[s_hist[sz] + atomicAdd(&s_cur[sz], 1u)] = (uint16_t)((tid << 4) | k);
  }
  __syncthreads();
  const uint32_t total = s_cur[0];  // s_cur[0] is only touched by size-0 chains, which do not exist
  // pass 2: one chain per lane and trip
  for (uint32_t e = tid; e < total; e += kChainThreads) {
    const uint32_t entry = s_list[e];
    setup_chain<NU, IB, ROT>(img, width, blocks_x, first_block, ws, sa_steps, seed, block_index_base,
                        tile * kChainThreads + (entry >> 4), (int)(entry & 15), s_w, s_rcp, s_acc, s_blk, s_pts, tid);
  }
}

// ------------------------------------------------------------------ annealing
// The sorted order is by descending key (see sort_key and the bins layout above): 4-bit-index
// chains first, within a precision class large clusters first, within a bin long chains first.
//
// bc7_anneal gives every CTA a HOME class, in proportion to the class's estimated work, so that
// the lanes of a warp build palettes of the same length (and, for the 4-bit class, walk the same
// 16 pixels); a lane whose home queue is dry steals from the other classes.
__global__ void bc7_bin_offsets(uint32_t *bins, uint32_t grid_ctas) {
  uint32_t off = 0;
  float work[3] = {0.0f, 0.0f, 0.0f};
  for (int k = kSortKeys - 1; k >= 0; k--) {
    int cls, lvl, n;
    sort_key_parts(k, cls, lvl, n);
    if (n == 16 && lvl == kLenLevels - 1) bins[kBinFetch + cls] = off;  // the class's region starts with its top key
    bins[kBinOffset + k] = off;
    off += bins[kBinCount + k];
    bins[kBinCursor + k] = 0;
    if (n == 0 && lvl == 0) bins[kBinEnd + cls] = off;
    // expected steps of the level x instructions per step ~ fixed part + palette entries + pixels
    work[cls] += (float)bins[kBinCount + k] * c_level_steps[lvl] * (300.0f + 9.0f * (float)(4 << cls) + 28.0f * (float)n);
  }
  bins[kBinTotal] = off;
  // measured correction of the model per class (share of CTA time each class's queue really took
  // on the SURVEY 8d image); lanes steal across classes once their queue is dry, so an imperfect
  // split costs warp uniformity, not idle time
  work[0] *= 1.3f; work[1] *= 0.62f; work[2] *= 0.67f;
  const float tot = work[0] + work[1] + work[2];
  // CTAs [0, b2) -> class 2, [b2, b1) -> class 1, [b1, grid) -> class 0
  uint32_t b2 = tot > 0.0f ? (uint32_t)(work[2] / tot * (float)grid_ctas + 0.5f) : 0;
  uint32_t b1 = tot > 0.0f ? (uint32_t)((work[2] + work[1]) / tot * (float)grid_ctas + 0.5f) : 0;
  if (b2 > grid_ctas) b2 = grid_ctas;
  if (b1 > grid_ctas) b1 = grid_ctas;
  if (b1 < b2) b1 = b2;
  bins[kBinHome + 2] = 0;
  bins[kBinHome + 1] = b2;
  bins[kBinHome + 0] = b1;
#ifdef FASTC_GPU_TAILSTATS
  printf("anneal classes: work %.3g / %.3g / %.3g -> CTAs %u / %u / %u; chains %u / %u / %u\n", work[0], work[1], work[2],
         grid_ctas - b1, b1 - b2, b2, bins[kBinEnd + 0] - bins[kBinFetch + 0], bins[kBinEnd + 1] - bins[kBinFetch + 1],
         bins[kBinEnd + 2] - bins[kBinFetch + 2]);
#endif
}

// Counting-sort scatter.  Ranks are taken in shared memory and each CTA reserves one range per
// key with a single global atomic (a few hot counters would otherwise serialise ~10 chains/block).
__global__ void __launch_bounds__(256) bc7_scatter(Ws ws, uint32_t num_blocks) {
  __shared__ uint32_t s_cnt[kSortKeys], s_base[kSortKeys];
  for (int k = threadIdx.x; k < kSortKeys; k += 256) s_cnt[k] = 0;
  __syncthreads();
  const uint32_t gid = blockIdx.x * 256 + threadIdx.x;
  int key = -1;
  uint32_t rank = 0;
  uint4 st0 = make_uint4(0, 0, 0, 0);
  if (gid < num_blocks * kSlots && ((ws.sa_mask[gid / kSlots] >> (gid % kSlots)) & 1u)) {
    st0 = *reinterpret_cast<const uint4 *>(ws.states + (size_t)gid * kStateWords);
    const uint32_t w0 = st0.x;
    {
      const int mode = (w0 >> 16) & 7, idx_mode = (w0 >> 21) & 1;
      const int ibits = idx_mode == 0 ? c_modes[mode].index_bits : c_modes[mode].alpha_index_bits;
      key = sort_key(ibits, (w0 >> 24) & 31, sort_error(ws, st0.w));
      rank = atomicAdd(&s_cnt[key], 1u);
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < kSortKeys; k += 256)
    if (s_cnt[k]) s_base[k] = ws.bins[kBinOffset + k] + atomicAdd(&ws.bins[kBinCursor + k], s_cnt[k]);
  __syncthreads();
  if (key >= 0) {
    // the annealing kernel reads its work in sorted order: one indirection less on its refill path
    uint4 st1 = *reinterpret_cast<const uint4 *>(ws.states + (size_t)gid * kStateWords + 4);
    st1.w = gid;
    uint4 *dst = ws.sorted + (size_t)(s_base[key] + rank) * 2;
    dst[0] = st0;
    dst[1] = st1;
  }
}

constexpr int kSaThreads = 128;
constexpr int kSaCtasPerSm = 8;
constexpr int kPixStride = kSaThreads + 1;  // odd row stride: a warp's cooperative column store is conflict-free

// OptimizeEndpointsForCluster (Compressor.cpp:538-630) as an all-integer state machine, one
// chain per lane.  Lanes fetch the next chain from the sorted list as soon as theirs ends,
// so a warp stays full until the list runs dry (the reference's chains have very uneven
// lengths: every new best restarts the schedule).
//
// Per evaluation (QuantizedError, RGBAEndpoints.cpp:190-310) the work is arranged as
//   palette[j] = interpolated colour of bucket j, two channels per 32-bit multiply
//   error(pixel, j) = sum of squared byte differences, VABSDIFF4 + DP4A   (exact integers)
// and the projection that picks the two candidate buckets uses one float multiply by a
// per-call reciprocal; whenever that product lands within 2^-16 of an integer (where the
// reference's own rounding could fall on the other side) the pixel is flagged and, after the
// branch-free main loop, re-evaluated with the reference's exact division sequence.
// The pixel loops run to the warp's largest cluster with per-lane predicates (lanes run in
// lock-step anyway), so their trip counts are warp-uniform and the loops unroll.
// Modes 4/5 project the ROTATED pixel with alpha forced to 255 (T16); instead of rotating
// every pixel the rotation is applied once per evaluation to the endpoint operands of the dot
// products:  pt . q = px . q' + 255 * q[3]  with  q' = q, alpha byte <- the rotated channel's byte,
// rotated channel's byte <- 0.
// Endpoint moves are byte-wise saturating adds (PickBestNeighboringEndpoints moves every channel
// by one grid step and clamps to [0, 255]); ToPixel's per-channel quantisation
// (RGBAEndpoints.cpp:126-177) is a 256-entry table per (precision, p-bit) built in shared memory
// from the same quantize_channel() the other kernels use.
struct SaConst {
  uint32_t stepb;        // per-channel step bytes
  uint32_t qkeep, qins;  // projection operand q' = (q & qkeep) | (((q >> qsh) << 24) & qins)
  int qsh;
  int calpha;            // 255 when the projected point's alpha is forced to 255 (modes 4/5), else 0
  int n, nbm1, woff;
  int xm, sh0;           // p-bit combos without branches: the flipped combo is combo ^ xm (shared 1, per
                         // endpoint 3, none 0); p-bit of endpoint 1 = (combo >> sh0) & 1, of endpoint 2 = combo & 1
  int tab_c, tab_a;      // quantisation table rows (precision class) of colour / alpha
};

// quantisation tables: row = class * 2 + pbit, class 0..4 = 4..8 kept bits, class 5 = "no bits
// kept" (the alpha of opaque modes: always 255).  The annealing loop only ever quantises with a
// p-bit of 0 or 1 (modes without p-bits evaluate with a zero p-bit, see fit_cluster).
constexpr int kQuantRows = 12;
__device__ __forceinline__ uint32_t sa_quantize(const uint8_t (*s_q)[256], const SaConst &K, uint32_t p, int pbit) {
  const uint8_t *tc = s_q[K.tab_c + pbit], *ta = s_q[K.tab_a + pbit];
  return (uint32_t)tc[p & 0xFF] | ((uint32_t)tc[(p >> 8) & 0xFF] << 8) | ((uint32_t)tc[(p >> 16) & 0xFF] << 16) |
         ((uint32_t)ta[p >> 24] << 24);
}

// one endpoint move: dir bit c set = channel c steps down.  With p-bits the step is taken only
// when it agrees with the p-bit flip (ChangePointForDirWithPbitChange, Compressor.cpp:364-418).
__device__ __forceinline__ uint32_t move_endpoint(uint32_t src, uint32_t dir, int old_pbit, int has_pbit, uint32_t stepb) {
  const uint32_t neg = (((dir & 15u) * 0x00204081u) & 0x01010101u) * 0xFFu;  // 0xFF in the bytes that step down
  uint32_t sub = stepb & neg, add = stepb & ~neg;
  if (has_pbit) {
    if (old_pbit == 0) add = 0;
    else sub = 0;
  }
  // Per byte exactly one of sub / add is non-zero, so the two saturating steps commute.  sm_100a has
  // no byte-wise saturating add (the compiler emulates __vaddus4 / __vsubus4 in ~11 instructions
  // each), but it has 16x2 min / max: spread the bytes over 16-bit lanes, where the add cannot
  // carry, clamp at 255, floor the subtraction at 0 with max(t, sub) - sub, and merge.
  const uint32_t lo = __byte_perm(src, 0u, 0x4240), hi = __byte_perm(src, 0u, 0x4341);
  const uint32_t alo = __byte_perm(add, 0u, 0x4240), ahi = __byte_perm(add, 0u, 0x4341);
  const uint32_t slo = __byte_perm(sub, 0u, 0x4240), shi = __byte_perm(sub, 0u, 0x4341);
  const uint32_t tlo = __viaddmin_u16x2(lo, alo, 0x00FF00FFu), thi = __viaddmin_u16x2(hi, ahi, 0x00FF00FFu);
  const uint32_t rlo = __vmaxu2(tlo, slo) - slo, rhi = __vmaxu2(thi, shi) - shi;
  return __byte_perm(rlo, rhi, 0x6240);
}

// The pixel loops of sa_eval.  UNI: every lane of the warp fits a cluster of nmax pixels (the
// usual case, the work list is sorted by cluster size), so no pixel needs a validity test.
template <bool UNI>
__device__ __forceinline__ void sa_pixels(uint32_t (*s_pix)[kPixStride], const uint2 *pal, int tid, int n, int nbm1,
                                          int nmax, uint32_t q1p, uint32_t q2p, uint32_t cq, float inv16, uint32_t k256,
                                          uint32_t &total, uint32_t &slow, uint32_t (&word)[2]) {
  const int n1 = n - 1, n2 = n - 2, n3 = n - 3;  // pixel i + k of a group of four is valid iff i < n - k
  // One pixel: adds its error unless it is flagged or past the lane's cluster, shifts the chosen
  // bucket into ACC from the top, leaves the flag in FLAG.
#define SA_PIXEL(PX, VALID, FLAG)                                                                          \
  {                                                                                                        \
    const uint32_t px_ = (PX);                                                                             \
    /* num + 0x4B400000: the bit pattern of the float 12582912 + num (|num| < 2^22), so the int -> float   \
       conversion is one exact FADD on the FMA pipe instead of an I2F on the (binding) ALU pipe */        \
    const int nb_ = (int)(__dp4a(px_, q2p, 0x4B400000u) - __dp4a(px_, q1p, cq));                           \
    const float fn_ = __fsub_rn(__int_as_float(nb_), 12582912.0f);                                         \
    /* vp = floor(num * inv16) + 1: the 16.16 fixed point bucket coordinate, plus one */                   \
    const int vp_ = __float2int_rd(__fmaf_rn(fn_, inv16, 1.0f));                                           \
    /* flagged: within 2^-16 of a bucket boundary (or past the cluster: those flags are masked off later) */ \
    const bool ok_ = (VALID) && (((uint32_t)vp_ & 0xFFFEu) != 0u);                                         \
    FLAG = !ok_;                                                                                           \
    const int ja_ = __vimin_s32_relu(vp_ >> 16, nbm1); /* floor, clamped (differs from it only when flagged) */ \
    const uint2 c_ = pal[ja_ * kSaThreads];                                                                \
    const uint32_t da_ = __vabsdiffu4(c_.x, px_), db_ = __vabsdiffu4(c_.y, px_);                           \
    /* keys = error << 8 | bucket: one min picks the error and the bucket (the lower bucket wins ties,     \
       like the reference's strict <), the sum of <= 16 keys carries sum(error) << 8 | sum(bucket) with    \
       no carry between the fields (errors < 2^18, buckets <= 15) */                                       \
    uint32_t ka_ = __dp4a(da_, da_, 0u) * k256 + (uint32_t)ja_;                                            \
    const uint32_t kb_ = __dp4a(db_, db_, 0u) * k256 + (uint32_t)ja_;                                      \
    /* a projection before endpoint 1 only tests bucket 0; past endpoint 2 both colours are bucket nbm1 */ \
    if (vp_ >= 1) ka_ = __viaddmin_u32(kb_, 1u, ka_);                                                      \
    if (!(VALID)) ka_ = 0u;                                                                                \
    total += ka_; /* flagged pixels included: the exact replay takes their provisional key back */         \
    acc = __funnelshift_r(acc, ka_, 4);                                                                    \
  }
#pragma unroll
  for (int half = 0; half < 2; half++) {
    const int i0 = 8 * half, i1 = min(nmax, i0 + 8);
    uint32_t acc = 0;
    int i = i0;
#pragma unroll 1
    for (; i + 4 <= i1; i += 4) {
      bool f0, f1, f2, f3;
      SA_PIXEL(s_pix[i][tid], UNI || i < n, f0)
      SA_PIXEL(s_pix[i + 1][tid], UNI || i < n1, f1)
      SA_PIXEL(s_pix[i + 2][tid], UNI || i < n2, f2)
      SA_PIXEL(s_pix[i + 3][tid], UNI || i < n3, f3)
      slow = (slow << 4) | (f0 ? 8u : 0u) | (f1 ? 4u : 0u) | (f2 ? 2u : 0u) | (f3 ? 1u : 0u);
    }
#pragma unroll 1
    for (; i < i1; i++) {
      bool f;
      SA_PIXEL(s_pix[i][tid], UNI || i < n, f)
      slow = slow + slow + (f ? 1u : 0u);
    }
    const int cnt = i1 - i0;
    word[half] = cnt > 0 ? acc >> (4 * (8 - cnt)) : 0u;
  }
}

// The same for a warp whose lanes all fit 16-pixel clusters (mode 6 and the mode 4/5 fits: two
// thirds of all annealing steps): no loop, no validity tests, no partial index words.
__device__ __forceinline__ void sa_pixels16(uint32_t (*s_pix)[kPixStride], const uint2 *pal, int tid, int nbm1,
                                            uint32_t q1p, uint32_t q2p, uint32_t cq, float inv16, uint32_t k256,
                                            uint32_t &total, uint32_t &slow, uint32_t (&word)[2]) {
#pragma unroll
  for (int half = 0; half < 2; half++) {
    uint32_t acc = 0;
#pragma unroll
    for (int g = 0; g < 2; g++) {
      const int i = 8 * half + 4 * g;
      bool f0, f1, f2, f3;
      SA_PIXEL(s_pix[i][tid], true, f0)
      SA_PIXEL(s_pix[i + 1][tid], true, f1)
      SA_PIXEL(s_pix[i + 2][tid], true, f2)
      SA_PIXEL(s_pix[i + 3][tid], true, f3)
      slow = (slow << 4) | (f0 ? 8u : 0u) | (f1 ? 4u : 0u) | (f2 ? 2u : 0u) | (f3 ? 1u : 0u);
    }
    word[half] = acc;
  }
}
#undef SA_PIXEL

// The palette of a warp whose lanes all use the same index precision (the usual case: the CTAs have
// a home precision class): unrolled with the weights as immediates; colour 0 and colour NB-1 are
// the endpoints themselves (weights 0 and 64), so only the NB - 2 inner colours are interpolated.
template <int NB>
__device__ __forceinline__ void sa_palette_uniform(uint2 (*s_pal)[kSaThreads], int tid, uint32_t q1, uint32_t q2,
                                                   uint32_t blo, uint32_t dlo, uint32_t bhi, uint32_t dhi) {
  uint32_t cur = q1;
#pragma unroll
  for (int j = 1; j <= NB - 2; j++) {
    const uint32_t w4 = 4u * (NB == 4 ? kSelWeights2[j] : (NB == 8 ? kSelWeights3[j] : kWeights4[j]));
    // the interpolated byte of each channel sits in byte 1 / 3 of its 16-bit lane (see sa_eval): two
    // multiply-adds and one byte permute per colour
    const uint32_t nxt = __byte_perm(blo + dlo * w4, bhi + dhi * w4, 0x7351);
    s_pal[j - 1][tid] = make_uint2(cur, nxt);
    cur = nxt;
  }
  s_pal[NB - 2][tid] = make_uint2(cur, q2);
  s_pal[NB - 1][tid] = make_uint2(q2, q2);
}

// Evaluate one cluster against quantised endpoints q1/q2: returns the total error and the
// chosen bucket of every pixel (4 bits each, cluster-local order) in idx_lo / idx_hi.
// nmax / nbmax: the warp's largest cluster size / bucket count - 1 (warp-uniform loop bounds).
// The palette is stored as PAIRS: row j of the lane's column holds (colour j, colour j + 1), the
// last row (colour nbm1, colour nbm1) -- the weight table is padded with 64 -- so the two
// candidate buckets of a pixel (floor and ceil of its projection) come from one 64-bit load.
// tid: the lane's own palette column; ptid: the s_pix column holding the cluster (the lane's own,
// except in the tail mode of bc7_anneal where a group of lanes evaluates one chain).
template <bool NU>
__device__ __forceinline__ uint32_t sa_eval(uint32_t (*s_pix)[kPixStride], uint2 (*s_pal)[kSaThreads],
                                            const uint8_t *__restrict__ s_w, int tid, int ptid, const SaConst &K, int nmax,
                                            int nbmax, int uflags, uint32_t q1, uint32_t q2, uint32_t k256, int rot,
                                            uint32_t &idx_lo, uint32_t &idx_hi) {
  const uint32_t d11 = __dp4a(q1, q1, 0u), d12 = __dp4a(q1, q2, 0u), d22 = __dp4a(q2, q2, 0u);
  const int den = (int)d22 - 2 * (int)d12 + (int)d11;       // |e2 - e1|^2
  {
    // ((64 - w) * e1 + w * e2 + 32) >> 6 per channel == (64 * e1 + 32 + w * (e2 - e1)) >> 6, channels
    // 0,2 and 1,3 in 16-bit lanes; the packed difference may borrow across lanes, the sum is
    // exact modulo 2^32 and the true value has no carries.  Everything is scaled by 4 (the largest
    // value, 4 * (64 * 255 + 32), still fits 16 bits) so that the >> 6 becomes "take byte 1 of the
    // lane": one byte permute packs the four channels.
    const uint32_t e1lo = q1 & 0x00FF00FFu, e1hi = (q1 >> 8) & 0x00FF00FFu;
    const uint32_t dlo = (q2 & 0x00FF00FFu) - e1lo, dhi = ((q2 >> 8) & 0x00FF00FFu) - e1hi;
    const uint32_t blo = e1lo * k256 + 0x00800080u, bhi = e1hi * k256 + 0x00800080u;
    if (uflags & 2) {  // every lane of the warp has nbm1 == nbmax
      if (nbmax == 3) sa_palette_uniform<4>(s_pal, tid, q1, q2, blo, dlo, bhi, dhi);
      else if (nbmax == 7) sa_palette_uniform<8>(s_pal, tid, q1, q2, blo, dlo, bhi, dhi);
      else sa_palette_uniform<16>(s_pal, tid, q1, q2, blo, dlo, bhi, dhi);
    } else {
      const uint8_t *wt = s_w + K.woff + 1;
      const uint32_t dlo4 = dlo * 4u, dhi4 = dhi * 4u;
      uint32_t cur = q1;  // colour 0 (weight 0) is endpoint 1 itself
#pragma unroll 4
      for (int j = 0; j <= nbmax; j++) {  // rows past this lane's bucket count are never read by it
        const uint32_t w = wt[j];
        const uint32_t nxt = __byte_perm(blo + dlo4 * w, bhi + dhi4 * w, 0x7351);
        s_pal[j][tid] = make_uint2(cur, nxt);
        cur = nxt;
      }
    }
  }
  const float fden = (float)den, fnb = (float)K.nbm1;
  // den == 0 (both endpoints equal): the zero reciprocal flags every pixel for the exact path
  const float inv16 = den ? __fdiv_rn(__fmul_rn(65536.0f, fnb), fden) : 0.0f;
  // projection operands with the rotation folded in (see above)
  const uint32_t q1p = (q1 & K.qkeep) | (((q1 >> K.qsh) << 24) & K.qins);
  const uint32_t q2p = (q2 & K.qkeep) | (((q2 >> K.qsh) << 24) & K.qins);
  // num = (pt - e1) . (e2 - e1) = px . q2' - (px . q1' + cq),  cq = e1 . (e2 - e1) - 255 * (e2 - e1)[alpha]
  const uint32_t cq = (uint32_t)((int)d12 - (int)d11 - K.calpha * ((int)(q2 >> 24) - (int)(q1 >> 24)));
  const int n = K.n, nbm1 = K.nbm1;
  const uint2 *pal = &s_pal[0][tid];
  if constexpr (NU) {
    // Non-uniform metric: float errors, summed in pixel order (the sum is not associative), so every
    // pixel is finished -- exact replay included -- before the next one is added.  Returns the
    // total's bit pattern.  (A niche setting: no warp-uniform fast paths here.)
    float w[4];
    nu_metric(rot, w);
    float total = 0.0f;
    uint32_t word[2] = {0u, 0u};
#pragma unroll 1
    for (int i = 0; i < n; i++) {
      const uint32_t px = s_pix[i][ptid];
      const int num = (int)(__dp4a(px, q2p, 0u) - __dp4a(px, q1p, cq));
      const int vp = __float2int_rd(__fmaf_rn((float)num, inv16, 1.0f));
      int ja = __vimin_s32_relu(vp >> 16, nbm1);
      bool two = vp >= 1;
      if (((uint32_t)vp & 0xFFFEu) == 0u) {  // within 2^-16 of a bucket boundary: the reference's own sequence
        ja = 0;
        two = false;
        if (den != 0) {
          const int k = (vp - 1 + 0x8000) >> 16;
          if (num * nbm1 == k * den) {
            ja = __vimin_s32_relu(k, nbm1);
          } else {
            const float t = __fmul_rn(__fdiv_rn((float)num, fden), fnb);
            const int x1 = min(max(0, (int)floorf(t)), nbm1), x2 = min((int)ceilf(t), nbm1);
            ja = x1;
            two = x1 + 1 <= x2;
          }
        }
      }
      const uint2 c = pal[ja * kSaThreads];
      float e = nu_error(c.x, px, w);
      uint32_t pick = (uint32_t)ja;
      if (two) {
        const float eb = nu_error(c.y, px, w);
        if (eb < e) { e = eb; pick = (uint32_t)ja + 1u; }  // (the last row repeats its colour: never taken there)
      }
      total = __fadd_rn(total, e);
      word[i >> 3] |= pick << (4 * (i & 7));
    }
    idx_lo = word[0];
    idx_hi = word[1];
    return __float_as_uint(total);
  }
  uint32_t total = 0, slow = 0, word[2];
  // `total` sums the pixels' keys (error << 8 | bucket, see SA_PIXEL)
  if ((uflags & 1) && nmax == 16) sa_pixels16(s_pix, pal, ptid, nbm1, q1p, q2p, cq, inv16, k256, total, slow, word);
  else if (uflags & 1) sa_pixels<true>(s_pix, pal, ptid, n, nbm1, nmax, q1p, q2p, cq, inv16, k256, total, slow, word);
  else sa_pixels<false>(s_pix, pal, ptid, n, nbm1, nmax, q1p, q2p, cq, inv16, k256, total, slow, word);
  // pixel i sits at bit nmax - 1 - i of `slow`; drop the flags of pixels past the lane's cluster
  slow &= 0xFFFFFFFFu << (nmax - n);
  // Flagged pixels: too close to a bucket boundary for the fast product.  Nearly all of them sit
  // EXACTLY on one (num * nbm1 == k * den): the reference then computes fl(fl(k / nbm1) * nbm1),
  // which is exactly k for every 0 <= k <= nbm1 and nbm1 in {3, 7, 15}, stays <= 0 for k <= 0 and
  // >= nbm1 for k >= nbm1 (tests/test_tables.py::test_projection_exact_buckets), so floor == ceil
  // and only bucket clamp(k) is tested.  The rest replay the reference's float sequence
  // (RGBAEndpoints.cpp:262-289); coinciding endpoints test bucket 0 (:226-251).
  // The main loop has already added a provisional key for them (both candidates of the floor
  // bucket's row, lower bucket on ties).  It is wrong only when the reference tests ONE bucket and
  // the row's other colour happened to be closer, or when the replayed floor differs: so an
  // exactly-on-boundary pixel whose provisional pick already is the boundary's bucket needs nothing,
  // and with coinciding endpoints (every colour equal, every pixel flagged) nothing ever changes.
  if (den == 0) slow = 0;
  while (slow) {
    const int bit = __ffs(slow) - 1;
    slow &= slow - 1u;
    const int i = nmax - 1 - bit;
    const uint32_t px = s_pix[i][ptid];
    const int num = (int)(__dp4a(px, q2p, 0u) - __dp4a(px, q1p, cq));
    const int vp = __float2int_rd(__fmaf_rn((float)num, inv16, 1.0f));
    const int sh = 4 * (i & 7);
    const uint32_t prov = ((i < 8 ? word[0] : word[1]) >> sh) & 15u;
    const int k = (vp - 1 + 0x8000) >> 16;
    int ja;
    bool two = false;
    if (num * nbm1 == k * den) {
      ja = __vimin_s32_relu(k, nbm1);
      if (prov == (uint32_t)ja) continue;
    } else {
      const float t = __fmul_rn(__fdiv_rn((float)num, fden), fnb);
      const int x1 = min(max(0, (int)floorf(t)), nbm1), x2 = min((int)ceilf(t), nbm1);
      ja = x1;
      two = x1 + 1 <= x2;
    }
    {  // take the provisional key back
      const int jp = __vimin_s32_relu(vp >> 16, nbm1);
      const uint2 c = pal[jp * kSaThreads];
      const uint32_t da = __vabsdiffu4(c.x, px), db = __vabsdiffu4(c.y, px);
      uint32_t ka = __dp4a(da, da, 0u) * 256u + (uint32_t)jp;
      const uint32_t kb = __dp4a(db, db, 0u) * 256u + (uint32_t)jp;
      if (vp >= 1) ka = __viaddmin_u32(kb, 1u, ka);
      total -= ka;
    }
    const uint2 c = pal[ja * kSaThreads];
    const uint32_t da = __vabsdiffu4(c.x, px), db = __vabsdiffu4(c.y, px);
    const uint32_t ea = __dp4a(da, da, 0u), eb = __dp4a(db, db, 0u);
    const bool up = two && eb < ea;
    const uint32_t pick = (uint32_t)ja + (up ? 1u : 0u);
    total += ((up ? eb : ea) << 8) | pick;
    const uint32_t clr = ~(0xFu << sh), ins = pick << sh;
    if (i < 8) word[0] = (word[0] & clr) | ins;
    else word[1] = (word[1] & clr) | ins;
  }
  idx_lo = word[0];
  idx_hi = word[1];
  return total >> 8;
}

// The CTA's shared tables (kSaThreads threads; the caller synchronises): quantisation rows, padded
// weight rows, and per (mode, index mode) the annealing constants of the mode
//   x: per-channel step bytes (opaque modes never move alpha, T3)
//   y: nbm1 [0:3] | weight row offset [4:11] | p-bit flip mask [12:13] | p-bit shift [14] |
//      colour / alpha quantisation rows [15:18] [19:22] | rotation [23]
__device__ __forceinline__ void sa_init_tables(uint8_t (*s_q)[256], uint8_t *s_w, uint2 *s_mode) {
  if (threadIdx.x < 80) {
    const int row = threadIdx.x >> 4, k = threadIdx.x & 15;  // row = index bits - 1
    s_w[threadIdx.x] = (row < 4 && k < (2 << row)) ? c_weight[threadIdx.x] : (uint8_t)64;
  }
  for (int e = threadIdx.x; e < kQuantRows * 256; e += kSaThreads) {
    const int row = e >> 8, cls = row >> 1, pbit = row & 1;
    const uint32_t mask = cls == 5 ? 0u : ((0xFF00u >> (cls + 4)) & 0xFFu);
    s_q[row][e & 255] = (uint8_t)quantize_channel((uint32_t)(e & 255), mask, pbit);
  }
  if (threadIdx.x < 16) {
    const int mode = threadIdx.x >> 1, idx_mode = threadIdx.x & 1;
    const ModeAttr A = c_modes[mode];
    const int ibits = max(1, idx_mode == 0 ? A.index_bits : A.alpha_index_bits);
    const uint32_t sc = 1u << (8 - A.color_bits), sa = (A.alpha_bits && mode >= 4) ? (1u << (8 - A.alpha_bits)) : 0u;
    const uint32_t xm = A.pbit == kPbitShared ? 1u : (A.pbit == kPbitPerEndpoint ? 3u : 0u);
    const uint32_t tab_c = (uint32_t)(A.color_bits - 4) * 2u, tab_a = (uint32_t)(A.alpha_bits ? A.alpha_bits - 4 : 5) * 2u;
    s_mode[threadIdx.x] = make_uint2(sc | (sc << 8) | (sc << 16) | (sa << 24),
                                     (uint32_t)((1 << ibits) - 1) | ((uint32_t)(16 * (ibits - 1)) << 4) | (xm << 12) |
                                         ((A.pbit == kPbitPerEndpoint ? 1u : 0u) << 14) | (tab_c << 15) | (tab_a << 19) |
                                         ((uint32_t)A.rotation << 23));
  }
}

// The constants of a chain from word 0 of its start state (see kStateWords) and its mode's table row.
__device__ __forceinline__ void sa_decode(uint32_t w0, const uint2 *s_mode, SaConst &K, int &rotation) {
  const int mode = (w0 >> 16) & 7, rot = (w0 >> 19) & 3, idx_mode = (w0 >> 21) & 1;
  const uint2 mt = s_mode[mode * 2 + idx_mode];
  K.n = (w0 >> 24) & 31;
  K.stepb = mt.x;
  K.nbm1 = mt.y & 15;
  K.woff = (mt.y >> 4) & 0xFF;
  K.xm = (mt.y >> 12) & 3;
  K.sh0 = (mt.y >> 14) & 1;
  K.tab_c = (mt.y >> 15) & 15;
  K.tab_a = (mt.y >> 19) & 15;
  rotation = (mt.y >> 23) & 1;
  K.qkeep = 0xFFFFFFFFu; K.qins = 0; K.qsh = 0; K.calpha = 0;
  if (rotation) {
    K.calpha = 255;
    K.qkeep = 0x00FFFFFFu;
    if (rot) { K.qsh = 8 * (rot - 1); K.qins = 0xFF000000u; K.qkeep &= ~(0xFFu << K.qsh); }
  }
}

// The state of one annealing chain: registers of the lane that runs it.
struct SaChain {
  uint32_t gid, cur1, cur2, best1, best2, cur_err, best_err, rng, best_lo, best_hi;
  int cur_combo, best_combo, energy, rotation;
  bool improved;
#ifdef FASTC_GPU_COUNTERS
  uint32_t ncalls, npbe;
#endif
};
// What a step needs besides the chain: the CTA's shared tables and the launch constants.
struct SaShared {
  uint32_t (*pix)[kPixStride];
  uint2 (*pal)[kSaThreads];
  const uint8_t (*q)[256];
  const uint8_t *w;
  uint32_t k256;
  float f_tm1, c_x;
  int sa_steps;
};

// One annealing step of a chain, in place.  tid: the lane's palette column, pcol: the s_pix column
// holding the cluster.  Returns true when the step was anything but a CLEAN REJECTION -- state
// unchanged, energy + 1, exactly three LCG draws (two moves, one Metropolis draw) -- which is what
// sa_tail speculates on.
template <bool NU>
__device__ __forceinline__ bool sa_step(SaChain &c, const SaConst &K, const SaShared &S, int tid, int pcol, int nmax,
                                        int nbmax, int uflags) {
    // PickBestNeighboringEndpoints (:426-498)
    const int has_pbit = K.xm != 0;
    const int ncombo = c.cur_combo ^ K.xm;  // the p-bit always flips (shared: 0 <-> 1, per endpoint: c <-> 3 - c)
    const int opb0 = (c.cur_combo >> K.sh0) & 1, opb1 = c.cur_combo & 1;
    uint32_t n1, n2;
    int guard = -1;
    bool visited;
    do {
      // pt = 0 moves endpoint 2 first and (as the reference does) tests p-bit [0] for it
      n2 = move_endpoint(c.cur2, lcg_next(c.rng), opb0, has_pbit, K.stepb);
      n1 = move_endpoint(c.cur1, lcg_next(c.rng), opb1, has_pbit, K.stepb);
      visited = (c.best1 == n1) && (c.best2 == n2) && (c.best_combo == ncombo);
    } while (visited && ++guard < 15);
    // modes without p-bits evaluate with a zero p-bit (reference quirk, see fit_cluster): their combo is 0
    const uint32_t q1 = sa_quantize(S.q, K, n1, (ncombo >> K.sh0) & 1), q2 = sa_quantize(S.q, K, n2, ncombo & 1);
    uint32_t ilo, ihi;
    const uint32_t err = sa_eval<NU>(S.pix, S.pal, S.w, tid, pcol, K, nmax, nbmax, uflags, q1, q2, S.k256,
                                     c.rotation ? (K.qins ? (K.qsh >> 3) + 1 : 0) : 0, ilo, ihi);
#ifdef FASTC_GPU_COUNTERS
    c.ncalls++; c.npbe += K.n;
#endif
    // AcceptNewEndpointError (:524-536)
    bool accept;
    if (err < c.cur_err) {
      accept = true;
    } else {
      const uint32_t r = lcg_next(c.rng) & 0xFFFF;
      const uint32_t m = ((r << 8) | (r >> 7)) & 0x7FFFFF;
      const float fr = __fsub_rn(__uint_as_float((127u << 23) | m), 1.0f);
      if (c.energy == 0) {
        accept = false;  // temp == 0: exp(-inf) = 0, exp(NaN) = NaN -> never accepted
      } else if (NU) {  // the errors are floats (bit patterns): the reference's double expression
        const float temp = __fdiv_rn((float)c.energy, S.f_tm1);
        const double x = ((double)0.1f * ((double)__uint_as_float(c.cur_err) - (double)__uint_as_float(err))) / (double)temp;
        accept = (double)fr < exp(x);
      } else {
        const float diff = (float)((int)c.cur_err - (int)err);  // exact (|.| < 2^24)
        // exp(0.1 * diff / temp), temp = c.energy / (steps - 1), through fast reciprocal / exponential
        const float pf = __expf(__fdividef(__fmul_rn(diff, S.c_x), (float)c.energy));
        if (fr < pf * (1.0f - 3e-5f)) accept = true;
        else if (fr > pf * (1.0f + 3e-5f)) accept = false;
        else {  // within the fast path's error band: the reference's double expression
          const float temp = __fdiv_rn((float)c.energy, S.f_tm1);
          const double x = ((double)0.1f * ((double)c.cur_err - (double)err)) / (double)temp;
          accept = (double)fr < exp(x);
        }
      }
    }
    if (accept) { c.cur_err = err; c.cur1 = n1; c.cur2 = n2; c.cur_combo = ncombo; }
    if (err < c.best_err) {
      c.best_err = err; c.best1 = n1; c.best2 = n2; c.best_combo = ncombo;
      c.best_lo = ilo; c.best_hi = ihi; c.improved = true;
      c.energy = 0;  // restart; the increment below makes it 1
    }
    c.energy++;
    return accept || guard >= 0;
}

// The chain has ended: its result record.
template <bool NU>
__device__ __forceinline__ void sa_finish(const SaChain &c, const SaConst &K, const Ws &ws) {
    // the indices belong to the evaluation that produced c.best_err: the start state's were
    // stored by bc7_setup, an c.improved state's were kept when it was found
    uint32_t *res = ws.results + (size_t)c.gid * kResWords;
    uint32_t o1 = c.best1, o2 = c.best2, alpha_err = 0;
    if (c.rotation) {
      const uint32_t *st = ws.states + (size_t)c.gid * kStateWords;
      const uint32_t abytes = st[6];
      alpha_err = st[5];
      o1 = (o1 & 0x00FFFFFFu) | ((abytes & 0xFF) << 24);
      o2 = (o2 & 0x00FFFFFFu) | (((abytes >> 8) & 0xFF) << 24);
    }
    *reinterpret_cast<uint4 *>(res) = make_uint4(c.best_err + alpha_err, o1, o2, (uint32_t)c.best_combo);
    if constexpr (NU) {  // the chain's error as the reference's double: colour fit (+ the alpha fit's, parked by bc7_setup)
      const double e = (double)__uint_as_float(c.best_err);
      ws.err64[c.gid] = c.rotation ? __dadd_rn(e, ws.err64[c.gid]) : e;
    }
    if (c.improved) {
      // nibbles past the cluster size come from the warp's longer loops: clear them
      const int n = K.n;
      const uint32_t mlo = n >= 8 ? 0xFFFFFFFFu : ((1u << (4 * n)) - 1u);
      const uint32_t mhi = n >= 16 ? 0xFFFFFFFFu : (n > 8 ? ((1u << (4 * (n - 8))) - 1u) : 0u);
      *reinterpret_cast<uint2 *>(res + 4) = make_uint2(c.best_lo & mlo, c.best_hi & mhi);
    }
}

#ifdef FASTC_GPU_TAILSTATS
// debug build: when did the queues run dry, when did the lanes / CTAs end (ns, globaltimer)
__device__ unsigned long long g_wid_steps[64];  // annealing steps executed by the lanes of hardware warp slot %warpid
__device__ unsigned long long g_tail[12];  // 0 start(min) 1 first dry(min) 2 last dry(max) 3 end(max) 4 - 5 CTAs 6..8 class c first seen dry
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__global__ void bc7_tail_reset() {
  g_tail[0] = g_tail[1] = g_tail[6] = g_tail[7] = g_tail[8] = ~0ull;
  g_tail[2] = g_tail[3] = g_tail[4] = g_tail[5] = g_tail[9] = g_tail[10] = 0;
  for (int k = 0; k < 64; k++) g_wid_steps[k] = 0;
}
__global__ void bc7_tail_report() {
  const double t0 = (double)g_tail[0];
  printf("anneal tail: all dry first seen %.3f ms, last %.3f ms, end %.3f ms; class queues dry at %.3f / %.3f / %.3f ms\n",
         ((double)g_tail[1] - t0) * 1e-6, ((double)g_tail[2] - t0) * 1e-6, ((double)g_tail[3] - t0) * 1e-6,
         ((double)g_tail[6] - t0) * 1e-6, ((double)g_tail[7] - t0) * 1e-6, ((double)g_tail[8] - t0) * 1e-6);
  printf("  last chain to finish: %llu steps, ran %llu us; longest chain: %llu steps, ran %llu us\n",
         ((g_tail[9] >> 14) & 0x3FFull) * 4ull, (g_tail[9] & 0xFFFull) * 16ull, g_tail[10] >> 32, g_tail[10] & 0xFFFFFFFFull);
  printf("  last chain: class %llu\n", (g_tail[9] >> 12) & 3ull);
  // the SM's schedulers do not share issue slots evenly among their eight resident warps: this is
  // what decides how long one (serial) chain takes, see DESIGN.md
  printf("  lane-steps by hardware warp slot (%%warpid):");
  for (int k = 0; k < 64; k++) printf(" %llu", g_wid_steps[k] >> 10);
  printf(" (x1024)\n");
}
#endif

#ifdef FASTC_GPU_CHAINSTATS
// debug build: chain length (annealing steps) against the start error per pixel, the predictor the
// work list is ordered by.  Row = floor(log2(error per pixel + 1)).
__device__ unsigned long long g_cs_count[32], g_cs_steps[32];
__device__ unsigned int g_cs_max[32], g_cs_hist[32][16];
__global__ void bc7_chainstats_reset() {
  const int t = threadIdx.x;
  if (t < 32) { g_cs_count[t] = 0; g_cs_steps[t] = 0; g_cs_max[t] = 0; for (int k = 0; k < 16; k++) g_cs_hist[t][k] = 0; }
}
__global__ void bc7_chainstats_report() {
  printf("chainstats: log2(err/px+1) | chains | mean steps | max steps | histogram of log2(steps)\n");
  for (int r = 0; r < 32; r++) {
    if (!g_cs_count[r]) continue;
    printf("cs %2d %10llu %8.1f %6u |", r, g_cs_count[r], (double)g_cs_steps[r] / (double)g_cs_count[r], g_cs_max[r]);
    for (int k = 0; k < 13; k++) printf(" %u", g_cs_hist[r][k]);
    printf("\n");
  }
}
#endif

// A chain of bc7_anneal parked for bc7_anneal_tail: its state goes into its own, now unused,
// start-state / result records, its id into the list.  Returns false when the list is full (the
// chain then stays with its lane).  Out of line: the caller's loop sits at the register limit.
__device__ __noinline__ bool sa_hand_over(Ws ws, uint32_t gid, uint32_t cur1, uint32_t cur2, uint32_t best1, uint32_t best2,
                                          uint32_t cur_err, uint32_t best_err, uint32_t rng, uint32_t packed, bool improved,
                                          int n, uint32_t best_lo, uint32_t best_hi) {
  const uint32_t slot = atomicAdd(&ws.bins[kBinTailCount], 1u);
  if (slot >= kTailCap) return false;
  ws.tail_list[slot] = gid;
  uint32_t *st = ws.states + (size_t)gid * kStateWords, *res = ws.results + (size_t)gid * kResWords;
  st[1] = cur1; st[2] = cur2; st[3] = best1; st[4] = best2;
  *reinterpret_cast<uint4 *>(res) = make_uint4(cur_err, best_err, rng, packed);
  if (improved) {  // as at the end of a chain: the start state's indices stand otherwise
    const uint32_t mlo = n >= 8 ? 0xFFFFFFFFu : ((1u << (4 * n)) - 1u);
    const uint32_t mhi = n >= 16 ? 0xFFFFFFFFu : (n > 8 ? ((1u << (4 * (n - 8))) - 1u) : 0u);
    *reinterpret_cast<uint2 *>(res + 4) = make_uint2(best_lo & mlo, best_hi & mhi);
  }
  return true;
}

// HAND: the warp's last chains go to bc7_anneal_tail (small submissions, where the kernel's tail counts).
template <bool NU, bool HAND>
__global__ void __launch_bounds__(kSaThreads, kSaCtasPerSm)
bc7_anneal(const uint32_t *__restrict__ img, uint32_t width, uint32_t blocks_x, uint32_t first_block, Ws ws,
           int sa_steps) {
  __shared__ uint2 s_pal[16][kSaThreads];
  __shared__ uint32_t s_pix[16][kPixStride];
  __shared__ uint8_t s_q[kQuantRows][256];
  __shared__ uint8_t s_w[80];  // c_weight with every slot past a row's last weight (and [64..79]) set to 64
  if (threadIdx.x < 80) {
    const int row = threadIdx.x >> 4, k = threadIdx.x & 15;  // row = index bits - 1
    s_w[threadIdx.x] = (row < 4 && k < (2 << row)) ? c_weight[threadIdx.x] : (uint8_t)64;
  }
  for (int e = threadIdx.x; e < kQuantRows * 256; e += kSaThreads) {
    const int row = e >> 8, cls = row >> 1, pbit = row & 1;
    const uint32_t mask = cls == 5 ? 0u : ((0xFF00u >> (cls + 4)) & 0xFFu);
    s_q[row][e & 255] = (uint8_t)quantize_channel((uint32_t)(e & 255), mask, pbit);
  }
  // per (mode, index mode): the annealing constants of the mode, decoded once per CTA
  //   x: per-channel step bytes (opaque modes never move alpha, T3)
  //   y: nbm1 [0:3] | weight row offset [4:11] | p-bit flip mask [12:13] | p-bit shift [14] |
  //      colour / alpha quantisation rows [15:18] [19:22] | rotation [23]
  __shared__ uint2 s_mode[16];
  __shared__ uint32_t s_end[3];
  if (threadIdx.x < 16) {
    const int mode = threadIdx.x >> 1, idx_mode = threadIdx.x & 1;
    const ModeAttr A = c_modes[mode];
    const int ibits = max(1, idx_mode == 0 ? A.index_bits : A.alpha_index_bits);
    const uint32_t sc = 1u << (8 - A.color_bits), sa = (A.alpha_bits && mode >= 4) ? (1u << (8 - A.alpha_bits)) : 0u;
    const uint32_t xm = A.pbit == kPbitShared ? 1u : (A.pbit == kPbitPerEndpoint ? 3u : 0u);
    const uint32_t tab_c = (uint32_t)(A.color_bits - 4) * 2u, tab_a = (uint32_t)(A.alpha_bits ? A.alpha_bits - 4 : 5) * 2u;
    s_mode[threadIdx.x] = make_uint2(sc | (sc << 8) | (sc << 16) | (sa << 24),
                                     (uint32_t)((1 << ibits) - 1) | ((uint32_t)(16 * (ibits - 1)) << 4) | (xm << 12) |
                                         ((A.pbit == kPbitPerEndpoint ? 1u : 0u) << 14) | (tab_c << 15) | (tab_a << 19) |
                                         ((uint32_t)A.rotation << 23));
  }
  if (threadIdx.x < 3) s_end[threadIdx.x] = ws.bins[kBinEnd + threadIdx.x];
  __syncthreads();
  const int tid = threadIdx.x, lane = tid & 31, wbase = tid & ~31;
  const unsigned full = 0xffffffffu;
  uint32_t drymask = 0;
#ifdef FASTC_GPU_TAILSTATS
  if (threadIdx.x == 0) atomicMin(&g_tail[0], gtime());
  uint32_t dbg_steps = 0, dbg_total = 0;
  unsigned long long dbg_start = 0;
#endif
  // 256 from a place the compiler cannot see through: `x * k256 + y` stays an IMAD (FMA pipe, which
  // has headroom) instead of a shift / LEA on the ALU pipe, the one that binds this kernel
  const uint32_t k256 = ws.bins[kBinWords - 1] + 256u;
  const float f_tm1 = (float)(sa_steps - 1);
  const float c_x = __fmul_rn(0.1f, f_tm1);  // fast Metropolis exponent: 0.1 * diff / (energy / (steps - 1))
  const int home = blockIdx.x >= ws.bins[kBinHome + 0] ? 0 : (blockIdx.x >= ws.bins[kBinHome + 1] ? 1 : 2);

  bool have = false;  // (drymask == 7: every queue has run dry for this lane)
  SaConst K = {0, 0xFFFFFFFFu, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  uint32_t gid = 0, cur1 = 0, cur2 = 0, best1 = 0, best2 = 0, cur_err = 0, best_err = 0, rng = 0;
  uint32_t best_lo = 0, best_hi = 0;
  int cur_combo = 0, best_combo = 0, energy = 0, rotation = 0;
  bool improved = false;
  int nmax = 0, nbmax = 0, uflags = 0;
#ifdef FASTC_GPU_COUNTERS
  uint32_t ncalls = 0, npbe = 0;
#endif
#ifdef FASTC_GPU_CHAINSTATS
  uint32_t cs_steps = 0;
  int cs_row = 0;
#endif

  for (;;) {
    const bool need = !have && drymask != 7u;
    if (__any_sync(full, need)) {
      // ---- refill: every idle lane takes the next chain of its home class's queue (then of the
      // others, longest steps first) ...
      uint32_t w0 = 0;
      bool got = false;
      if (need) {
        uint32_t pos = 0;
#pragma unroll
        for (int a = 0; a < 3 && !got; a++) {
          const int cls = a == 0 ? home : (2 - (a - 1) - ((2 - (a - 1)) <= home ? 1 : 0));
          if (cls < 0) break;
          if (!((drymask >> cls) & 1)) {  // a class this lane has seen run dry stays dry
            pos = atomicAdd(&ws.bins[kBinFetch + cls], 1u);
            got = pos < s_end[cls];
            if (!got) drymask |= 1u << cls;
#ifdef FASTC_GPU_TAILSTATS
            if (!got) {
              const unsigned long long t = gtime();
              atomicMin(&g_tail[6 + cls], t);
              if (drymask == 7u) { atomicMin(&g_tail[1], t); atomicMax(&g_tail[2], t); }
            }
#endif
          }
        }
        if (got) {
          // the chain's start state (sorted copy: two independent 16-byte loads) and the constants
          // of its mode (one table row)
          const uint4 s0 = ws.sorted[(size_t)pos * 2], s1 = ws.sorted[(size_t)pos * 2 + 1];
          gid = s1.w;
          w0 = s0.x;
          const int mode = (w0 >> 16) & 7, rot = (w0 >> 19) & 3, idx_mode = (w0 >> 21) & 1;
          const uint2 mt = s_mode[mode * 2 + idx_mode];
          K.n = (w0 >> 24) & 31;
          K.stepb = mt.x;
          K.nbm1 = mt.y & 15;
          K.woff = (mt.y >> 4) & 0xFF;
          K.xm = (mt.y >> 12) & 3;
          K.sh0 = (mt.y >> 14) & 1;
          K.tab_c = (mt.y >> 15) & 15;
          K.tab_a = (mt.y >> 19) & 15;
          rotation = (mt.y >> 23) & 1;
          K.qkeep = 0xFFFFFFFFu; K.qins = 0; K.qsh = 0; K.calpha = 0;
          if (rotation) {
            K.calpha = 255;
            K.qkeep = 0x00FFFFFFu;
            if (rot) { K.qsh = 8 * (rot - 1); K.qins = 0xFF000000u; K.qkeep &= ~(0xFFu << K.qsh); }
          }
          cur1 = best1 = s0.y; cur2 = best2 = s0.z;
          cur_err = best_err = s0.w;
          rng = s1.x;
          cur_combo = best_combo = (w0 >> 22) & 3;
          energy = 0;
          improved = false;
          have = true;
#ifdef FASTC_GPU_CHAINSTATS
          cs_steps = 0;
          cs_row = 31 - __clz((s0.w / max((w0 >> 24) & 31u, 1u)) + 1u);
#endif
#ifdef FASTC_GPU_TAILSTATS
          dbg_steps = 0;
          dbg_start = gtime();
#endif
        }
      }
      // ... and the warp loads the new chains' pixels together: lane i < 16 fetches pixel i of the
      // block and stores it at its rank within the subset mask, into the owner lane's column
      __syncwarp();  // (the columns written below were read by their owners' last steps)
      unsigned gm = __ballot_sync(full, got);
      while (gm) {
        const int owner = __ffs(gm) - 1;
        gm &= gm - 1;
        const uint32_t g = __shfl_sync(full, gid, owner), m = __shfl_sync(full, w0, owner);
        if (lane < 16 && ((m >> lane) & 1)) {
          const uint32_t bi = first_block + g / kSlots;
          const uint32_t *base = img + (size_t)(bi / blocks_x) * 4 * width + (size_t)(bi % blocks_x) * 4;
          s_pix[__popc(m & ((1u << lane) - 1u))][wbase + owner] = __ldg(base + (size_t)(lane >> 2) * width + (lane & 3));
        }
      }
      __syncwarp();
      // warp-uniform loop bounds of the evaluations: they only change here (a lane that ends its
      // chain with the queue dry leaves them larger than needed, which is harmless)
      nmax = __reduce_max_sync(full, have ? K.n : 0);
      nbmax = __reduce_max_sync(full, have ? K.nbm1 : 0);
      // bit 0: every lane's cluster has nmax pixels, bit 1: every lane has nbmax + 1 buckets
      uflags = (__all_sync(full, !have || K.n == nmax) ? 1 : 0) | (__all_sync(full, !have || K.nbm1 == nbmax) ? 2 : 0);
      // ---- hand-over to bc7_anneal_tail (see there).  Every lane passes here when it goes idle with
      // the queues dry (that is when it sees the last class run dry), so the warp notices the moment
      // it is down to kTailChains chains: they are parked in their own, now unused, start-state /
      // result records and listed; the warp is done.
      if (HAND && __any_sync(full, drymask == 7u) && __popc(__ballot_sync(full, have)) <= kTailChains) {
        if (have)
          have = !sa_hand_over(ws, gid, cur1, cur2, best1, best2, cur_err, best_err, rng,
                               (uint32_t)cur_combo | ((uint32_t)best_combo << 2) | ((uint32_t)energy << 8), improved, K.n,
                               best_lo, best_hi);
      }
    }
    if (!__any_sync(full, have)) break;
    if (!have) continue;

    // ---- one annealing step.  (The step and the chain's end are spelled out here although sa_step /
    // sa_finish hold the same code for bc7_anneal_tail: calling them from this loop, which sits at
    // the register limit of 8 CTAs / SM, measured 3 % slower.)
    bool done = !(best_err > 0 && energy < sa_steps);
    if (!done) {
      // PickBestNeighboringEndpoints (:426-498)
      const int has_pbit = K.xm != 0;
      const int ncombo = cur_combo ^ K.xm;  // the p-bit always flips (shared: 0 <-> 1, per endpoint: c <-> 3 - c)
      const int opb0 = (cur_combo >> K.sh0) & 1, opb1 = cur_combo & 1;
      uint32_t n1, n2;
      int guard = -1;
      bool visited;
      do {
        // pt = 0 moves endpoint 2 first and (as the reference does) tests p-bit [0] for it
        n2 = move_endpoint(cur2, lcg_next(rng), opb0, has_pbit, K.stepb);
        n1 = move_endpoint(cur1, lcg_next(rng), opb1, has_pbit, K.stepb);
        visited = (best1 == n1) && (best2 == n2) && (best_combo == ncombo);
      } while (visited && ++guard < 15);
      // modes without p-bits evaluate with a zero p-bit (reference quirk, see fit_cluster): their combo is 0
      const uint32_t q1 = sa_quantize(s_q, K, n1, (ncombo >> K.sh0) & 1), q2 = sa_quantize(s_q, K, n2, ncombo & 1);
      uint32_t ilo, ihi;
      const uint32_t err = sa_eval<NU>(s_pix, s_pal, s_w, tid, tid, K, nmax, nbmax, uflags, q1, q2, k256,
                                       rotation ? (K.qins ? (K.qsh >> 3) + 1 : 0) : 0, ilo, ihi);
#ifdef FASTC_GPU_COUNTERS
      ncalls++; npbe += K.n;
#endif
      // AcceptNewEndpointError (:524-536)
      bool accept;
      if (err < cur_err) {
        accept = true;
      } else {
        const uint32_t r = lcg_next(rng) & 0xFFFF;
        const uint32_t m = ((r << 8) | (r >> 7)) & 0x7FFFFF;
        const float fr = __fsub_rn(__uint_as_float((127u << 23) | m), 1.0f);
        if (energy == 0) {
          accept = false;  // temp == 0: exp(-inf) = 0, exp(NaN) = NaN -> never accepted
        } else if (NU) {  // the errors are floats (bit patterns): the reference's double expression
          const float temp = __fdiv_rn((float)energy, f_tm1);
          const double x = ((double)0.1f * ((double)__uint_as_float(cur_err) - (double)__uint_as_float(err))) / (double)temp;
          accept = (double)fr < exp(x);
        } else {
          const float diff = (float)((int)cur_err - (int)err);  // exact (|.| < 2^24)
          // exp(0.1 * diff / temp), temp = energy / (steps - 1), through fast reciprocal / exponential
          const float pf = __expf(__fdividef(__fmul_rn(diff, c_x), (float)energy));
          if (fr < pf * (1.0f - 3e-5f)) accept = true;
          else if (fr > pf * (1.0f + 3e-5f)) accept = false;
          else {  // within the fast path's error band: the reference's double expression
            const float temp = __fdiv_rn((float)energy, f_tm1);
            const double x = ((double)0.1f * ((double)cur_err - (double)err)) / (double)temp;
            accept = (double)fr < exp(x);
          }
        }
      }
      if (accept) { cur_err = err; cur1 = n1; cur2 = n2; cur_combo = ncombo; }
      if (err < best_err) {
        best_err = err; best1 = n1; best2 = n2; best_combo = ncombo;
        best_lo = ilo; best_hi = ihi; improved = true;
        energy = 0;  // restart; the increment below makes it 1
      }
      energy++;
      done = !(best_err > 0 && energy < sa_steps);
    }
#ifdef FASTC_GPU_TAILSTATS
    dbg_steps++;
    dbg_total++;
    if (done) {
      const unsigned long long t = gtime();
      const unsigned long long dur_us = (t - dbg_start) / 1000ull;
      // finish time (ns, 40 bits) | steps (12 bits, saturated) | duration us (12 bits, saturated)
      const unsigned long long cls_ = K.nbm1 == 3 ? 0ull : (K.nbm1 == 7 ? 1ull : 2ull);
      const unsigned long long rec = ((t & 0xFFFFFFFFFFull) << 24) | ((unsigned long long)min(dbg_steps / 4u, 1023u) << 14) |
                                     (cls_ << 12) | (unsigned long long)min(dur_us / 16ull, 4095ull);
      atomicMax(&g_tail[9], rec);
      atomicMax(&g_tail[10], ((unsigned long long)dbg_steps << 32) | (unsigned long long)(uint32_t)dur_us);
    }
#endif
#ifdef FASTC_GPU_CHAINSTATS
    cs_steps++;
    if (done) {
      atomicAdd(&g_cs_count[cs_row], 1ull);
      atomicAdd(&g_cs_steps[cs_row], (unsigned long long)cs_steps);
      atomicMax(&g_cs_max[cs_row], cs_steps);
      atomicAdd(&g_cs_hist[cs_row][min(31 - __clz(cs_steps), 15)], 1u);
    }
#endif
    if (done) {
      // the indices belong to the evaluation that produced best_err: the start state's were
      // stored by bc7_setup, an improved state's were kept when it was found
      uint32_t *res = ws.results + (size_t)gid * kResWords;
      uint32_t o1 = best1, o2 = best2, alpha_err = 0;
      if (rotation) {
        const uint32_t *st = ws.states + (size_t)gid * kStateWords;
        const uint32_t abytes = st[6];
        alpha_err = st[5];
        o1 = (o1 & 0x00FFFFFFu) | ((abytes & 0xFF) << 24);
        o2 = (o2 & 0x00FFFFFFu) | (((abytes >> 8) & 0xFF) << 24);
      }
      *reinterpret_cast<uint4 *>(res) = make_uint4(best_err + alpha_err, o1, o2, (uint32_t)best_combo);
      if constexpr (NU) {  // the chain's error as the reference's double: colour fit (+ the alpha fit's, parked by bc7_setup)
        const double e = (double)__uint_as_float(best_err);
        ws.err64[gid] = rotation ? __dadd_rn(e, ws.err64[gid]) : e;
      }
      if (improved) {
        // nibbles past the cluster size come from the warp's longer loops: clear them
        const int n = K.n;
        const uint32_t mlo = n >= 8 ? 0xFFFFFFFFu : ((1u << (4 * n)) - 1u);
        const uint32_t mhi = n >= 16 ? 0xFFFFFFFFu : (n > 8 ? ((1u << (4 * (n - 8))) - 1u) : 0u);
        *reinterpret_cast<uint2 *>(res + 4) = make_uint2(best_lo & mlo, best_hi & mhi);
      }
      have = false;
    }
  }
#ifdef FASTC_GPU_COUNTERS
  atomicAdd(&ws.counters[0], (unsigned long long)ncalls);
  atomicAdd(&ws.counters[1], (unsigned long long)npbe);
#endif
#ifdef FASTC_GPU_TAILSTATS
  {
    uint32_t warp_slot;
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(warp_slot));
    atomicAdd(&g_wid_steps[warp_slot & 63], (unsigned long long)dbg_total);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long t = gtime();
    atomicMax(&g_tail[3], t);
    atomicAdd(&g_tail[4], t);
    atomicAdd(&g_tail[5], 1ull);
  }
#endif
}

// The tail of the annealing.  bc7_anneal ends when its longest remaining chain does -- a fixed cost per
// launch that weighs on small submissions (a 1/8 shard, a 2048^2 texture): chain lengths are very
// uneven and a chain is serial.  But nearly all of its steps are clean rejections (see sa_step), so
// the last chains of every warp (handed over when the queues are dry and the warp is down to
// kTailChains) are finished here by a WARP each: lane d evaluates step t + d of the chain on the
// assumption that steps t .. t + d - 1 are clean rejections (energy + d, LCG jumped 3 d draws
// ahead).  The warp then commits up to its first step that is NOT a clean rejection: every lane
// takes the state that lane left behind -- exactly the state the serial chain has after those
// steps -- and the later lanes' work is discarded.  Results are bit-identical to the
// one-lane-per-chain schedule; the tail gets several times shorter.
template <bool NU>
__global__ void __launch_bounds__(kSaThreads, 4)
bc7_anneal_tail(const uint32_t *__restrict__ img, uint32_t width, uint32_t blocks_x, uint32_t first_block, Ws ws,
                int sa_steps) {
  __shared__ uint2 s_pal[16][kSaThreads];
  __shared__ uint32_t s_pix[16][kPixStride];  // (one column per warp is used: column `wbase`)
  __shared__ uint8_t s_q[kQuantRows][256];
  __shared__ uint8_t s_w[80];
  __shared__ uint2 s_mode[16];
  sa_init_tables(s_q, s_w, s_mode);
  __syncthreads();
  const int tid = threadIdx.x, lane = tid & 31, wbase = tid & ~31;
  const unsigned full = 0xffffffffu;
  const uint32_t k256 = ws.bins[kBinWords - 1] + 256u;
  const float f_tm1 = (float)(sa_steps - 1);
  const SaShared S = {s_pix, s_pal, s_q, s_w, k256, f_tm1, __fmul_rn(0.1f, f_tm1), sa_steps};
  const uint32_t count = min(ws.bins[kBinTailCount], kTailCap);
  // LCG jump over 3 * lane draws: x -> jmp_a * x + jmp_c
  uint32_t jmp_a = 1u, jmp_c = 0u;
  for (int i = 0; i < 3 * lane; i++) { jmp_c = 214013u * jmp_c + 2531011u; jmp_a *= 214013u; }
  for (;;) {
    uint32_t slot = 0;
    if (lane == 0) slot = atomicAdd(&ws.bins[kBinTailFetch], 1u);
    slot = __shfl_sync(full, slot, 0);
    if (slot >= count) break;
    SaChain c = {};
    SaConst K;
    c.gid = ws.tail_list[slot];
    const uint32_t *st = ws.states + (size_t)c.gid * kStateWords, *res = ws.results + (size_t)c.gid * kResWords;
    const uint32_t w0 = st[0];
    sa_decode(w0, s_mode, K, c.rotation);
    c.cur1 = st[1]; c.cur2 = st[2]; c.best1 = st[3]; c.best2 = st[4];
    const uint4 r0 = *reinterpret_cast<const uint4 *>(res);
    const uint2 r1 = *reinterpret_cast<const uint2 *>(res + 4);
    c.cur_err = r0.x; c.best_err = r0.y; c.rng = r0.z;
    c.cur_combo = r0.w & 3; c.best_combo = (r0.w >> 2) & 3; c.energy = (int)(r0.w >> 8);
    c.best_lo = r1.x; c.best_hi = r1.y;
    c.improved = true;  // the record's indices are the best state's (they are rewritten as they are)
    __syncwarp();       // (the previous chain's pixels are no longer read)
    if (lane < 16 && ((w0 >> lane) & 1)) {
      const uint32_t bi = first_block + c.gid / kSlots;
      const uint32_t *base = img + (size_t)(bi / blocks_x) * 4 * width + (size_t)(bi % blocks_x) * 4;
      s_pix[__popc(w0 & ((1u << lane) - 1u))][wbase] = __ldg(base + (size_t)(lane >> 2) * width + (lane & 3));
    }
    __syncwarp();
    const int nbmax = K.nbm1;
    for (;;) {
      // every lane holds the chain's state: move to this lane's speculative position
      c.energy += lane;
      c.rng = jmp_a * c.rng + jmp_c;
      bool stepped = false, changed = false;
      if (c.best_err > 0 && c.energy < sa_steps) {
        stepped = true;
        changed = sa_step<NU>(c, K, S, tid, wbase, K.n, nbmax, 3);
      }
      // commit: the first step that is not a clean rejection, else the last step taken (the steps
      // taken are a prefix of the lanes: lane d steps iff energy + d < sa_steps)
      const unsigned chg = __ballot_sync(full, stepped && changed), stp = __ballot_sync(full, stepped);
      const int cl = chg ? __ffs(chg) - 1 : 31 - __clz(stp);
#define SA_TAKE(v) v = __shfl_sync(full, v, cl)
      SA_TAKE(c.cur1); SA_TAKE(c.cur2); SA_TAKE(c.best1); SA_TAKE(c.best2); SA_TAKE(c.cur_err); SA_TAKE(c.best_err); SA_TAKE(c.rng);
      SA_TAKE(c.best_lo); SA_TAKE(c.best_hi); SA_TAKE(c.cur_combo); SA_TAKE(c.best_combo); SA_TAKE(c.energy);
#undef SA_TAKE
      if (!(c.best_err > 0 && c.energy < sa_steps)) break;
    }
    if (lane == 0) sa_finish<NU>(c, K, ws);
#ifdef FASTC_GPU_COUNTERS  // evaluations executed, the discarded speculative ones included
    atomicAdd(&ws.counters[0], (unsigned long long)c.ncalls);
    atomicAdd(&ws.counters[1], (unsigned long long)c.npbe);
#endif
  }
}

// ------------------------------------------------------------------ pack
struct BitWriter {
  uint32_t w[4] = {0, 0, 0, 0};
  int pos = 0;
  __device__ __forceinline__ void write(uint32_t v, int n) {  // LSB first (BitStream.h:65-94)
    if (n == 0) return;
    v &= (n >= 32) ? 0xFFFFFFFFu : ((1u << n) - 1);
    const int word = pos >> 5, off = pos & 31;
    w[word] |= v << off;
    if (off + n > 32 && word < 3) w[word + 1] |= v >> (32 - off);
    pos += n;
  }
};

template <bool NU>
__global__ void __launch_bounds__(128)
bc7_pack(const uint32_t *__restrict__ img, uint32_t width, uint32_t blocks_x, uint32_t first_block,
         uint32_t num_blocks, Ws ws, uint8_t *__restrict__ out) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= num_blocks) return;
  const uint32_t bi = first_block + t;
  const uint32_t selw = ws.sel[t];
  const uint32_t type = selw >> 24;
  BitWriter s;
  if (type == kTypeSolid) {
    // CompressOptimalColorBC7 (Compressor.cpp:1424-1458) + watermark order (T1): word index =
    // number of solid blocks before this one = tile prefix + solid blocks earlier in the tile.
    const uint32_t px = __ldg(img + (size_t)(bi / blocks_x) * 4 * width + (size_t)(bi % blocks_x) * 4);
    uint32_t before = *ws.wm_running + ws.tile_count[t / kTile];
    for (uint32_t j = (t / kTile) * kTile; j < t; j++) before += ((ws.sel[j] >> 24) == kTypeSolid);
    s.write(1u << 5, 6);
    s.write(0, 2);
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
      const int v = chan(px, ch);
      s.write(c_opt7[2 * v], 7);
      s.write(c_opt7[2 * v + 1], 7);
    }
    s.write(px >> 24, 8);
    s.write(px >> 24, 8);
    s.write(0xaaaaaaabu, 31);
    s.write(c_wm[before % 9], 31);
    reinterpret_cast<uint4 *>(out)[bi] = make_uint4(s.w[0], s.w[1], s.w[2], s.w[3]);
    if (ws.stats) {  // the statistics variant's records (Compressor.cpp:2012-2019): mode 5, path 0
      double *st = ws.stats + (size_t)t * kStatDoubles;
      st[0] = 5.0; st[1] = 0.0;
      for (int m = 0; m < 8; m++) st[2 + m] = -1.0;
    }
    return;
  }
  if (type == kTypeTransparent) {  // WriteTransparentBlock (:1416-1421)
    reinterpret_cast<uint4 *>(out)[bi] = make_uint4(1u << 6, 0, 0, 0);
    if (ws.stats) {  // (:2036-2043): mode 6, path 1
      double *st = ws.stats + (size_t)t * kStatDoubles;
      st[0] = 6.0; st[1] = 1.0;
      for (int m = 0; m < 8; m++) st[2 + m] = -1.0;
    }
    return;
  }

  // ---- CompressClusters' selection (:1771-1808): modes in the order {0,2,1,3,7,4,5,6},
  // shape slots ascending, strict <.
  const uint32_t *res = ws.results + (size_t)t * kSlots * kResWords;
  const bool layout_b = (selw >> 22) & 1;
  // errors: exact integers under the uniform metric; the reference's doubles under the non-uniform
  // one (sums over the subsets in subset order)
  using Err = typename std::conditional<NU, double, unsigned long long>::type;
  Err best_err = NU ? (Err)DBL_MAX : (Err)~0ull;
  int best_mode = -1, best_si = 0, best_first = 0;  // best_first: first slot of the winning candidate
  // all sixteen chain errors at once: independent loads, static register indices below (slots
  // without a live chain hold stale words, which no live candidate reads)
  Err e[kSlots];
#pragma unroll
  for (int k = 0; k < kSlots; k++) {
    if constexpr (NU) e[k] = ws.err64[(size_t)t * kSlots + k];
    else e[k] = res[k * kResWords];
  }
  // candidate (mode, shape slot si): its first chain slot, its error, the slot Pack reads
  double mode_err[8] = {-1.0, -1.0, -1.0, -1.0, -1.0, -1.0, -1.0, -1.0};  // statistics: errors[mode] (Compressor.cpp:1797-1798)
  auto consider = [&](int mode, int si, int first, Err err, int pick) {
    if (!decode_chain(selw, first).active) return;
    if (ws.stats) {
#pragma unroll
      for (int m = 0; m < 8; m++)
        if (m == mode && (mode_err[m] < 0.0 || (double)err < mode_err[m])) mode_err[m] = (double)err;
    }
    if (err < best_err) { best_err = err; best_mode = mode; best_si = si; best_first = pick; }
  };
  if (!layout_b) {  // opaque layout, reference order {0, 2, 1, 3, 7, (4, 5: none), 6}
    consider(0, 1, 0, e[0] + e[1] + e[2], 0);
    consider(2, 1, 3, e[3] + e[4] + e[5], 3);
    consider(1, 0, 6, e[6] + e[7], 6);
    consider(3, 0, 8, e[8] + e[9], 8);
    consider(7, 0, 10, e[10] + e[11], 10);
    consider(6, 0, 12, e[12], 12);
    consider(6, 1, 13, e[13], 13);
  } else {          // alpha layout: {7, 4, 5, 6}
    consider(7, 0, 13, e[13] + e[14], 13);
    // modes 4 / 5: best rotation / index mode, strict < in loop order (:1320-1348)
    Err be = e[0];
    int pick = 0;
#pragma unroll
    for (int k = 1; k < 8; k++)
      if (e[k] < be) { be = e[k]; pick = k; }
    consider(4, 0, 0, be, pick);
    be = e[8];
    pick = 8;
#pragma unroll
    for (int k = 9; k < 12; k++)
      if (e[k] < be) { be = e[k]; pick = k; }
    consider(5, 0, 8, be, pick);
    consider(6, 0, 12, e[12], 12);
  }
  if (ws.stats) {  // path 2: a shape estimate hit zero (early-out), 3: full search (Compressor.cpp:2110, :2164, :2172)
    double *st = ws.stats + (size_t)t * kStatDoubles;
    st[0] = (double)best_mode;
    st[1] = ((selw >> 23) & 1) ? 2.0 : 3.0;
#pragma unroll
    for (int m = 0; m < 8; m++) st[2 + m] = mode_err[m];
  }
  if (best_mode < 0) {  // unreachable with the default mode mask
    reinterpret_cast<uint4 *>(out)[bi] = make_uint4(0, 0, 0, 0);
    return;
  }

  // ---- CompressionMode::Pack (:1096-1298)
  const ModeAttr A = c_modes[best_mode];
  const int shape = A.subsets == 3 ? ((selw >> 6) & 63) : (A.subsets == 2 ? (selw & 63) : (best_si ? ((selw >> 6) & 63) : (selw & 63)));
  int rot = 0, idx_mode = 0;
  if (best_mode == 4) { rot = best_first >> 1; idx_mode = best_first & 1; }
  if (best_mode == 5) { rot = best_first - 8; }
  const uint32_t qm = quant_mask(A);

  uint32_t px1[3], px2[3];
  int combo[3];
  uint8_t idx[16], aidx[16];
#pragma unroll
  for (int i = 0; i < 16; i++) { idx[i] = 0; aidx[i] = 0; }
  for (int sub = 0; sub < A.subsets; sub++) {
    const uint32_t *r = res + (best_first + sub) * kResWords;
    int pb0, pb1;
    combo[sub] = (int)r[3];
    pbit_combo(A.pbit, combo[sub], pb0, pb1);
    px1[sub] = to_pixel_b(r[1], qm, pb0);
    px2[sub] = to_pixel_b(r[2], qm, pb1);
    const unsigned long long ci = (unsigned long long)r[4] | ((unsigned long long)r[5] << 32);
    const unsigned long long ai = (unsigned long long)r[6] | ((unsigned long long)r[7] << 32);
    int k = 0;
    for (int i = 0; i < 16; i++)
      if (subset_of(i, shape, A.subsets) == sub) {
        idx[i] = (ci >> (4 * k)) & 15;
        if (A.rotation) aidx[i] = (ai >> (4 * k)) & 15;
        k++;
      }
  }
  const int nib = idx_mode == 0 ? A.index_bits : A.alpha_index_bits;
  const int nab = idx_mode == 0 ? A.alpha_index_bits : A.index_bits;
  for (int sub = 0; sub < A.subsets; sub++) {
    const int anchor = anchor_of(sub, shape, A.subsets);
    if (idx[anchor] >> (nib - 1)) {
      const uint32_t tmp = px1[sub]; px1[sub] = px2[sub]; px2[sub] = tmp;
      for (int i = 0; i < 16; i++)
        if (subset_of(i, shape, A.subsets) == sub) idx[i] = (uint8_t)(((1 << nib) - 1) - idx[i]);
      if (A.rotation)
        for (int i = 0; i < 16; i++) aidx[i] = (uint8_t)(((1 << nab) - 1) - aidx[i]);
    }
    if (A.rotation && (aidx[anchor] >> (nab - 1))) {
      const uint32_t a1 = px1[sub] & 0xFF000000u, a2 = px2[sub] & 0xFF000000u;
      px1[sub] = (px1[sub] & 0x00FFFFFFu) | a2;
      px2[sub] = (px2[sub] & 0x00FFFFFFu) | a1;
      for (int i = 0; i < 16; i++) aidx[i] = (uint8_t)(((1 << nab) - 1) - aidx[i]);
    }
  }
  s.write(1u << best_mode, best_mode + 1);
  s.write((uint32_t)shape, A.partition_bits);
  if (A.rotation) s.write((uint32_t)rot, 2);
  if (A.idx_mode) s.write((uint32_t)idx_mode, 1);
  for (int ch = 0; ch < 3; ch++)
    for (int sub = 0; sub < A.subsets; sub++) {
      s.write(chan(px1[sub], ch) >> (8 - A.color_bits), A.color_bits);
      s.write(chan(px2[sub], ch) >> (8 - A.color_bits), A.color_bits);
    }
  if (A.alpha_bits)
    for (int sub = 0; sub < A.subsets; sub++) {
      s.write((px1[sub] >> 24) >> (8 - A.alpha_bits), A.alpha_bits);
      s.write((px2[sub] >> 24) >> (8 - A.alpha_bits), A.alpha_bits);
    }
  if (A.pbit != kPbitNone)
    for (int sub = 0; sub < A.subsets; sub++) {
      int pb0, pb1;
      pbit_combo(A.pbit, combo[sub], pb0, pb1);  // T17: original order even if the endpoints were swapped
      s.write((uint32_t)pb0, 1);
      if (A.pbit != kPbitShared) s.write((uint32_t)pb1, 1);
    }
  if (A.idx_mode && idx_mode == 1) {
    for (int i = 0; i < 16; i++) s.write(aidx[i], i == 0 ? 1 : 2);
    for (int i = 0; i < 16; i++) s.write(idx[i], i == 0 ? 2 : 3);
  } else {
    for (int i = 0; i < 16; i++) {
      const int sub = subset_of(i, shape, A.subsets);
      s.write(idx[i], i == anchor_of(sub, shape, A.subsets) ? nib - 1 : nib);
    }
    if (A.rotation)
      for (int i = 0; i < 16; i++) s.write(aidx[i], i == 0 ? nab - 1 : nab);
  }
  reinterpret_cast<uint4 *>(out)[bi] = make_uint4(s.w[0], s.w[1], s.w[2], s.w[3]);
}

// ------------------------------------------------------------------ host side
// CompressSingleColor's per-channel search (Compressor.cpp:268-332), run once on
// the host for every (mode, index mode, p-bit combo, channel class, value).
void build_single_table(uint32_t *tab) {
  static const int kAttr[8][9] = {{4, 3, 3, 0, 4, 0, 0, 0, 1}, {6, 2, 3, 0, 6, 0, 0, 0, 0}, {6, 3, 2, 0, 5, 0, 0, 0, 2},
                                  {6, 2, 2, 0, 7, 0, 0, 0, 1}, {0, 1, 2, 3, 5, 6, 1, 1, 2}, {0, 1, 2, 2, 7, 8, 1, 0, 2},
                                  {0, 1, 4, 0, 7, 7, 0, 0, 1}, {6, 2, 2, 0, 5, 5, 0, 0, 1}};
  static const int kPB[4][2] = {{0, 0}, {0, 1}, {1, 0}, {1, 1}};
  for (int mode = 0; mode < 8; mode++)
    for (int im = 0; im < 2; im++)
      for (int pbi = 0; pbi < 4; pbi++)
        for (int al = 0; al < 2; al++) {
          const int pbt = kAttr[mode][8];
          const int *combo = pbt == kPbitShared ? (pbi ? kPB[3] : kPB[0]) : (pbt == kPbitPerEndpoint ? kPB[pbi] : kPB[0]);
          int nbits = al ? kAttr[mode][5] : kAttr[mode][4];
          const int ibits = im == 0 ? kAttr[mode][2] : kAttr[mode][3];
          uint32_t *row = tab + ((((mode * 2 + im) * 4 + pbi) * 2 + al) << 8);
          if (nbits == 0 || ibits == 0) {
            for (int v = 0; v < 256; v++) row[v] = 0xFF | (0xFF << 8) | ((uint32_t)(0xFF - v) << 16);
            continue;
          }
          const int nvals = 1 << nbits;
          const bool have = pbt != kPbitNone;
          if (have) nbits++;
          int vh[256], vl[256];
          for (int i = 0; i < nvals; i++) {
            int h = i, l = i;
            if (have) { h = (h << 1) | combo[1]; l = (l << 1) | combo[0]; }
            vh[i] = h << (8 - nbits); vh[i] |= vh[i] >> nbits;
            vl[i] = l << (8 - nbits); vl[i] |= vl[i] >> nbits;
          }
          const uint32_t w1 = bc7tab::kWeight[(ibits - 1) * 16 + 1], w0 = 64 - w1;
          for (int v = 0; v < 256; v++) {
            uint32_t best = 0xFF, b1 = 0xFF, b2 = 0xFF;  // memset(0xFF) leaves 0xFFFFFFFF; only the low byte is ever used
            for (int i = 0; best > 0 && i < nvals; i++)
              for (int j = 0; best > 0 && j < nvals; j++) {
                const uint32_t c = (w0 * vl[i] + w1 * vh[j] + 32) >> 6;
                const uint32_t e = c > (uint32_t)v ? c - v : v - c;
                if (e < best) { best = e; b1 = vl[i]; b2 = vh[j]; }
              }
            row[v] = (b1 & 0xFF) | ((b2 & 0xFF) << 8) | (best << 16);
          }
        }
}

size_t ws_bytes(uint32_t nblocks, bool nu) {
  const size_t ntiles = (nblocks + kTile - 1) / kTile;
  size_t b = 0;
  b += ((size_t)nblocks * 4 + 255) & ~(size_t)255;               // sel
  b += ((ntiles + 1) * 4 + 255) & ~(size_t)255;                  // tile counts
  b += 256;                                                      // total
  b += (size_t)nblocks * kSlots * kResWords * 4;                 // results
  b += (size_t)nblocks * kSlots * 8 * 4;                         // states
  b += ((size_t)nblocks * 4 + 255) & ~(size_t)255;               // start-state masks
  b += (size_t)nblocks * kSlots * kStateWords * 4;               // sorted states
  b += kBinWords * 4;                                            // bins
  b += (size_t)kTailCap * 4;                                     // hand-over list of the annealing tail
  if (nu) b += (size_t)nblocks * kSlots * 8;                     // chain errors as doubles (non-uniform metric)
  return b;
}

Ws carve(void *base, uint32_t nblocks, bool nu) {
  const size_t ntiles = (nblocks + kTile - 1) / kTile;
  uint8_t *p = static_cast<uint8_t *>(base);
  Ws w;
  w.sel = reinterpret_cast<uint32_t *>(p); p += ((size_t)nblocks * 4 + 255) & ~(size_t)255;
  w.tile_count = reinterpret_cast<uint32_t *>(p); p += ((ntiles + 1) * 4 + 255) & ~(size_t)255;
  w.total_solid = reinterpret_cast<uint32_t *>(p); p += 256;
  w.counters = nullptr;
  w.wm_running = nullptr;
  w.results = reinterpret_cast<uint32_t *>(p); p += (size_t)nblocks * kSlots * kResWords * 4;
  w.states = reinterpret_cast<uint32_t *>(p); p += (size_t)nblocks * kSlots * 8 * 4;
  w.sa_mask = reinterpret_cast<uint32_t *>(p); p += ((size_t)nblocks * 4 + 255) & ~(size_t)255;
  w.sorted = reinterpret_cast<uint4 *>(p); p += (size_t)nblocks * kSlots * kStateWords * 4;
  w.bins = reinterpret_cast<uint32_t *>(p); p += kBinWords * 4;
  w.tail_list = reinterpret_cast<uint32_t *>(p); p += (size_t)kTailCap * 4;
  w.err64 = nu ? reinterpret_cast<double *>(p) : nullptr;
  w.stats = nullptr;
  return w;
}

cudaError_t ensure_ws(Bc7Workspace &ws, size_t bytes) {
  if (!ws.host_count) {
    cudaError_t e = cudaMallocHost(reinterpret_cast<void **>(&ws.host_count), 64);
    if (e != cudaSuccess) return e;
    e = cudaMalloc(reinterpret_cast<void **>(&ws.wm_running), 64);
    if (e != cudaSuccess) return e;
    e = cudaMalloc(reinterpret_cast<void **>(&ws.counters), 64);
    if (e != cudaSuccess) return e;
    e = cudaMemset(ws.counters, 0, 64);
    if (e != cudaSuccess) return e;
  }
  if (ws.bytes >= bytes) return cudaSuccess;
  if (ws.base) {
    cudaError_t e = cudaFree(ws.base);
    if (e != cudaSuccess) return e;
    ws.base = nullptr; ws.bytes = 0;
  }
  cudaError_t e = cudaMalloc(&ws.base, bytes);
  if (e != cudaSuccess) return e;
  ws.bytes = bytes;
  return cudaSuccess;
}

}  // namespace

cudaError_t bc7_upload_tables() {
  using namespace bc7tab;
  cudaError_t e;
#define UP(sym, src) if ((e = cudaMemcpyToSymbol(sym, src, sizeof(src))) != cudaSuccess) return e
  UP(c_shape2, kShape2); UP(c_shape3, kShape3); UP(c_anchor2, kAnchor2); UP(c_anchor3a, kAnchor3a);
  UP(c_anchor3b, kAnchor3b); UP(c_weight, kWeight); UP(c_opt7, kOpt7Mode5); UP(c_opt6, kOpt6Dxt1);
  UP(c_wm, kWatermark);
#undef UP
  {  // bc7_select's compile-time weights must be the table's 3-bit / 2-bit rows
    static const uint32_t w3[8] = {0, 9, 18, 27, 37, 46, 55, 64}, w2[4] = {0, 21, 43, 64};
    for (int j = 0; j < 8; j++)
      if (kWeight[32 + j] != w3[j]) return cudaErrorInvalidValue;
    for (int j = 0; j < 4; j++)
      if (kWeight[16 + j] != w2[j]) return cudaErrorInvalidValue;
    static const uint32_t w4[16] = {0, 4, 9, 13, 17, 21, 26, 30, 34, 38, 43, 47, 51, 55, 60, 64};
    for (int j = 0; j < 16; j++)
      if (kWeight[48 + j] != w4[j]) return cudaErrorInvalidValue;
  }
  {  // kErrorMetrics[eErrorMetric_Nonuniform] (Compressor.cpp:205-208), rounded like the reference's initialiser
    const float w[4] = {sqrtf(0.3f), sqrtf(0.56f), sqrtf(0.11f), 1.0f};
    if ((e = cudaMemcpyToSymbol(c_nu_weights, w, sizeof(w))) != cudaSuccess) return e;
  }
  bc7_build_quant_table<<<1, 256>>>();
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if ((e = cudaStreamSynchronize(0)) != cudaSuccess) return e;  // callers use non-blocking streams
  static uint32_t host_single[8 * 2 * 4 * 2 * 256];
  static bool built = false;
  if (!built) { build_single_table(host_single); built = true; }
  return cudaMemcpyToSymbol(g_single, host_single, sizeof(host_single));
}

void bc7_free_workspace(Bc7Workspace &ws) {
  if (ws.base) cudaFree(ws.base);
  if (ws.host_count) cudaFreeHost(ws.host_count);
  if (ws.wm_running) cudaFree(ws.wm_running);
  if (ws.counters) cudaFree(ws.counters);
  for (auto &row : ws.ev)
    for (auto &evn : row)
      if (evn) { cudaEventDestroy(evn); evn = nullptr; }
  for (auto &evn : ws.ev_mid)
    if (evn) { cudaEventDestroy(evn); evn = nullptr; }
  ws.base = nullptr; ws.bytes = 0; ws.host_count = nullptr; ws.wm_running = nullptr; ws.counters = nullptr;
}

// Blocks per internal chunk: bounds the scratch (512 B of chain results per block).
constexpr uint32_t kChunkBlocks = 1u << 22;

uint32_t bc7_max_submission() { return kChunkBlocks; }
uint32_t bc7_stat_doubles() { return kStatDoubles; }

// One submission of <= kChunkBlocks blocks, in two halves so that a caller can learn the
// submission's solid-block count (and those of earlier submissions on other streams / GPUs)
// before the watermark base is needed:
//   bc7_front  classify + scan, shape selection, fits, annealing; the solid count is copied to
//              wsp.host_count and `count_ready` (if given) recorded right after the scan
//   bc7_back   sets the watermark base and packs
cudaError_t bc7_front(Bc7Workspace &wsp, const void *rgba_dev, uint32_t width, uint32_t first_block,
                      uint32_t num_blocks, const EncodeParams &prm, uint32_t block_index_base, cudaStream_t stream,
                      cudaEvent_t count_ready, uint32_t *launches) {
  if (num_blocks == 0 || num_blocks > kChunkBlocks) return cudaErrorInvalidValue;
  const bool nu = prm.error_metric != 0;
  wsp.nu = nu;
  cudaError_t e = ensure_ws(wsp, ws_bytes(num_blocks, nu) + 256);
  if (e != cudaSuccess) return e;
  const uint32_t *img = static_cast<const uint32_t *>(rgba_dev);
  const uint32_t bx = width / 4, nb = num_blocks, fb = first_block;
  Ws ws = carve(wsp.base, nb, nu);
  ws.wm_running = wsp.wm_running;
  ws.counters = wsp.counters;
#ifdef FASTC_GPU_COUNTERS
  cudaMemsetAsync(wsp.counters, 0, 16, stream);
#endif
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const uint32_t sa_grid = (uint32_t)sms * kSaCtasPerSm;  // persistent lanes: a multiple of the SM count
  const uint32_t ntiles = (nb + kTile - 1) / kTile;
  uint32_t n = 0;
  cudaEvent_t *ev = nullptr;
  if (wsp.timing && wsp.timed_chunks < Bc7Workspace::kMaxTimedChunks) {
    ev = wsp.ev[wsp.timed_chunks++];
    for (int k = 0; k < 5; k++)
      if (!ev[k] && (e = cudaEventCreate(&ev[k])) != cudaSuccess) return e;
    if (!wsp.ev_mid[wsp.timed_chunks - 1] && (e = cudaEventCreate(&wsp.ev_mid[wsp.timed_chunks - 1])) != cudaSuccess) return e;
    cudaEventRecord(ev[0], stream);
  }
  wsp.cur_ev = ev;
  bc7_classify<<<ntiles, kTile, 0, stream>>>(img, width, bx, fb, nb, ws.sel, ws.tile_count);
  bc7_wm_scan<<<1, 1024, 0, stream>>>(ws.tile_count, ntiles, ws.total_solid);
  n += 2;
  if (count_ready) {
    if ((e = cudaMemcpyAsync(wsp.host_count, ws.total_solid, 4, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
    if ((e = cudaEventRecord(count_ready, stream)) != cudaSuccess) return e;
  }
  if (ev) cudaEventRecord(ev[1], stream);
  if (nu)
    bc7_select<true><<<(nb + kSelWarps - 1) / kSelWarps, kSelWarps * 32, 0, stream>>>(img, width, bx, fb, nb, ws.sel,
                                                                                    prm.block_modes);
  else
    bc7_select<false><<<(nb + kSelWarps - 1) / kSelWarps, kSelWarps * 32, 0, stream>>>(img, width, bx, fb, nb, ws.sel,
                                                                                     prm.block_modes);
  if (ev) cudaEventRecord(ev[2], stream);
  const uint64_t nthreads = (uint64_t)nb * kSlots;
  cudaMemsetAsync(ws.bins, 0, kBinWords * 4, stream);
  cudaMemsetAsync(ws.sa_mask, 0, (size_t)nb * 4, stream);  // no start states yet
  {
    const uint32_t tiles = (nb + kChainThreads - 1) / kChainThreads;
    // the 16-bucket fits (mode 6, whole blocks) are the longest chains: first
#define FASTC_SETUP_LAUNCH(NU_, IB_, ROT_) \
  bc7_setup<NU_, IB_, ROT_><<<tiles, kChainThreads, 0, stream>>>(img, width, bx, fb, nb, ws, prm.quality, prm.seed, block_index_base)
    if (nu) {
      FASTC_SETUP_LAUNCH(true, 4, false);
      FASTC_SETUP_LAUNCH(true, 3, true);
      FASTC_SETUP_LAUNCH(true, 2, true);
      FASTC_SETUP_LAUNCH(true, 3, false);
      FASTC_SETUP_LAUNCH(true, 2, false);
    } else {
      FASTC_SETUP_LAUNCH(false, 4, false);
      FASTC_SETUP_LAUNCH(false, 3, true);
      FASTC_SETUP_LAUNCH(false, 2, true);
      FASTC_SETUP_LAUNCH(false, 3, false);
      FASTC_SETUP_LAUNCH(false, 2, false);
    }
#undef FASTC_SETUP_LAUNCH
  }
  n += 6;
  if (prm.quality > 0) {
    bc7_bin_offsets<<<1, 1, 0, stream>>>(ws.bins, sa_grid);
    bc7_scatter<<<(uint32_t)((nthreads + 255) / 256), 256, 0, stream>>>(ws, nb);
    if (ev) cudaEventRecord(wsp.ev_mid[wsp.timed_chunks - 1], stream);
#ifdef FASTC_GPU_TAILSTATS
    bc7_tail_reset<<<1, 1, 0, stream>>>();
#endif
#ifdef FASTC_GPU_CHAINSTATS
    bc7_chainstats_reset<<<1, 32, 0, stream>>>();
#endif
    // The hand-over to bc7_anneal_tail saves most of the persistent kernel's tail (0.5 - 1 ms per launch)
    // but any edit of that kernel's loop costs ~1 % of its time (profiles/r02_tail_notes.md): it pays
    // below ~1 M blocks per submission (measured: -7 % at 2048^2, -1 % on a 1/8 shard of 8192^2, +0.7 % on a half).
    const bool hand = nb < kTailMaxBlocks;
    if (nu) {
      if (hand) bc7_anneal<true, true><<<sa_grid, kSaThreads, 0, stream>>>(img, width, bx, fb, ws, prm.quality);
      else bc7_anneal<true, false><<<sa_grid, kSaThreads, 0, stream>>>(img, width, bx, fb, ws, prm.quality);
      if (hand) bc7_anneal_tail<true><<<(uint32_t)sms * 4, kSaThreads, 0, stream>>>(img, width, bx, fb, ws, prm.quality);
    } else {
      if (hand) bc7_anneal<false, true><<<sa_grid, kSaThreads, 0, stream>>>(img, width, bx, fb, ws, prm.quality);
      else bc7_anneal<false, false><<<sa_grid, kSaThreads, 0, stream>>>(img, width, bx, fb, ws, prm.quality);
      if (hand) bc7_anneal_tail<false><<<(uint32_t)sms * 4, kSaThreads, 0, stream>>>(img, width, bx, fb, ws, prm.quality);
    }
#ifdef FASTC_GPU_CHAINSTATS
    bc7_chainstats_report<<<1, 1, 0, stream>>>();
#endif
#ifdef FASTC_GPU_TAILSTATS
    bc7_tail_report<<<1, 1, 0, stream>>>();
#endif
    n += hand ? 4 : 3;
  } else if (ev) {
    cudaEventRecord(wsp.ev_mid[wsp.timed_chunks - 1], stream);
  }
  if (ev) cudaEventRecord(ev[3], stream);
  if (launches) *launches += n;
  return cudaGetLastError();
}

// wm_base_dev: device word holding the base (chained submissions of launch_bc7), or NULL: `wm_base`.
cudaError_t bc7_back(Bc7Workspace &wsp, const void *rgba_dev, uint32_t width, uint32_t first_block,
                     uint32_t num_blocks, void *out_dev, uint32_t wm_base, bool base_on_device, cudaStream_t stream,
                     uint32_t *launches, double *stats_dev) {
  if (num_blocks == 0 || !wsp.base) return cudaErrorInvalidValue;
  const uint32_t *img = static_cast<const uint32_t *>(rgba_dev);
  Ws ws = carve(wsp.base, num_blocks, wsp.nu);
  ws.wm_running = wsp.wm_running;
  ws.counters = wsp.counters;
  ws.stats = stats_dev;
  uint32_t n = 0;
  if (!base_on_device) {
    bc7_set_u32<<<1, 1, 0, stream>>>(wsp.wm_running, wm_base);
    n++;
  }
  if (wsp.nu)
    bc7_pack<true><<<(num_blocks + 127) / 128, 128, 0, stream>>>(img, width, width / 4, first_block, num_blocks, ws,
                                                                 static_cast<uint8_t *>(out_dev));
  else
    bc7_pack<false><<<(num_blocks + 127) / 128, 128, 0, stream>>>(img, width, width / 4, first_block, num_blocks, ws,
                                                                  static_cast<uint8_t *>(out_dev));
  n++;
  if (wsp.cur_ev) cudaEventRecord(wsp.cur_ev[4], stream);
  wsp.cur_ev = nullptr;
  if (launches) *launches += n;
  return cudaGetLastError();
}

cudaError_t launch_bc7(Bc7Workspace &wsp, const void *rgba_dev, uint32_t width, uint32_t height,
                       uint32_t first_block, uint32_t num_blocks, void *out_dev, const EncodeParams &prm,
                       uint32_t wm_base, uint32_t block_index_base, cudaStream_t stream, uint32_t *launches) {
  (void)height;
  if (num_blocks == 0) return cudaSuccess;
  const uint32_t chunk = num_blocks < kChunkBlocks ? num_blocks : kChunkBlocks;
  cudaError_t e = ensure_ws(wsp, ws_bytes(chunk, prm.error_metric != 0) + 256);
  if (e != cudaSuccess) return e;
  // The running watermark base lives on the device, so chunks chain without a host sync.
  uint32_t n = 0;
  bc7_set_u32<<<1, 1, 0, stream>>>(wsp.wm_running, wm_base);
  n++;
  wsp.timed_chunks = 0;
  for (uint32_t off = 0; off < num_blocks; off += chunk) {
    const uint32_t nb = num_blocks - off < chunk ? num_blocks - off : chunk;
    const uint32_t fb = first_block + off;
    if ((e = bc7_front(wsp, rgba_dev, width, fb, nb, prm, block_index_base, stream, nullptr, &n)) != cudaSuccess) return e;
    if ((e = bc7_back(wsp, rgba_dev, width, fb, nb, out_dev, 0, true, stream, &n, nullptr)) != cudaSuccess) return e;
    if (off + chunk < num_blocks) {
      Ws ws = carve(wsp.base, nb, wsp.nu);
      bc7_add_u32<<<1, 1, 0, stream>>>(wsp.wm_running, ws.total_solid);
      n++;
    }
  }
  if (launches) *launches = n;
  return cudaGetLastError();
}

cudaError_t bc7_count_solid(Bc7Workspace &wsp, const void *rgba_dev, uint32_t width, uint32_t first_block,
                            uint32_t num_blocks, cudaStream_t stream, uint32_t *count_out) {
  *count_out = 0;
  if (num_blocks == 0) return cudaSuccess;
  const uint32_t ntiles = (num_blocks + kTile - 1) / kTile;
  cudaError_t e = ensure_ws(wsp, ((size_t)(ntiles + 1) * 4 + 511) & ~(size_t)255);
  if (e != cudaSuccess) return e;
  uint32_t *tiles = static_cast<uint32_t *>(wsp.base);
  uint32_t *total = tiles + ntiles;
  bc7_classify<<<ntiles, kTile, 0, stream>>>(static_cast<const uint32_t *>(rgba_dev), width, width / 4, first_block,
                                              num_blocks, nullptr, tiles);
  bc7_wm_scan<<<1, 1024, 0, stream>>>(tiles, ntiles, total);
  e = cudaMemcpyAsync(wsp.host_count, total, 4, cudaMemcpyDeviceToHost, stream);
  if (e != cudaSuccess) return e;
  e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) return e;
  *count_out = *wsp.host_count;
  return cudaGetLastError();
}

cudaError_t bc7_debug_dump(Bc7Workspace &wsp, uint32_t nblocks, uint32_t *sel_out, uint32_t *results_out) {
  if (!wsp.base) return cudaErrorInvalidValue;
  Ws ws = carve(wsp.base, nblocks, wsp.nu);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return e;
  e = cudaMemcpy(sel_out, ws.sel, (size_t)nblocks * 4, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return e;
  return cudaMemcpy(results_out, ws.results, (size_t)nblocks * kSlots * kResWords * 4, cudaMemcpyDeviceToHost);
}

cudaError_t bc7_stage_timing(Bc7Workspace &wsp, int enable, double *ms6) {
  // ms6 = {classify+scan, select, setup(+sort), anneal, pack, total}
  if (ms6) {
    for (int k = 0; k < 6; k++) ms6[k] = 0.0;
    for (int c = 0; c < wsp.timed_chunks; c++) {
      cudaError_t e = cudaEventSynchronize(wsp.ev[c][4]);
      if (e != cudaSuccess) return e;
      cudaEvent_t seq[6] = {wsp.ev[c][0], wsp.ev[c][1], wsp.ev[c][2], wsp.ev_mid[c], wsp.ev[c][3], wsp.ev[c][4]};
      for (int k = 0; k < 5; k++) {
        float ms = 0;
        e = cudaEventElapsedTime(&ms, seq[k], seq[k + 1]);
        if (e != cudaSuccess) return e;
        ms6[k] += ms;
        ms6[5] += ms;
      }
    }
  }
  wsp.timing = enable != 0;
  return cudaSuccess;
}

cudaError_t bc7_read_counters(Bc7Workspace &wsp, uint64_t *qe_calls, uint64_t *pbe) {
  *qe_calls = 0;
  *pbe = 0;
  if (!wsp.counters) return cudaSuccess;
  unsigned long long h[2] = {0, 0};
  cudaError_t e = cudaMemcpy(h, wsp.counters, sizeof(h), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return e;
  *qe_calls = h[0];
  *pbe = h[1];
  return cudaSuccess;
}

}  // namespace fastc
