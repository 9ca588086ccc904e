// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Thin extern "C" harness around the UNMODIFIED reference (GammaUNC/FasTC,
// compiled from /root/reference by oracle/Makefile into oracle/_ref/).  Only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load the resulting libfastc_ref.so.
//
// What it exposes:
//   fastc_ref_compress    -> the reference's own CompressImageData
//                            (Core/src/TexComp.cpp:427-525), i.e. exactly what
//                            `tc -f <fmt> -q <q> -t <threads> -j <job>` runs.
//   fastc_ref_decompress  -> CompressedImage::DecompressImage
//                            (Core/src/CompressedImage.cpp:86-119).
//   fastc_ref_psnr        -> FasTC::Image<Pixel>::ComputePSNR
//                            (Base/src/Image.cpp:205-255).
//   fastc_ref_set_state   -> pokes the reference's two process-global
//                            variables (BPTCEncoder/src/Compressor.cpp:140
//                            gWMVal, :358 g_seed).  They are file-static in the
//                            reference; oracle/Makefile makes the two symbols
//                            global with `objcopy --globalize-symbol` on the
//                            compiled object, so no reference source is edited.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <unistd.h>
#include <fcntl.h>

#include "FasTC/TexComp.h"
#include "FasTC/ImageFile.h"
#include "rg_etc1.h"
#include "FasTC/BPTCCompressor.h"
#include "FasTC/CompressionJob.h"
#include "FasTC/CompressedImage.h"
#include "FasTC/CompressionFormat.h"
#include "FasTC/Image.h"
#include "FasTC/Pixel.h"

// BPTCC::g_seed / BPTCC::gWMVal (internal linkage in the reference; globalized
// post-compile by the Makefile).
extern uint32_t fastc_ref_g_seed asm("_ZN5BPTCCL6g_seedE");
extern uint32_t fastc_ref_g_wm asm("_ZN5BPTCCL6gWMValE");

namespace {
FasTC::ECompressionFormat to_format(int f) {
  switch (f) {
    case 0: return FasTC::eCompressionFormat_DXT1;
    case 1: return FasTC::eCompressionFormat_DXT5;
    case 2: return FasTC::eCompressionFormat_ETC1;
    case 4: return FasTC::eCompressionFormat_PVRTC4;
    default: return FasTC::eCompressionFormat_BPTC;
  }
}
}  // namespace

extern "C" {

// format: 0=DXT1 1=DXT5 2=ETC1 3=BPTC 4=PVRTC4 (same numbering as include/fastc_gpu.h)
// Returns 0 on success. *ms receives the wall time of the CompressImageData call.
int fastc_ref_compress(int format, const uint8_t *rgba, uint32_t width,
                       uint32_t height, uint8_t *out, uint32_t out_size,
                       int quality, int threads, int job_size, double *ms) {
  SCompressionSettings s;
  s.format = to_format(format);
  s.bUseSIMD = false;
  s.iNumThreads = threads < 1 ? 1 : threads;
  s.iQuality = quality;
  s.iNumCompressions = 1;
  s.iJobSize = job_size;
  s.bUseAtomics = false;
  s.bUsePVRTexLib = false;
  s.bUseNVTT = false;
  s.logStream = NULL;

  // The reference prints "Compression time: ..." to stdout; keep the caller's
  // stdout clean (bench.py prints exactly one JSON line).
  fflush(stdout);
  int saved = dup(1);
  int devnull = open("/dev/null", O_WRONLY);
  if (devnull >= 0) { dup2(devnull, 1); close(devnull); }

  auto t0 = std::chrono::steady_clock::now();
  bool ok = CompressImageData(rgba, width, height, out, out_size, s);
  auto t1 = std::chrono::steady_clock::now();

  fflush(stdout);
  if (saved >= 0) { dup2(saved, 1); close(saved); }
  if (ms) *ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
  return ok ? 0 : 1;
}

int fastc_ref_decompress(int format, const uint8_t *cmp, uint32_t width,
                         uint32_t height, uint8_t *rgba_out) {
  CompressedImage ci(width, height, to_format(format), cmp);
  return ci.DecompressImage(rgba_out, width * height * 4) ? 0 : 1;
}

double fastc_ref_psnr(const uint8_t *a, const uint8_t *b, uint32_t width,
                      uint32_t height) {
  const uint32_t n = width * height;
  FasTC::Pixel *pa = new FasTC::Pixel[n];
  FasTC::Pixel *pb = new FasTC::Pixel[n];
  const uint32_t *ua = reinterpret_cast<const uint32_t *>(a);
  const uint32_t *ub = reinterpret_cast<const uint32_t *>(b);
  for (uint32_t i = 0; i < n; i++) { pa[i].Unpack(ua[i]); pb[i].Unpack(ub[i]); }
  FasTC::Image<FasTC::Pixel> ia(width, height, pa);
  FasTC::Image<FasTC::Pixel> ib(width, height, pb);
  delete[] pa;
  delete[] pb;
  return ia.ComputePSNR(&ib);
}

// BPTCC::Compress(job, settings) over the whole image (BPTCEncoder/src/Compressor.cpp:1473): the
// reference's per-format operator entry with non-default CompressionSettings -- m_BlockModes and
// m_ErrorMetric are not reachable through CompressImageData.
int fastc_ref_bptc_compress_settings(const uint8_t *rgba, uint32_t width, uint32_t height, uint8_t *out,
                                     int quality, uint32_t block_modes, int error_metric) {
  BPTCC::CompressionSettings s;
  s.m_NumSimulatedAnnealingSteps = (uint32_t)quality;
  s.m_BlockModes = block_modes;
  s.m_ErrorMetric = error_metric ? BPTCC::eErrorMetric_Nonuniform : BPTCC::eErrorMetric_Uniform;
  FasTC::CompressionJob cj(FasTC::eCompressionFormat_BPTC, rgba, out, width, height);
  BPTCC::Compress(cj, s);
  return 0;
}

// rg_etc1::pack_etc1_block over the whole image at a given quality (FasTC's ETCC::Compress_RG,
// ETCEncoder/src/Compressor.cpp:26-54, hard-codes cLowQuality; this is its block loop with the
// library's other two levels).
int fastc_ref_etc1_compress_quality(const uint8_t *rgba, uint32_t width, uint32_t height, uint8_t *out, int quality) {
  rg_etc1::etc1_pack_params params;
  params.m_quality = quality == 0 ? rg_etc1::cLowQuality : (quality == 1 ? rg_etc1::cMediumQuality : rg_etc1::cHighQuality);
  params.m_dithering = false;
  rg_etc1::pack_etc1_block_init();
  const uint32_t *in = reinterpret_cast<const uint32_t *>(rgba);
  for (uint32_t j = 0; j < height; j += 4)
    for (uint32_t i = 0; i < width; i += 4) {
      uint32_t pixels[16];
      for (int r = 0; r < 4; r++) memcpy(pixels + 4 * r, in + (size_t)(j + r) * width + i, 16);
      rg_etc1::pack_etc1_block(out, pixels, params);
      out += 8;
    }
  return 0;
}

// The reference's own KTX writer (IO/src/ImageWriterKTX.cpp:69-160) through ImageFile::Write.
// NOTE the writer keeps the payload size in function-local statics, so one process must use one
// image size per format family (BPTC / DXT5, DXT1, RGBA).
int fastc_ref_write_ktx(int format, const uint8_t *cmp, uint32_t width, uint32_t height, const char *path) {
  CompressedImage ci(width, height, to_format(format), cmp);
  ImageFile f(path, eFileFormat_KTX, ci);
  return f.Write() ? 0 : 1;
}

uint32_t fastc_ref_compressed_size(int format, uint32_t width, uint32_t height) {
  return CompressedImage::GetCompressedSize(width, height, to_format(format));
}

// seed: value of the reference's LCG state before the next compression;
// wm: number of solid-colour BC7 blocks "already written" (0 = fresh process).
void fastc_ref_set_state(uint32_t seed, uint32_t wm_count) {
  fastc_ref_g_seed = seed;
  fastc_ref_g_wm = wm_count == 0 ? 0xFFFFFFFFu : ((wm_count - 1) % 9);
}

uint32_t fastc_ref_get_seed(void) { return fastc_ref_g_seed; }

}  // extern "C"
