// Host side of the compression path: FasTC's Core API over the C ABI of libfastc_gpu.so.
//
// Mirrors, for the GPU path, reference Core/src/TexComp.cpp (CompressImageData :427-525,
// CompressImage<> :367-420, ChooseFuncFromSettings :110-155, the Serial / ThreadGroup /
// WorkerQueue / Atomics drivers :161-365) and Base/src/CompressionJob.cpp.  Where the
// reference picks a CompressionFunc and a pthread scheduler, this hands the whole job to
// fastc_gpu_compress, which shards the block rows over GPUs and pipeline chunks
// (fastc_b200/csrc/capi.cu).  Same error messages on stderr ("TexComp -- ..."), same
// "Compression time: %0.3f ms" line on stdout, bool returns, no exceptions.
#include "FasTC/TexComp.h"

#include <cstdio>
#include <cstring>
#include <ctime>
#include <ostream>
#include <sched.h>
#include <vector>

#include "FasTC/BlockCompressors.h"
#include "FasTC/Image.h"
#include "FasTC/ImageFile.h"
#include "FasTC/Pixel.h"
#include "fastc_gpu.h"

using FasTC::CompressionJob;
using FasTC::CompressionJobList;
using FasTC::ECompressionFormat;

namespace {

void ReportError(const char *msg) { fprintf(stderr, "TexComp -- %s\n", msg); }

// FasTC format -> C ABI format, or -1 when the GPU path has no encoder for it.
int GpuFormat(ECompressionFormat f) {
  switch (f) {
    case FasTC::eCompressionFormat_DXT1: return FASTC_GPU_DXT1;
    case FasTC::eCompressionFormat_DXT5: return FASTC_GPU_DXT5;
    case FasTC::eCompressionFormat_ETC1: return FASTC_GPU_ETC1;
    case FasTC::eCompressionFormat_BPTC: return FASTC_GPU_BPTC;
    case FasTC::eCompressionFormat_PVRTC4: return FASTC_GPU_PVRTC4;
    default: return -1;
  }
}

// One CompressionFunc call: the job's raster block range on the current device.
void RunJob(const CompressionJob &cj, int quality, unsigned long long seed, const fastc_gpu_options *opt = NULL) {
  const int fmt = GpuFormat(cj.Format());
  if (fmt < 0) {
    ReportError("Could not find adequate compression function for specified settings");
    return;
  }
  const uint32 n = cj.NumBlocks();
  if (n == 0) return;
  if (fastc_gpu_compress_opt(fmt, cj.InBuf(), cj.Width(), cj.Height(), cj.FirstBlock(), n, cj.OutBuf(), quality, seed, 0,
                             1, NULL, opt) != 0)
    ReportError(fastc_gpu_last_error());
}

}  // namespace

// ---------------------------------------------------------------------------------------
// The reference's per-format operator entry points (FasTC/BlockCompressors.h).
namespace BPTCC {
void Compress(const CompressionJob &cj, CompressionSettings settings) {
  if (settings.m_ShapeSelectionFn) {
    ReportError("BPTC shape-selection callbacks cannot run on the GPU path");
    return;
  }
  // m_BlockModes restricts the mode search, m_ErrorMetric weights the channel errors
  // (reference Compressor.cpp:1848-1857, :205-208)
  fastc_gpu_options opt = FASTC_GPU_OPTIONS_INIT;
  opt.bptc_block_modes = settings.m_BlockModes;
  opt.bptc_error_metric = (int)settings.m_ErrorMetric;
  RunJob(cj, (int)settings.m_NumSimulatedAnnealingSteps, 0, &opt);
}
void Decompress(const FasTC::DecompressionJob &dj) {
  if (fastc_gpu_decompress(FASTC_GPU_BPTC, dj.InBuf(), dj.Width(), dj.Height(), dj.OutBuf(), NULL) != 0)
    ReportError(fastc_gpu_last_error());
}
}  // namespace BPTCC

namespace DXTC {
void CompressImageDXT1(const CompressionJob &cj) { RunJob(cj, 0, 0); }
void CompressImageDXT5(const CompressionJob &cj) { RunJob(cj, 0, 0); }
void DecompressDXT1(const FasTC::DecompressionJob &dj) {
  if (fastc_gpu_decompress(FASTC_GPU_DXT1, dj.InBuf(), dj.Width(), dj.Height(), dj.OutBuf(), NULL) != 0)
    ReportError(fastc_gpu_last_error());
}
void DecompressDXT5(const FasTC::DecompressionJob &dj) {
  if (fastc_gpu_decompress(FASTC_GPU_DXT5, dj.InBuf(), dj.Width(), dj.Height(), dj.OutBuf(), NULL) != 0)
    ReportError(fastc_gpu_last_error());
}
}  // namespace DXTC

namespace ETCC {
void Compress_RG(const CompressionJob &cj) { RunJob(cj, 0, 0); }
void Decompress(const FasTC::DecompressionJob &dj) {
  if (fastc_gpu_decompress(FASTC_GPU_ETC1, dj.InBuf(), dj.Width(), dj.Height(), dj.OutBuf(), NULL) != 0)
    ReportError(fastc_gpu_last_error());
}
}  // namespace ETCC

// ---------------------------------------------------------------------------------------
SCompressionSettings::SCompressionSettings()
    : format(FasTC::eCompressionFormat_BPTC), bUseSIMD(false), iNumThreads(1), iQuality(50), iNumCompressions(1),
      iJobSize(0), bUseAtomics(false), bUsePVRTexLib(false), bUseNVTT(false), logStream(NULL), iNumGPUs(1),
      uSeed(0) {}

namespace {
// Checks shared by CompressImageData and CompressImageList (reference TexComp.cpp:436-496).
bool ValidateRequest(const SCompressionSettings &settings, uint32 width, uint32 height, uint32 cmpDataSz,
                     bool haveSize) {
  if (settings.bUseSIMD) {
    ReportError("Platform does not support SIMD!\n");
    return false;
  }
  if ((uint64)width * height == 0) {
    ReportError("No data sent to compress!");
    return false;
  }
  uint32 blockDims[2];
  FasTC::GetBlockDimensions(settings.format, blockDims);
  if ((width % blockDims[0]) != 0 || (height % blockDims[1]) != 0) {
    ReportError("ERROR - CompressImageData: width or height is not multiple of block dimension");
    return false;
  }
  if (settings.format >= FasTC::kNumCompressionFormats) {
    ReportError("Unknown compression format");
    return false;
  }
  if (haveSize && CompressedImage::GetCompressedSize(width, height, settings.format) > cmpDataSz) {
    ReportError("Not enough space for compressed data!");
    return false;
  }
  if (GpuFormat(settings.format) < 0 || settings.bUsePVRTexLib || settings.bUseNVTT) {
    // PVRTC2 and ASTC have no encoder in FasTC either; the external libraries are not linked
    ReportError("Could not find adequate compression function for specified settings");
    return false;
  }
  if (settings.format == FasTC::eCompressionFormat_PVRTC4 &&
      (width != height || (width & (width - 1)) != 0 || width < 8)) {  // reference TexComp.cpp:477-482
    ReportError("ERROR - CompressImageData: PVRTC4 images must be square and power-of-two.");
    return false;
  }
  if (settings.iQuality < 0) {
    ReportError("Quality must not be negative");
    return false;
  }
  return true;
}
}  // namespace

namespace {
// The per-block lines of the reference's BPTCC::CompressWithStats log ("<block>: <stat> -- <value>",
// BPTCEncoder/src/Compressor.cpp:106-131, :1947-1951, :1982-1992): path, mode, and per mode the
// estimate and the error.  The GPU selection kernel keeps no per-mode estimates: they are logged as -1,
// the value the reference logs for "not computed".
void WriteBlockStats(std::ostream &os, const std::vector<fastc_gpu_bptc_block_stat> &stats) {
  static const char *kModeName[8] = {"Zero", "One", "Two", "Three", "Four", "Five", "Six", "Seven"};
  for (size_t i = 0; i < stats.size(); i++) {
    const fastc_gpu_bptc_block_stat &b = stats[i];
    os << i << ": BlockStat_Path -- " << (int)b.path << std::endl;
    os << i << ": BlockStat_Mode -- " << (int)b.mode << std::endl;
    for (int m = 0; m < 8; m++) {
      os << i << ": BlockStat_Mode" << kModeName[m] << "Estimate -- " << -1.0 << std::endl;
      os << i << ": BlockStat_Mode" << kModeName[m] << "Error -- " << b.mode_error[m] << std::endl;
    }
  }
}
}  // namespace

bool CompressImageData(const unsigned char *data, const unsigned int width, const unsigned int height,
                       unsigned char *cmpData, const unsigned int cmpDataSz, const SCompressionSettings &settings) {
  if (!ValidateRequest(settings, width, height, cmpDataSz, true)) return false;
  if (!data || !cmpData) {
    ReportError("No data sent to compress!");
    return false;
  }
  // Scheduler knobs (reference TexComp.cpp:502-514): iNumThreads / bUseAtomics choose a pthread
  // scheduler there; here the sharder is the GPU library.  iJobSize keeps its meaning of
  // "blocks handed out at a time" as the pipeline chunk size.
  const int reps = settings.iNumCompressions > 0 ? settings.iNumCompressions : 0;
  double total_ms = 0.0;
  // logStream + BPTC: the reference switches to its statistics-gathering encoder
  // (ChooseFuncFromSettings, TexComp.cpp:110-155); here the pack kernel fills per-block records
  const bool wantStats = settings.logStream != NULL && settings.format == FasTC::eCompressionFormat_BPTC;
  std::vector<fastc_gpu_bptc_block_stat> stats(wantStats ? (size_t)(width / 4) * (height / 4) : 0);
  fastc_gpu_options opt = FASTC_GPU_OPTIONS_INIT;
  opt.bptc_block_stats = wantStats ? stats.data() : NULL;
  for (int i = 0; i < reps; i++) {
    fastc_gpu_timing tm;
    if (fastc_gpu_compress_opt(GpuFormat(settings.format), data, width, height, 0, 0, cmpData, settings.iQuality,
                               settings.uSeed, settings.iJobSize > 0 ? (uint32)settings.iJobSize : 0,
                               settings.iNumGPUs, &tm, &opt) != 0) {
      ReportError(fastc_gpu_last_error());
      return false;
    }
    total_ms += tm.total_ms;
  }
  if (wantStats && reps) WriteBlockStats(*settings.logStream, stats);
  fprintf(stdout, "Compression time: %0.3f ms\n", reps ? total_ms / reps : 0.0);
  return true;
}

bool CompressImageList(const CompressionJobList &jobs, const SCompressionSettings &settings) {
  const uint32 n = jobs.GetNumJobs();
  std::vector<fastc_gpu_job> list(n);
  for (uint32 i = 0; i < n; i++) {
    const CompressionJob *cj = jobs.GetJob(i);
    if (!cj || cj->Format() != settings.format) {
      ReportError("Job format does not match the compression settings");
      return false;
    }
    if (!ValidateRequest(settings, cj->Width(), cj->Height(), 0, false)) return false;
    if (cj->FirstBlock() != 0 || cj->NumBlocks() != (cj->Width() / 4) * (cj->Height() / 4)) {
      ReportError("Batch submissions take whole-image jobs");
      return false;
    }
    list[i].rgba_host = cj->InBuf();
    list[i].out_host = cj->OutBuf();
    list[i].width = cj->Width();
    list[i].height = cj->Height();
  }
  fastc_gpu_timing tm;
  if (n && fastc_gpu_compress_batch(GpuFormat(settings.format), list.data(), n, settings.iQuality, settings.uSeed,
                                    settings.iNumGPUs, &tm) != 0) {
    ReportError(fastc_gpu_last_error());
    return false;
  }
  for (uint32 i = 0; i < n; i++) *jobs.GetFinishedFlag(i) = 1;
  fprintf(stdout, "Compression time: %0.3f ms\n", n ? tm.total_ms : 0.0);
  return true;
}

template <typename PixelType>
CompressedImage *CompressImage(FasTC::Image<PixelType> *img, const SCompressionSettings &settings) {
  if (!img) return NULL;
  uint32 width = img->GetWidth(), height = img->GetHeight();
  uint32 blockDims[2];
  FasTC::GetBlockDimensions(settings.format, blockDims);
  if ((width % blockDims[0]) != 0 || (height % blockDims[1]) != 0) {
    ReportError("WARNING - Image size is not a multiple of block size. Padding with zeros...");
    width = ((width + blockDims[0] - 1) / blockDims[0]) * blockDims[0];
    height = ((height + blockDims[1] - 1) / blockDims[1]) * blockDims[1];
  }
  std::vector<uint32> data((size_t)width * height, 0u);
  img->ComputePixels();
  for (uint32 j = 0; j < img->GetHeight(); j++)
    for (uint32 i = 0; i < img->GetWidth(); i++) data[(size_t)j * width + i] = (*img)(i, j).Pack();
  const uint32 cmpDataSz = CompressedImage::GetCompressedSize(width, height, settings.format);
  std::vector<uint8> cmpData(cmpDataSz);
  if (!CompressImageData(reinterpret_cast<const uint8 *>(data.data()), width, height, cmpData.data(), cmpDataSz,
                         settings))
    return NULL;
  return new CompressedImage(width, height, settings.format, cmpData.data());
}
template CompressedImage *CompressImage(FasTC::Image<FasTC::Pixel> *, const SCompressionSettings &settings);

// Declared by the reference (Core/include/FasTC/TexComp.h:93) and never defined there; here it is
// the PSNR `tc` prints: the compressed image decoded on the GPU against the file's pixels.
double ComputePSNR(const CompressedImage &ci, const ImageFile &file) {
  FasTC::Image<> *img = file.GetImage();
  if (!img) return -1.0;
  CompressedImage copy(ci);
  return img->ComputePSNR(&copy);
}

void YieldThread() { sched_yield(); }

// ---------------------------------------------------------------------------------------
// CompressionJobList (reference Base/src/CompressionJob.cpp:29-106)
namespace FasTC {

CompressionJobList::CompressionJobList(const uint32 nJobs)
    : m_Jobs(static_cast<CompressionJob *>(operator new(sizeof(CompressionJob) * (nJobs ? nJobs : 1)))), m_NumJobs(0),
      m_TotalNumJobs(nJobs), m_FinishedFlags(new uint32[nJobs ? nJobs : 1]()), m_CurrentJobIndex(0),
      m_CurrentBlockIndex(0) {}

CompressionJobList::~CompressionJobList() {
  operator delete(m_Jobs);  // CompressionJob is trivially destructible
  delete[] m_FinishedFlags;
}

CompressionJobList::CompressionJobList(const CompressionJobList &o)
    : m_Jobs(static_cast<CompressionJob *>(operator new(sizeof(CompressionJob) * (o.m_TotalNumJobs ? o.m_TotalNumJobs : 1)))),
      m_NumJobs(o.m_NumJobs), m_TotalNumJobs(o.m_TotalNumJobs),
      m_FinishedFlags(new uint32[o.m_TotalNumJobs ? o.m_TotalNumJobs : 1]()), m_CurrentJobIndex(o.m_CurrentJobIndex),
      m_CurrentBlockIndex(o.m_CurrentBlockIndex) {
  for (uint32 i = 0; i < m_NumJobs; i++) {
    new (&m_Jobs[i]) CompressionJob(o.m_Jobs[i]);
    m_FinishedFlags[i] = o.m_FinishedFlags[i];
  }
}

CompressionJobList &CompressionJobList::operator=(const CompressionJobList &o) {
  if (this == &o) return *this;
  CompressionJobList tmp(o);
  CompressionJob *j = m_Jobs; m_Jobs = tmp.m_Jobs; tmp.m_Jobs = j;
  uint32 *f = m_FinishedFlags; m_FinishedFlags = tmp.m_FinishedFlags; tmp.m_FinishedFlags = f;
  m_NumJobs = tmp.m_NumJobs;
  m_TotalNumJobs = tmp.m_TotalNumJobs;
  m_CurrentJobIndex = tmp.m_CurrentJobIndex;
  m_CurrentBlockIndex = tmp.m_CurrentBlockIndex;
  return *this;
}

bool CompressionJobList::AddJob(const CompressionJob &cj) {
  if (m_NumJobs == m_TotalNumJobs) return false;
  new (&m_Jobs[m_NumJobs++]) CompressionJob(cj);
  return true;
}

const CompressionJob *CompressionJobList::GetJob(uint32 idx) const { return idx < m_NumJobs ? &m_Jobs[idx] : NULL; }

uint32 *CompressionJobList::GetFinishedFlag(uint32 idx) const { return idx < m_NumJobs ? &m_FinishedFlags[idx] : NULL; }

}  // namespace FasTC
