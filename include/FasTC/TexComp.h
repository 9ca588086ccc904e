// The public compression API -- drop-in for reference Core/include/FasTC/TexComp.h:30-96.
// Global-namespace symbols exactly as in the reference; the block encoders
// behind it are the sm_100a kernels of libfastc_gpu.so (include/fastc_gpu.h).
#ifndef FASTC_B200_TEXCOMP_H_
#define FASTC_B200_TEXCOMP_H_

#include <iosfwd>

#include "FasTC/CompressedImage.h"
#include "FasTC/CompressionJob.h"
#include "FasTC/ImageFwd.h"

class ImageFile;

struct SCompressionSettings {
  SCompressionSettings();  // every field initialised (the reference leaves five unset)

  FasTC::ECompressionFormat format;  // default BPTC
  bool bUseSIMD;          // requesting it fails: "Platform does not support SIMD!" 
  int iNumThreads;        // accepted; placement is decided by the GPU sharder
  int iQuality;           // BPTC: simulated-annealing steps per endpoint fit (default 50)
  int iNumCompressions;   // repeat the compression, report the mean time
  int iJobSize;           // blocks per pipeline chunk (0 = automatic)
  bool bUseAtomics;       // accepted (the GPU path has no separate "atomics" scheduler)
  bool bUsePVRTexLib;     // PVRTC is not supported on the GPU path
  bool bUseNVTT;          // no NVTT back end
  std::ostream *logStream;  // BPTC: per-block statistics lines (path, mode, error of every mode tried) are written here

  // ---- extensions (appended, so reference call sites compile unchanged) ----
  int iNumGPUs;                  // devices to shard block rows over (0 = all visible, default 1)
  unsigned long long uSeed;      // keys the per-block annealing RNG streams (reference: time(NULL))
};

template <typename PixelType>
extern CompressedImage *CompressImage(FasTC::Image<PixelType> *img, const SCompressionSettings &settings);

extern bool CompressImageData(const unsigned char *data, const unsigned int width, const unsigned int height,
                              unsigned char *cmpData, const unsigned int cmpDataSz,
                              const SCompressionSettings &settings);

// Batch submission: every job of the list is compressed (whole textures are dealt round-robin
// to the GPUs).  The GPU counterpart of BPTCC::CompressAtomic over a CompressionJobList
// (reference BPTCEncoder/src/Compressor.cpp:1542-1574), for every supported format.
extern bool CompressImageList(const FasTC::CompressionJobList &jobs, const SCompressionSettings &settings);

extern double ComputePSNR(const CompressedImage &ci, const ImageFile &file);
extern void YieldThread();

#endif
