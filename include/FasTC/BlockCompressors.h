// The per-format operator interface the reference's schedulers call:
//   typedef void (*CompressionFunc)(const FasTC::CompressionJob &)   (reference Core/src/CompressionFuncs.h:29)
// with the reference's own entry-point names, so a host that keeps its Serial / ThreadGroup /
// WorkerQueue code can swap the block encoders alone:
//   BPTCC::Compress           reference BPTCEncoder/include/FasTC/BPTCCompressor.h:168-169
//   DXTC::CompressImageDXT1/5 reference DXTEncoder/include/FasTC/DXTCompressor.h:24-25
//   ETCC::Compress_RG         reference ETCEncoder/include/FasTC/ETCCompressor.h:35
// Each call encodes the job's raster block range on the current CUDA device and returns when
// the bytes are in cj.OutBuf().  They are re-entrant (no global RNG / watermark state: the
// watermark order is derived from the block index).  Errors cannot be returned
// through this signature: they are printed ("TexComp -- ...") and the range is left untouched.
#ifndef FASTC_B200_BLOCKCOMPRESSORS_H_
#define FASTC_B200_BLOCKCOMPRESSORS_H_

#include "FasTC/CompressionJob.h"

typedef void (*CompressionFunc)(const FasTC::CompressionJob &);

namespace BPTCC {
struct CompressionSettings {
  // reference BPTCCompressor.h:123-158.  m_ShapeSelectionFn is a host callback and cannot run
  // on the device: it must stay NULL (a call with it set is rejected).  m_BlockModes (bit m = mode
  // m allowed) and m_ErrorMetric (0 = eErrorMetric_Uniform, 1 = eErrorMetric_Nonuniform) are
  // honoured like the reference's (Compressor.cpp:1857, :205-208).
  void *m_ShapeSelectionFn;
  const void *m_ShapeSelectionUserData;
  uint32 m_BlockModes;
  int m_ErrorMetric;
  uint32 m_NumSimulatedAnnealingSteps;
  CompressionSettings()
      : m_ShapeSelectionFn(0), m_ShapeSelectionUserData(0), m_BlockModes(0xFF), m_ErrorMetric(0),
        m_NumSimulatedAnnealingSteps(50) {}
};
void Compress(const FasTC::CompressionJob &, CompressionSettings settings = CompressionSettings());
void Decompress(const FasTC::DecompressionJob &);
}  // namespace BPTCC

namespace DXTC {
void CompressImageDXT1(const FasTC::CompressionJob &);
void CompressImageDXT5(const FasTC::CompressionJob &);
void DecompressDXT1(const FasTC::DecompressionJob &);
void DecompressDXT5(const FasTC::DecompressionJob &);
}  // namespace DXTC

namespace ETCC {
void Compress_RG(const FasTC::CompressionJob &);
void Decompress(const FasTC::DecompressionJob &);
}  // namespace ETCC

#endif
