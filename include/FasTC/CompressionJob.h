// FasTC::CompressionJob / DecompressionJob / CompressionJobList.
// API mirror of reference Base/include/FasTC/CompressionJob.h:40-209 and
// Base/src/CompressionJob.cpp:29-106: a job is a raster-order block range
// [CoordsToBlockIdx(XStart,YStart), CoordsToBlockIdx(XEnd,YEnd)) of a Width x Height image;
// block i is written at OutBuf + i * GetBlockSize(format).
#ifndef FASTC_B200_COMPRESSIONJOB_H_
#define FASTC_B200_COMPRESSIONJOB_H_

#include "FasTC/CompressionFormat.h"
#include "FasTC/TexCompTypes.h"

namespace FasTC {

class CompressionJob {
 public:
  CompressionJob(ECompressionFormat fmt, const uint8 *inBuf, unsigned char *outBuf, const uint32 width,
                 const uint32 height)
      : m_Format(fmt), m_InBuf(inBuf), m_OutBuf(outBuf), m_Width(width), m_Height(height),
        m_XStart(0), m_XEnd(width), m_YStart(0), m_YEnd(height) {}

  CompressionJob(ECompressionFormat fmt, const uint8 *inBuf, unsigned char *outBuf, const uint32 width,
                 const uint32 height, const uint32 xOffset, const uint32 yOffset)
      : m_Format(fmt), m_InBuf(inBuf), m_OutBuf(outBuf), m_Width(width), m_Height(height),
        m_XStart(xOffset), m_XEnd(width), m_YStart(yOffset), m_YEnd(height) {}

  CompressionJob(ECompressionFormat fmt, const uint8 *inBuf, unsigned char *outBuf, const uint32 width,
                 const uint32 height, const uint32 xOffset, const uint32 yOffset, const uint32 xEndpoint,
                 const uint32 yEndpoint)
      : m_Format(fmt), m_InBuf(inBuf), m_OutBuf(outBuf), m_Width(width), m_Height(height),
        m_XStart(xOffset), m_XEnd(xEndpoint), m_YStart(yOffset), m_YEnd(yEndpoint) {}

  ECompressionFormat Format() const { return m_Format; }
  const uint8 *InBuf() const { return m_InBuf; }
  uint8 *OutBuf() const { return m_OutBuf; }
  uint32 Width() const { return m_Width; }
  uint32 Height() const { return m_Height; }
  uint32 XStart() const { return m_XStart; }
  uint32 XEnd() const { return m_XEnd; }
  uint32 YStart() const { return m_YStart; }
  uint32 YEnd() const { return m_YEnd; }

  // Pixel coordinates of the top-left corner of block `blockIdx`.
  void BlockIdxToCoords(uint32 blockIdx, uint32 (&out)[2]) const {
    uint32 dim[2];
    GetBlockDimensions(m_Format, dim);
    const uint32 blocksX = m_Width / dim[0];
    out[0] = (blockIdx % blocksX) * dim[0];
    out[1] = (blockIdx / blocksX) * dim[1];
  }

  // Raster index of the block containing pixel (x, y).
  uint32 CoordsToBlockIdx(uint32 x, uint32 y) const {
    uint32 dim[2];
    GetBlockDimensions(m_Format, dim);
    return (y / dim[1]) * (m_Width / dim[0]) + x / dim[0];
  }

  // First block and block count of the job's range (what the reference's encoder loops
  // iterate: BPTCEncoder/src/Compressor.cpp:1476-1519 and its DXT / ETC twins).  A job whose
  // end point is (Width, Height) or (0, Height) runs to the end of the image.
  uint32 FirstBlock() const { return CoordsToBlockIdx(m_XStart, m_YStart); }
  uint32 NumBlocks() const {
    uint32 dim[2];
    GetBlockDimensions(m_Format, dim);
    const uint32 total = (m_Width / dim[0]) * (m_Height / dim[1]);
    uint32 end = (m_YEnd >= m_Height) ? total : CoordsToBlockIdx(m_XEnd >= m_Width ? 0 : m_XEnd, m_YEnd) +
                                                    (m_XEnd >= m_Width ? m_Width / dim[0] : 0);
    if (end > total) end = total;
    const uint32 first = FirstBlock();
    return end > first ? end - first : 0;
  }

 private:
  ECompressionFormat m_Format;
  const uint8 *m_InBuf;
  uint8 *m_OutBuf;
  uint32 m_Width, m_Height;
  uint32 m_XStart, m_XEnd;
  uint32 m_YStart, m_YEnd;
};

class DecompressionJob {
 public:
  DecompressionJob(ECompressionFormat fmt, const uint8 *inBuf, uint8 *outBuf, uint32 width, uint32 height)
      : m_Format(fmt), m_InBuf(inBuf), m_OutBuf(outBuf), m_Width(width), m_Height(height) {}
  const uint8 *InBuf() const { return m_InBuf; }
  uint8 *OutBuf() const { return m_OutBuf; }
  uint32 Width() const { return m_Width; }
  uint32 Height() const { return m_Height; }
  ECompressionFormat Format() const { return m_Format; }

 private:
  const ECompressionFormat m_Format;
  const uint8 *m_InBuf;
  uint8 *m_OutBuf;
  const uint32 m_Width, m_Height;
};

// Fixed-capacity list of textures to compress in one submission.
class CompressionJobList {
 public:
  explicit CompressionJobList(const uint32 nJobs);
  ~CompressionJobList();
  CompressionJobList(const CompressionJobList &);
  CompressionJobList &operator=(const CompressionJobList &);

  bool AddJob(const CompressionJob &);  // false when the list is full
  uint32 GetTotalNumJobs() const { return m_TotalNumJobs; }
  uint32 GetNumJobs() const { return m_NumJobs; }
  // Unlike the reference (which returns the *current* job whatever idx is) this
  // returns job idx, or NULL when idx is out of range.
  const CompressionJob *GetJob(uint32 idx) const;
  uint32 *GetFinishedFlag(uint32 idx) const;

 private:
  CompressionJob *m_Jobs;
  uint32 m_NumJobs;
  uint32 m_TotalNumJobs;
  uint32 *m_FinishedFlags;

 public:
  uint32 m_CurrentJobIndex;
  uint32 m_CurrentBlockIndex;
};

}  // namespace FasTC
#endif
