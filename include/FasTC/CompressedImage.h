// CompressedImage: owns compressed bytes and decodes them on demand.
// API mirror of reference Core/include/FasTC/CompressedImage.h:26-72 and
// Core/src/CompressedImage.cpp:86-149; decoding runs on the GPU (fastc_gpu_decompress).
#ifndef FASTC_B200_COMPRESSEDIMAGE_H_
#define FASTC_B200_COMPRESSEDIMAGE_H_

#include "FasTC/CompressionFormat.h"
#include "FasTC/Image.h"
#include "FasTC/TexCompTypes.h"

class CompressedImage : public FasTC::Image<FasTC::Pixel> {
 public:
  CompressedImage(const CompressedImage &);
  CompressedImage &operator=(const CompressedImage &);
  // `data` holds GetCompressedSize(width, height, format) bytes and is copied.
  CompressedImage(const uint32 width, const uint32 height, const FasTC::ECompressionFormat format,
                  const uint8 *data);
  virtual ~CompressedImage();

  virtual FasTC::Image<FasTC::Pixel> *Clone() const { return new CompressedImage(*this); }
  virtual void ComputePixels();

  static uint32 GetCompressedSize(uint32 width, uint32 height, FasTC::ECompressionFormat format);
  uint32 GetCompressedSize() const { return GetCompressedSize(GetWidth(), GetHeight(), m_Format); }
  uint32 GetUncompressedSize() const { return GetWidth() * GetHeight() * sizeof(uint32); }

  // Decodes into outBuf (width*height*4 bytes).  false on a short buffer, an unsupported
  // format or a GPU error.
  bool DecompressImage(uint8 *outBuf, uint32 outBufSz) const;

  const uint8 *GetCompressedData() const { return m_CompressedData; }
  FasTC::ECompressionFormat GetFormat() const { return m_Format; }

 private:
  FasTC::ECompressionFormat m_Format;
  uint8 *m_CompressedData;
};
#endif
