// CompressedImage -- a texture held in its block-compressed form.
//
// Source-compatible with the class FasTC applications use (reference
// Core/include/FasTC/CompressedImage.h:26-72, behaviour of Core/src/CompressedImage.cpp:86-149):
// it owns a copy of the compressed payload, knows its format, reports the payload size, and can
// expand itself into RGBA8 pixels.  What differs is where the work happens: ComputePixels() and
// DecompressImage() run the GPU decoders behind fastc_gpu_decompress (include/fastc_gpu.h), which
// are bit-identical to the reference's CPU decoders.
#ifndef FASTC_B200_COMPRESSEDIMAGE_H_
#define FASTC_B200_COMPRESSEDIMAGE_H_

#include "FasTC/CompressionFormat.h"
#include "FasTC/Image.h"
#include "FasTC/TexCompTypes.h"

class CompressedImage : public FasTC::Image<FasTC::Pixel> {
  typedef FasTC::Image<FasTC::Pixel> Base;

 public:
  // ---- construction / copying -------------------------------------------------------------
  // Takes a private copy of GetCompressedSize(width, height, format) bytes starting at `data`.
  CompressedImage(const uint32 width, const uint32 height, const FasTC::ECompressionFormat format,
                  const uint8 *data);
  CompressedImage(const CompressedImage &other);
  CompressedImage &operator=(const CompressedImage &other);
  virtual ~CompressedImage();
  virtual Base *Clone() const { return new CompressedImage(*this); }

  // ---- what is stored ---------------------------------------------------------------------
  FasTC::ECompressionFormat GetFormat() const { return m_Format; }
  const uint8 *GetCompressedData() const { return m_CompressedData; }
  // ceil(w / 4) * ceil(h / 4) blocks of 8 (DXT1, ETC1) or 16 (DXT5, BPTC) bytes.
  static uint32 GetCompressedSize(uint32 width, uint32 height, FasTC::ECompressionFormat format);
  uint32 GetCompressedSize() const { return GetCompressedSize(GetWidth(), GetHeight(), m_Format); }
  // Bytes DecompressImage() writes: one RGBA8 word per pixel.
  uint32 GetUncompressedSize() const { return GetWidth() * GetHeight() * sizeof(uint32); }

  // ---- decoding ---------------------------------------------------------------------------
  // Fills the inherited pixel array from the payload (Image<>::GetPixels() calls this lazily).
  virtual void ComputePixels();
  // Expands the payload into outBuf.  Returns false when outBufSz is below
  // GetUncompressedSize(), when the format has no GPU decoder, or when the GPU call fails.
  bool DecompressImage(uint8 *outBuf, uint32 outBufSz) const;

 private:
  FasTC::ECompressionFormat m_Format;
  uint8 *m_CompressedData;
};

#endif  // FASTC_B200_COMPRESSEDIMAGE_H_
