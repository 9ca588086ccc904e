// FasTC::Image<Pixel> and CompressedImage (reference Base/src/Image.cpp:205-255,
// Core/src/CompressedImage.cpp:29-149).  Decoding and the PSNR reduction run on the GPU.
#include "FasTC/Image.h"

#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "FasTC/CompressedImage.h"
#include "fastc_gpu.h"

namespace FasTC {

template <typename P>
Image<P>::Image(uint32 width, uint32 height) : m_Width(width), m_Height(height), m_Pixels(new P[(size_t)width * height]()) {}

template <typename P>
Image<P>::Image(uint32 width, uint32 height, const P *pixels)
    : m_Width(width), m_Height(height), m_Pixels(new P[(size_t)width * height]) {
  if (pixels) memcpy(m_Pixels, pixels, sizeof(P) * (size_t)width * height);
}

template <typename P>
Image<P>::Image(uint32 width, uint32 height, const uint32 *rgba)
    : m_Width(width), m_Height(height), m_Pixels(new P[(size_t)width * height]) {
  for (size_t i = 0; i < (size_t)width * height; i++) m_Pixels[i].Unpack(rgba[i]);
}

template <typename P>
Image<P>::Image(const Image<P> &o) : m_Width(o.m_Width), m_Height(o.m_Height), m_Pixels(0) {
  if (o.m_Pixels) {
    m_Pixels = new P[(size_t)m_Width * m_Height];
    memcpy(m_Pixels, o.m_Pixels, sizeof(P) * (size_t)m_Width * m_Height);
  }
}

template <typename P>
Image<P> &Image<P>::operator=(const Image<P> &o) {
  if (this == &o) return *this;
  P *np = 0;
  if (o.m_Pixels) {
    np = new P[(size_t)o.m_Width * o.m_Height];
    memcpy(np, o.m_Pixels, sizeof(P) * (size_t)o.m_Width * o.m_Height);
  }
  delete[] m_Pixels;
  m_Pixels = np;
  m_Width = o.m_Width;
  m_Height = o.m_Height;
  return *this;
}

template <typename P>
Image<P>::~Image() { delete[] m_Pixels; }

template <typename P>
void Image<P>::SetImageData(uint32 width, uint32 height, P *data) {
  delete[] m_Pixels;
  m_Pixels = data;
  m_Width = width;
  m_Height = height;
}

template <typename P>
double Image<P>::ComputePSNR(Image<P> *other) {
  if (!other) return -1.0;
  if (GetWidth() != other->GetWidth() || GetHeight() != other->GetHeight()) return -1.0;
  ComputePixels();
  other->ComputePixels();
  const size_t n = GetNumPixels();
  if (!GetPixels() || !other->GetPixels() || n == 0) return -1.0;
  std::vector<uint32> a(n), b(n);
  for (size_t i = 0; i < n; i++) {
    a[i] = GetPixels()[i].Pack();
    b[i] = other->GetPixels()[i].Pack();
  }
  double psnr = -1.0;
  if (fastc_gpu_psnr(reinterpret_cast<const uint8 *>(a.data()), reinterpret_cast<const uint8 *>(b.data()), GetWidth(),
                     GetHeight(), &psnr) != 0) {
    fprintf(stderr, "TexComp -- %s\n", fastc_gpu_last_error());
    return -1.0;
  }
  return psnr;
}

template class Image<Pixel>;

}  // namespace FasTC

// ---------------------------------------------------------------------------------------
namespace {
int GpuFormatOf(FasTC::ECompressionFormat f) {
  switch (f) {
    case FasTC::eCompressionFormat_DXT1: return FASTC_GPU_DXT1;
    case FasTC::eCompressionFormat_DXT5: return FASTC_GPU_DXT5;
    case FasTC::eCompressionFormat_ETC1: return FASTC_GPU_ETC1;
    case FasTC::eCompressionFormat_BPTC: return FASTC_GPU_BPTC;
    case FasTC::eCompressionFormat_PVRTC4: return FASTC_GPU_PVRTC4;
    default: return -1;
  }
}
}  // namespace

CompressedImage::CompressedImage(const uint32 width, const uint32 height, const FasTC::ECompressionFormat format,
                                 const uint8 *data)
    : FasTC::Image<FasTC::Pixel>(width, height, static_cast<const FasTC::Pixel *>(0)), m_Format(format),
      m_CompressedData(0) {
  const uint32 sz = GetCompressedSize(width, height, format);
  m_CompressedData = new uint8[sz ? sz : 1];
  if (data) memcpy(m_CompressedData, data, sz);
}

CompressedImage::CompressedImage(const CompressedImage &o)
    : FasTC::Image<FasTC::Pixel>(o), m_Format(o.m_Format), m_CompressedData(0) {
  const uint32 sz = o.GetCompressedSize();
  m_CompressedData = new uint8[sz ? sz : 1];
  memcpy(m_CompressedData, o.m_CompressedData, sz);
}

CompressedImage &CompressedImage::operator=(const CompressedImage &o) {
  if (this == &o) return *this;
  FasTC::Image<FasTC::Pixel>::operator=(o);
  const uint32 sz = o.GetCompressedSize();
  uint8 *nd = new uint8[sz ? sz : 1];
  memcpy(nd, o.m_CompressedData, sz);
  delete[] m_CompressedData;
  m_CompressedData = nd;
  m_Format = o.m_Format;
  return *this;
}

CompressedImage::~CompressedImage() { delete[] m_CompressedData; }

uint32 CompressedImage::GetCompressedSize(uint32 width, uint32 height, FasTC::ECompressionFormat format) {
  uint32 dim[2];
  FasTC::GetBlockDimensions(format, dim);
  return ((width + dim[0] - 1) / dim[0]) * ((height + dim[1] - 1) / dim[1]) * FasTC::GetBlockSize(format);
}

bool CompressedImage::DecompressImage(uint8 *outBuf, uint32 outBufSz) const {
  if (outBufSz < GetUncompressedSize()) return false;
  const int fmt = GpuFormatOf(m_Format);
  if (fmt < 0) {
    fprintf(stderr, "Have not implemented decompression method.\n");
    return false;
  }
  if (fastc_gpu_decompress(fmt, m_CompressedData, GetWidth(), GetHeight(), outBuf, NULL) != 0) {
    fprintf(stderr, "TexComp -- %s\n", fastc_gpu_last_error());
    return false;
  }
  return true;
}

void CompressedImage::ComputePixels() {
  const size_t n = (size_t)GetWidth() * GetHeight();
  std::vector<uint32> buf(n);
  if (!DecompressImage(reinterpret_cast<uint8 *>(buf.data()), (uint32)(n * 4))) return;
  FasTC::Pixel *px = new FasTC::Pixel[n];
  for (size_t i = 0; i < n; i++) px[i].Unpack(buf[i]);
  SetImageData(GetWidth(), GetHeight(), px);
}
