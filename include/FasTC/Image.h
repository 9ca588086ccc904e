// FasTC::Image<PixelType>: the slice of reference Base/include/FasTC/Image.h that sits on
// either side of the compression path -- a W x H pixel container with lazily computed
// pixels (CompressedImage decodes on demand) and the reference's PSNR
// (Base/src/Image.cpp:205-255).  The analysis utilities (SSIM, entropy, DCT, filters) are
// out of scope.
#ifndef FASTC_B200_IMAGE_H_
#define FASTC_B200_IMAGE_H_

#include "FasTC/ImageFwd.h"
#include "FasTC/Pixel.h"
#include "FasTC/TexCompTypes.h"

namespace FasTC {

template <typename PixelType>
class Image {
 public:
  Image() : m_Width(0), m_Height(0), m_Pixels(0) {}
  Image(uint32 width, uint32 height);
  Image(uint32 width, uint32 height, const PixelType *pixels);
  Image(uint32 width, uint32 height, const uint32 *rgba);
  Image(const Image<PixelType> &other);
  Image<PixelType> &operator=(const Image<PixelType> &other);
  virtual ~Image();

  virtual Image<PixelType> *Clone() const { return new Image<PixelType>(*this); }

  PixelType &operator()(uint32 i, uint32 j) { return m_Pixels[j * m_Width + i]; }
  const PixelType &operator()(uint32 i, uint32 j) const { return m_Pixels[j * m_Width + i]; }

  uint32 GetWidth() const { return m_Width; }
  uint32 GetHeight() const { return m_Height; }
  uint32 GetNumPixels() const { return m_Width * m_Height; }
  const PixelType *GetPixels() const { return m_Pixels; }

  // Materialise the RGBA pixels (no-op here; CompressedImage decodes).
  virtual void ComputePixels() {}

  // The reference's PSNR against `other` (-1.0 on size mismatch); evaluated on the GPU.
  double ComputePSNR(Image<PixelType> *other);

 protected:
  void SetImageData(uint32 width, uint32 height, PixelType *data);  // takes ownership

 private:
  uint32 m_Width, m_Height;
  PixelType *m_Pixels;
};

}  // namespace FasTC
#endif
