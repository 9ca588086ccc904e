#ifndef FASTC_B200_IMAGEFWD_H_
#define FASTC_B200_IMAGEFWD_H_
namespace FasTC {
class Pixel;
template <typename PixelType = Pixel> class Image;
}  // namespace FasTC
#endif
