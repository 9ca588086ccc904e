// ImageFile: the file formats either side of the compression path.
// API mirror of the used part of reference IO/include/FasTC/ImageFile.h: Load() / Write() /
// GetImage() / DetectFileFormat().  Readers: PNG (8-bit, non-interlaced), TGA (uncompressed + RLE true colour, 24/32 bit),
// KTX (RGBA8 or a BPTC / DXT1 / DXT5 / ETC1 payload).  Writers: TGA, KTX (compressed
// payload at byte 96 like reference IO/src/ImageWriterKTX.cpp:69-160, plus ETC1 which the
// reference cannot write) and PNG (zlib).
#ifndef FASTC_B200_IMAGEFILE_H_
#define FASTC_B200_IMAGEFILE_H_

#include "FasTC/CompressedImage.h"
#include "FasTC/Image.h"

enum EImageFileFormat { eFileFormat_PNG, eFileFormat_PVR, eFileFormat_TGA, eFileFormat_KTX, eFileFormat_ASTC, kNumImageFileFormats };

class ImageFile {
 public:
  explicit ImageFile(const char *filename);
  ImageFile(const char *filename, EImageFileFormat format);
  ImageFile(const char *filename, EImageFileFormat format, const FasTC::Image<> &image);
  ~ImageFile();

  static EImageFileFormat DetectFileFormat(const CHAR *filename);

  bool Load();
  bool Write();
  FasTC::Image<> *GetImage() const { return m_Image; }
  uint32 GetWidth() const { return m_Image ? m_Image->GetWidth() : 0; }
  uint32 GetHeight() const { return m_Image ? m_Image->GetHeight() : 0; }

 private:
  ImageFile(const ImageFile &);
  ImageFile &operator=(const ImageFile &);
  char m_Filename[512];
  EImageFileFormat m_FileFormat;
  FasTC::Image<> *m_Image;
};
#endif
