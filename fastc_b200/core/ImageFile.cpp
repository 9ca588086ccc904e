// File formats either side of the compression path.  Behaviour follows
// reference IO/src: TGA rows are flipped on load and the image descriptor is ignored
// (ImageLoaderTGA.cpp:36-48 + ImageLoader.cpp:122-149), the KTX writer emits the compressed
// payload at byte 96 with the "KTXorientation" key (ImageWriterKTX.cpp:69-160).  Beyond the
// reference: ETC1 payloads can be written to KTX, compressed KTX files can be loaded back,
// and PNG input / output needs only zlib (no libpng).
#include "FasTC/ImageFile.h"

#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

bool ReadAll(const char *path, std::vector<uint8> &out) {
  FILE *f = fopen(path, "rb");
  if (!f) {
    fprintf(stderr, "Error opening file for reading: %s\n", path);
    return false;
  }
  fseek(f, 0, SEEK_END);
  const long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  out.resize(sz > 0 ? (size_t)sz : 0);
  const size_t got = out.empty() ? 0 : fread(out.data(), 1, out.size(), f);
  fclose(f);
  return got == out.size();
}

bool WriteAll(const char *path, const std::vector<uint8> &data) {
  FILE *f = fopen(path, "wb");
  if (!f) {
    fprintf(stderr, "Error opening file for writing: %s\n", path);
    return false;
  }
  const size_t put = fwrite(data.data(), 1, data.size(), f);
  fclose(f);
  return put == data.size();
}

void Put32(std::vector<uint8> &v, uint32 x) {
  for (int i = 0; i < 4; i++) v.push_back((uint8)(x >> (8 * i)));
}
uint32 Get32(const uint8 *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32)p[3] << 24); }

// OpenGL enums (reference IO/src/GLDefines.h)
enum {
  kGL_BYTE = 0x1400, kGL_UNSIGNED_BYTE = 0x1401, kGL_RGB = 0x1907, kGL_RGBA = 0x1908, kGL_RGBA8 = 0x8058,
  kGL_DXT1 = 0x83F0, kGL_DXT5 = 0x83F3, kGL_BPTC = 0x8E8C, kGL_ETC1 = 0x8D64,
  kGL_PVRTC4 = 0x8C02  // GL_COMPRESSED_RGBA_PVRTC_4BPPV1_IMG (IO/src/GLDefines.h:70-72)
};
const uint8 kKtxId[12] = {0xAB, 0x4B, 0x54, 0x58, 0x20, 0x31, 0x31, 0xBB, 0x0D, 0x0A, 0x1A, 0x0A};

// ---- TGA ------------------------------------------------------------------------------
FasTC::Image<> *LoadTGA(const std::vector<uint8> &d) {
  if (d.size() < 18) return NULL;
  const int idLen = d[0], cmap = d[1], type = d[2], bpp = d[16];
  const uint32 w = d[12] | (d[13] << 8), h = d[14] | (d[15] << 8);
  if (cmap != 0 || (type != 2 && type != 10) || (bpp != 24 && bpp != 32) || w == 0 || h == 0) {
    fprintf(stderr, "Unsupported TGA variant (type %d, %d bpp, colour map %d)\n", type, bpp, cmap);
    return NULL;
  }
  const int bytes = bpp / 8;
  size_t pos = 18 + (size_t)idLen;
  std::vector<uint32> px((size_t)w * h);
  auto fetch = [&](size_t at) -> uint32 {  // BGR(A) -> R | G << 8 | B << 16 | A << 24
    return (uint32)d[at + 2] | ((uint32)d[at + 1] << 8) | ((uint32)d[at] << 16) |
           ((bytes == 4 ? (uint32)d[at + 3] : 255u) << 24);
  };
  size_t n = 0;
  if (type == 2) {
    if (d.size() < pos + px.size() * bytes) return NULL;
    for (; n < px.size(); n++, pos += bytes) px[n] = fetch(pos);
  } else {
    while (n < px.size()) {
      if (pos >= d.size()) return NULL;
      const int hdr = d[pos++];
      const int count = (hdr & 127) + 1;
      if (hdr & 128) {
        if (pos + bytes > d.size()) return NULL;
        const uint32 v = fetch(pos);
        pos += bytes;
        for (int k = 0; k < count && n < px.size(); k++) px[n++] = v;
      } else {
        if (pos + (size_t)count * bytes > d.size()) return NULL;
        for (int k = 0; k < count && n < px.size(); k++, pos += bytes) px[n++] = fetch(pos);
      }
    }
  }
  std::vector<uint32> flipped(px.size());
  for (uint32 j = 0; j < h; j++) memcpy(&flipped[(size_t)j * w], &px[(size_t)(h - 1 - j) * w], (size_t)w * 4);
  return new FasTC::Image<>(w, h, flipped.data());
}

bool WriteTGA(const char *path, FasTC::Image<> &img) {
  img.ComputePixels();
  const uint32 w = img.GetWidth(), h = img.GetHeight();
  if (!img.GetPixels() || w > 0xFFFF || h > 0xFFFF) return false;
  std::vector<uint8> out(18, 0);
  out[2] = 2;
  out[12] = w & 0xFF; out[13] = w >> 8; out[14] = h & 0xFF; out[15] = h >> 8;
  out[16] = 32;
  out[17] = 8;  // 8 alpha bits, bottom-left origin (rows stored bottom-up, like the loader expects)
  out.reserve(18 + (size_t)w * h * 4);
  for (uint32 j = 0; j < h; j++)
    for (uint32 i = 0; i < w; i++) {
      const FasTC::Pixel &p = img(i, h - 1 - j);
      out.push_back(p.B()); out.push_back(p.G()); out.push_back(p.R()); out.push_back(p.A());
    }
  return WriteAll(path, out);
}

// ---- KTX ------------------------------------------------------------------------------
bool WriteKTX(const char *path, FasTC::Image<> &img) {
  std::vector<uint8> out(kKtxId, kKtxId + 12);
  Put32(out, 0x04030201);
  const char *key = "KTXorientation", *val = "S=r,T=d";
  const uint32 kvSz = (uint32)strlen(key) + 1 + (uint32)strlen(val) + 1;
  const uint32 tkvSz = (kvSz + 4 + 3) & ~3u;
  CompressedImage *ci = dynamic_cast<CompressedImage *>(&img);
  uint32 imageSize;
  const uint8 *payload;
  if (ci) {
    uint32 internal, base;
    switch (ci->GetFormat()) {
      case FasTC::eCompressionFormat_BPTC: internal = kGL_BPTC; base = kGL_RGBA; break;
      case FasTC::eCompressionFormat_DXT1: internal = kGL_DXT1; base = kGL_RGB; break;
      case FasTC::eCompressionFormat_DXT5: internal = kGL_DXT5; base = kGL_RGBA; break;
      case FasTC::eCompressionFormat_PVRTC4: internal = kGL_PVRTC4; base = kGL_RGBA; break;
      case FasTC::eCompressionFormat_ETC1: internal = kGL_ETC1; base = kGL_RGB; break;  // not writable by the reference
      default:
        fprintf(stderr, "Unsupported KTX compressed format: %d\n", ci->GetFormat());
        return false;
    }
    Put32(out, 0); Put32(out, 1); Put32(out, 0); Put32(out, internal); Put32(out, base);
    imageSize = ci->GetCompressedSize();
    payload = ci->GetCompressedData();
  } else {
    img.ComputePixels();
    Put32(out, kGL_BYTE); Put32(out, 1); Put32(out, kGL_RGBA); Put32(out, kGL_RGBA8); Put32(out, kGL_RGBA);
    imageSize = img.GetWidth() * img.GetHeight() * 4;
    payload = NULL;
  }
  Put32(out, img.GetWidth()); Put32(out, img.GetHeight());
  Put32(out, 0); Put32(out, 0); Put32(out, 1); Put32(out, 1);
  Put32(out, tkvSz); Put32(out, kvSz);
  out.insert(out.end(), key, key + strlen(key) + 1);
  out.insert(out.end(), val, val + strlen(val) + 1);
  out.insert(out.end(), key, key + (tkvSz - kvSz - 4));  // padding bytes, as the reference writes them
  Put32(out, imageSize);
  if (payload) {
    out.insert(out.end(), payload, payload + imageSize);
  } else {
    for (uint32 i = 0; i < img.GetNumPixels(); i++) Put32(out, img.GetPixels()[i].Pack());
  }
  return WriteAll(path, out);
}

FasTC::Image<> *LoadKTX(const std::vector<uint8> &d) {
  if (d.size() < 68 || memcmp(d.data(), kKtxId, 12) != 0 || Get32(&d[12]) != 0x04030201) {
    fprintf(stderr, "Not a little-endian KTX 1.1 file\n");
    return NULL;
  }
  const uint32 glType = Get32(&d[16]), internal = Get32(&d[28]);
  const uint32 w = Get32(&d[36]), h = Get32(&d[40]), kvBytes = Get32(&d[60]);
  const size_t at = 64 + (size_t)kvBytes;
  if (d.size() < at + 4) return NULL;
  const uint32 imageSize = Get32(&d[at]);
  if (d.size() < at + 4 + imageSize) return NULL;
  const uint8 *payload = &d[at + 4];
  FasTC::ECompressionFormat fmt;
  switch (internal) {
    case kGL_BPTC: fmt = FasTC::eCompressionFormat_BPTC; break;
    case kGL_DXT1: fmt = FasTC::eCompressionFormat_DXT1; break;
    case kGL_DXT5: fmt = FasTC::eCompressionFormat_DXT5; break;
    case kGL_ETC1: fmt = FasTC::eCompressionFormat_ETC1; break;
    case kGL_PVRTC4: fmt = FasTC::eCompressionFormat_PVRTC4; break;  // (ImageLoaderKTX.cpp:247)
    default:
      if ((glType == kGL_BYTE || glType == kGL_UNSIGNED_BYTE) && imageSize >= (uint64)w * h * 4)
        return new FasTC::Image<>(w, h, reinterpret_cast<const uint32 *>(payload));
      fprintf(stderr, "Unsupported KTX internal format 0x%x\n", internal);
      return NULL;
  }
  if (imageSize < CompressedImage::GetCompressedSize(w, h, fmt)) return NULL;
  return new CompressedImage(w, h, fmt, payload);
}

// ---- PNG (8-bit RGBA, zlib) -----------------------------------------------------------
void PngChunk(std::vector<uint8> &out, const char *tag, const std::vector<uint8> &body) {
  const uint32 n = (uint32)body.size();
  for (int i = 3; i >= 0; i--) out.push_back((uint8)(n >> (8 * i)));
  const size_t start = out.size();
  out.insert(out.end(), tag, tag + 4);
  out.insert(out.end(), body.begin(), body.end());
  const uint32 crc = (uint32)crc32(0L, &out[start], (uInt)(out.size() - start));
  for (int i = 3; i >= 0; i--) out.push_back((uint8)(crc >> (8 * i)));
}

bool WritePNG(const char *path, FasTC::Image<> &img) {
  img.ComputePixels();
  const uint32 w = img.GetWidth(), h = img.GetHeight();
  if (!img.GetPixels()) return false;
  std::vector<uint8> raw;
  raw.reserve((size_t)h * (1 + (size_t)w * 4));
  for (uint32 j = 0; j < h; j++) {
    raw.push_back(0);  // filter: none
    for (uint32 i = 0; i < w; i++) {
      const FasTC::Pixel &p = img(i, j);
      raw.push_back(p.R()); raw.push_back(p.G()); raw.push_back(p.B()); raw.push_back(p.A());
    }
  }
  uLongf zlen = compressBound((uLong)raw.size());
  std::vector<uint8> z(zlen);
  if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return false;
  z.resize(zlen);
  static const uint8 sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  std::vector<uint8> out(sig, sig + 8), ihdr;
  for (int i = 3; i >= 0; i--) ihdr.push_back((uint8)(w >> (8 * i)));
  for (int i = 3; i >= 0; i--) ihdr.push_back((uint8)(h >> (8 * i)));
  ihdr.push_back(8); ihdr.push_back(6); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
  PngChunk(out, "IHDR", ihdr);
  PngChunk(out, "IDAT", z);
  PngChunk(out, "IEND", std::vector<uint8>());
  return WriteAll(path, out);
}

// PNG loader with the coverage of the reference's libpng loader (IO/src/ImageLoaderPNG.cpp:58-260):
// bit depth 8 only ("Only 8-bit images currently supported."), colour types grey, RGB, palette
// (opaque, like the reference: tRNS is ignored), grey + alpha, RGBA; rows in file order.  Adam7
// interlacing is rejected (the reference reads rows without interlace handling).
FasTC::Image<> *LoadPNG(const std::vector<uint8> &d) {
  static const uint8 sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  if (d.size() < 8 || memcmp(d.data(), sig, 8) != 0) {
    fprintf(stderr, "Incorrect PNG signature\n");
    return NULL;
  }
  auto be32 = [&](size_t at) { return ((uint32)d[at] << 24) | ((uint32)d[at + 1] << 16) | ((uint32)d[at + 2] << 8) | d[at + 3]; };
  uint32 w = 0, h = 0;
  int depth = 0, ctype = -1, interlace = 0;
  std::vector<uint8> idat, plte;
  for (size_t pos = 8; pos + 12 <= d.size();) {
    const uint32 len = be32(pos);
    if (pos + 12 + (size_t)len > d.size()) break;
    const uint8 *tag = &d[pos + 4], *body = &d[pos + 8];
    if (!memcmp(tag, "IHDR", 4) && len >= 13) {
      w = be32(pos + 8); h = be32(pos + 12);
      depth = body[8]; ctype = body[9]; interlace = body[12];
    } else if (!memcmp(tag, "PLTE", 4)) {
      plte.assign(body, body + len);
    } else if (!memcmp(tag, "IDAT", 4)) {
      idat.insert(idat.end(), body, body + len);
    } else if (!memcmp(tag, "IEND", 4)) {
      break;
    }
    pos += 12 + (size_t)len;
  }
  if (w == 0 || h == 0 || ctype < 0) {
    fprintf(stderr, "Could not read PNG header\n");
    return NULL;
  }
  if (depth != 8) {
    fprintf(stderr, "Only 8-bit images currently supported.\n");
    return NULL;
  }
  int channels = 0;
  switch (ctype) {
    case 0: channels = 1; break;
    case 2: channels = 3; break;
    case 3: channels = 1; break;
    case 4: channels = 2; break;
    case 6: channels = 4; break;
    default: fprintf(stderr, "PNG color type unsupported\n"); return NULL;
  }
  if (interlace != 0) {
    fprintf(stderr, "Interlaced PNG images are not supported\n");
    return NULL;
  }
  if (ctype == 3 && plte.size() < 3) {
    fprintf(stderr, "Couldn't find PLTE chunk\n");
    return NULL;
  }
  const size_t stride = (size_t)w * channels;
  std::vector<uint8> raw((size_t)h * (stride + 1));
  uLongf rawLen = (uLongf)raw.size();
  if (idat.empty() || uncompress(raw.data(), &rawLen, idat.data(), (uLong)idat.size()) != Z_OK || rawLen != raw.size()) {
    fprintf(stderr, "Could not decode PNG image data\n");
    return NULL;
  }
  // undo the per-row filters in place (PNG specification, filter types 0-4)
  for (uint32 j = 0; j < h; j++) {
    uint8 *row = &raw[(size_t)j * (stride + 1) + 1];
    const uint8 *up = j ? row - (stride + 1) : NULL;
    const int filter = row[-1];
    for (size_t i = 0; i < stride; i++) {
      const int a = i >= (size_t)channels ? row[i - channels] : 0, b = up ? up[i] : 0,
                c = (up && i >= (size_t)channels) ? up[i - channels] : 0;
      int pred = 0;
      switch (filter) {
        case 0: pred = 0; break;
        case 1: pred = a; break;
        case 2: pred = b; break;
        case 3: pred = (a + b) >> 1; break;
        case 4: {
          const int pp = a + b - c, pa = abs(pp - a), pb = abs(pp - b), pc = abs(pp - c);
          pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
          break;
        }
        default: fprintf(stderr, "Could not decode PNG image data\n"); return NULL;
      }
      row[i] = (uint8)(row[i] + pred);
    }
  }
  std::vector<uint32> px((size_t)w * h);
  const size_t npal = plte.size() / 3;
  for (uint32 j = 0; j < h; j++) {
    const uint8 *row = &raw[(size_t)j * (stride + 1) + 1];
    for (uint32 i = 0; i < w; i++) {
      uint32 r, g, b, a = 255;
      switch (ctype) {
        case 0: r = g = b = row[i]; break;
        case 2: r = row[3 * i]; g = row[3 * i + 1]; b = row[3 * i + 2]; break;
        case 3: {
          const size_t e = std::min<size_t>(row[i], npal - 1);
          r = plte[3 * e]; g = plte[3 * e + 1]; b = plte[3 * e + 2];
          break;
        }
        case 4: r = g = b = row[2 * i]; a = row[2 * i + 1]; break;
        default: r = row[4 * i]; g = row[4 * i + 1]; b = row[4 * i + 2]; a = row[4 * i + 3]; break;
      }
      px[(size_t)j * w + i] = r | (g << 8) | (b << 16) | (a << 24);
    }
  }
  return new FasTC::Image<>(w, h, px.data());
}

}  // namespace

ImageFile::ImageFile(const char *filename) : m_FileFormat(DetectFileFormat(filename)), m_Image(NULL) {
  snprintf(m_Filename, sizeof(m_Filename), "%s", filename);
}
ImageFile::ImageFile(const char *filename, EImageFileFormat format) : m_FileFormat(format), m_Image(NULL) {
  snprintf(m_Filename, sizeof(m_Filename), "%s", filename);
}
ImageFile::ImageFile(const char *filename, EImageFileFormat format, const FasTC::Image<> &image)
    : m_FileFormat(format), m_Image(image.Clone()) {
  snprintf(m_Filename, sizeof(m_Filename), "%s", filename);
}
ImageFile::~ImageFile() { delete m_Image; }

EImageFileFormat ImageFile::DetectFileFormat(const CHAR *filename) {
  const char *dot = strrchr(filename, '.');
  if (!dot) {
    fprintf(stderr, "Unknown file format: %s (no extension)\n", filename);
    return kNumImageFileFormats;
  }
  std::string ext(dot + 1);
  for (size_t i = 0; i < ext.size(); i++) ext[i] = (char)tolower(ext[i]);
  if (ext == "png") return eFileFormat_PNG;
  if (ext == "pvr") return eFileFormat_PVR;
  if (ext == "tga") return eFileFormat_TGA;
  if (ext == "ktx") return eFileFormat_KTX;
  if (ext == "astc") return eFileFormat_ASTC;
  return kNumImageFileFormats;
}

bool ImageFile::Load() {
  delete m_Image;
  m_Image = NULL;
  std::vector<uint8> d;
  if (!ReadAll(m_Filename, d)) return false;
  switch (m_FileFormat) {
    case eFileFormat_TGA: m_Image = LoadTGA(d); break;
    case eFileFormat_KTX: m_Image = LoadKTX(d); break;
    case eFileFormat_PNG: m_Image = LoadPNG(d); break;
    default:
      fprintf(stderr, "Unable to load image: unsupported input file format (PNG, TGA and KTX are).\n");
      return false;
  }
  if (!m_Image) fprintf(stderr, "Unable to load image!\n");
  return m_Image != NULL;
}

bool ImageFile::Write() {
  if (!m_Image) return false;
  switch (m_FileFormat) {
    case eFileFormat_TGA: return WriteTGA(m_Filename, *m_Image);
    case eFileFormat_KTX: return WriteKTX(m_Filename, *m_Image);
    case eFileFormat_PNG: return WritePNG(m_Filename, *m_Image);
    default:
      fprintf(stderr, "Unable to write image: unknown file format.\n");
      return false;
  }
}
