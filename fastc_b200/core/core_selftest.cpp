// Test driver for the C++ host layer (run by tests/test_core_cpp.py).
//   core_selftest api                         no GPU needed: job arithmetic, job list, sizes, error paths
//   core_selftest gpu <rgba.raw> <w> <h> <fmt> <quality> <seed> <out-prefix>
//       writes <prefix>.whole   CompressImageData on the whole image
//              <prefix>.split   the same image through the per-format CompressionFunc, as three
//                               CompressionJobs with ThreadGroup-style block ranges
//              <prefix>.list0/1 CompressImageList over two jobs (the image and its top half)
//              <prefix>.dec     CompressedImage::DecompressImage of .whole
//       and prints "PSNR: %.6f" of .dec against the input.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "FasTC/BlockCompressors.h"
#include "FasTC/CompressedImage.h"
#include "FasTC/ImageFile.h"
#include "FasTC/TexComp.h"

static int g_fail = 0;
#define CHECK(cond)                                                      \
  do {                                                                   \
    if (!(cond)) { printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond); g_fail++; } \
  } while (0)

static FasTC::ECompressionFormat ParseFormat(const char *s) {
  if (!strcmp(s, "DXT1")) return FasTC::eCompressionFormat_DXT1;
  if (!strcmp(s, "DXT5")) return FasTC::eCompressionFormat_DXT5;
  if (!strcmp(s, "ETC1")) return FasTC::eCompressionFormat_ETC1;
  if (!strcmp(s, "PVRTC4")) return FasTC::eCompressionFormat_PVRTC4;
  return FasTC::eCompressionFormat_BPTC;
}

static bool Dump(const std::string &path, const void *p, size_t n) {
  FILE *f = fopen(path.c_str(), "wb");
  if (!f) return false;
  const bool ok = fwrite(p, 1, n, f) == n;
  fclose(f);
  return ok;
}

static int TestApi() {
  using namespace FasTC;
  // block geometry
  uint32 d[2];
  GetBlockDimensions(eCompressionFormat_BPTC, d); CHECK(d[0] == 4 && d[1] == 4);
  GetBlockDimensions(eCompressionFormat_PVRTC2, d); CHECK(d[0] == 8 && d[1] == 4);
  GetBlockDimensions(eCompressionFormat_ASTC12x10, d); CHECK(d[0] == 12 && d[1] == 10);
  CHECK(GetBlockSize(eCompressionFormat_DXT1) == 8 && GetBlockSize(eCompressionFormat_ETC1) == 8);
  CHECK(GetBlockSize(eCompressionFormat_DXT5) == 16 && GetBlockSize(eCompressionFormat_BPTC) == 16);
  CHECK(GetBlockSize(eCompressionFormat_ASTC8x8) == 16 && GetBlockSize(eCompressionFormat_PVRTC4) == 8);
  CHECK(CompressedImage::GetCompressedSize(256, 128, eCompressionFormat_DXT1) == 64 * 32 * 8);
  CHECK(CompressedImage::GetCompressedSize(10, 6, eCompressionFormat_BPTC) == 3 * 2 * 16);

  // job arithmetic (reference CompressionJob.h:114-139 and the encoders' block loops)
  uint8 in[1], out[1];
  CompressionJob whole(eCompressionFormat_DXT5, in, out, 64, 32);
  CHECK(whole.FirstBlock() == 0 && whole.NumBlocks() == 16 * 8);
  uint32 c[2];
  whole.BlockIdxToCoords(37, c); CHECK(c[0] == 20 && c[1] == 8);
  CHECK(whole.CoordsToBlockIdx(20, 8) == 37 && whole.CoordsToBlockIdx(23, 11) == 37);
  // ThreadGroup-style split (reference ThreadGroup.cpp:146-188): [startBlock, endBlock) via coords
  uint32 s[2], e[2];
  whole.BlockIdxToCoords(37, s); whole.BlockIdxToCoords(90, e);
  CompressionJob mid(eCompressionFormat_DXT5, in, out, 64, 32, s[0], s[1], e[0], e[1]);
  CHECK(mid.FirstBlock() == 37 && mid.NumBlocks() == 53);
  whole.BlockIdxToCoords(128, e);  // one past the last block -> (0, Height)
  CompressionJob tail(eCompressionFormat_DXT5, in, out, 64, 32, s[0], s[1], e[0], e[1]);
  CHECK(tail.FirstBlock() == 37 && tail.NumBlocks() == 128 - 37);
  CompressionJob from(eCompressionFormat_DXT5, in, out, 64, 32, 8, 4);
  CHECK(from.FirstBlock() == 18 && from.NumBlocks() == 128 - 18);

  // job list
  CompressionJobList list(2);
  CHECK(list.GetTotalNumJobs() == 2 && list.GetNumJobs() == 0 && list.GetJob(0) == NULL);
  CHECK(list.AddJob(whole) && list.AddJob(mid) && !list.AddJob(tail));
  CHECK(list.GetNumJobs() == 2 && list.GetJob(1)->FirstBlock() == 37 && list.GetJob(2) == NULL);
  CompressionJobList copy(list);
  CHECK(copy.GetNumJobs() == 2 && copy.GetJob(1)->NumBlocks() == 53 && *copy.GetFinishedFlag(0) == 0);

  // settings defaults (every field initialised)
  SCompressionSettings st;
  CHECK(st.format == eCompressionFormat_BPTC && !st.bUseSIMD && st.iNumThreads == 1 && st.iQuality == 50);
  CHECK(st.iNumCompressions == 1 && st.iJobSize == 0 && !st.bUseAtomics && !st.bUsePVRTexLib && !st.bUseNVTT);
  CHECK(st.logStream == NULL && st.iNumGPUs == 1 && st.uSeed == 0);

  // error paths that need no GPU (reference TexComp.cpp:436-496); messages go to stderr
  std::vector<uint8> img(8 * 8 * 4, 0), cmp(64, 0);
  st.format = eCompressionFormat_DXT1;
  st.bUseSIMD = true;
  CHECK(!CompressImageData(img.data(), 8, 8, cmp.data(), 64, st));
  st.bUseSIMD = false;
  CHECK(!CompressImageData(img.data(), 6, 8, cmp.data(), 64, st));
  CHECK(!CompressImageData(img.data(), 8, 8, cmp.data(), 8, st));
  CHECK(!CompressImageData(img.data(), 0, 8, cmp.data(), 64, st));
  st.format = eCompressionFormat_PVRTC2;  // no encoder in FasTC
  CHECK(!CompressImageData(img.data(), 8, 8, cmp.data(), 64, st));
  {  // PVRTC4: square power-of-two images only (reference TexComp.cpp:477-482)
    st.format = eCompressionFormat_PVRTC4;
    std::vector<uint8> wide(16 * 8 * 4, 0), wcmp(64, 0);
    CHECK(!CompressImageData(wide.data(), 16, 8, wcmp.data(), 64, st));
  }
  st.format = eCompressionFormat_ASTC4x4;
  CHECK(!CompressImageData(img.data(), 8, 8, cmp.data(), 64, st));
  CHECK(CompressImage<FasTC::Pixel>(NULL, st) == NULL);

  // file format detection + a TGA / KTX / PNG write-read round trip of an uncompressed image
  CHECK(ImageFile::DetectFileFormat("a/b.c/x.TGA") == eFileFormat_TGA);
  CHECK(ImageFile::DetectFileFormat("x.ktx") == eFileFormat_KTX && ImageFile::DetectFileFormat("x.png") == eFileFormat_PNG);
  std::vector<uint32> px(12 * 8);
  for (size_t i = 0; i < px.size(); i++) px[i] = (uint32)(i * 2654435761u);
  FasTC::Image<> im(12, 8, px.data());
  const char *tmp = getenv("TMPDIR") ? getenv("TMPDIR") : "/tmp";
  for (const char *ext : {"tga", "ktx"}) {
    const std::string path = std::string(tmp) + "/fastc_core_selftest." + ext;
    ImageFile w(path.c_str(), ImageFile::DetectFileFormat(path.c_str()), im);
    CHECK(w.Write());
    ImageFile r(path.c_str());
    CHECK(r.Load() && r.GetWidth() == 12 && r.GetHeight() == 8);
    if (r.GetImage())
      for (uint32 j = 0; j < 8; j++)
        for (uint32 i = 0; i < 12; i++) CHECK((*r.GetImage())(i, j).Pack() == px[j * 12 + i]);
    remove(path.c_str());
  }
  {
    const std::string path = std::string(tmp) + "/fastc_core_selftest.png";
    ImageFile w(path.c_str(), eFileFormat_PNG, im);
    CHECK(w.Write());
    remove(path.c_str());
  }
  printf(g_fail ? "api: %d failures\n" : "api: ok\n", g_fail);
  return g_fail ? 1 : 0;
}

static int TestGpu(int argc, char **argv) {
  if (argc < 9) return 2;
  const uint32 w = (uint32)atoi(argv[3]), h = (uint32)atoi(argv[4]);
  const FasTC::ECompressionFormat fmt = ParseFormat(argv[5]);
  const int quality = atoi(argv[6]);
  const unsigned long long seed = strtoull(argv[7], NULL, 0);
  const std::string prefix = argv[8];
  std::vector<uint8> img((size_t)w * h * 4);
  FILE *f = fopen(argv[2], "rb");
  if (!f || fread(img.data(), 1, img.size(), f) != img.size()) return 2;
  fclose(f);

  SCompressionSettings st;
  st.format = fmt;
  st.iQuality = quality;
  st.uSeed = seed;
  st.iNumThreads = 4;  // accepted, ignored for placement
  const uint32 sz = CompressedImage::GetCompressedSize(w, h, fmt);
  std::vector<uint8> whole(sz, 0xEE), split(sz, 0xEE);
  CHECK(CompressImageData(img.data(), w, h, whole.data(), sz, st));
  CHECK(Dump(prefix + ".whole", whole.data(), sz));

  // the reference's ThreadGroup split (ceil(nBlocks / nThreads) contiguous blocks per thread)
  // driving the per-format CompressionFunc
  FasTC::CompressionJob cj(fmt, img.data(), split.data(), w, h);
  const uint32 total = cj.NumBlocks(), per = (total + 2) / 3;
  for (uint32 t = 0; t < 3; t++) {
    uint32 s[2], e[2];
    const uint32 a = std::min(total, t * per), b = std::min(total, (t + 1) * per);
    cj.BlockIdxToCoords(a, s);
    cj.BlockIdxToCoords(b, e);
    FasTC::CompressionJob part(fmt, img.data(), split.data(), w, h, s[0], s[1], e[0], e[1]);
    CHECK(part.FirstBlock() == a && part.NumBlocks() == b - a);
    switch (fmt) {
      case FasTC::eCompressionFormat_DXT1: DXTC::CompressImageDXT1(part); break;
      case FasTC::eCompressionFormat_DXT5: DXTC::CompressImageDXT5(part); break;
      case FasTC::eCompressionFormat_ETC1: ETCC::Compress_RG(part); break;
      default: {
        BPTCC::CompressionSettings bs;
        bs.m_NumSimulatedAnnealingSteps = quality;
        BPTCC::Compress(part, bs);
      }
    }
  }
  CHECK(Dump(prefix + ".split", split.data(), sz));

  FasTC::CompressionJobList list(2);
  const uint32 hh = h >= 8 ? (h / 8) * 4 : h;
  const uint32 sz1 = CompressedImage::GetCompressedSize(w, hh, fmt);
  std::vector<uint8> l0(sz, 0), l1(sz1, 0);
  list.AddJob(FasTC::CompressionJob(fmt, img.data(), l0.data(), w, h));
  list.AddJob(FasTC::CompressionJob(fmt, img.data(), l1.data(), w, hh));
  CHECK(CompressImageList(list, st));
  CHECK(*list.GetFinishedFlag(0) == 1 && *list.GetFinishedFlag(1) == 1);
  CHECK(Dump(prefix + ".list0", l0.data(), sz) && Dump(prefix + ".list1", l1.data(), sz1));

  CompressedImage ci(w, h, fmt, whole.data());
  std::vector<uint8> dec((size_t)w * h * 4);
  CHECK(ci.DecompressImage(dec.data(), (uint32)dec.size()));
  CHECK(Dump(prefix + ".dec", dec.data(), dec.size()));
  FasTC::Image<> orig(w, h, reinterpret_cast<const uint32 *>(img.data()));
  printf("PSNR: %.6f\n", orig.ComputePSNR(&ci));
  printf(g_fail ? "gpu: %d failures\n" : "gpu: ok\n", g_fail);
  return g_fail ? 1 : 0;
}

// `convert in out`: ImageFile::Load + ImageFile::Write (no GPU): exercises the loaders / writers.
static int Convert(const char *in, const char *out) {
  ImageFile src(in);
  if (!src.Load()) return 1;
  ImageFile dst(out, ImageFile::DetectFileFormat(out), *src.GetImage());
  return dst.Write() ? 0 : 1;
}

// `writektx fmt w h payload.bin out.ktx`: a compressed payload wrapped by the KTX writer (no GPU).
static int WriteKtx(const char *fmt, const char *w, const char *h, const char *payload, const char *out) {
  const uint32 width = (uint32)atoi(w), height = (uint32)atoi(h);
  const FasTC::ECompressionFormat f = ParseFormat(fmt);
  std::vector<uint8> data(CompressedImage::GetCompressedSize(width, height, f));
  FILE *fp = fopen(payload, "rb");
  if (!fp || fread(data.data(), 1, data.size(), fp) != data.size()) return 2;
  fclose(fp);
  CompressedImage ci(width, height, f, data.data());
  ImageFile dst(out, eFileFormat_KTX, ci);
  return dst.Write() ? 0 : 1;
}

int main(int argc, char **argv) {
  if (argc >= 7 && !strcmp(argv[1], "writektx")) return WriteKtx(argv[2], argv[3], argv[4], argv[5], argv[6]);
  if (argc >= 2 && !strcmp(argv[1], "api")) return TestApi();
  if (argc >= 4 && !strcmp(argv[1], "convert")) return Convert(argv[2], argv[3]);
  if (argc >= 2 && !strcmp(argv[1], "gpu")) return TestGpu(argc, argv);
  fprintf(stderr, "usage: core_selftest api | convert <in> <out> | writektx <fmt> <w> <h> <payload> <out.ktx> | gpu <rgba.raw> <w> <h> <fmt> <quality> <seed> <out-prefix>\n");
  return 2;
}
