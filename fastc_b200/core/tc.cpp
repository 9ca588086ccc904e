// `tc`: the command-line front end.  Flag surface, defaults, stdout lines and exit codes of
// reference CLTool/src/tc.cpp:40-346 (`-f -q -n -d -nd -t -j -a -l -v -simd -h`), driving
// the GPU library through CompressImage.  Additions: `-g N` shards over N GPUs (0 = all),
// `-s SEED` pins the annealing RNG key.  Differences: the *Lib formats (PVRTexLib, NVTT) are rejected
// (external libraries), `-l` and `-v` statistics are accepted and ignored, input is PNG (8-bit, non-interlaced) / TGA / KTX.
#include <algorithm>
#include <cstdio>
#include <fstream>
#include <cstdlib>
#include <cstring>

#include "FasTC/CompressedImage.h"
#include "FasTC/Image.h"
#include "FasTC/ImageFile.h"
#include "FasTC/TexComp.h"

static void PrintUsage() {
  fprintf(stderr, "Usage: tc [OPTIONS] imagefile\n");
  fprintf(stderr, "\n");
  fprintf(stderr, "\t-v\t\tVerbose mode (accepted; image statistics are not computed)\n");
  fprintf(stderr, "\t-f <fmt>\tFormat to use. Either \"BPTC\", \"ETC1\", \"DXT1\", \"DXT5\" or \"PVRTC\".\n");
  fprintf(stderr, "\t\t\tDefault: BPTC\n");
  fprintf(stderr, "\t-l\t\tSave an output log (<basename>.log: BPTC per-block path / mode / errors).\n");
  fprintf(stderr, "\t-d <file>\tSpecify decompressed output (default: basename-<fmt>.png); .ktx stores the compressed payload\n");
  fprintf(stderr, "\t-nd\t\tSuppress decompressed output\n");
  fprintf(stderr, "\t-q <quality>\tSet compression quality level. Default: 50\n");
  fprintf(stderr, "\t-n <num>\tCompress the image num times and give the average time and PSNR. Default: 1\n");
  fprintf(stderr, "\t-simd\t\tUse SIMD compression path (not supported)\n");
  fprintf(stderr, "\t-t <num>\tCompress the image using <num> threads (accepted; the GPU sharder decides placement). Default: 1\n");
  fprintf(stderr, "\t-a \t\tCompress the image using synchronization via atomic operations (accepted). Default: Off\n");
  fprintf(stderr, "\t-j <num>\tUse <num> blocks for each work item (pipeline chunk). Default: automatic\n");
  fprintf(stderr, "\t-g <num>\tShard block rows over <num> GPUs (0 = all visible). Default: 1\n");
  fprintf(stderr, "\t-s <seed>\tKey of the per-block annealing random streams. Default: 0\n");
}

static void ExtractBasename(const char *filename, char *buf, size_t bufSz) {
  const char *end = filename + strlen(filename);
  const char *dot = end;
  const char *p = end;
  while (p != filename && *(p - 1) != '/' && *(p - 1) != '\\') {
    --p;
    if (*p == '.' && dot == end) dot = p;
  }
  const size_t n = std::min<size_t>(bufSz - 1, (size_t)(dot - p));
  memcpy(buf, p, n);
  buf[n] = '\0';
}

int main(int argc, char **argv) {
  int fileArg = 1;
  if (fileArg == argc) {
    PrintUsage();
    exit(1);
  }
  char decompressedOutput[256] = "";
  bool bDecompress = true, bUseSIMD = false, bUseAtomics = false, bFormatOk = true, bSaveLog = false;
  int numJobs = 0, quality = 50, numThreads = 1, numCompressions = 1, numGPUs = 1;
  unsigned long long seed = 0;
  FasTC::ECompressionFormat format = FasTC::eCompressionFormat_BPTC;

  // every option that takes a value shares one pattern: missing value or value below `lo` -> usage, exit 1
  auto intArg = [&](int &dst, int lo) {
    fileArg++;
    if (fileArg == argc || (dst = atoi(argv[fileArg])) < lo) {
      PrintUsage();
      exit(1);
    }
    fileArg++;
  };
  bool known = true;
  while (known && fileArg < argc) {
    const char *a = argv[fileArg];
    known = true;
    if (!strcmp(a, "-n")) intArg(numCompressions, 0);
    else if (!strcmp(a, "-t")) intArg(numThreads, 1);
    else if (!strcmp(a, "-q")) intArg(quality, 0);
    else if (!strcmp(a, "-j")) intArg(numJobs, 0);
    else if (!strcmp(a, "-g")) intArg(numGPUs, 0);
    else if (!strcmp(a, "-s")) {
      fileArg++;
      if (fileArg == argc) { PrintUsage(); exit(1); }
      seed = strtoull(argv[fileArg++], NULL, 0);
    } else if (!strcmp(a, "-f")) {
      fileArg++;
      if (fileArg == argc) { PrintUsage(); exit(1); }
      const char *f = argv[fileArg++];
      if (!strcmp(f, "ETC1")) format = FasTC::eCompressionFormat_ETC1;
      else if (!strcmp(f, "DXT1")) format = FasTC::eCompressionFormat_DXT1;
      else if (!strcmp(f, "DXT5")) format = FasTC::eCompressionFormat_DXT5;
      else if (!strcmp(f, "BPTC")) format = FasTC::eCompressionFormat_BPTC;
      else if (!strcmp(f, "PVRTC")) format = FasTC::eCompressionFormat_PVRTC4;
      else if (!strcmp(f, "PVRTCLib") || !strcmp(f, "BPTCLib")) bFormatOk = false;
      // any other string silently keeps the current format, like the reference (tc.cpp:121-148)
    } else if (!strcmp(a, "-h") || !strcmp(a, "--help")) {
      PrintUsage();
      exit(0);
    } else if (!strcmp(a, "-d")) {
      fileArg++;
      if (fileArg == argc) { PrintUsage(); exit(1); }
      snprintf(decompressedOutput, sizeof(decompressedOutput), "%s", argv[fileArg++]);
    } else if (!strcmp(a, "-nd")) { fileArg++; bDecompress = false; }
    else if (!strcmp(a, "-l")) { bSaveLog = true; fileArg++; }
    else if (!strcmp(a, "-v")) fileArg++;
    else if (!strcmp(a, "-simd")) { fileArg++; bUseSIMD = true; }
    else if (!strcmp(a, "-a")) { fileArg++; bUseAtomics = true; }
    else known = false;
  }
  if (fileArg == argc) {
    PrintUsage();
    exit(1);
  }
  if (!bFormatOk) {
    fprintf(stderr, "TexComp -- the external-library encoders (PVRTCLib, BPTCLib) are not supported on the GPU path\n");
    return 1;
  }

  char basename[256];
  ExtractBasename(argv[fileArg], basename, sizeof(basename));

  ImageFile file(argv[fileArg]);
  if (!file.Load()) return 1;
  file.GetImage()->ComputePixels();  // a compressed KTX input decodes here
  FasTC::Image<> img(*file.GetImage());

  SCompressionSettings settings;
  settings.format = format;
  settings.bUseSIMD = bUseSIMD;
  settings.bUseAtomics = bUseAtomics;
  settings.iNumThreads = numThreads;
  settings.iQuality = quality;
  settings.iNumCompressions = numCompressions;
  settings.iJobSize = numJobs;
  settings.iNumGPUs = numGPUs;
  settings.uSeed = seed;
  // -l: "<basename>.log" like the reference (CLTool/src/tc.cpp:268-291)
  std::ofstream logFile;
  if (bSaveLog) {
    char logname[300];
    snprintf(logname, sizeof(logname), "%s.log", basename);
    logFile.open(logname);
    settings.logStream = &logFile;
  }

  CompressedImage *ci = CompressImage(&img, settings);
  if (NULL == ci) return 1;

  if (ci->GetWidth() != img.GetWidth() || ci->GetHeight() != img.GetHeight()) {
    fprintf(stderr, "Cannot compute image metrics: compressed and uncompressed dimensions differ.\n");
  } else {
    const double PSNR = img.ComputePSNR(ci);
    if (PSNR > 0.0) fprintf(stdout, "PSNR: %.3f\n", PSNR);
    else fprintf(stderr, "Error computing PSNR\n");
  }

  int rc = 0;
  if (bDecompress) {
    char outname[512];
    if (decompressedOutput[0] != '\0') {
      snprintf(outname, sizeof(outname), "%s", decompressedOutput);
    } else {
      const char *suffix = format == FasTC::eCompressionFormat_BPTC   ? "-bptc.png"
                           : format == FasTC::eCompressionFormat_DXT1 ? "-dxt1.png"
                           : format == FasTC::eCompressionFormat_DXT5 ? "-dxt5.png"
                           : format == FasTC::eCompressionFormat_PVRTC4 ? "-pvrtc-4bpp.png"
                                                                      : "-etc1.png";
      snprintf(outname, sizeof(outname), "%s%s", basename, suffix);
    }
    ImageFile out(outname, ImageFile::DetectFileFormat(outname), *ci);
    if (!out.Write()) rc = 1;
  }
  delete ci;
  return rc;
}
