// FasTC::ECompressionFormat and the per-format block geometry.
// API mirror of reference Base/include/FasTC/CompressionFormat.h:24-170: same enumerators
// in the same order (so integer values passed across a binary boundary agree), same
// GetBlockDimensions / GetBlockSize results.  Only DXT1, DXT5, ETC1 and BPTC have an
// encoder in this library.
#ifndef FASTC_B200_COMPRESSIONFORMAT_H_
#define FASTC_B200_COMPRESSIONFORMAT_H_

#include "FasTC/TexCompTypes.h"

namespace FasTC {

enum ECompressionFormat {
  eCompressionFormat_DXT1,
  eCompressionFormat_DXT5,
  eCompressionFormat_ETC1,
  eCompressionFormat_BPTC,

  eCompressionFormat_PVRTC2,
  eCompressionFormat_PVRTC4,
  COMPRESSION_FORMAT_PVRTC_BEGIN = eCompressionFormat_PVRTC2,
  COMPRESSION_FORMAT_PVRTC_END = eCompressionFormat_PVRTC4,

  eCompressionFormat_ASTC4x4,
  eCompressionFormat_ASTC5x4,
  eCompressionFormat_ASTC5x5,
  eCompressionFormat_ASTC6x5,
  eCompressionFormat_ASTC6x6,
  eCompressionFormat_ASTC8x5,
  eCompressionFormat_ASTC8x6,
  eCompressionFormat_ASTC8x8,
  eCompressionFormat_ASTC10x5,
  eCompressionFormat_ASTC10x6,
  eCompressionFormat_ASTC10x8,
  eCompressionFormat_ASTC10x10,
  eCompressionFormat_ASTC12x10,
  eCompressionFormat_ASTC12x12,
  COMPRESSION_FORMAT_ASTC_BEGIN = eCompressionFormat_ASTC4x4,
  COMPRESSION_FORMAT_ASTC_END = eCompressionFormat_ASTC12x12,

  kNumCompressionFormats
};

// Block footprint in pixels (x, y).
inline static void GetBlockDimensions(ECompressionFormat fmt, uint32 (&outSz)[2]) {
  static const uint8 kAstc[14][2] = {{4, 4}, {5, 4}, {5, 5}, {6, 5}, {6, 6}, {8, 5}, {8, 6},
                                     {8, 8}, {10, 5}, {10, 6}, {10, 8}, {10, 10}, {12, 10}, {12, 12}};
  outSz[0] = 4;
  outSz[1] = 4;
  if (fmt == eCompressionFormat_PVRTC2) {
    outSz[0] = 8;
  } else if (fmt >= COMPRESSION_FORMAT_ASTC_BEGIN && fmt <= COMPRESSION_FORMAT_ASTC_END) {
    outSz[0] = kAstc[fmt - COMPRESSION_FORMAT_ASTC_BEGIN][0];
    outSz[1] = kAstc[fmt - COMPRESSION_FORMAT_ASTC_BEGIN][1];
  }
}

// Compressed bytes per block.
inline static uint32 GetBlockSize(ECompressionFormat fmt) {
  switch (fmt) {
    case eCompressionFormat_DXT1:
    case eCompressionFormat_ETC1:
    case eCompressionFormat_PVRTC2:
    case eCompressionFormat_PVRTC4:
      return 8;
    default:
      return fmt < kNumCompressionFormats ? 16 : 8;
  }
}

}  // namespace FasTC
#endif
