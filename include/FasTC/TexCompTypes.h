// Integer typedefs of the FasTC API (mirrors reference Base/include/FasTC/TexCompTypes.h).
#ifndef FASTC_B200_TEXCOMPTYPES_H_
#define FASTC_B200_TEXCOMPTYPES_H_
#include <stdint.h>
typedef int8_t int8;
typedef uint8_t uint8;
typedef int16_t int16;
typedef uint16_t uint16;
typedef int32_t int32;
typedef uint32_t uint32;
typedef int64_t int64;
typedef uint64_t uint64;
typedef char CHAR;
#endif
