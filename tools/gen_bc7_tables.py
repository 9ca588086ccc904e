#!/usr/bin/env python3
"""Generates the BC7 constant tables used by the oracle (oracle/bc7_tables.h) and
by the CUDA encoder (fastc_b200/csrc/bc7_tables.cuh).

Everything here is either BC7-specification data (partition sets, anchor
indices, interpolation weights, per-mode bit layout) or derived by the search
programs the reference documents in comments next to its lookup tables
(reference/BPTCEncoder/src/BCLookupTables.h:44-96 and :364-449).  The only
reference-specific constants are the nine "watermark" words the reference
writes into the unused alpha-index field of solid-colour blocks
(reference/BPTCEncoder/src/Compressor.cpp:135-140); bit-identical output
requires the same words.

tests/test_tables.py re-derives the reference's tables from its headers (when
/root/reference is present) and checks they equal what this script emits.
"""
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

# --- BC7 spec: 2-subset partitions, bit i = subset of pixel i (raster order) ---
P2 = [
    0xcccc, 0x8888, 0xeeee, 0xecc8, 0xc880, 0xfeec, 0xfec8, 0xec80,
    0xc800, 0xffec, 0xfe80, 0xe800, 0xffe8, 0xff00, 0xfff0, 0xf000,
    0xf710, 0x008e, 0x7100, 0x08ce, 0x008c, 0x7310, 0x3100, 0x8cce,
    0x088c, 0x3110, 0x6666, 0x366c, 0x17e8, 0x0ff0, 0x718e, 0x399c,
    0xaaaa, 0xf0f0, 0x5a5a, 0x33cc, 0x3c3c, 0x55aa, 0x9696, 0xa55a,
    0x73ce, 0x13c8, 0x324c, 0x3bdc, 0x6996, 0xc33c, 0x9966, 0x0660,
    0x0272, 0x04e4, 0x4e40, 0x2720, 0xc936, 0x936c, 0x39c6, 0x639c,
    0x9336, 0x9cc6, 0x817e, 0xe718, 0xccf0, 0x0fcc, 0x7744, 0xee22,
]

# --- BC7 spec: 3-subset partitions as (mask of pixels in subset 1 or 2, mask of subset 2) ---
P3 = [
    (0xfecc, 0xf600), (0xffc8, 0x7300), (0xff90, 0x3310), (0xecce, 0x00ce),
    (0xff00, 0xcc00), (0xcccc, 0xcc00), (0xffcc, 0x00cc), (0xffcc, 0x3300),
    (0xff00, 0xf000), (0xfff0, 0xf000), (0xfff0, 0xff00), (0xcccc, 0x8888),
    (0xeeee, 0x8888), (0xeeee, 0xcccc), (0xffec, 0xec80), (0x739c, 0x7310),
    (0xfec8, 0xc800), (0x39ce, 0x3100), (0xfff0, 0xccc0), (0xfccc, 0x0ccc),
    (0xeeee, 0xee00), (0xff88, 0x7700), (0xeec0, 0xcc00), (0x7730, 0x3300),
    (0x0cee, 0x00cc), (0xffcc, 0xfc88), (0x6ff6, 0x0660), (0xff60, 0x6600),
    (0xcbbc, 0xc88c), (0xf966, 0xf900), (0xceec, 0x0cc0), (0xff10, 0x7310),
    (0xff80, 0xec80), (0xccce, 0x08ce), (0xeccc, 0xec80), (0x6666, 0x4444),
    (0x0ff0, 0x0f00), (0x6db6, 0x4924), (0x6bd6, 0x4294), (0xcf3c, 0x0c30),
    (0xc3fc, 0x03c0), (0xffaa, 0xff00), (0xff00, 0x5500), (0xfcfc, 0xcccc),
    (0xcccc, 0x0c0c), (0xf6f6, 0x6666), (0xaffa, 0x0ff0), (0xfff0, 0x5550),
    (0xfaaa, 0xf000), (0xeeee, 0x0e0e), (0xf8f8, 0x8888), (0xfff0, 0x9990),
    (0xeeee, 0xe00e), (0x8ff8, 0x8888), (0xf666, 0xf000), (0xff00, 0x9900),
    (0xff66, 0xff00), (0xcccc, 0xc00c), (0xcffc, 0xcccc), (0xf000, 0x9000),
    (0x8888, 0x0808), (0xfefe, 0xeeee), (0xfffa, 0xfff0), (0x7bde, 0x7310),
]

# --- BC7 spec: anchor ("fix-up") index of the 2nd subset (2-subset sets) ---
A2 = [
    15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15,
    15, 2, 8, 2, 2, 8, 8, 15, 2, 8, 2, 2, 8, 8, 2, 2,
    15, 15, 6, 8, 2, 8, 15, 15, 2, 8, 2, 2, 2, 15, 15, 6,
    6, 2, 6, 8, 15, 15, 2, 2, 15, 15, 15, 15, 15, 2, 2, 15,
]
# --- anchor index of the 2nd and 3rd subset (3-subset sets) ---
A3A = [
    3, 3, 15, 15, 8, 3, 15, 15, 8, 8, 6, 6, 6, 5, 3, 3,
    3, 3, 8, 15, 3, 3, 6, 10, 5, 8, 8, 6, 8, 5, 15, 15,
    8, 15, 3, 5, 6, 10, 8, 15, 15, 3, 15, 5, 15, 15, 15, 15,
    3, 15, 5, 5, 5, 8, 5, 10, 5, 10, 8, 13, 15, 12, 3, 3,
]
A3B = [
    15, 8, 8, 3, 15, 15, 3, 8, 15, 15, 15, 15, 15, 15, 15, 8,
    15, 8, 15, 3, 15, 8, 15, 8, 3, 15, 6, 10, 15, 15, 10, 8,
    15, 3, 15, 10, 10, 8, 9, 10, 6, 15, 8, 15, 3, 6, 6, 8,
    15, 3, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 3, 15, 15, 8,
]

# --- BC7 spec: interpolation weights for 2/3/4-bit indices ---
W = {
    2: [0, 21, 43, 64],
    3: [0, 9, 18, 27, 37, 46, 55, 64],
    4: [0, 4, 9, 13, 17, 21, 26, 30, 34, 38, 43, 47, 51, 55, 60, 64],
}

# mode: (partition bits, subsets, index bits, alpha index bits, colour bits, alpha bits,
#        rotation, idx-mode bit, p-bit type [0 shared, 1 per-endpoint, 2 none])
MODES = [
    (4, 3, 3, 0, 4, 0, 0, 0, 1),
    (6, 2, 3, 0, 6, 0, 0, 0, 0),
    (6, 3, 2, 0, 5, 0, 0, 0, 2),
    (6, 2, 2, 0, 7, 0, 0, 0, 1),
    (0, 1, 2, 3, 5, 6, 1, 1, 2),
    (0, 1, 2, 2, 7, 8, 1, 0, 2),
    (0, 1, 4, 0, 7, 7, 0, 0, 1),
    (6, 2, 2, 0, 5, 5, 0, 0, 1),
]

# reference/BPTCEncoder/src/Compressor.cpp:135-138
WM = [0x32b92180, 0x32ba3080, 0x31103200, 0x28103c80, 0x32bb3080, 0x25903600, 0x3530b900, 0x3b32b180,
      0x34b5b98]


def expand_vals(nbits):
    vals, last = [], -1
    for i in range(256):
        num = ((i >> (8 - nbits)) << (8 - nbits)) | (i >> nbits)
        if num != last:
            last = num
            vals.append(num)
    return vals


def opt7_mode5():
    """BCLookupTables.h:44-96 -- endpoints (7 bit) whose index-1 interpolant hits i."""
    vals = expand_vals(7)
    out = []
    for i in range(256):
        best, bj, bk = 1 << 32, 0, 0
        for j, vj in enumerate(vals):
            for k, vk in enumerate(vals):
                d = abs(i - ((43 * vj + 21 * vk + 32) >> 6))
                if d < best:
                    best, bj, bk = d, j, k
        assert best == 0
        out.append((vals[bj] >> 1, vals[bk] >> 1))
    return out


def opt_dxt1(nbits):
    """BCLookupTables.h:364-449 -- best midpoint / one-third pair per value."""
    vals = expand_vals(nbits)
    out = []
    for i in range(256):
        def search(f):
            best, bj, bk = 1 << 32, 0, 0
            for j in range(len(vals)):
                for k in range(j, len(vals)):
                    d = abs(i - f(vals[j], vals[k]))
                    if d < best:
                        best, bj, bk = d, j, k
            return best, vals[bj] >> (8 - nbits), vals[bk] >> (8 - nbits)
        d0, a0, b0 = search(lambda a, b: (a + b) // 2)
        d1, a1, b1 = search(lambda a, b: (2 * a + b) // 3)
        e0, e1 = (1, a0, b0), (0, a1, b1)
        out.append((e0, e1) if d1 > d0 else (e1, e0))
    return out


def p3_packed():
    """2 bits per pixel: subset index of pixel i in bits [2i, 2i+1]."""
    res = []
    for m0, m1 in P3:
        v = 0
        for i in range(16):
            s = (1 + ((m1 >> i) & 1)) if (m0 >> i) & 1 else 0
            v |= s << (2 * i)
        res.append(v)
    return res


def arr(name, ctype, vals, per=8, fmt="0x%x", qual="static const"):
    s = f"{qual} {ctype} {name}[{len(vals)}] = {{\n"
    for i in range(0, len(vals), per):
        s += "  " + ", ".join(fmt % v for v in vals[i:i + per]) + ",\n"
    return s + "};\n"


def emit(qual, banner):
    o7 = opt7_mode5()
    o6 = opt_dxt1(6)
    s = banner
    s += arr("kShape2", "uint16_t", P2, fmt="0x%04x", qual=qual)
    s += arr("kShape3", "uint32_t", p3_packed(), per=4, fmt="0x%08xu", qual=qual)
    s += arr("kAnchor2", "uint8_t", A2, per=16, fmt="%d", qual=qual)
    s += arr("kAnchor3a", "uint8_t", A3A, per=16, fmt="%d", qual=qual)
    s += arr("kAnchor3b", "uint8_t", A3B, per=16, fmt="%d", qual=qual)
    # weights indexed [bits-1][index]; row 0 (1-bit) exists in the reference but is never used
    flat = []
    for bits in (1, 2, 3, 4):
        w = W.get(bits, [0, 31, 64])
        flat += [w[i] if i < len(w) else 255 for i in range(16)]
    s += arr("kWeight", "uint8_t", flat, per=16, fmt="%d", qual=qual)
    s += arr("kOpt7Mode5", "uint8_t", [v for p in o7 for v in p], per=16, fmt="%d", qual=qual)
    # mode-4 constant alpha: only [a][0][0], [a][0][1], [a][0][2], [a][1][1], [a][1][2] are read
    s += arr("kOpt6Dxt1", "uint8_t", [v for e in o6 for t in e for v in t], per=12, fmt="%d", qual=qual)
    s += arr("kWatermark", "uint32_t", WM, per=4, fmt="0x%08xu", qual=qual)
    return s


ORACLE_BANNER = """// GENERATED by tools/gen_bc7_tables.py -- do not edit.
// TEST INFRASTRUCTURE (oracle side).  BC7 specification tables + the lookup
// tables the reference derives (see the generator for provenance).
#pragma once
#include <stdint.h>
namespace bc7t {
"""

ORACLE_TAIL = """
enum { kPbitShared = 0, kPbitPerEndpoint = 1, kPbitNone = 2 };
struct ModeAttr {
  int partition_bits, subsets, index_bits, alpha_index_bits, color_bits, alpha_bits;
  int has_rotation, has_idx_mode, pbit_type;
};
static const ModeAttr kModes[8] = {
%s};
// kInterp[bits-1][index][0|1] = (weight of endpoint 0, weight of endpoint 1)
struct InterpTable {
  uint32_t v[4][16][2];
  InterpTable() {
    for (int b = 0; b < 4; b++)
      for (int i = 0; i < 16; i++) {
        int w = kWeight[b * 16 + i];
        v[b][i][0] = w == 255 ? 0 : 64 - w;
        v[b][i][1] = w == 255 ? 0 : w;
      }
    // reference row 0 is {64,0},{33,31},{0,64} (Compressor.cpp:160); never used.
    v[0][1][0] = 33; v[0][1][1] = 31;
  }
  const uint32_t (*operator[](int b) const)[2] { return v[b]; }
};
static const InterpTable kInterp;
static inline int subset_of(int idx, int shape, int nsubsets) {
  if (nsubsets == 2) return (kShape2[shape] >> idx) & 1;
  if (nsubsets == 3) return (kShape3[shape] >> (2 * idx)) & 3;
  return 0;
}
static inline int anchor_of(int subset, int shape, int nsubsets) {
  if (subset == 0) return 0;
  if (subset == 1) return nsubsets == 2 ? kAnchor2[shape] : kAnchor3a[shape];
  return kAnchor3b[shape];
}
}  // namespace bc7t
"""

CUDA_BANNER = """// GENERATED by tools/gen_bc7_tables.py -- do not edit.
// BC7 specification tables + derived lookup tables (host copies; bc7.cu uploads
// them to __constant__ / global memory).
#pragma once
#include <stdint.h>
namespace fastc {
namespace bc7tab {
"""


def main():
    modes = "".join("  {%s},\n" % ", ".join(str(v) for v in m) for m in MODES)
    (ROOT / "oracle" / "bc7_tables.h").write_text(emit("static const", ORACLE_BANNER) + ORACLE_TAIL % modes)
    (ROOT / "fastc_b200" / "csrc" / "bc7_tables.cuh").write_text(
        emit("static const", CUDA_BANNER) + "}  // namespace bc7tab\n}  // namespace fastc\n")


if __name__ == "__main__":
    main()
