"""CPU: the constant tables this repo GENERATES (tools/gen_bc7_tables.py for BC7; the
derivation rule in oracle/etc1_oracle.cpp / fastc_b200/csrc/etc1.cu for rg_etc1's
solid-colour lists) equal the reference's literal arrays.  Needs the reference tree
(build container only); the oracle-vs-golden tests cover the same tables indirectly on
the GPU box."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")

pytestmark = pytest.mark.skipif(not (REF / "BPTCEncoder").exists(), reason="reference tree not present")


def _array(text: str, name: str):
    """All integer literals of the C array definition `name[...] = { ... };`."""
    m = re.search(re.escape(name) + r"\s*(\[[^\]]*\]\s*)+=\s*\{", text)
    assert m, name
    end = text.index("};", m.end())
    body = re.sub(r"//[^\n]*", "", text[m.end():end])
    return [int(x, 0) for x in re.findall(r"0[xX][0-9a-fA-F]+|\d+", body)]


def _ours(name: str, path: Path = ROOT / "oracle" / "bc7_tables.h"):
    return _array(path.read_text().replace("u,", ",").replace("u\n", "\n"), name)


def test_bc7_tables_equal_reference():
    shapes = (REF / "BPTCEncoder/include/FasTC/Shapes.h").read_text()
    anchors = (REF / "BPTCEncoder/src/AnchorTables.h").read_text()
    luts = (REF / "BPTCEncoder/src/BCLookupTables.h").read_text()
    comp = (REF / "BPTCEncoder/src/Compressor.cpp").read_text()
    assert _ours("kShape2") == _array(shapes, "kShapeMask2")
    s3 = _array(shapes, "kShapeMask3")
    want3 = []
    for a, b in zip(s3[0::2], s3[1::2]):  # (subset 1 or 2, subset 2) masks -> 2 bits per pixel
        v = 0
        for i in range(16):
            sub = (1 + ((b >> i) & 1)) if (a >> i) & 1 else 0
            v |= sub << (2 * i)
        want3.append(v)
    assert _ours("kShape3") == want3
    assert _ours("kAnchor2") == _array(anchors, "kAnchorIdx2")
    a3 = _array(anchors, "kAnchorIdx3")
    assert _ours("kAnchor3a") == a3[:64] and _ours("kAnchor3b") == a3[64:]
    assert _ours("kOpt7Mode5") == _array(luts, "Optimal7CompressBC7Mode5")
    assert _ours("kOpt6Dxt1") == _array(luts, "Optimal6CompressDXT1")
    assert _ours("kWatermark") == _array(comp, "kWMValues")
    # the CUDA side's copy is generated from the same script
    cu = ROOT / "fastc_b200" / "csrc" / "bc7_tables.cuh"
    for name in ("kShape2", "kShape3", "kAnchor2", "kAnchor3a", "kAnchor3b", "kWeight", "kOpt7Mode5", "kOpt6Dxt1",
                 "kWatermark"):
        assert _ours(name) == _ours(name, cu), name


def test_etc1_solid_colour_lists_equal_reference():
    src = (REF / "ETCEncoder/src/rg_etc1.cpp").read_text()
    t0 = _array(src, "g_color8_to_etc_block_config_0_255")
    t1 = _array(src, "g_color8_to_etc_block_config_1_to_254")
    want = {}
    lists, cur = [], []
    for v in t0 + t1:
        if v == 0xFFFF:
            lists.append(cur)
            cur = []
        else:
            cur.append(v)
    assert len(lists) == 256
    want[0], want[255] = lists[0], lists[1]
    for c in range(1, 255):
        want[c] = lists[1 + c]
    lib = C.CDLL(str(ROOT / "oracle" / "libfastc_oracle.so"))
    lib.fastc_oracle_etc1_tables.restype = C.c_uint32
    buf = (C.c_uint16 * 64)()
    inv = (C.c_uint16 * (64 * 256))()
    for c in range(256):
        n = lib.fastc_oracle_etc1_tables(c, buf, inv)
        assert list(buf[:n]) == want[c], c
    # inverse lookup: spot-check its defining property (pack_etc1_block_init, rg_etc1.cpp:1905-1936)
    inten = [[-8, -2, 2, 8], [-17, -5, 5, 17], [-29, -9, 9, 29], [-42, -13, 13, 42], [-60, -18, 18, 60],
             [-80, -24, 24, 80], [-106, -33, 33, 106], [-183, -47, 47, 183]]
    inv = np.frombuffer(inv, dtype=np.uint16).reshape(64, 256)
    for idx in (0, 1, 0x1f, 0x2e, 0x3f):
        diff, it, sel = idx & 1, (idx >> 1) & 7, idx >> 4
        for v in (0, 1, 100, 254, 255):
            errs = []
            for pc in range(32 if diff else 16):
                c = ((pc >> 2) | (pc << 3)) if diff else (pc | (pc << 4))
                errs.append(abs(min(255, max(0, c + inten[it][sel])) - v))
            assert inv[idx, v] >> 8 == min(errs) and inv[idx, v] & 0xFF == errs.index(min(errs))


def test_projection_exact_buckets():
    """bc7_anneal's exact-bucket shortcut (fastc_b200/csrc/bc7.cu: sa_eval slow path): when a pixel's
    projection num/den * nbm1 is EXACTLY an integer k, the reference's float sequence
    fl(fl(num / den) * nbm1) (RGBAEndpoints.cpp:262-268) yields exactly k for 0 <= k <= nbm1, stays
    <= 0 for k <= 0 and >= nbm1 for k >= nbm1, so floor == ceil after clamping and one bucket is tested."""
    f = np.float32
    for m in (3, 7, 15):
        for k in range(-70000, 70001):   # |num * nbm1 / den| <= 15 * 4 * 255^2 / 1 in principle; the clamp only needs the sign / >= nbm1
            t = f(f(k) / f(m)) * f(m)
            assert isinstance(t, np.float32)
            if 0 <= k <= m:
                assert t == f(k), (m, k, t)
            elif k < 0:
                assert t < 0, (m, k, t)
            else:
                assert t > f(m) - f(0.5) and np.ceil(t) >= m and np.floor(t) >= m, (m, k, t)


def test_small_division_exact():
    """bc7_setup's div_small (fastc_b200/csrc/bc7.cu): q = RN(a * rc), r = fma(-q, c, a), RN(fma(r, rc, q))
    with rc = RN(1 / c) equals the reference's correctly rounded float division a / c
    (k-means centroid = bucket sum / count, Compressor.cpp:1008-1013) for every integer bucket sum
    a in [0, 16 * 255] and count c in [1, 16].  Exact rational arithmetic, float32 rounding to nearest even."""
    from fractions import Fraction

    def rn(fr):
        if fr == 0:
            return np.float32(0)
        y = np.float32(float(fr))
        cands = [y, np.nextafter(y, np.float32(np.inf)), np.nextafter(y, np.float32(-np.inf))]
        return min(cands, key=lambda c: (abs(Fraction(float(c)) - fr), int(np.float32(c).view(np.uint32)) & 1))

    for c in range(1, 17):
        rc = Fraction(float(rn(Fraction(1, c))))
        for a in range(0, 16 * 255 + 1):
            q = Fraction(float(rn(a * rc)))
            r = Fraction(float(rn(a - q * c)))          # the FMA's exact product, rounded once
            assert r == a - q * c                       # ... and that remainder is exactly representable
            assert rn(q + r * rc) == rn(Fraction(a, c)), (a, c)


def test_fast_projection_guard_band():
    """bc7_anneal / bc7_select pick a pixel's two candidate buckets from v = floor(float(num) * inv16),
    inv16 = RN(65536 * nbm1 / den), a 16.16 fixed-point bucket coordinate, and only trust it when v is
    not within one unit of a multiple of 65536 (otherwise the reference's own float sequence is replayed).
    bc7_anneal computes vp = floor(fma(float(num), inv16, 1)) = v + 1 in one rounding and tests vp instead
    (flag: vp mod 65536 in {0, 1}; bucket: vp >> 16; "not before endpoint 1": vp >= 1): checked too.
    Check, in float32 emulation, that every UNFLAGGED v gives the reference's candidates
    (RGBAEndpoints.cpp:262-289): j1 = clamp(floor(t)), j2 = min(ceil(t), nbm1), t = RN(RN(num / den) * nbm1)."""
    f = np.float32
    rng = np.random.default_rng(42)
    for nbm1 in (3, 7, 15):
        den = np.concatenate([rng.integers(1, 260101, 400000), rng.integers(1, 2000, 200000)]).astype(np.int64)
        num = (rng.random(den.size) * 1.2 - 0.1) * den            # mostly inside [0, den], some outside
        num = np.rint(num).astype(np.int64)
        # adversarial: projections that land next to a bucket boundary
        k = rng.integers(0, nbm1 + 1, den.size)
        near = (k * den) // nbm1 + rng.integers(-1, 2, den.size)
        num = np.where(rng.random(den.size) < 0.5, near, num)
        fden, fnum = den.astype(f), num.astype(f)
        inv16 = (f(65536.0) * f(nbm1)) / fden                      # __fdiv_rn(__fmul_rn(65536, nbm1), fden)
        v = np.floor((fnum * inv16).astype(f)).astype(np.int64)
        flagged = ((v + 1) & 0xFFFE) == 0
        t = ((fnum / fden).astype(f) * f(nbm1)).astype(f)
        j1 = np.clip(np.floor(t).astype(np.int64), 0, nbm1)
        j2 = np.minimum(np.ceil(t).astype(np.int64), nbm1)
        ref_two = (j1 + 1) <= j2
        ja = np.clip(v >> 16, 0, nbm1)
        gpu_two = (v >= 0) & (ja < nbm1)                           # the last palette row repeats its colour
        ok = ~flagged
        assert ok.mean() > 0.5
        assert (ja[ok] == j1[ok]).all(), nbm1
        assert (gpu_two[ok] == ref_two[ok]).all(), nbm1
        # the fused form: one rounding of the exact product plus one
        exact = fnum.astype(np.float64) * inv16.astype(np.float64) + 1.0   # exact in float64 (24 x 24 bit product)
        vp = np.floor(exact.astype(f)).astype(np.int64)
        flagged_p = (vp & 0xFFFE) == 0
        jp = np.clip(vp >> 16, 0, nbm1)
        two_p = (vp >= 1) & (jp < nbm1)
        okp = ~flagged_p
        assert okp.mean() > 0.5
        assert (jp[okp] == j1[okp]).all(), nbm1
        assert (two_p[okp] == ref_two[okp]).all(), nbm1


def test_dxt_quantisation_identities():
    """dxt_block.cuh replaces stb_dxt's Mul8Bit / Expand tables by multiply-shifts and spreads a row's four
    2-bit indices to the byte weights of stb__RefineBlock's w1Tab: every input checked here."""
    def mul8bit(a, b):
        t = a * b + 128
        return (t + (t >> 8)) >> 8
    for x in range(256):
        assert mul8bit(x, 31) == (x * 7967 + 32896) >> 16
        assert mul8bit(x, 63) == (x * 16191 + 32896) >> 16
    for q in range(32):
        assert ((q << 3) | (q >> 2)) == (q * 33) >> 2
    for q in range(64):
        assert ((q << 2) | (q >> 4)) == (q * 65) >> 4
    w1tab = [3, 0, 2, 1]
    for m in range(256):
        s = m
        s = (s | (s << 12)) & 0x000F000F
        s = (s | (s << 6)) & 0x03030303
        s0, s1 = s & 0x01010101, (s >> 1) & 0x01010101
        w1 = ((s0 ^ 0x01010101) << 1) | (s0 ^ s1 ^ 0x01010101)
        for k in range(4):
            assert (w1 >> (8 * k)) & 0xFF == w1tab[(m >> (2 * k)) & 3]
