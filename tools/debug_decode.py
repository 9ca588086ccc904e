import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from _checkers import Oracle
from fastc_b200 import ECompressionFormat as F, lib
g, o = lib(), Oracle()
rng = np.random.default_rng(3)
w, h = 256, 128
rnd = rng.integers(0, 256, (w // 4) * (h // 4) * 16, dtype=np.uint8)
blk = rnd.reshape(-1, 16)
m = rng.integers(0, 8, len(blk))
blk[:, 0] = (blk[:, 0] & ~((1 << (m + 1)) - 1).astype(np.uint8)) | (1 << m).astype(np.uint8)
got = g.decompress(F.BPTC, rnd, w, h); want = o.decode("BPTC", rnd, w, h)
gb = got.reshape(h // 4, 4, w // 4, 4, 4).transpose(0, 2, 1, 3, 4).reshape(-1, 16, 4)
wb = want.reshape(h // 4, 4, w // 4, 4, 4).transpose(0, 2, 1, 3, 4).reshape(-1, 16, 4)
bad = np.nonzero((gb != wb).any((1, 2)))[0]
print("bad blocks", len(bad), "modes of bad", np.bincount(m[bad], minlength=8), "all", np.bincount(m, minlength=8))
for b in bad[:3]:
    print("block", b, "mode", m[b], "bytes", blk[b].tolist())
    print(" got ", gb[b].tolist()); print(" want", wb[b].tolist())
