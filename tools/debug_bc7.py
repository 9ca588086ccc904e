"""Debug helper (GPU box): categorise BC7 mismatches between the CUDA path and the oracle."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from collections import Counter
from _checkers import Oracle
from fastc_b200 import ECompressionFormat as F, lib
from fastc_b200.synth import synth_rgba

q = int(sys.argv[1]) if len(sys.argv) > 1 else 0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
w = h = 256
img = synth_rgba(w, h, seed)
g, o = lib(), Oracle()
got, _ = g.compress(F.BPTC, img, quality=q, seed=5)
want, _ = o.compress("BPTC", img, quality=q, rng_mode=1, seed=5)
G = got.reshape(-1, 16); W = want.reshape(-1, 16)
bad = np.nonzero((G != W).any(1))[0]
def mode(b):
    v = int(b[0]); k = 0
    while k < 8 and not (v >> k) & 1: k += 1
    return k
print("mismatches", len(bad), "of", len(G))
print("oracle mode hist", Counter(mode(b) for b in W))
print("(gpu mode, oracle mode) of mismatches:", sorted(Counter((mode(G[k]), mode(W[k])) for k in bad).items()))
blocks = img.reshape(h // 4, 4, w // 4, 4, 4).transpose(0, 2, 1, 3, 4).reshape(-1, 16, 4)
for k in bad[:12]:
    print(k, "gpu", G[k].tobytes().hex(), "orc", W[k].tobytes().hex(), "amin", blocks[k][:, 3].min())
dec_g = o.decode("BPTC", got, w, h); dec_w = o.decode("BPTC", want, w, h)
print("psnr gpu", o.psnr(img, dec_g), "oracle", o.psnr(img, dec_w))
