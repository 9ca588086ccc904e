"""Debug helper (GPU box): compare per-chain fit errors, GPU vs oracle, block by block."""
import sys, ctypes as C
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
from collections import Counter
from _checkers import Oracle
from fastc_b200 import ECompressionFormat as F, lib
from fastc_b200.synth import synth_rgba

q = int(sys.argv[1]) if len(sys.argv) > 1 else 0
w = h = 128
img = synth_rgba(w, h, 1)
g, o = lib(), Oracle()
nb = (w // 4) * (h // 4)
d_in = torch.from_numpy(img).cuda(); d_out = torch.zeros(nb * 16, dtype=torch.uint8, device="cuda")
g.compress_device(F.BPTC, d_in, d_out, width=w, height=h, quality=q, seed=5)
torch.cuda.synchronize()
sel = np.zeros(nb, np.uint32); res = np.zeros((nb, 16, 8), np.uint32)
g.check(g.cdll.fastc_gpu_debug_bc7_dump(nb, sel.ctypes.data, res.ctypes.data))
o.lib.fastc_oracle_bc7_chain_errors.argtypes = [C.POINTER(C.c_double)]
blocks = img.reshape(h // 4, 4, w // 4, 4, 4).transpose(0, 2, 1, 3, 4).reshape(-1, 4, 4, 4)
def slot_chain(selw, slot):
    layout_b = (selw >> 22) & 1
    if not layout_b:
        if slot < 3: return 0 * 8 + 4 + slot
        if slot < 6: return 2 * 8 + 4 + slot - 3
        if slot < 8: return 1 * 8 + slot - 6
        if slot < 10: return 3 * 8 + slot - 8
        if slot < 12: return 7 * 8 + slot - 10
        if slot < 14: return 6 * 8 + (slot - 12) * 4
        return None
    if slot < 8: return 32 + slot
    if slot < 12: return 40 + slot - 8
    if slot == 12: return 48
    if slot < 15: return 56 + slot - 13
    return None
bad = Counter(); shown = 0
for k in range(nb):
    if (sel[k] >> 24) != 0: continue
    blk = np.ascontiguousarray(blocks[k])
    # oracle on the single block but keyed like block k
    out = np.zeros(16, np.uint8)
    o.lib.fastc_oracle_bc7_keyed(blk.ctypes.data_as(C.POINTER(C.c_uint8)), 4, 4, 0, 1, out.ctypes.data_as(C.POINTER(C.c_uint8)), q, C.c_uint64(5), 0, k)
    errs = (C.c_double * 64)(); o.lib.fastc_oracle_bc7_chain_errors(errs)
    for slot in range(16):
        cid = slot_chain(int(sel[k]), slot)
        if cid is None: continue
        oe = errs[cid]
        if oe < 0: continue
        ge = float(res[k, slot, 0])
        if ge != oe:
            bad[cid // 8] += 1
            if shown < 15:
                shown += 1
                print("block", k, "sel", hex(int(sel[k])), "slot", slot, "chain", cid, "gpu err", ge, "oracle err", oe,
                      "gpu p1/p2 %08x %08x combo %d" % (res[k, slot, 1], res[k, slot, 2], res[k, slot, 3]))
print("chains with differing error, by mode:", sorted(bad.items()))
