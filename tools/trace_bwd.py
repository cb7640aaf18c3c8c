"""Timeline of one CTA pair of the backward kernel (stall-accounting build): per ring box, when the producer saw the
slot free / issued the TMA, when the MMA thread started / stopped waiting for it and committed its MMAs."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("DCLIP_B200_LIB", "build/prof/libdclip_b200_prof.so")
import numpy as np
import torch
import tools.time_bwd as tb

B, d = 4096, 768
tb.run(B, d, n=3)
lib = tb._lib.load_library()
buf = (ctypes.c_uint * (8 * 2048))()
lib.dclip_debug_pair_trace(buf, 8 * 2048)
tr = np.frombuffer(buf, dtype=np.uint32).reshape(8, 2048).astype(np.int64)
mma = tr[6]
t0 = mma[0]
rel = lambda x: int((x - t0) & 0xffffffff)
print("job  tile ph idx | prod: slot-free  issued | mma: wait-start  full-seen (waited)  committed | issue->full")
for t in range(0, 4):
    for j in range(24):
        q = t * 24 + j                      # MMA job index (P1 boxes 0-11; P2 pairs 12-17)
        if j >= 18: continue
        ws, fs, cm = mma[3 * q], mma[3 * q + 1], mma[3 * q + 2]
        # ring box index in consumption order: P1(0) boxes 0..11; then per tile t: P1(t+1) 12 boxes, P2(t) 12 boxes
        if j < 12:
            box = j if t == 0 else 12 + (t - 1) * 24 + j
            ph = "P1"
        else:
            box = 12 + t * 24 + 12 + (j - 12) * 2 if t < 15 else None
            ph = "P2"
        # NOTE MMA order: P1(0), then t: P1(t+1), P2(t).  Trace index uses (t, phase) of the tile itself.
        if box is None: continue
        w, k = box % 6, box // 6
        pe, pi = tr[w][2 * k], tr[w][2 * k + 1]
        print(f"{q:4d} {t:4d} {ph} {j:3d} | {rel(pe):10d} {rel(pi):8d} | {rel(ws):10d} {rel(fs):10d} ({int(fs-ws):5d}) {rel(cm):10d} | {int(fs-pi):6d}")
g = tr[7]
for t in range(4):
    print(f"tile {t}: g_full wait {rel(g[2*t])} -> {rel(g[2*t+1])} ({int(g[2*t+1]-g[2*t])});  s_empty wait {rel(g[64+2*t])} -> {rel(g[64+2*t+1])} ({int(g[64+2*t+1]-g[64+2*t])})")
