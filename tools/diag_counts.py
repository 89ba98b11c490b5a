import sys, torch
sys.path.insert(0, ".")
from structuredetector_b200 import ops
from structuredetector_b200.synth import CONFIGS, make_raw, split_outputs
cfg = CONFIGS["cfg3"]
for mode in ("noise", "blobs"):
    raw = make_raw(cfg, mode, batch=32).cuda()
    outs = split_outputs(raw, cfg.labels, cfg.parts)
    for wk in (False, True):
        pk = ops.decode_packed(outs, 100, 100, 0.4, 0.1, warp_kernel=wk)
        d = pk.diag.cpu().float()
        print(mode, "warp" if wk else "tma", "candidates per plane: mean %.0f min %.0f max %.0f" % (d[:,0].mean(), d[:,0].min(), d[:,0].max()))
# how many pixels / rows beat the exact plane threshold (best case for any one-pass pruning)
import numpy as np
raw = make_raw(cfg, "noise", batch=2)
x = raw[:, :3].numpy()
for b in range(2):
    for c in range(3):
        pl = x[b, c]
        # peaks in logit space
        from oracle.sdnet_oracle import window_max
        mx = window_max(pl[None, None])[0, 0]
        peaks = np.sort(pl[pl == mx])[::-1]
        thr = peaks[99]
        # online: rows processed top to bottom, floor = 100th best peak seen so far (plane-wide, ideal)
        H, W = pl.shape
        seen = []
        slow_rows = 0; emitted = 0
        import heapq
        heap = []
        for r in range(H):
            floor = heap[0] if len(heap) >= 100 else -np.inf
            row = pl[r]
            for p0 in range(0, W, 128):
                if (row[p0:p0+128] > floor).any(): slow_rows += 1
            pk = row[(row == mx[r]) & (row > floor)]
            emitted += len(pk)
            for v in pk:
                if len(heap) < 100: heapq.heappush(heap, v)
                elif v > heap[0]: heapq.heapreplace(heap, v)
        print(f"plane {b},{c}: ideal-floor slow panel-rows {slow_rows}/{H*5} = {slow_rows/(H*5):.3f}, emitted {emitted}, final thr {thr:.3f}")
