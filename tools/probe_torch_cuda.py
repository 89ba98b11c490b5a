"""GPU-box probe: how do the stock torch CUDA ops the reference calls break ties?

Writes gpurun_out/probe_torch_cuda.json.  Test infrastructure only (not product code).
"""
import json, os, sys, time
import torch

out = {"torch": torch.__version__, "device": torch.cuda.get_device_name(0)}
dev = "cuda"
g = torch.Generator().manual_seed(7)

def canon_topk(v, k):
    # (value desc, index asc) via stable sort
    idx = torch.sort(v, dim=-1, descending=True, stable=True).indices[..., :k]
    return torch.gather(v, -1, idx), idx

res = {}
for name, n, k, levels in [("plane313k_q64", 313344, 100, 64), ("plane313k_q8", 313344, 100, 8),
                           ("plane16k_q16", 16384, 100, 16), ("plane65k_k500_q32", 65536, 500, 32),
                           ("stage2_300_q8", 300, 100, 8), ("stage2_15000_k500_q64", 15000, 500, 64),
                           ("allequal16k", 16384, 10, 1), ("plane313k_sparse", 313344, 100, 0)]:
    B = 6
    if levels == 0:
        v = torch.zeros(B, n)
        pos = torch.randint(0, n, (B, 37), generator=g)
        v.scatter_(1, pos, torch.randint(1, 5, (B, 37), generator=g).float() / 8)
    elif levels == 1:
        v = torch.full((B, n), 0.5)
    else:
        v = torch.randint(0, levels, (B, n), generator=g).float() / levels
    vc = v.to(dev)
    tv, ti = torch.topk(vc, k)
    cv, ci = canon_topk(vc, k)
    res[name] = {"values_equal": bool(torch.equal(tv, cv)), "indices_equal_canonical": bool(torch.equal(ti, ci)),
                 "first_row_head": ti[0, :12].tolist(), "canon_head": ci[0, :12].tolist()}
    # 3-D view like the reference: (B, C, HW)
    v3 = vc.view(2, 3, n)
    tv3, ti3 = torch.topk(v3, k)
    res[name]["indices_equal_canonical_3d"] = bool(torch.equal(ti3.view(B, k), ci))
out["topk"] = res

# min(dim=1) tie rule
d = torch.randint(0, 4, (4, 100, 100), generator=g).float().to(dev)
mv, mi = d.min(dim=1)
first = torch.argmax((d == mv.unsqueeze(1)).int(), dim=1)
out["min_dim1_first_index"] = bool(torch.equal(mi, first))

# sigmoid monotonicity + clamp over all fp32 in [-20, 20]
def sweep(lo_bits, hi_bits, negative):
    bad = 0; total = 0
    step = 1 << 26
    prev_last = None
    for s in range(lo_bits, hi_bits, step):
        e = min(s + step, hi_bits)
        bits = torch.arange(s, e, device=dev, dtype=torch.int64).to(torch.int32)
        x = bits.view(torch.float32)
        y = torch.clamp(torch.sigmoid(x), min=1e-6, max=1 - 1e-6)
        dy = y[1:] - y[:-1]
        # for negative floats, increasing bits => decreasing x
        bad += int(((dy > 0) if negative else (dy < 0)).sum())
        if prev_last is not None:
            dd = float(y[0] - prev_last)
            bad += int(dd > 0) if negative else int(dd < 0)
        prev_last = y[-1].clone()
        total += e - s
    return bad, total
import struct
f2b = lambda f: struct.unpack("<I", struct.pack("<f", f))[0]
bp, tp = sweep(0, f2b(20.0) + 1, False)
bn, tn = sweep(f2b(-0.0) - (1 << 32), f2b(-20.0) - (1 << 32) + 1, True) if False else (None, None)
# negative side: bits 0x80000000.. as int64 then wrap
def sweep_neg():
    bad = 0; total = 0; step = 1 << 26; prev_last = None
    lo, hi = 0x80000000, f2b(-20.0) + 1
    for s in range(lo, hi, step):
        e = min(s + step, hi)
        bits = (torch.arange(s, e, device=dev, dtype=torch.int64) - (1 << 32)).to(torch.int32)
        x = bits.view(torch.float32)
        y = torch.clamp(torch.sigmoid(x), min=1e-6, max=1 - 1e-6)
        dy = y[1:] - y[:-1]
        bad += int((dy > 0).sum())
        if prev_last is not None:
            bad += int(float(y[0] - prev_last) > 0)
        prev_last = y[-1].clone(); total += e - s
    return bad, total
bn, tn = sweep_neg()
out["sigmoid_monotone"] = {"pos_violations": bp, "pos_total": tp, "neg_violations": bn, "neg_total": tn}
xs = torch.tensor([13.0, 13.5, 13.8, 13.81, 13.82, 13.9, 14.0, 20.0, -13.8, -13.81, -13.82, -13.9, -14, 0.0], device=dev)
out["sigmoid_samples"] = {str(float(a)): float(b) for a, b in zip(xs, torch.clamp(torch.sigmoid(xs), 1e-6, 1 - 1e-6))}
out["clamp_bits"] = {"lo": f2b(float(torch.tensor(1e-6, dtype=torch.float32))), "hi": f2b(float(torch.tensor(1 - 1e-6, dtype=torch.float32)))}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe_torch_cuda.json", "w"), indent=1)
print(json.dumps(out, indent=1))
