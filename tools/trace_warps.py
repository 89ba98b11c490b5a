"""Per-warp timeline of the peaks tile kernel (diagnostics build: tools/xbuild.sh trace -DSDNET_X_TRACE=1).
usage: python tools/trace_warps.py [images] [mode] [ENV=VAL ...]"""
import ctypes, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
for kv in sys.argv[3:]:
    k, v = kv.split("="); os.environ[k] = v
import numpy as np, torch
from structuredetector_b200 import _native, ops
from structuredetector_b200.synth import CONFIGS, make_raw, split_outputs
images = int(sys.argv[1]) if len(sys.argv) > 1 else 128
mode = sys.argv[2] if len(sys.argv) > 2 else "noise"
lib = _native.load_from(ROOT / "structuredetector_b200/csrc/exp/lib_trace.so")
cfg = CONFIGS["cfg5"]; dev = torch.device("cuda:0")
uniq = make_raw(cfg, mode, batch=32).to(dev)
raw = uniq[torch.arange(images, device=dev) % 32].contiguous()
o = split_outputs(raw, 2, 1)
plan = ops.DecodePlan(dev, images, 2, 1, 512, 612, 100, 100, lib=lib)
call = (o["anchor_hm"], o["part_hm"], o["offsets"], o["embeddings"], float(np.float32(0.4)), float(np.float32(51.2)))
for _ in range(3): plan.run(*call)
torch.cuda.synchronize()
ms = plan.run_timed(*call)
sched = plan.schedule(*call[:4])
nw = sched["ctas"] * sched["warps_per_cta"]
buf = (ctypes.c_ulonglong * (4 * nw))()
lib.sdnet_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert lib.sdnet_debug_trace(buf, nw) == 0
t = np.frombuffer(buf, dtype=np.uint64).reshape(nw, 4).astype(np.int64)
t0 = t[:, 0].min()
start, first, end = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3, (t[:, 2] - t0) / 1e3
groups, units = t[:, 3] >> 32, t[:, 3] & 0xffffffff
q = lambda a: [round(float(x), 1) for x in np.percentile(a, [0, 5, 50, 95, 100])]
print(f"{images} images {mode}: kernel {ms[0]*1e3:.1f} us (events), warps {nw}, sched {sched['units']} units chunk {sched['chunk_groups']}")
print("start  us  p0/5/50/95/100", q(start))
print("first tile landed       ", q(first - start), "(after start)")
print("end    us               ", q(end))
print("busy   us (end - start) ", q(end - start))
print("groups per warp         ", q(groups), "units", q(units))
print("us per group            ", q((end - first) / np.maximum(groups, 1)))
print(f"mean busy {np.mean(end-start):.1f} us vs span {end.max():.1f} us -> balance {np.mean(end-start)/end.max():.3f}")
