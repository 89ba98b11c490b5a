"""Time the dense sigmoid+NMS kernel (RawDecoder) on cfg5-shaped maps. usage: python tools/time_suppress.py [images]"""
import sys, torch
sys.path.insert(0, ".")
from structuredetector_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
x = torch.randn(B, 3, 512, 612, device="cuda") * 2 - 3
for dt in (torch.float32, torch.float16, torch.bfloat16):
    y = x.to(dt)
    ops.suppress_maps(y); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops._suppress_op(y, 2)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gb = y.numel() * (y.element_size() + 4) / 1e9
    print(dt, f"{ms:.3f} ms for {B} images, {gb/ms*1e3:.0f} GB/s (read + fp32 write), {B/ms*1e3:.0f} img/s")
