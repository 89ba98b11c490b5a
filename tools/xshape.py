"""Time the three decode kernels on noise of an arbitrary shape (kernel experiments).
usage: python tools/xshape.py B C_total H W [reps]   (2 anchor classes, C_total-6 part kinds... see below)"""
import sys
import torch
sys.path.insert(0, ".")
from structuredetector_b200 import ops

B, H, W = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
M, N, K, P = 2, 1, 100, 100
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
nu = min(B, 32)
raw = (torch.randn(nu, M + N + 4, H, W, device=dev, generator=g) * 2.0 - 3.0)
raw = raw.repeat((B + nu - 1) // nu, 1, 1, 1)[:B].contiguous()
a, p_, o, e = raw[:, :M], raw[:, M:M + N], raw[:, M + N:M + N + 2], raw[:, M + N + 2:]
plan = ops.DecodePlan(dev, B, M, N, H, W, K, P, torch.float32)
ts = []
for _ in range(reps + 3):
    ts.append(plan.run_timed(a, p_, o, e, 0.4, 0.1 * min(H, W)))
ts = ts[3:]
pk = sum(t[0] for t in ts) / len(ts)
gb = B * (M + N) * H * W * 4 / 1e9
print(f"B={B} H={H} W={W}: peaks {pk:.4f} ms  {gb/pk:.0f} GB/s  tail {sum(t[2] for t in ts)/len(ts):.4f}")
