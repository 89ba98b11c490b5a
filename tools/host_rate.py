"""Is a small shard's step bound by the host's launch rate?  Host issue time per step against device time per step for
plain launches, and the same decodes replayed from CUDA graphs (one graph per in-flight slot, replayed round-robin on the
slots' streams).  usage: python tools/host_rate.py [images] [depth]   (under torchrun: the fused gather plans)"""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch, torch.distributed as dist
from structuredetector_b200 import ops
from structuredetector_b200.synth import CONFIGS, make_raw, split_outputs

images = int(sys.argv[1]) if len(sys.argv) > 1 else 128
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 6
world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
cfg = CONFIGS["cfg5"]
for mode in ("noise", "blobs"):
    uniq = make_raw(cfg, mode, batch=32).to(dev)
    raw = uniq[(torch.arange(images, device=dev) + rank * images) % 32].contiguous()
    o = split_outputs(raw, 2, 1)
    call = (o["anchor_hm"], o["part_hm"], o["offsets"], o["embeddings"], float(np.float32(0.4)), float(np.float32(51.2)))
    if world > 1:
        from structuredetector_b200.parallel import FusedGatherPlan
        mk = lambda _i: FusedGatherPlan(dev, images * world, 2, 1, 512, 612, 100, 100)
    else:
        mk = lambda _i: ops.DecodePlan(dev, images, 2, 1, 512, 612, 100, 100)
    pipe = ops.DecodePipeline(dev, depth, mk)

    def fence():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier(); torch.cuda.synchronize(dev)

    def timed(fn, n=400):
        for _ in range(30): fn()
        fence()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); t = time.perf_counter()
        for _ in range(n): fn()
        host = time.perf_counter() - t
        pipe.drain(); e1.record(); fence()
        return host / n * 1e6, e0.elapsed_time(e1) / n * 1e3

    h, d = timed(lambda: pipe.submit(*call))
    line = f"[{rank}] {mode} {images} img x{world} depth {depth}: plain host {h:.1f} us/step, device {d:.1f} us/step"
    try:
        graphs = []
        for k in range(depth):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=pipe.streams[k]):
                pipe.plans[k].run(*call, stream=pipe.streams[k])
            graphs.append(g)
        fence()
        nxt = [0]
        def replay():
            k = nxt[0]; nxt[0] = (k + 1) % depth
            with torch.cuda.stream(pipe.streams[k]):
                graphs[k].replay()
        h, d = timed(replay)
        line += f" | graphs host {h:.1f} us/step, device {d:.1f} us/step"
    except Exception as exc:  # noqa: BLE001
        line += f" | graph capture failed: {exc!r}"[:400]
    print(line, flush=True)
    del pipe
if world > 1:
    dist.destroy_process_group()
