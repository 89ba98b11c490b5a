"""Small-batch decode rate and latency (cfg1: 1 image, cfg2: 64 images of 128x128 maps): plain launches vs a CUDA graph replay.
usage: python tools/latency.py"""
import sys, time
import torch
sys.path.insert(0, ".")
from structuredetector_b200 import ops
from structuredetector_b200.synth import CONFIGS, make_raw, split_outputs

dev = torch.device("cuda:0")
for name in ("cfg1", "cfg2", "cfg3"):
    cfg = CONFIGS[name]
    raw = make_raw(cfg, "blobs").to(dev)
    o = split_outputs(raw, cfg.labels, cfg.parts)
    plan = ops.DecodePlan(dev, cfg.batch, cfg.labels, cfg.parts, cfg.height, cfg.width, cfg.max_objects, cfg.max_parts)
    conf, dist = ops._f32(cfg.conf_threshold), ops._f32(cfg.dist_thresh * min(cfg.width, cfg.height))
    run = lambda: plan.run(o["anchor_hm"], o["part_hm"], o["offsets"], o["embeddings"], conf, dist)
    for _ in range(20): run()
    torch.cuda.synchronize()
    n = 2000
    t = time.perf_counter()
    for _ in range(n): run()
    t_issue = time.perf_counter() - t
    torch.cuda.synchronize(); t_all = time.perf_counter() - t
    # one-at-a-time latency
    t = time.perf_counter()
    for _ in range(200):
        run(); torch.cuda.synchronize()
    lat = (time.perf_counter() - t) / 200
    want = [t.clone() for t in (plan.out.anchor_inds, plan.out.part_inds, plan.out.anchor_out, plan.out.part_out, plan.out.assign)]
    line = f"{name}: B={cfg.batch} back-to-back {t_all/n*1e6:.1f} us/decode (host issue {t_issue/n*1e6:.1f} us), synchronous latency {lat*1e6:.1f} us"
    try:
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            run(); torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                plan.run(o["anchor_hm"], o["part_hm"], o["offsets"], o["embeddings"], conf, dist, stream=s)
        plan.out.blob.zero_()
        g.replay(); torch.cuda.synchronize()
        ok = all(torch.equal(a, b) for a, b in zip(want, (plan.out.anchor_inds, plan.out.part_inds, plan.out.anchor_out, plan.out.part_out, plan.out.assign)))
        t = time.perf_counter()
        for _ in range(n): g.replay()
        torch.cuda.synchronize(); tg = (time.perf_counter() - t) / n
        t = time.perf_counter()
        for _ in range(200):
            g.replay(); torch.cuda.synchronize()
        lg = (time.perf_counter() - t) / 200
        line += f" | graph replay {tg*1e6:.1f} us/decode, latency {lg*1e6:.1f} us, same result {ok}"
    except Exception as exc:  # noqa: BLE001
        line += f" | graph capture failed: {exc!r}"[:300]
    print(line, flush=True)
