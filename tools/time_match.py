"""Time the evaluator path on the device: decode cfg5-shaped noise, then match against seeded ground truth.
usage: python tools/time_match.py [images]"""
import sys, time
from types import SimpleNamespace
import numpy as np, torch
sys.path.insert(0, ".")
from structuredetector_b200 import ImageAnnotation, Keypoint, Object, ops
from structuredetector_b200.evaluator import Evaluator
from structuredetector_b200.synth import CONFIGS, make_raw, split_outputs

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = CONFIGS["cfg5"]
dev = torch.device("cuda:0")
raw = make_raw(cfg, "blobs", batch=32).to(dev)
raw = raw.repeat(B // 32, 1, 1, 1)
outs = split_outputs(raw, cfg.labels, cfg.parts)
packed = ops.decode_packed(outs, cfg.max_objects, cfg.max_parts, cfg.conf_threshold, cfg.dist_thresh)
rng = np.random.default_rng(0)
a = packed.anchor_out.cpu().numpy()
anns = []
for b in range(B):
    objs = [Object(f"label{int(c)}", Keypoint("a", 4 * x + rng.normal(0, 4), 4 * y + rng.normal(0, 4)),
                   [Keypoint("part0", 4 * x + rng.normal(0, 30), 4 * y + rng.normal(0, 30)) for _ in range(3)])
            for x, y, s, c in a[b] if s > 0.4]
    anns.append(ImageAnnotation("gt", objs, img_size=(2448, 2048)))
args = SimpleNamespace(labels={"label0": 0, "label1": 1}, parts={"part0": 0}, width=2448, height=2048, dist_threshold=0.05,
                       conf_threshold=0.4, down_ratio=4.0)
ev = Evaluator(args)
ev.accumulate_packed(packed, anns, (cfg.width, cfg.height))
torch.cuda.synchronize()
t = time.perf_counter(); ev.reset(); ev.accumulate_packed(packed, anns, (cfg.width, cfg.height)); torch.cuda.synchronize(); dt = time.perf_counter() - t
print(f"accumulate_packed: {B} images in {dt*1e3:.2f} ms ({B/dt:.0f} img/s), gt objects/img {np.mean([len(x.objects) for x in anns]):.1f}")
print(ev.anchor_eval.reduce(), "|", ev.part_eval.reduce())
# kernel alone
import ctypes
from structuredetector_b200 import _native
(gt_a, n_a, wa), (gt_p, n_p, wp), scale = ev._pack_ground_truth(anns, dev)
M, N, K, P = 2, 1, cfg.max_objects, cfg.max_parts
bufs = [torch.empty(B, M, 3, dtype=torch.int32, device=dev), torch.empty(B, N, 3, dtype=torch.int32, device=dev),
        torch.empty(B, K, dtype=torch.float64, device=dev), torch.empty(B, P, dtype=torch.float64, device=dev)]
prm = _native.SdnetMatchParams(); prm.struct_size = ctypes.sizeof(prm)
prm.B, prm.M, prm.N, prm.K, prm.P, prm.max_gt_anchors, prm.max_gt_parts = B, M, N, K, P, wa, wp
prm.conf, prm.sx, prm.sy = 0.4, 4.0, 4.0
prm.anchor_out, prm.part_out, prm.image_scale = packed.anchor_out.data_ptr(), packed.part_out.data_ptr(), scale.data_ptr()
prm.gt_anchors, prm.n_gt_anchors, prm.gt_parts, prm.n_gt_parts = gt_a.data_ptr(), n_a.data_ptr(), gt_p.data_ptr(), n_p.data_ptr()
prm.anchor_stats, prm.part_stats, prm.anchor_acc, prm.part_acc = (x.data_ptr() for x in bufs)
lib = _native.load(); st = torch.cuda.current_stream().cuda_stream
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3): lib.sdnet_match_launch(ctypes.byref(prm), ctypes.c_void_p(st))
e0.record()
for _ in range(20): lib.sdnet_match_launch(ctypes.byref(prm), ctypes.c_void_p(st))
e1.record(); torch.cuda.synchronize()
print(f"sdnet_match_kernel: {e0.elapsed_time(e1)/20*1e3:.1f} us for {B} images")
# the reference-style Python loop on the same data (our oracle restatement), 8 images
from oracle import evaluator_oracle as EO
p_out = packed.part_out.cpu().numpy()
t = time.perf_counter()
for b in range(8):
    objs = [(f"label{int(c)}", 4.0 * x, 4.0 * y, float(s)) for x, y, s, c in a[b] if s > 0.4]
    parts = [("part0", 4.0 * r[0], 4.0 * r[1], float(r[2])) for r in p_out[b] if not r[2] < 0.4]
    gts = [(o.name, o.anchor.x, o.anchor.y, [(k.kind, k.x, k.y) for k in o.parts]) for o in anns[b].objects]
    EO.evaluate_image(objs, parts, gts, (2448, 2048), (2448, 2048), 0.05, ["label0", "label1"], ["part0"])
print(f"python loop (restatement of evaluator.py:244-334): {8/(time.perf_counter()-t):.0f} img/s")
