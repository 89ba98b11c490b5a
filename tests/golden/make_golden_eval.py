"""Golden vectors for the evaluator matching, made by EXECUTING the unmodified reference:
its `Decoder` on the stored raw inputs, then its `Evaluator.accumulate(prediction, annotation, raw_parts)`
(reference: src/sdnet/model/evaluator.py:225-334) against seeded ground truth.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_eval.py      -> tests/golden/eval.json

Ground truth per image: ~80 % of the predicted objects with jittered anchors and ~80 % of their parts
(jittered), some parts re-assigned, plus spurious objects; image sizes with different x / y ratios so
the resize factors differ.  Runs only where /root/reference exists.
"""
from __future__ import annotations

import json
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference/src")

from sdnet.data.decoders import Decoder  # noqa: E402  (the reference itself)
from sdnet.model.evaluator import Evaluator  # noqa: E402
from sdnet.utils import ImageAnnotation, Keypoint, Object  # noqa: E402

from structuredetector_b200.synth import split_outputs  # noqa: E402

CASES = ["ladder_small", "ladder_wide", "blobs_cfg1", "noise_defaults", "noise_multiclass"]


def main():
    index = json.loads((HERE / "index.json").read_text())
    out = {}
    for ci, name in enumerate(CASES):
        meta = index[name]
        b, m, n, h, w = meta["shape"]
        raw = torch.from_numpy(np.load(HERE / f"{name}.npz")["raw"])
        labels = {f"label{i}": i for i in range(m)}
        parts = {f"part{i}": i for i in range(n)}
        args = SimpleNamespace(
            _r_labels={i: k for k, i in labels.items()}, _r_parts={i: k for k, i in parts.items()}, labels=labels, parts=parts,
            anchor_name="stem", down_ratio=4.0, max_objects=meta["K"], max_parts=meta["P"], conf_threshold=meta["conf"],
            decoder_dist_thresh=meta["dist"], width=4 * w, height=4 * h, dist_threshold=0.05, csi_threshold=0.75)
        data = Decoder(args)({k: v.clone() for k, v in split_outputs(raw, m, n).items()}, return_metadata=True)
        rng = np.random.default_rng(7000 + ci)
        evaluator = Evaluator(args)
        images = []
        for bi in range(b):
            pred, raw_parts = data["annotation"][bi], data["raw_parts"][bi]
            img_size = (3 * args.width + 7 + bi, 2 * args.height + 3)
            objects = []
            for obj in pred.objects:
                if rng.random() < 0.8:
                    kps = [Keypoint(p.kind, p.x + rng.normal(0, 5), p.y + rng.normal(0, 5)) for p in obj.parts
                           if rng.random() < 0.8]
                    objects.append(Object(obj.name, Keypoint("stem", obj.anchor.x + rng.normal(0, 6),
                                                             obj.anchor.y + rng.normal(0, 6)), kps))
            for kp in raw_parts[:: 3]:  # parts the grouping may have dropped, attached to a spurious object
                if rng.random() < 0.3:
                    objects.append(Object(f"label{int(rng.integers(0, m))}",
                                          Keypoint("stem", float(rng.uniform(0, args.width)), float(rng.uniform(0, args.height))),
                                          [Keypoint(kp.kind, kp.x + rng.normal(0, 4), kp.y + rng.normal(0, 4))]))
            if objects and rng.random() < 0.5:  # a duplicated ground truth: two targets at the same distance
                o = objects[0]
                objects.append(Object(o.name, Keypoint("stem", o.anchor.x, o.anchor.y), []))
            gt = ImageAnnotation(f"batch_{bi}", objects, img_size=img_size)
            before = {k: (v.tp, len(v.acc)) for ev in (evaluator.anchor_eval, evaluator.part_eval) for k, v in ev.items()}
            evaluator.accumulate(pred, gt, raw_parts)
            images.append({
                "img_size": list(img_size),
                "gt": [[o.name, o.anchor.x, o.anchor.y, [[p.kind, p.x, p.y] for p in o.parts]] for o in objects],
            })
        result = {}
        for key, evals in (("anchor", evaluator.anchor_eval), ("part", evaluator.part_eval)):
            result[key] = {label: {"tp": int(e.tp), "npos": int(e.npos), "ndet": int(e.ndet), "acc": [float(a) for a in e.acc]}
                           for label, e in evals.items()}
        out[name] = {"width": args.width, "height": args.height, "dist_threshold": args.dist_threshold, "images": images,
                     "result": result, "torch": torch.__version__}
        print(name, {k: {l: (r["tp"], r["npos"], r["ndet"]) for l, r in v.items()} for k, v in result.items()})
    if "--check" in sys.argv:
        committed = json.loads((HERE / "eval.json").read_text())
        for name in CASES:
            assert committed[name]["result"] == out[name]["result"], f"{name}: the live reference disagrees with the fixture"
            assert committed[name]["images"] == out[name]["images"], f"{name}: ground truth drifted"
        print("fixture matches the live reference")
        return
    (HERE / "eval.json").write_text(json.dumps(out))


if __name__ == "__main__":
    main()
