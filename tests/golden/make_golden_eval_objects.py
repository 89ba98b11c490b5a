"""Golden vectors for the OBJECT-level evaluator metrics, made by EXECUTING the unmodified reference: its `Decoder` on
the stored raw inputs, then `Evaluator.accumulate(prediction, annotation, raw_parts, True, True)` -- what `evaluate`
calls (reference: src/sdnet/cli/evaluate.py:43-45; src/sdnet/model/evaluator.py:380-474, 539-581) -- against the seeded
ground truth of tests/golden/eval.json.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_eval_objects.py   -> tests/golden/eval_objects.json

The reference's classification labels are hard-coded to "bean_<n>" / "maize_<n>", so the two-label cases are evaluated
with label0 -> bean, label1 -> maize (stored as `rename`); the multi-class case keeps its names (its classification table is
all zeros, as the reference's would be).  Runs only where /root/reference exists.
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


def main():
    index = json.loads((HERE / "index.json").read_text())
    cases = json.loads((HERE / "eval.json").read_text())
    out = {}
    for name, case in cases.items():
        meta = index[name]
        b, m, n, h, w = meta["shape"]
        rename = {"label0": "bean", "label1": "maize"} if m == 2 else {}
        raw = torch.from_numpy(np.load(HERE / f"{name}.npz")["raw"])
        labels = {rename.get(f"label{i}", f"label{i}"): i for i in range(m)}
        parts = {f"part{i}": i for i in range(n)}
        args = SimpleNamespace(
            _r_labels={i: k for k, i in labels.items()}, _r_parts={i: k for k, i in parts.items()}, labels=labels, parts=parts,
            anchor_name="stem", down_ratio=4.0, max_objects=meta["K"], max_parts=meta["P"], conf_threshold=meta["conf"],
            decoder_dist_thresh=meta["dist"], width=case["width"], height=case["height"], dist_threshold=case["dist_threshold"],
            csi_threshold=0.5)
        data = Decoder(args)({k: v.clone() for k, v in split_outputs(raw, m, n).items()}, return_metadata=True)
        evaluator = Evaluator(args)
        for bi, image in enumerate(case["images"]):
            objects = [Object(rename.get(o[0], o[0]), Keypoint("stem", o[1], o[2]), [Keypoint(k, x, y) for k, x, y in o[3]])
                       for o in image["gt"]]
            gt = ImageAnnotation(f"batch_{bi}", objects, img_size=tuple(image["img_size"]))
            evaluator.accumulate(data["annotation"][bi], gt, data["raw_parts"][bi], True, True)
        result = {}
        for key, evals in (("csi", evaluator.csi_eval), ("classification", evaluator.classification_eval)):
            result[key] = {label: {"tp": int(e.tp), "npos": int(e.npos), "ndet": int(e.ndet), "acc": [float(a) for a in e.acc]}
                           for label, e in evals.items()}
        out[name] = {"rename": rename, "csi_threshold": args.csi_threshold, "result": result, "torch": torch.__version__}
        print(name, {k: {l: (r["tp"], r["npos"], r["ndet"]) for l, r in v.items() if r["npos"] or r["ndet"]} for k, v in result.items()})
    if "--check" in sys.argv:
        committed = json.loads((HERE / "eval_objects.json").read_text())
        for name in out:
            assert committed[name]["result"] == out[name]["result"], f"{name}: the live reference disagrees with the fixture"
        print("fixture matches the live reference")
        return
    (HERE / "eval_objects.json").write_text(json.dumps(out))


if __name__ == "__main__":
    main()
