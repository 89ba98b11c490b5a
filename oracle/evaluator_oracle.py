"""CPU restatement of the reference evaluator's location matching -- TEST INFRASTRUCTURE ONLY
(imported by tests/ and nothing else; the product path is structuredetector_b200/evaluator.py on the GPU).

Follows src/sdnet/model/evaluator.py:244-284 (``eval_anchor``) and :286-334 (``eval_part``) on plain
tuples, in Python floats like the reference.  Pinned against tests/golden/eval.json, which
tests/golden/make_golden_eval.py produced by executing the reference's own ``Evaluator.accumulate``.
"""
from __future__ import annotations

import sys

import numpy as np


def _match(preds, gts, dist_thresh, norm):
    """preds: [(x, y, score)] of one label; gts: [(x, y)] of that label, annotation order.
    evaluator.py:262-282: stable sort by score (descending), nearest ground truth with a strict '<',
    true positive if under the threshold and not yet visited."""
    order = sorted(range(len(preds)), key=lambda i: preds[i][2], reverse=True)
    visited = [False] * len(gts)
    tp, acc = 0, []
    for i in order:
        px, py, _ = preds[i]
        min_dist, j_min = sys.float_info.max, None
        for j, (gx, gy) in enumerate(gts):
            dist = np.hypot(px - gx, py - gy)  # utils.py:31-32
            if dist < min_dist:
                min_dist, j_min = dist, j
        if min_dist < dist_thresh and not visited[j_min]:
            visited[j_min] = True
            tp += 1
            acc.append(float(min_dist / norm))
    return tp, acc


def evaluate_image(objects, raw_parts, gt_objects, img_size, net_size, dist_threshold, labels, kp_labels):
    """objects: [(name, x, y, score)] predicted anchors (decoder frame = network-input pixels);
    raw_parts: [(kind, x, y, score)]; gt_objects: [(name, x, y, [(kind, x, y), ...])] in the same frame.
    Returns {"anchor": {label: (tp, npos, ndet, acc)}, "part": {...}} for this image."""
    rx, ry = img_size[0] / net_size[0], img_size[1] / net_size[1]  # resized(): utils.py:19-26
    dist_thresh = min(img_size) * dist_threshold  # evaluator.py:250
    norm = min(img_size)
    out = {"anchor": {}, "part": {}}
    for label in labels:
        preds = [(x * rx, y * ry, s) for name, x, y, s in objects if name == label]
        gts = [(x * rx, y * ry) for name, x, y, _ in gt_objects if name == label]
        tp, acc = _match(preds, gts, dist_thresh, norm)
        out["anchor"][label] = (tp, len(gts), len(preds), acc)
    gt_parts = [kp for _, _, _, kps in gt_objects for kp in kps]
    for label in kp_labels:
        preds = [(x * rx, y * ry, s) for kind, x, y, s in raw_parts if kind == label]
        gts = [(x * rx, y * ry) for kind, x, y in gt_parts if kind == label]
        tp, acc = _match(preds, gts, dist_thresh, norm)
        out["part"][label] = (tp, len(gts), len(preds), acc)
    return out


def evaluate_batch(annotation_plain, raw_parts_plain, eval_case, labels, kp_labels):
    """Accumulate a whole golden case: `annotation_plain` / `raw_parts_plain` as stored in
    tests/golden/index.json, `eval_case` one entry of tests/golden/eval.json."""
    total = {"anchor": {l: [0, 0, 0, []] for l in labels}, "part": {l: [0, 0, 0, []] for l in kp_labels}}
    net_size = (eval_case["width"], eval_case["height"])
    for image, objs, parts in zip(eval_case["images"], annotation_plain, raw_parts_plain):
        objects = [(name, anchor[1], anchor[2], anchor[3]) for name, anchor, _ in objs]
        gts = [(name, x, y, [tuple(kp) for kp in kps]) for name, x, y, kps in image["gt"]]
        res = evaluate_image(objects, [tuple(p) for p in parts], gts, tuple(image["img_size"]), net_size,
                             eval_case["dist_threshold"], labels, kp_labels)
        for key in ("anchor", "part"):
            for label, (tp, npos, ndet, acc) in res[key].items():
                t = total[key][label]
                t[0] += tp
                t[1] += npos
                t[2] += ndet
                t[3] += acc
    return total
