"""CPU restatement of the reference evaluator's matching -- TEST INFRASTRUCTURE ONLY
(imported by tests/ and nothing else; the product path is structuredetector_b200/evaluator.py on the GPU).

Follows src/sdnet/model/evaluator.py:244-284 (``eval_anchor``), :286-334 (``eval_part``), :380-420 + :539-581
(``eval_csi`` / ``compute_csi``) and :429-474 (``eval_classif``) on plain tuples, in Python floats like the
reference.  Pinned against tests/golden/eval.json and tests/golden/eval_objects.json, which
tests/golden/make_golden_eval.py and make_golden_eval_objects.py produced by executing the reference's own
``Evaluator.accumulate``.
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


# ---------------------------------------------------------------------------------------------------------------
# object-level metrics: predicted objects = (name, x, y, score, [(kind, x, y, score)...]) in score order, ground truth
# = (name, x, y, [(kind, x, y)...]) in annotation order; every coordinate already in the evaluator's frame.
CLASSIFICATION_LABELS = [f"bean_{i}" for i in range(10)] + [f"maize_{i}" for i in range(10)]  # evaluator.py:422-427


def pair_csi(pred, gt, dist_thresh):
    """compute_csi (evaluator.py:539-581): critical success index of one prediction / ground-truth pair."""
    if pred[0] != gt[0]:
        return 0.0
    tp = int(np.hypot(pred[1] - gt[1], pred[2] - gt[2]) < dist_thresh)
    npos, ndet = 1 + len(gt[3]), 1 + len(pred[4])
    for kind in {kp[0] for kp in gt[3]} | {kp[0] for kp in pred[4]}:
        preds = sorted((kp for kp in pred[4] if kp[0] == kind), key=lambda kp: kp[3], reverse=True)  # stable
        gts = [kp for kp in gt[3] if kp[0] == kind]
        visited = [False] * len(gts)
        for kp in preds:
            min_dist, j_min = sys.float_info.max, None
            for j, target in enumerate(gts):
                dist = np.hypot(kp[1] - target[1], kp[2] - target[2])
                if dist < min_dist:
                    min_dist, j_min = dist, j
            if min_dist < dist_thresh and not visited[j_min]:
                visited[j_min] = True
                tp += 1
    denominator = npos + ndet - tp
    return tp / denominator if denominator != 0 else 1


def eval_csi_image(preds, gts, labels, dist_thresh, csi_threshold):
    """eval_csi (evaluator.py:380-420) -> {label: (tp, npos, ndet, acc)}."""
    out = {}
    for label in labels:
        p_lab = sorted((p for p in preds if p[0] == label), key=lambda p: p[3], reverse=True)
        g_lab = [g for g in gts if g[0] == label]
        visited = [False] * len(g_lab)
        tp, acc = 0, []
        for pred in p_lab:
            best, idx = 0.0, None
            for j, gt in enumerate(g_lab):
                csi = pair_csi(pred, gt, dist_thresh)
                if csi > best:
                    best, idx = csi, j
            if idx is not None and best >= csi_threshold and not visited[idx]:
                visited[idx] = True
                tp += 1
                acc.append(float(best))
        out[label] = (tp, len(g_lab), len(p_lab), acc)
    return out


def eval_classif_image(preds, gts, dist_thresh, norm):
    """eval_classif (evaluator.py:429-474): objects keyed by "<name>_<number of parts>" -> {label: (tp, npos, ndet, acc)}."""
    out = {}
    for label in CLASSIFICATION_LABELS:
        p_lab = sorted((p for p in preds if f"{p[0]}_{len(p[4])}" == label), key=lambda p: p[3], reverse=True)
        g_lab = [g for g in gts if f"{g[0]}_{len(g[3])}" == label]
        visited = [False] * len(g_lab)
        tp, acc = 0, []
        for pred in p_lab:
            best, idx = sys.float_info.max, None
            for j, gt in enumerate(g_lab):
                dist = np.hypot(pred[1] - gt[1], pred[2] - gt[2])
                if dist < best:
                    best, idx = dist, j
            if idx is not None and best <= dist_thresh and not visited[idx]:
                visited[idx] = True
                tp += 1
                acc.append(float(best / norm))
        out[label] = (tp, len(g_lab), len(p_lab), acc)
    return out


def evaluate_objects_batch(annotation_plain, eval_case, labels, rename=None):
    """Accumulate eval_csi + eval_classif over a golden case.  `annotation_plain` as stored in tests/golden/index.json
    (decoder frame), `eval_case` one entry of tests/golden/eval_objects.json; `rename` maps the stored label names to
    the names the case was evaluated under (e.g. label0 -> bean)."""
    rename = rename or {}
    net_w, net_h = eval_case["width"], eval_case["height"]
    csi = {rename.get(l, l): [0, 0, 0, []] for l in labels}
    classif = {l: [0, 0, 0, []] for l in CLASSIFICATION_LABELS}
    for image, objs in zip(eval_case["images"], annotation_plain):
        img_w, img_h = image["img_size"]
        rx, ry = img_w / net_w, img_h / net_h
        thresh, norm = min(img_w, img_h) * eval_case["dist_threshold"], min(img_w, img_h)
        preds = [(rename.get(name, name), a[1] * rx, a[2] * ry, a[3], [(k[0], k[1] * rx, k[2] * ry, k[3]) for k in kps])
                 for name, a, kps in objs]
        gts = [(rename.get(name, name), x * rx, y * ry, [(k, px * rx, py * ry) for k, px, py in kps]) for name, x, y, kps in image["gt"]]
        for total, res in ((csi, eval_csi_image(preds, gts, list(csi), thresh, eval_case["csi_threshold"])),
                           (classif, eval_classif_image(preds, gts, thresh, norm))):
            for label, (tp, npos, ndet, acc) in res.items():
                t = total[label]
                t[0] += tp
                t[1] += npos
                t[2] += ndet
                t[3] += acc
    return {"csi": csi, "classification": classif}
