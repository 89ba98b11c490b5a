"""Location metrics of the reference evaluator, computed on the device from the packed detections.

Mirrors ``Evaluation`` / ``Evaluations`` / ``Evaluator`` of the reference (reference:
src/sdnet/model/evaluator.py:13-120, 123-206, 209-334): same attribute and property names, same
formulas, so tables and CSV code written against them keep working.  What moves to the GPU is the
matching of ``eval_anchor`` (:244-284) and ``eval_part`` (:286-334) -- per image and label, detections in
score order against their nearest ground truth -- through ``sdnet_match_launch`` on the tensors
``sdnet_decode_launch`` wrote; no Python object is built for a prediction.  The CSI and classification
tables (``eval_csi``, ``eval_classif``) are not covered; the objects our ``Decoder`` returns feed the
reference's own ``Evaluator`` unchanged for those.
"""
from __future__ import annotations

import ctypes
from functools import reduce

import numpy as np
import torch

from . import _native, ops

__all__ = ["Evaluation", "Evaluations", "Evaluator"]


class Evaluation:
    """tp / npos / ndet counters plus the localisation errors of the true positives."""

    def __init__(self, tp=0, npos=0, ndet=0, acc=None, counts=None):
        assert tp >= 0 and ndet >= 0 and npos >= 0, "tp, npos and ndet should be positive"
        assert tp <= ndet, "tp must be lower than or equal to ndet"
        assert tp <= npos, "tp must be lower than or equal to npos"
        self.tp, self.npos, self.ndet = tp, npos, ndet
        self.acc = acc or []
        self.count_errors = counts or []

    def reset(self):
        self.__init__()

    def __iadd__(self, other):
        self.tp += other.tp
        self.npos += other.npos
        self.ndet += other.ndet
        self.acc += other.acc
        self.count_errors += other.count_errors
        return self

    def __add__(self, other):
        total = Evaluation(self.tp, self.npos, self.ndet, list(self.acc), list(self.count_errors))
        total += other
        return total

    fp = property(lambda self: self.ndet - self.tp)
    fn = property(lambda self: self.npos - self.tp)

    @property
    def csi(self):
        union = self.npos + self.ndet - self.tp
        return self.tp / union if union != 0 else 1

    @property
    def precision(self):
        return self.tp / self.ndet if self.ndet != 0 else 1 if self.npos == 0 else 0

    @property
    def recall(self):
        return self.tp / self.npos if self.npos != 0 else 1 if self.ndet == 0 else 0

    @property
    def f1_score(self):
        total = self.npos + self.ndet
        return 2 * self.tp / total if total != 0 else 1

    @property
    def avg_acc(self):
        return np.mean(self.acc) if len(self.acc) != 0 else float("nan")

    @property
    def acc_err(self):
        return np.std(self.acc) / np.sqrt(len(self.acc)) if len(self.acc) != 0 else float("nan")

    def stats(self):
        return (f"{self.npos}", f"{self.ndet}", f"{self.recall:.2%}", f"{self.precision:.2%}", f"{self.f1_score:.2%}",
                f"{self.avg_acc:.4%}", f"{self.acc_err:.4%}")

    def __repr__(self):
        return (f"f1: {self.f1_score:.2%}, rec: {self.recall:.2%}, prec: {self.precision:.2%}, npos: {self.npos}, "
                f"ndet: {self.ndet}, tp/fp/fn: {self.tp}/{self.fp}/{self.fn}, avg_acc: {self.avg_acc:.2}")


class Evaluations:
    """One ``Evaluation`` per label."""

    def __init__(self, labels=None):
        self.evals = {label: Evaluation() for label in labels} if labels else {}

    def reset(self):
        for evaluation in self.evals.values():
            evaluation.reset()

    labels = property(lambda self: self.evals.keys())

    def items(self):
        return self.evals.items()

    def __getitem__(self, label):
        return self.evals[label]

    def __setitem__(self, label, evaluation):
        self.evals[label] = evaluation

    def __len__(self):
        return len(self.evals)

    def __iadd__(self, other):
        assert self.labels == other.labels, "The Evaluations should have the same labels"
        for label, evaluation in other.items():
            self.evals[label] += evaluation
        return self

    def __add__(self, other):
        assert self.labels == other.labels, "The Evaluations should have the same labels"
        total = Evaluations()
        total.evals = {label: self.evals[label] + evaluation for label, evaluation in other.items()}
        return total

    def __or__(self, other):
        merged = Evaluations()
        merged.evals = {label: self[label] + other[label] for label in self.labels & other.labels}
        merged.evals.update({label: self[label] for label in self.labels - other.labels})
        merged.evals.update({label: other[label] for label in other.labels - self.labels})
        return merged

    def reduce(self):
        return reduce(Evaluation.__iadd__, self.evals.values(), Evaluation())

    def __repr__(self):
        lines = [f"total: {self.reduce()}"] if len(self) > 1 else []
        return "\n".join(lines + [f"{label}: {evaluation}" for label, evaluation in self.items()])


class Evaluator:
    """``Evaluator(args)`` reads ``args.labels`` / ``args.parts`` (name -> class index, args.py),
    ``args.width`` / ``args.height`` (network input size), ``args.dist_threshold``, and for the decoder
    side ``args.conf_threshold`` and ``args.down_ratio`` -- the reference evaluator's fields
    (evaluator.py:210-213,245-250) plus the two the decoder applied before it."""

    def __init__(self, args):
        self.args = args
        self.labels = args.labels.keys()
        self.kp_labels = args.parts.keys()
        self._label_index = dict(args.labels)
        self._part_index = dict(args.parts)
        self._label_names = {index: name for name, index in args.labels.items()}
        self._part_names = {index: name for name, index in args.parts.items()}
        self.lib = _native.load()
        self.reset()

    def reset(self):
        self.anchor_eval = Evaluations(self.labels)
        self.part_eval = Evaluations(self.kp_labels)

    @property
    def kps_eval(self):
        return self.anchor_eval | self.part_eval

    # -- ground truth -> tensors -------------------------------------------------------------
    def _pack_ground_truth(self, annotations, device):
        rows_a = [[(obj.anchor.x, obj.anchor.y, self._label_index.get(obj.name, -1)) for obj in ann.objects]
                  for ann in annotations]
        rows_p = [[(kp.x, kp.y, self._part_index.get(kp.kind, -1)) for obj in ann.objects for kp in obj.parts]
                  for ann in annotations]

        def pack(rows):
            width = max(1, max((len(r) for r in rows), default=0))
            if width > _native.MAX_GT:
                raise ValueError(f"{width} ground-truth points in one image; the matcher takes at most {_native.MAX_GT}")
            data = np.zeros((len(rows), width, 3), dtype=np.float64)
            for b, row in enumerate(rows):
                if row:
                    data[b, : len(row)] = row
            counts = np.array([len(r) for r in rows], dtype=np.int32)
            return torch.from_numpy(data).to(device), torch.from_numpy(counts).to(device), width

        scale = np.empty((len(annotations), 4), dtype=np.float64)
        for b, ann in enumerate(annotations):
            img_w, img_h = ann.img_size
            scale[b] = (img_w / self.args.width, img_h / self.args.height, min(ann.img_size) * self.args.dist_threshold,
                        min(ann.img_size))
        return pack(rows_a), pack(rows_p), torch.from_numpy(scale).to(device)

    # -- the batched equivalent of accumulate(prediction, annotation, raw_parts) -------------------
    def accumulate_packed(self, packed: ops.PackedDetections, annotations, out_size, conf_thresh=None):
        """Add one decoded batch.  ``packed`` = ``ops.decode_packed(...)`` (device tensors),
        ``annotations`` = the batch's ground truth (``ImageAnnotation`` with ``img_size``, coordinates in the
        network-input frame, as the reference's dataset yields them), ``out_size`` = (W, H) of the heat maps.
        Equivalent to ``accumulate(prediction, annotation, raw_parts)`` of the reference per image."""
        conf = self.args.conf_threshold if conf_thresh is None else conf_thresh
        B, K = packed.anchor_inds.shape
        P = packed.part_inds.shape[1]
        assert len(annotations) == B
        device = packed.anchor_out.device
        M, N = len(self._label_index), len(self._part_index)
        (gt_a, n_a, wa), (gt_p, n_p, wp), scale = self._pack_ground_truth(annotations, device)
        a_stats = torch.empty(B, M, 3, dtype=torch.int32, device=device)
        p_stats = torch.empty(B, N, 3, dtype=torch.int32, device=device)
        a_acc = torch.empty(B, K, dtype=torch.float64, device=device)
        p_acc = torch.empty(B, P, dtype=torch.float64, device=device)
        out_w, out_h = out_size
        in_w, in_h = int(self.args.down_ratio * out_w), int(self.args.down_ratio * out_h)  # decoders.py:37-38
        prm = _native.SdnetMatchParams()
        prm.struct_size = ctypes.sizeof(_native.SdnetMatchParams)
        prm.B, prm.M, prm.N, prm.K, prm.P = B, M, N, K, P
        prm.max_gt_anchors, prm.max_gt_parts = wa, wp
        prm.conf, prm.sx, prm.sy = float(conf), in_w / out_w, in_h / out_h
        prm.anchor_out, prm.part_out = packed.anchor_out.data_ptr(), packed.part_out.data_ptr()
        prm.image_scale = scale.data_ptr()
        prm.gt_anchors, prm.n_gt_anchors = gt_a.data_ptr(), n_a.data_ptr()
        prm.gt_parts, prm.n_gt_parts = gt_p.data_ptr(), n_p.data_ptr()
        prm.anchor_stats, prm.part_stats = a_stats.data_ptr(), p_stats.data_ptr()
        prm.anchor_acc, prm.part_acc = a_acc.data_ptr(), p_acc.data_ptr()
        stream = torch.cuda.current_stream(device).cuda_stream
        _native.check(self.lib.sdnet_match_launch(ctypes.byref(prm), ctypes.c_void_p(stream)), "sdnet_match_launch")
        a_cls = packed.anchor_out[:, :, 3].to(torch.int32)
        p_cls = packed.part_out[:, :, 3].to(torch.int32)
        host = [t.cpu().numpy() for t in (a_stats, p_stats, a_acc, p_acc, a_cls, p_cls)]
        self._absorb(self.anchor_eval, self._label_names, host[0], host[2], host[4])
        self._absorb(self.part_eval, self._part_names, host[1], host[3], host[5])

    @staticmethod
    def _absorb(target: Evaluations, names, stats, acc, cls):
        totals = stats.sum(axis=0)  # (classes, 3): ndet, npos, tp
        for index, name in names.items():
            res = target[name]
            res.ndet += int(totals[index, 0])
            res.npos += int(totals[index, 1])
            res.tp += int(totals[index, 2])
            hit = (cls == index) & ~np.isnan(acc)
            res.acc += acc[hit].tolist()  # image-major, slot (= score) order: the order the reference appends in

    def pretty_print(self):
        for title, evals in (("Anchor Location", self.anchor_eval), ("Part Location", self.part_eval),
                             ("All Kps Location", self.kps_eval)):
            print(title)
            for label, evaluation in evals.items():
                print(" ", label, *evaluation.stats())
            if len(evals) > 1:
                print("  Total", *evals.reduce().stats())
