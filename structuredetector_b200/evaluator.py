"""The reference evaluator's metrics, computed on the device from the packed detections.

Mirrors ``Evaluation`` / ``Evaluations`` / ``Evaluator`` of the reference (reference:
src/sdnet/model/evaluator.py:13-120, 123-206, 209-646): same attribute and property names, same
formulas, so tables and CSV code written against them keep working.  What moves to the GPU is the matching:
``eval_anchor`` (:244-284) and ``eval_part`` (:286-334) -- per image and label, detections in score order
against their nearest ground truth -- through ``sdnet_match_launch``; ``eval_csi`` with ``compute_csi``
(:380-420, 539-581) and ``eval_classif`` (:429-474) -- predicted objects (anchor + grouped parts) against
ground-truth objects -- through ``sdnet_match_objects_launch``; both straight on the tensors
``sdnet_decode_launch`` wrote: no Python object is built for a prediction.
"""
from __future__ import annotations

import ctypes
from functools import reduce
from pathlib import Path

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

    COLUMNS = ("Gts.", "Preds.", "Rec.", "Prec.", "F1 Score", "L. Acc.", "L. Err.")

    @staticmethod
    def columns():
        """Column headers of ``stats()`` (evaluator.py:88-98; ``rich`` Column objects when rich is installed)."""
        try:
            from rich.table import Column
        except ImportError:
            return Evaluation.COLUMNS
        return tuple(Column(name, justify="right", **({"style": "green"} if name == "F1 Score" else {}))
                     for name in Evaluation.COLUMNS)

    def pretty_print(self):
        single = Evaluations()
        single.evals = {"": self}
        _print_table(None, single)

    def save_conf_matrix(self):
        """One 10 x 10 (expected parts, predicted parts) count matrix per label, as ``conf_mat_<label>.npy``
        (evaluator.py:107-113)."""
        by_label = {}
        for label, predicted, expected in self.count_errors:
            by_label.setdefault(label, []).append((predicted, expected))
        for label, pairs in by_label.items():
            mat = np.zeros((10, 10))
            for predicted, expected in pairs:
                mat[expected, predicted] += 1
            np.save(f"conf_mat_{label}.npy", mat)

    def __repr__(self):
        return (f"f1: {self.f1_score:.2%}, rec: {self.recall:.2%}, prec: {self.precision:.2%}, npos: {self.npos}, "
                f"ndet: {self.ndet}, tp/fp/fn: {self.tp}/{self.fp}/{self.fn}, avg_acc: {self.avg_acc:.2}")


def _print_table(title, evaluations):
    """A ``rich`` table like the reference prints (evaluator.py:190-196, 583-604); plain text without rich."""
    rows = [(label, *evaluation.stats()) for label, evaluation in evaluations.items()]
    total = ("Total", *evaluations.reduce().stats()) if len(evaluations) > 1 else None
    try:
        from rich import print as rprint
        from rich.table import Column, Table
    except ImportError:
        if title:
            print(title)
        for row in rows + ([total] if total else []):
            print(" ", *row)
        return
    table = Table(Column("Label", style="bold"), *Evaluation.columns(), title=title)
    for row in rows:
        table.add_row(*row)
    if total:
        table.add_row(*total, style="bold")
    rprint(table)


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

    def __ior__(self, other):
        """In-place union (evaluator.py:180-185; the reference's version recurses into itself -- this does what it
        is written to mean): labels only ``other`` has are adopted, shared labels are summed."""
        for label, evaluation in other.items():
            self.evals[label] = self.evals[label] + evaluation if label in self.evals else evaluation
        return self

    def reduce(self):
        return reduce(Evaluation.__iadd__, self.evals.values(), Evaluation())

    def pretty_print(self, table_name=None):
        _print_table(table_name, self)

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
        self.csi_eval = Evaluations(self.labels)
        self.classification_eval = Evaluations(Evaluator.get_classification_labels())

    @property
    def kps_eval(self):
        return self.anchor_eval | self.part_eval

    @staticmethod
    def get_classification_labels():
        """The reference's hard-coded (label, number of parts) classes (evaluator.py:422-427)."""
        return [f"bean_{index}" for index in range(10)] + [f"maize_{index}" for index in range(10)]

    # -- ground truth -> tensors -------------------------------------------------------------
    def _pack_ground_truth(self, annotations, device):
        rows_a = [[(obj.anchor.x, obj.anchor.y, self._label_index.get(obj.name, -1)) for obj in ann.objects]
                  for ann in annotations]
        rows_p = [[(kp.x, kp.y, self._part_index.get(kp.kind, -1)) for obj in ann.objects for kp in obj.parts]
                  for ann in annotations]

        owners = [[j for j, obj in enumerate(ann.objects) for _ in obj.parts] for ann in annotations]
        if any(len(obj.parts) > 64 for ann in annotations for obj in ann.objects):
            raise ValueError("a ground-truth object has more than 64 parts; the object matcher tracks them in a 64-bit mask")

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
        owner = np.zeros((len(annotations), max(1, max((len(o) for o in owners), default=0))), dtype=np.int32)
        for b, row in enumerate(owners):
            owner[b, : len(row)] = row
        return pack(rows_a), pack(rows_p), torch.from_numpy(scale).to(device), torch.from_numpy(owner).to(device)

    # -- the batched equivalent of accumulate(prediction, annotation, raw_parts) -------------------
    def accumulate_packed(self, packed: ops.PackedDetections, annotations, out_size, conf_thresh=None, eval_csi=False,
                          eval_classif=False):
        """Add one decoded batch.  ``packed`` = ``ops.decode_packed(...)`` (device tensors),
        ``annotations`` = the batch's ground truth (``ImageAnnotation`` with ``img_size``, coordinates in the
        network-input frame, as the reference's dataset yields them), ``out_size`` = (W, H) of the heat maps.
        Equivalent to ``accumulate(prediction, annotation, raw_parts, eval_csi, eval_classif)`` of the reference per
        image (evaluator.py:226-242; ``evaluate`` passes True, True: cli/evaluate.py:43-45)."""
        conf = self.args.conf_threshold if conf_thresh is None else conf_thresh
        B, K = packed.anchor_inds.shape
        P = packed.part_inds.shape[1]
        assert len(annotations) == B
        device = packed.anchor_out.device
        M, N = len(self._label_index), len(self._part_index)
        (gt_a, n_a, wa), (gt_p, n_p, wp), scale, owner = self._pack_ground_truth(annotations, device)
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
        if not (eval_csi or eval_classif):
            return
        # ---- object-level metrics: eval_csi / compute_csi and eval_classif (evaluator.py:380-474, 539-581)
        groups = torch.tensor([{"bean": 0, "maize": 1}.get(self._label_names.get(i), -1) for i in range(M)], dtype=torch.int32,
                              device=device)
        c_stats = torch.empty(B, M, 3, dtype=torch.int32, device=device)
        k_stats = torch.empty(B, 20, 3, dtype=torch.int32, device=device)
        c_acc = torch.empty(B, K, dtype=torch.float64, device=device)
        k_acc = torch.empty(B, K, dtype=torch.float64, device=device)
        n_parts = torch.empty(B, K, dtype=torch.int32, device=device)
        op = _native.SdnetObjectMatchParams()
        op.struct_size = ctypes.sizeof(_native.SdnetObjectMatchParams)
        op.B, op.M, op.N, op.K, op.P = B, M, N, K, P
        op.max_gt_objects, op.max_gt_parts = wa, wp
        op.conf, op.sx, op.sy = float(conf), in_w / out_w, in_h / out_h
        op.csi_threshold = float(getattr(self.args, "csi_threshold", 0.75))
        op.anchor_out, op.part_out, op.assign = packed.anchor_out.data_ptr(), packed.part_out.data_ptr(), packed.assign.data_ptr()
        op.image_scale = scale.data_ptr()
        op.gt_objects, op.n_gt_objects = gt_a.data_ptr(), n_a.data_ptr()
        op.gt_parts, op.gt_part_owner, op.n_gt_parts = gt_p.data_ptr(), owner.data_ptr(), n_p.data_ptr()
        op.cls_group = groups.data_ptr()
        op.csi_stats, op.csi_acc = c_stats.data_ptr(), c_acc.data_ptr()
        op.classif_stats, op.classif_acc, op.pred_parts = k_stats.data_ptr(), k_acc.data_ptr(), n_parts.data_ptr()
        _native.check(self.lib.sdnet_match_objects_launch(ctypes.byref(op), ctypes.c_void_p(stream)), "sdnet_match_objects_launch")
        c_stats, k_stats, c_acc, k_acc, n_parts = (t.cpu().numpy() for t in (c_stats, k_stats, c_acc, k_acc, n_parts))
        if eval_csi:
            self._absorb(self.csi_eval, self._label_names, c_stats, c_acc, host[4])
        if eval_classif:
            group_np = groups.cpu().numpy()
            key = np.where((n_parts >= 0) & (n_parts <= 9), group_np[np.clip(host[4], 0, M - 1)] * 10 + n_parts, -1)
            key = np.where(group_np[np.clip(host[4], 0, M - 1)] >= 0, key, -1)
            self._absorb(self.classification_eval, dict(enumerate(Evaluator.get_classification_labels())), k_stats, k_acc, key)

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

    def _results(self):
        return {"Anchor Location": self.anchor_eval, "Part Location": self.part_eval, "All Kps Location": self.kps_eval,
                "CSI": self.csi_eval, "Classification": self.classification_eval}

    def pretty_print(self):
        for title, evals in self._results().items():
            _print_table(title, evals)

    def save_kps_csv(self, path):
        """label, recall, precision, f1, average localisation error per keypoint label (evaluator.py:606-626)."""
        evals = self.kps_eval
        lines = [",".join((label, str(evals[label].recall), str(evals[label].precision), str(evals[label].f1_score),
                           str(evals[label].avg_acc))) for label in sorted(evals.labels)]
        Path(path).write_text("\n".join(lines))

    def __repr__(self):
        text = ""
        for title, evals in self._results().items():
            text += f"{title}\n"
            if len(evals) > 1:
                text += f"  total: {evals.reduce()}\n"
            for label, evaluation in sorted(evals.items(), key=lambda item: item[0]):
                text += f"  {label}: {evaluation}\n"
        return text
