"""Drop-in decoders: same constructor, call signature and return structure as the reference
(reference: src/sdnet/data/decoders.py:17-179 ``Decoder``, 182-342 ``CoreMLDecoder``,
345-423 ``KeypointDecoder``), with the tensor half running in the sm_100a kernels.

What stays in Python is what the reference also does on the host: turning K+P packed
rows per image into ``ImageAnnotation`` objects.  The reference reads every scalar with
``.item()`` (~1,000 device syncs per image); here the packed result comes back in ONE
device->host copy and the objects are built from ``.tolist()`` rows.
"""
from __future__ import annotations

import contextlib
import gc

import numpy as np
import torch

from . import ops
from .annotations import ImageAnnotation, Keypoint, Object

try:
    from . import _fastobj  # C object assembly (csrc/fastobj.c), built by structuredetector_b200.build
except ImportError as exc:  # no silent slow path: say how to get it
    raise ImportError("structuredetector_b200._fastobj is missing: run `python -m structuredetector_b200.build` "
                      "(gcc, CPython headers) to build the object-assembly extension") from exc

__all__ = ["Decoder", "CoreMLDecoder", "KeypointDecoder", "RawDecoder"]


@contextlib.contextmanager
def _gc_paused():
    """Building ~300 small acyclic objects per image makes the cyclic collector re-scan an ever larger
    young generation (measured: 1.3 k -> 4.0 k images/s on dense outputs with it paused)."""
    was_enabled = gc.isenabled()
    gc.disable()
    try:
        yield
    finally:
        if was_enabled:
            gc.enable()


class _DecoderBase:
    _pre_activated = False

    def __init__(self, args):
        # the eight fields the reference decoder reads (decoders.py:19-26, 33-38)
        self.label_map = args._r_labels
        self.part_map = args._r_parts
        self.anchor_name = args.anchor_name
        self.args = args
        self.down_ratio = args.down_ratio
        self.max_objects = args.max_objects  # K
        self.max_parts = args.max_parts  # P

    # -- helpers ---------------------------------------------------------------------------
    def _sizes(self, outputs):
        out_h, out_w = outputs["anchor_hm"].shape[2:]
        in_h, in_w = int(self.down_ratio * out_h), int(self.down_ratio * out_w)
        return (out_w, out_h), (in_w, in_h)

    @staticmethod
    def _to_host(packed: ops.PackedDetections) -> ops.PackedDetections:
        """One device->host copy of the whole packed result."""
        blob = packed.blob.cpu()
        b, k = packed.anchor_inds.shape
        p = packed.part_inds.shape[1]
        return ops._carve(blob, b, k, p, packed.diag.shape[0] // b)


class Decoder(_DecoderBase):
    """``decoder(outputs, conf_thresh=None, dist_thresh=None, return_metadata=False)``."""

    def __call__(self, outputs, conf_thresh=None, dist_thresh=None, return_metadata=False):
        conf_thresh = self.args.conf_threshold if conf_thresh is None else conf_thresh
        dist_thresh = self.args.decoder_dist_thresh if dist_thresh is None else dist_thresh
        out_size, in_size = self._sizes(outputs)

        packed = ops.decode_packed(outputs, self.max_objects, self.max_parts, conf_thresh, dist_thresh,
                                   pre_activated=self._pre_activated)
        host = self._to_host(packed)
        with _gc_paused():
            annotations = self._assemble(host, conf_thresh, out_size, in_size)
        if not return_metadata:
            return annotations

        meta = {"annotation": annotations}
        if not self._pre_activated:
            meta["anchor_hm_sig"] = ops.activate_maps(outputs["anchor_hm"])
            meta["part_hm_sig"] = ops.activate_maps(outputs["part_hm"])
        a, p = packed.anchor_out, packed.part_out
        # the reference re-binds the score tensors to their masked (-1) versions before
        # returning them (decoders.py:79,84 then 166-173): kept, oddity included
        in_dtype = outputs["anchor_hm"].dtype
        conf32 = torch.tensor(conf_thresh, dtype=in_dtype, device=a.device).float()  # compared in the scores' dtype
        mask = lambda s: torch.where(s > conf32, s, torch.full_like(s, -1.0))
        meta["embeddings"] = packed.part_emb if in_dtype == torch.float32 else packed.part_emb.to(in_dtype)
        meta["topk_anchor"] = (mask(a[..., 2]), packed.anchor_inds, a[..., 3], a[..., 1], a[..., 0])
        meta["topk_kp"] = (mask(p[..., 2]), packed.part_inds, p[..., 3], p[..., 1], p[..., 0])
        with _gc_paused():
            meta["raw_parts"] = self._raw_parts(host, conf_thresh, out_size, in_size)
        meta["raw_embeddings"] = outputs["embeddings"]
        meta["raw_offsets"] = outputs["offsets"]
        return meta

    # reference: decoders.py:103-139
    def _assemble(self, host: ops.PackedDetections, conf_thresh, out_size, in_size):
        """Python objects from the packed rows, built by the C loop in ``csrc/fastobj.c``: float32 coordinates
        widened to double and multiplied by ``in/out`` in double (the reference multiplies ``.item()`` doubles),
        ``score > conf`` in double, parts bucketed by anchor slot in part-slot order."""
        sx, sy = in_size[0] / out_size[0], in_size[1] / out_size[1]
        a, p, assign = host.anchor_out, host.part_out, host.assign
        B, K = a.shape[:2]
        return _fastobj.assemble(a.contiguous().numpy(), p.contiguous().numpy(), assign.contiguous().numpy(), B, K,
                                 p.shape[1], float(conf_thresh), sx, sy, self._names(self.label_map),
                                 self._names(self.part_map), self.anchor_name, Keypoint, Object, ImageAnnotation)

    @staticmethod
    def _names(mapping):
        """Class index -> name as a list (dict lookups with int keys, hoisted out of the loops)."""
        if not mapping:
            return []
        table = [None] * (max(mapping) + 1)
        for index, name in mapping.items():
            table[index] = name
        return table

    # reference: decoders.py:142-159
    def _raw_parts(self, host: ops.PackedDetections, conf_thresh, out_size, in_size):
        sx, sy = in_size[0] / out_size[0], in_size[1] / out_size[1]
        p = host.part_out
        return _fastobj.keypoints(p.contiguous().numpy(), p.shape[0], p.shape[1], p.shape[2], float(conf_thresh), sx, sy,
                                  self._names(self.part_map), Keypoint)


class CoreMLDecoder(Decoder):
    """Variant whose heat maps already went through sigmoid + NMS inside the exported model
    (reference: decoders.py:182-342).  Same outputs minus the two ``*_hm_sig`` maps."""

    _pre_activated = True


class KeypointDecoder(_DecoderBase):
    """Keypoints only, no grouping (reference: decoders.py:345-423): returns, per image, the
    anchors then the parts whose fp32 score is not below ``args.conf_threshold``."""

    def __call__(self, outputs):
        conf_thresh = self.args.conf_threshold
        (out_w, out_h), (in_w, in_h) = self._sizes(outputs)
        r_h, r_w = np.float32(in_h / out_h), np.float32(in_w / out_w)
        packed = ops.decode_packed(outputs, self.max_objects, self.max_parts, conf_thresh, 0.0, group=False)
        host = self._to_host(packed)
        with _gc_paused():
            return self._keypoints(host, np.float32(conf_thresh), r_w, r_h)

    def _keypoints(self, host, conf32, r_w, r_h):
        annotations = []
        anchors, parts = host.anchor_out.numpy(), host.part_out.numpy()
        for b in range(anchors.shape[0]):
            keypoints = []
            for rows, names in ((anchors[b], self.label_map), (parts[b], self.part_map)):
                keep = ~(rows[:, 2] < conf32)
                xs = (rows[:, 0] * r_w)[keep].tolist()  # fp32 multiply, as on the reference's tensors
                ys = (rows[:, 1] * r_h)[keep].tolist()
                for x, y, score, cls in zip(xs, ys, rows[keep, 2].tolist(), rows[keep, 3].tolist()):
                    keypoints.append(Keypoint(kind=names[int(cls)], x=x, y=y, score=score))
            annotations.append(keypoints)
        return annotations


class RawDecoder:
    """``RawDecoder(nb_hms)(raw)``: the network's raw ``(B, M+N+4, H, W)`` output with its first ``nb_hms``
    channels replaced by ``nms(clamped_sigmoid(.))`` -- what the reference bakes into its exported model
    (reference: src/sdnet/cli/convert_coreml.py:12-19) and ``CoreMLDecoder`` then consumes."""

    def __init__(self, nb_hms: int) -> None:
        self.nb_hms = nb_hms

    def __call__(self, input: torch.Tensor) -> torch.Tensor:
        heatmaps = ops.suppress_maps(input[:, : self.nb_hms])
        return torch.cat(tensors=(heatmaps, input[:, self.nb_hms:]), dim=1)
