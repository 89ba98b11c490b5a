"""Shared test plumbing: plain-tuple views of decoder outputs, sigmoid adapters, arg namespaces."""
from __future__ import annotations

import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

REFERENCE_SRC = Path("/root/reference/src")
GOLDEN_DIR = Path(__file__).resolve().parent / "golden"


def make_args(cfg, conf=None, dist=None, anchor_name="stem", down_ratio=4.0, max_objects=None, max_parts=None):
    labels = {i: f"label{i}" for i in range(cfg.labels)}
    parts = {i: f"part{i}" for i in range(cfg.parts)}
    if cfg.labels == 2 and cfg.parts == 1:  # reference labels.json
        labels, parts = {0: "bean", 1: "maize"}, {0: "leaf"}
    return SimpleNamespace(
        _r_labels=labels, _r_parts=parts, anchor_name=anchor_name, down_ratio=down_ratio,
        max_objects=cfg.max_objects if max_objects is None else max_objects,
        max_parts=cfg.max_parts if max_parts is None else max_parts,
        conf_threshold=cfg.conf_threshold if conf is None else conf,
        decoder_dist_thresh=cfg.dist_thresh if dist is None else dist,
    )


def plain(annotations):
    """ImageAnnotation list (ours or the reference's) -> nested tuples for equality checks."""
    return [[(o.name, (o.anchor.kind, o.anchor.x, o.anchor.y, o.anchor.score),
              [(p.kind, p.x, p.y, p.score) for p in o.parts]) for o in ann.objects] for ann in annotations]


def plain_keypoints(images):
    return [[(k.kind, k.x, k.y, k.score) for k in kps] for kps in images]


def torch_sigmoid_fn(device="cpu"):
    """numpy->numpy sigmoid evaluated by torch on ``device`` (bit-identical to what the
    reference computes there)."""
    def fn(x):
        t = torch.from_numpy(np.ascontiguousarray(x)).to(device)
        return torch.sigmoid(t).cpu().numpy()
    return fn


def np_inputs(outputs):
    return [outputs[k].detach().cpu().numpy() for k in ("anchor_hm", "part_hm", "offsets", "embeddings")]


def reference_available() -> bool:
    return (REFERENCE_SRC / "sdnet" / "data" / "decoders.py").exists()


def import_reference():
    """Import the unmodified reference package read-only (never writes bytecode)."""
    sys.dont_write_bytecode = True
    if str(REFERENCE_SRC) not in sys.path:
        sys.path.insert(0, str(REFERENCE_SRC))
    import sdnet.data.decoders as ref_decoders  # noqa: WPS433

    return ref_decoders


def packed_np(packed):
    return {k: v.detach().cpu().numpy() for k, v in packed.as_dict().items()}


def assert_packed_equal(got: dict, want: dict, *, what=""):
    """Bit-exact comparison of every packed field the oracle also produces."""
    for key in ("anchor_inds", "part_inds", "assign", "counts"):
        if key in want and key in got:
            np.testing.assert_array_equal(got[key], want[key], err_msg=f"{what}: {key}")
    for key in ("anchor_out", "part_out", "part_emb"):
        if key in want and key in got:
            g, w = np.ascontiguousarray(got[key]).view(np.uint32), np.ascontiguousarray(want[key]).view(np.uint32)
            bad = np.argwhere(g != w)
            assert bad.size == 0, f"{what}: {key} differs at {bad[:5].tolist()} got {got[key][tuple(bad[0])]} want {want[key][tuple(bad[0])]}"
