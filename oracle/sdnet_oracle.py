"""CPU oracle for the SDNet decoding hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this module.  The shipped decoder (``structuredetector_b200``) never does: it
fails loudly when its CUDA library is missing.

What it restates
----------------
A numpy restatement of ``Decoder.__call__`` of laclouis5/StructureDetector
(reference: src/sdnet/data/decoders.py:33-177) and of the five tensor helpers it
calls (reference: src/sdnet/utils/utils.py:342-361, 422-467).  The arithmetic of
those helpers lives in a third-party dependency, PyTorch ATen (pinned torch 2.5.1 in
the reference's uv.lock:970-971; 2.11.0+cu128 in this image): sigmoid, clamp,
max_pool2d, topk, gather, min.  Their published semantics are restated here in
numpy float32.

Parity pinning
--------------
The reference has no tests, fixtures or golden vectors of its own (SURVEY.md
section 4), so this oracle is pinned against *outputs of the reference itself*:
``tests/golden/make_golden.py`` imports the unmodified reference from
``/root/reference/src`` and stores its results for the committed inputs under
``tests/golden/``; ``tests/test_oracle.py`` checks this file against them (and,
when ``/root/reference`` is present, against the live reference).

Tie rule
--------
``torch.topk`` leaves the order of equal values unspecified.  This oracle fixes the
canonical rule **(score descending, flat index ascending)**, which is what
torch's CUDA ``topk`` produces for k > 32 (probed on a B200,
``profiles/r01_probe_torch_cuda.json``); for k <= 32 torch-CUDA's final bitonic
sort permutes equal-score runs and torch-CPU's order is a libstdc++ artefact, so
comparisons against those go through :func:`canonicalise_ties`.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
CLAMP_LO = F32(1e-6)  # reference: src/sdnet/utils/utils.py:361
CLAMP_HI = F32(1 - 1e-6)  # double 0.999999 rounded to fp32 = 0.99999899
MASKED_FAR = F32(1e6)  # reference: src/sdnet/data/decoders.py:80-86


# --------------------------------------------------------------------------- helpers
def sigmoid_correctly_rounded(x: np.ndarray) -> np.ndarray:
    """fp32 sigmoid evaluated in float64 and rounded once.

    ATen's CPU (Sleef) and CUDA (``1/(1+expf(-x))``) kernels differ from this by at
    most a couple of ulps; tests that need bit-identical scores pass the device's own
    sigmoid through ``sigmoid_fn``.
    """
    x64 = np.asarray(x, dtype=np.float64)
    with np.errstate(over="ignore"):
        return (1.0 / (1.0 + np.exp(-x64))).astype(F32)


def clamped_sigmoid(x: np.ndarray, sigmoid_fn=None) -> np.ndarray:
    """reference: src/sdnet/utils/utils.py:355-361 (sigmoid, then clamp to [1e-6, 1-1e-6])."""
    s = (sigmoid_fn or sigmoid_correctly_rounded)(np.asarray(x, dtype=F32))
    return np.clip(np.asarray(s, dtype=F32), CLAMP_LO, CLAMP_HI).astype(F32)


def window_max(hm: np.ndarray, radius: int = 2) -> np.ndarray:
    """(2r+1)^2 stride-1 max filter with -inf padding (max_pool2d semantics)."""
    b, c, h, w = hm.shape
    pad = np.full((b, c, h + 2 * radius, w + 2 * radius), -np.inf, dtype=hm.dtype)
    pad[:, :, radius : radius + h, radius : radius + w] = hm
    # separable: rows then columns
    rows = pad[:, :, :, radius : radius + w].copy()
    for d in range(-radius, radius + 1):
        np.maximum(rows, pad[:, :, :, radius + d : radius + d + w], out=rows)
    out = rows[:, :, radius : radius + h].copy()
    for d in range(-radius, radius + 1):
        np.maximum(out, rows[:, :, radius + d : radius + d + h], out=out)
    return out


def nms(hm: np.ndarray, radius: int = 2) -> np.ndarray:
    """reference: src/sdnet/utils/utils.py:441-443 -- keep a pixel iff it equals its
    5x5 window maximum (whole plateaus survive), else 0.0."""
    mx = window_max(hm, radius)
    return np.where(hm == mx, hm, F32(0.0)).astype(F32)


def _topk_desc_stable(values: np.ndarray, k: int):
    """Top-k along the last axis under (value desc, index asc)."""
    n = values.shape[-1]
    if k > n:
        raise RuntimeError("selected index k out of range")  # torch.topk's error text
    order = np.argsort(-values, axis=-1, kind="stable")[..., :k]
    return np.take_along_axis(values, order, axis=-1), order.astype(np.int64)


def topk(scores: np.ndarray, k: int):
    """reference: src/sdnet/utils/utils.py:447-467 -- per-channel top-k, then top-k over
    the channel-major concatenation; class = floor(pos / k)."""
    b, c, h, w = scores.shape
    if h * w >= (1 << 24):
        raise ValueError("float index arithmetic of the reference is only exact below 2**24 pixels")
    s1, i1 = _topk_desc_stable(scores.reshape(b, c, h * w), k)  # (B, C, k)
    ys1 = (i1 // w).astype(F32)  # floor(float(ind) / W), exact below 2**24
    xs1 = (i1 % w).astype(F32)
    s2, pos = _topk_desc_stable(s1.reshape(b, c * k), k)  # (B, k)
    cls = (pos // k).astype(F32)
    inds = np.take_along_axis(i1.reshape(b, c * k), pos, axis=1)
    ys = np.take_along_axis(ys1.reshape(b, c * k), pos, axis=1)
    xs = np.take_along_axis(xs1.reshape(b, c * k), pos, axis=1)
    return s2.astype(F32), inds, cls, ys, xs


def transpose_and_gather(feat: np.ndarray, ind: np.ndarray) -> np.ndarray:
    """reference: src/sdnet/utils/utils.py:347-351 -- (B, Cf, H, W), (B, n) -> (B, n, Cf)."""
    b, cf = feat.shape[:2]
    flat = feat.reshape(b, cf, -1)
    idx = np.broadcast_to(ind[:, None, :], (b, cf, ind.shape[1]))
    return np.take_along_axis(flat, idx, axis=2).transpose(0, 2, 1)


def hypot(d: np.ndarray) -> np.ndarray:
    """reference: src/sdnet/utils/utils.py:422-437 -- square, sum over the last axis of
    size 2, sqrt; every step rounded to fp32, no fused multiply-add."""
    d = d.astype(F32)
    sq = (d * d).astype(F32)
    return np.sqrt((sq[..., 0] + sq[..., 1]).astype(F32)).astype(F32)


# --------------------------------------------------------------------------- decode
def decode_packed(
    anchor_hm,
    part_hm,
    offsets,
    embeddings,
    max_objects: int,
    max_parts: int,
    conf_thresh: float,
    dist_thresh: float,
    *,
    sigmoid_fn=None,
    pre_activated: bool = False,
    radius: int = 2,
    group: bool = True,
    activation_fn=None,
    conf_cmp=None,
):
    """Tensor half of the decoder (reference: src/sdnet/data/decoders.py:40-100).

    ``pre_activated`` skips sigmoid+NMS (the CoreMLDecoder variant, decoders.py:211,226).
    Returns a dict of numpy arrays laid out like the C-ABI outputs in
    ``include/sdnet_decode.h``.

    Reduced-precision inputs (fp16 / bf16, what the reference decodes under ``--amp``): pass the maps
    converted exactly to float32, ``activation_fn`` = the dtype's whole clamped sigmoid (ATen rounds
    the sigmoid and the clamp to the tensor dtype; its results are exactly representable in float32)
    and ``conf_cmp`` = the threshold rounded to that dtype (``scores > conf`` compares in the scores'
    dtype).  Everything after the activation is dtype-independent.
    """
    anchor_hm = np.asarray(anchor_hm, dtype=F32)
    part_hm = np.asarray(part_hm, dtype=F32)
    offsets = np.asarray(offsets, dtype=F32)
    embeddings = np.asarray(embeddings, dtype=F32)
    b, _, h, w = anchor_hm.shape
    k, p = int(max_objects), int(max_parts)

    if pre_activated:
        a_sig, p_sig, a_nms, p_nms = anchor_hm, part_hm, anchor_hm, part_hm
    else:
        act = (lambda m: np.asarray(activation_fn(m), dtype=F32)) if activation_fn else (lambda m: clamped_sigmoid(m, sigmoid_fn))
        a_sig, p_sig = act(anchor_hm), act(part_hm)
        a_nms, p_nms = nms(a_sig, radius), nms(p_sig, radius)

    a_score, a_ind, a_cls, a_ys, a_xs = topk(a_nms, k)
    a_off = transpose_and_gather(offsets, a_ind)
    a_x = (a_xs + a_off[..., 0]).astype(F32)
    a_y = (a_ys + a_off[..., 1]).astype(F32)
    anchor_out = np.stack((a_x, a_y, a_score, a_cls), axis=2).astype(F32)

    p_score, p_ind, p_cls, p_ys, p_xs = topk(p_nms, p)
    p_off = transpose_and_gather(offsets, p_ind)
    p_emb = transpose_and_gather(embeddings, p_ind).astype(F32)
    p_x = (p_xs + p_off[..., 0]).astype(F32)
    p_y = (p_ys + p_off[..., 1]).astype(F32)
    o_x = (p_x + p_emb[..., 0]).astype(F32)
    o_y = (p_y + p_emb[..., 1]).astype(F32)
    part_out = np.stack((p_x, p_y, p_score, p_cls, o_x, o_y), axis=2).astype(F32)

    out = {
        "anchor_out": anchor_out,
        "part_out": part_out,
        "anchor_inds": a_ind,
        "part_inds": p_ind,
        "part_emb": p_emb,
        "anchor_sig": a_sig,
        "part_sig": p_sig,
    }
    if not group:
        return out

    conf32 = F32(conf_thresh if conf_cmp is None else conf_cmp)  # tensor > python-float compares in the tensor's dtype
    one = F32(1.0)
    p_mask = (p_score > conf32).astype(F32)
    a_mask = (a_score > conf32).astype(F32)
    out["part_scores_masked"] = (-(one - p_mask) + p_mask * p_score).astype(F32)
    out["anchor_scores_masked"] = (-(one - a_mask) + a_mask * a_score).astype(F32)
    ori_x = (-MASKED_FAR * (one - p_mask) + p_mask * o_x).astype(F32)
    ori_y = (-MASKED_FAR * (one - p_mask) + p_mask * o_y).astype(F32)
    pos_x = (MASKED_FAR * (one - a_mask) + a_mask * a_x).astype(F32)
    pos_y = (MASKED_FAR * (one - a_mask) + a_mask * a_y).astype(F32)
    # (B, K, P, 2): origins - anchor_pos
    dx = (ori_x[:, None, :] - pos_x[:, :, None]).astype(F32)
    dy = (ori_y[:, None, :] - pos_y[:, :, None]).astype(F32)
    dist = hypot(np.stack((dx, dy), axis=-1))  # (B, K, P)
    min_inds = np.argmin(dist, axis=1)  # first minimum wins
    min_vals = np.take_along_axis(dist, min_inds[:, None, :], axis=1)[:, 0, :]
    gate = F32(float(dist_thresh) * min(w, h))  # product in double, compare in fp32
    ok = min_vals < gate
    out["min_inds"] = min_inds.astype(np.int64)
    out["min_vals"] = min_vals.astype(F32)
    out["assign"] = np.where(ok, min_inds, -1).astype(np.int32)
    out["counts"] = np.stack((a_mask.sum(axis=1), p_mask.sum(axis=1)), axis=1).astype(np.int32)
    return out


def assemble(packed, label_map, part_map, anchor_name, conf_thresh, out_size, in_size):
    """Object half of the decoder (reference: src/sdnet/data/decoders.py:103-139).

    Returns, per image, a list of ``(label, (anchor_name, x, y, score), [(kind, x, y, score)...])``
    in anchor-slot order with parts in part-slot order; coordinates already resized by
    ``in_size/out_size`` in double precision (reference: src/sdnet/utils/utils.py:19-26).
    """
    (ow, oh), (iw, ih) = out_size, in_size
    rx, ry = iw / ow, ih / oh
    anchors = packed["anchor_out"].astype(np.float64)
    parts = packed["part_out"].astype(np.float64)
    assign = packed["assign"]
    images = []
    for b in range(anchors.shape[0]):
        buckets = {}
        for i, slot in enumerate(assign[b].tolist()):
            if slot >= 0:
                buckets.setdefault(slot, []).append(i)
        objects = []
        for a_i in range(anchors.shape[1]):
            ax, ay, score, cls = anchors[b, a_i].tolist()
            if score <= conf_thresh:  # compared in double
                continue
            kps = []
            for i in buckets.get(a_i, ()):
                px, py, ps, pc = parts[b, i, :4].tolist()
                kps.append((part_map[int(pc)], px * rx, py * ry, ps))
            objects.append((label_map[int(cls)], (anchor_name, ax * rx, ay * ry, score), kps))
        images.append(objects)
    return images


def raw_parts(packed, part_map, conf_thresh, out_size, in_size):
    """reference: src/sdnet/data/decoders.py:142-159 (note ``score < conf`` skip, in double)."""
    (ow, oh), (iw, ih) = out_size, in_size
    rx, ry = iw / ow, ih / oh
    parts = packed["part_out"].astype(np.float64)
    images = []
    for b in range(parts.shape[0]):
        kept = []
        for px, py, ps, pc in parts[b, :, :4].tolist():
            if ps < conf_thresh:
                continue
            kept.append((part_map[int(pc)], px * rx, py * ry, ps))
        images.append(kept)
    return images


def keypoint_decode(anchor_hm, part_hm, offsets, max_objects, max_parts, conf_thresh, down_ratio,
                    label_map, part_map, *, sigmoid_fn=None, radius: int = 2):
    """reference: src/sdnet/data/decoders.py:345-423 (KeypointDecoder: no grouping, coordinates
    scaled by r_w/r_h in fp32, ``score < conf`` skip evaluated on fp32 tensors)."""
    pk = decode_packed(anchor_hm, part_hm, offsets, np.zeros_like(np.asarray(offsets, dtype=F32)),
                       max_objects, max_parts, conf_thresh, 0.0, sigmoid_fn=sigmoid_fn, radius=radius, group=False)
    h, w = np.asarray(anchor_hm).shape[2:]
    in_h, in_w = int(down_ratio * h), int(down_ratio * w)
    r_h, r_w = F32(in_h / h), F32(in_w / w)
    conf32 = F32(conf_thresh)
    images = []
    for b in range(pk["anchor_out"].shape[0]):
        kps = []
        for arr, names in ((pk["anchor_out"][b], label_map), (pk["part_out"][b], part_map)):
            for row in arr:
                x, y, score, cls = F32(row[0] * r_w), F32(row[1] * r_h), row[2], row[3]
                if score < conf32:
                    continue
                kps.append((names[int(cls)], float(x), float(y), float(score)))
        images.append(kps)
    return images


# --------------------------------------------------------------------------- comparison
def canonicalise_ties(scores: np.ndarray, cls: np.ndarray, inds: np.ndarray):
    """Permutation that re-orders each row's equal-score runs by (class asc, index asc).

    Used to compare against torch builds whose ``topk`` orders equal scores
    differently (CPU always; CUDA for k <= 32).  Returns (B, k) gather indices.
    """
    b, k = scores.shape
    perm = np.empty((b, k), dtype=np.int64)
    for r in range(b):
        perm[r] = np.lexsort((inds[r], cls[r], -scores[r].astype(np.float64)))
    return perm


def boundary_is_unambiguous(nms_map: np.ndarray, k: int) -> np.ndarray:
    """Per (B, C): True when the k-th and (k+1)-th largest values differ, i.e. the top-k
    *set* does not depend on the tie rule."""
    b, c = nms_map.shape[:2]
    flat = nms_map.reshape(b, c, -1)
    if flat.shape[-1] <= k:
        return np.ones((b, c), dtype=bool)
    part = -np.partition(-flat, k, axis=-1)[..., : k + 1]
    part.sort(axis=-1)
    return part[..., 0] != part[..., 1]
