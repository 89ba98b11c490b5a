"""Generate golden vectors by EXECUTING the unmodified reference decoder.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Runs only where /root/reference exists (the build container).  For each case it stores the
input tensor and what the reference's `Decoder`, `CoreMLDecoder` and `KeypointDecoder`
(reference: src/sdnet/data/decoders.py) returned on CPU with this image's torch build, as one
compressed .npz + a JSON side-car holding the Python-object outputs.  The fixtures are small
(a few hundred KB in total) and are what `-m "not gpu"` tests pin the oracle against and what
`-m gpu` tests pin the CUDA path against on the GPU box, where the reference does not exist.

Cases are chosen so that the CPU result does not depend on torch.topk's unspecified tie order
where it matters: `ladder` inputs are tie-free by construction; for the others the script
records, per top-k list, how many leading entries are tie-free and unambiguous
(`*_stable` counts) so consumers compare exactly that prefix.
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

from sdnet.data.decoders import CoreMLDecoder, Decoder, KeypointDecoder  # noqa: E402  (the reference itself)
from sdnet.utils import clamped_sigmoid, nms  # noqa: E402

from structuredetector_b200.synth import DecodeConfig, make_raw, split_outputs  # noqa: E402

CASES = [
    # name, (B, M, N, H, W, K, P), mode, conf, dist
    ("ladder_small", (2, 2, 1, 40, 52, 30, 40), "ladder", 0.4, 0.1),
    ("ladder_wide", (1, 2, 1, 72, 160, 100, 100), "ladder", 0.4, 0.1),
    ("blobs_cfg1", (1, 2, 1, 128, 128, 100, 100), "blobs", 0.4, 0.1),
    ("noise_defaults", (2, 2, 1, 64, 96, 20, 40), "noise", 0.5, 0.1),  # the CLI's default K/P/conf (args.py:103-146)
    ("noise_multiclass", (1, 5, 3, 48, 64, 60, 50), "noise", 0.3, 0.25),
    ("ties_small", (2, 2, 2, 32, 44, 40, 40), "ties", 0.4, 0.1),
]


def args_for(cfg, conf, dist):
    return SimpleNamespace(
        _r_labels={i: f"label{i}" for i in range(cfg.labels)}, _r_parts={i: f"part{i}" for i in range(cfg.parts)},
        anchor_name="stem", down_ratio=4.0, max_objects=cfg.max_objects, max_parts=cfg.max_parts,
        conf_threshold=conf, decoder_dist_thresh=dist)


def plain(annotations):
    return [[[o.name, [o.anchor.kind, o.anchor.x, o.anchor.y, o.anchor.score],
              [[p.kind, p.x, p.y, p.score] for p in o.parts]] for o in ann.objects] for ann in annotations]


def stable_prefix(scores: np.ndarray) -> list:
    """Per row: length of the leading run in which every score is strictly greater than the next."""
    out = []
    for row in scores:
        n = 0
        while n + 1 < len(row) and row[n] > row[n + 1]:
            n += 1
        out.append(n)  # entries [0, n) are ordered unambiguously
    return out


def main():
    torch.manual_seed(0)
    index = {}
    for i, (name, (b, m, n, h, w, k, p), mode, conf, dist) in enumerate(CASES):
        cfg = DecodeConfig(name, b, m, n, h, w, k, p, conf, dist, cfg_id=100 + i)
        raw = make_raw(cfg, mode)
        outs = split_outputs(raw, m, n)
        args = args_for(cfg, conf, dist)
        clone = lambda: {key: val.clone() for key, val in outs.items()}
        meta = Decoder(args)(clone(), return_metadata=True)
        kp = KeypointDecoder(args)(clone())
        # CoreMLDecoder consumes maps that already went through sigmoid + nms inside the exported model
        pre = clone()
        pre["anchor_hm"] = nms(clamped_sigmoid(pre["anchor_hm"]))
        pre["part_hm"] = nms(clamped_sigmoid(pre["part_hm"]))
        cm = CoreMLDecoder(args)({key: val.clone() for key, val in pre.items()}, return_metadata=True)

        ta, tk = meta["topk_anchor"], meta["topk_kp"]
        arrays = {
            "raw": raw.numpy(),
            "a_scores_masked": ta[0].numpy(), "a_inds": ta[1].numpy(), "a_labels": ta[2].numpy(),
            "a_ys": ta[3].numpy(), "a_xs": ta[4].numpy(),
            "p_scores_masked": tk[0].numpy(), "p_inds": tk[1].numpy(), "p_labels": tk[2].numpy(),
            "p_ys": tk[3].numpy(), "p_xs": tk[4].numpy(),
            "embeddings": meta["embeddings"].numpy(),
            "coreml_a_inds": cm["topk_anchor"][1].numpy(), "coreml_p_inds": cm["topk_kp"][1].numpy(),
        }
        if name == "ladder_small":  # the full sigmoid maps are kept for one small case only
            arrays["anchor_sig"] = meta["anchor_hm_sig"].numpy()
            arrays["part_sig"] = meta["part_hm_sig"].numpy()
        np.savez_compressed(HERE / f"{name}.npz", **arrays)
        # unmasked scores (for the stable-prefix bookkeeping) straight from the reference helpers
        a_nms = nms(clamped_sigmoid(outs["anchor_hm"]))
        p_nms = nms(clamped_sigmoid(outs["part_hm"]))
        a_top = torch.topk(torch.topk(a_nms.view(b, m, -1), k)[0].view(b, -1), k)[0].numpy()
        p_top = torch.topk(torch.topk(p_nms.view(b, n, -1), p)[0].view(b, -1), p)[0].numpy()
        index[name] = {
            "shape": [b, m, n, h, w], "K": k, "P": p, "mode": mode, "conf": conf, "dist": dist,
            "anchor_name": "stem", "down_ratio": 4.0,
            "annotation": plain(meta["annotation"]),
            "raw_parts": [[[q.kind, q.x, q.y, q.score] for q in img] for img in meta["raw_parts"]],
            "coreml_annotation": plain(cm["annotation"]),
            "keypoints": [[[q.kind, q.x, q.y, q.score] for q in img] for img in kp],
            "anchor_stable": stable_prefix(a_top), "part_stable": stable_prefix(p_top),
            "torch": torch.__version__,
        }
        print(name, "objects", [len(a) for a in index[name]["annotation"]], "stable", index[name]["anchor_stable"],
              index[name]["part_stable"])
    # reduced-precision inputs (the --amp validation path): ties everywhere, so only tie-independent
    # facts are stored -- the sorted score lists and the number of objects per image
    for name, dtype in (("half_f16", torch.float16), ("half_bf16", torch.bfloat16)):
        cfg = DecodeConfig(name, 2, 2, 1, 48, 64, 40, 40, 0.4, 0.1, cfg_id=150)
        raw = make_raw(cfg, "blobs").to(dtype)
        outs = split_outputs(raw, 2, 1)
        args = args_for(cfg, 0.4, 0.1)
        meta = Decoder(args)({key: val.clone() for key, val in outs.items()}, return_metadata=True)
        np.savez_compressed(HERE / f"{name}.npz", raw=raw.float().numpy(),
                            a_scores_masked=meta["topk_anchor"][0].float().numpy(),
                            p_scores_masked=meta["topk_kp"][0].float().numpy(),
                            anchor_sig=meta["anchor_hm_sig"].float().numpy())
        index[name] = {"shape": [2, 2, 1, 48, 64], "K": 40, "P": 40, "mode": "blobs", "conf": 0.4, "dist": 0.1,
                       "anchor_name": "stem", "down_ratio": 4.0, "dtype": str(dtype).split(".")[-1],
                       "objects_per_image": [len(a) for a in meta["annotation"]],
                       "parts_per_image": [a.nb_parts for a in meta["annotation"]], "torch": torch.__version__}
        print(name, index[name]["objects_per_image"], index[name]["parts_per_image"])
    (HERE / "index.json").write_text(json.dumps(index))
    print("wrote", len(CASES), "cases")


if __name__ == "__main__":
    main()
