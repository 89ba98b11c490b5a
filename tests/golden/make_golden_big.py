"""Golden vectors at the FULL map sizes of BASELINE configs 3 and 4, by EXECUTING the unmodified reference.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_big.py

Same idea as make_golden.py, for shapes whose inputs are too large to commit (2 x 7 x 512 x 612 and
2 x 34 x 256 x 256 floats): the fixture stores what is needed to REGENERATE the input bit for bit (the
synthetic mode, shape and cfg_id that seed ``structuredetector_b200.synth.make_raw``, plus a float64 sum and a
CRC of the bytes so that a drifting generator is noticed) and what the reference's ``Decoder`` returned for it on
CPU: the top-k index / label / score tensors and the Python objects.  cfg3 = 512 x 612 maps, 2 labels + 1 part,
K = P = 100 (tie-free `ladder` and realistic `blobs`); cfg4 = 256 x 256 maps, 20 labels + 10 parts, K = P = 500,
dense `noise`.  Runs only where /root/reference exists.
"""
from __future__ import annotations

import json
import sys
import zlib
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference/src")

from sdnet.data.decoders import Decoder  # noqa: E402  (the reference itself)
from sdnet.utils import clamped_sigmoid, nms  # noqa: E402

from structuredetector_b200.synth import DecodeConfig, make_raw, split_outputs  # noqa: E402
from tests.golden.make_golden import args_for, plain, stable_prefix  # noqa: E402

CASES = [
    # name, (B, M, N, H, W, K, P), mode, conf, dist, cfg_id
    ("big_cfg3_ladder", (2, 2, 1, 512, 612, 100, 100), "ladder", 0.4, 0.1, 303),
    ("big_cfg3_blobs", (2, 2, 1, 512, 612, 100, 100), "blobs", 0.4, 0.1, 304),
    ("big_cfg4_noise", (2, 20, 10, 256, 256, 500, 500), "noise", 0.4, 0.1, 404),
]


def fingerprint(raw: torch.Tensor) -> dict:
    data = raw.numpy()
    return {"sum": float(data.astype(np.float64).sum()), "crc32": zlib.crc32(data.tobytes())}


def main():
    index = {}
    for name, (b, m, n, h, w, k, p), mode, conf, dist, cfg_id in CASES:
        cfg = DecodeConfig(name, b, m, n, h, w, k, p, conf, dist, cfg_id=cfg_id)
        raw = make_raw(cfg, mode)
        outs = split_outputs(raw, m, n)
        args = args_for(cfg, conf, dist)
        meta = Decoder(args)({key: val.clone() for key, val in outs.items()}, return_metadata=True)
        ta, tk = meta["topk_anchor"], meta["topk_kp"]
        np.savez_compressed(
            HERE / f"{name}.npz",
            a_scores_masked=ta[0].numpy(), a_inds=ta[1].numpy().astype(np.int32), a_labels=ta[2].numpy().astype(np.int16),
            a_ys=ta[3].numpy(), a_xs=ta[4].numpy(),
            p_scores_masked=tk[0].numpy(), p_inds=tk[1].numpy().astype(np.int32), p_labels=tk[2].numpy().astype(np.int16),
            p_ys=tk[3].numpy(), p_xs=tk[4].numpy(), embeddings=meta["embeddings"].numpy())
        a_nms = nms(clamped_sigmoid(outs["anchor_hm"]))
        p_nms = nms(clamped_sigmoid(outs["part_hm"]))
        a_top = torch.topk(torch.topk(a_nms.view(b, m, -1), k)[0].view(b, -1), k)[0].numpy()
        p_top = torch.topk(torch.topk(p_nms.view(b, n, -1), p)[0].view(b, -1), p)[0].numpy()
        index[name] = {
            "shape": [b, m, n, h, w], "K": k, "P": p, "mode": mode, "conf": conf, "dist": dist, "cfg_id": cfg_id,
            "anchor_name": "stem", "down_ratio": 4.0, "input": fingerprint(raw),
            "annotation": plain(meta["annotation"]),
            "raw_parts_per_image": [len(img) for img in meta["raw_parts"]],
            "anchor_stable": stable_prefix(a_top), "part_stable": stable_prefix(p_top),
            "torch": torch.__version__,
        }
        print(name, "objects", [len(a) for a in index[name]["annotation"]], "stable", index[name]["anchor_stable"],
              index[name]["part_stable"], index[name]["input"])
    (HERE / "index_big.json").write_text(json.dumps(index))


if __name__ == "__main__":
    main()
