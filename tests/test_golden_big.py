"""Reference-executed goldens at the full map sizes of BASELINE configs 3 and 4 (tests/golden/make_golden_big.py):
512 x 612 maps with K = P = 100, and 30 planes of 256 x 256 with K = P = 500, dense.  The inputs are not stored --
they are regenerated from the recorded seed and checked against the recorded fingerprint -- the reference's
outputs are.  CPU: the torch port and the numpy oracle reproduce them; GPU (-m gpu): the CUDA path does."""
import json
import zlib

import numpy as np
import pytest
import torch

from oracle import sdnet_oracle as O
from oracle import torch_port as TP
from structuredetector_b200.synth import DecodeConfig, make_raw, split_outputs
from tests.helpers import GOLDEN_DIR, assert_objects_close, golden_args, listify, plain, torch_sigmoid_fn

INDEX = json.loads((GOLDEN_DIR / "index_big.json").read_text())
NAMES = sorted(INDEX)


def regenerate(name):
    meta = INDEX[name]
    b, m, n, h, w = meta["shape"]
    cfg = DecodeConfig(name, b, m, n, h, w, meta["K"], meta["P"], meta["conf"], meta["dist"], cfg_id=meta["cfg_id"])
    raw = make_raw(cfg, meta["mode"])
    data = raw.numpy()
    assert zlib.crc32(data.tobytes()) == meta["input"]["crc32"], "synthetic generator drifted: regenerate the goldens"
    return meta, dict(np.load(GOLDEN_DIR / f"{name}.npz")), raw


def check_topk(meta, arr, got, *, score_atol):
    """got: dict with a_inds, p_inds, a_scores_masked, p_scores_masked, a_labels, p_labels (numpy).  Indices and labels
    on the tie-free prefix the reference's CPU order is unambiguous on; scores everywhere (sorted lists agree whatever
    the order inside a run of equal scores)."""
    for who, stable in (("a", meta["anchor_stable"]), ("p", meta["part_stable"])):
        np.testing.assert_allclose(got[f"{who}_scores_masked"], arr[f"{who}_scores_masked"], rtol=0, atol=score_atol)
        for b, n in enumerate(stable):
            np.testing.assert_array_equal(got[f"{who}_inds"][b, :n], arr[f"{who}_inds"][b, :n], err_msg=f"{who} inds image {b}")
            np.testing.assert_array_equal(got[f"{who}_labels"][b, :n], arr[f"{who}_labels"][b, :n])


@pytest.mark.parametrize("name", NAMES)
def test_torch_port_reproduces_the_reference(name):
    meta, arr, raw = regenerate(name)
    b, m, n, h, w = meta["shape"]
    pk = TP.decode_tensors(split_outputs(raw, m, n), meta["K"], meta["P"], meta["conf"], meta["dist"])
    got = {"a_inds": pk["anchor_inds"].numpy(), "p_inds": pk["part_inds"].numpy(),
           "a_scores_masked": pk["anchor_scores_masked"].numpy(), "p_scores_masked": pk["part_scores_masked"].numpy(),
           "a_labels": pk["anchor_out"][..., 3].numpy(), "p_labels": pk["part_out"][..., 3].numpy()}
    check_topk(meta, arr, got, score_atol=0.0)


@pytest.mark.parametrize("name", [n for n in NAMES if "cfg3" in n])
def test_oracle_reproduces_the_reference_at_full_map_size(name):
    meta, arr, raw = regenerate(name)
    b, m, n, h, w = meta["shape"]
    r = raw.numpy()
    pk = O.decode_packed(r[:, :m], r[:, m:m + n], r[:, m + n:m + n + 2], r[:, m + n + 2:], meta["K"], meta["P"], meta["conf"],
                         meta["dist"], sigmoid_fn=torch_sigmoid_fn("cpu"))
    np.testing.assert_array_equal(pk["anchor_inds"], arr["a_inds"])
    np.testing.assert_array_equal(pk["part_inds"], arr["p_inds"])
    np.testing.assert_array_equal(pk["anchor_scores_masked"], arr["a_scores_masked"])
    args = golden_args(meta)
    objs = O.assemble(pk, args._r_labels, args._r_parts, args.anchor_name, meta["conf"], (w, h), (4 * w, 4 * h))
    assert listify(objs) == meta["annotation"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_path_matches_the_reference_at_full_map_size(cuda_device, name):
    """Indices / labels exact on the tie-free prefix, scores within 1e-6 (CPU vs CUDA sigmoid), objects --
    structure, names, grouping exact, coordinates within 1e-5 relative -- wherever the reference's order is unambiguous."""
    from structuredetector_b200 import Decoder

    meta, arr, raw = regenerate(name)
    b, m, n, h, w = meta["shape"]
    outs = split_outputs(raw.to(cuda_device), m, n)
    dec = Decoder(golden_args(meta))
    out = dec(outs, return_metadata=True)
    ta, tk = out["topk_anchor"], out["topk_kp"]
    got = {"a_inds": ta[1].cpu().numpy(), "p_inds": tk[1].cpu().numpy(), "a_scores_masked": ta[0].cpu().numpy(),
           "p_scores_masked": tk[0].cpu().numpy(), "a_labels": ta[2].cpu().numpy(), "p_labels": tk[2].cpu().numpy()}
    check_topk(meta, arr, got, score_atol=1e-6)
    anns = listify(plain(out["annotation"]))
    assert [len(a) for a in anns] == [len(a) for a in meta["annotation"]]
    assert [len(r) for r in out["raw_parts"]] == meta["raw_parts_per_image"]
    if min(meta["anchor_stable"]) >= meta["K"] - 1 and min(meta["part_stable"]) >= meta["P"] - 1:  # tie-free: everything
        assert_objects_close(anns, meta["annotation"], score_atol=1e-6, coord_rtol=1e-5, what=name)
    else:  # dense ties (cfg4): the objects of the unambiguous prefix, anchors only (which tied parts make the top-P, and their order, is the reference's CPU artefact)
        for b_i, n_ok in enumerate(meta["anchor_stable"]):
            for g, wnt in zip(anns[b_i][:n_ok], meta["annotation"][b_i][:n_ok]):
                assert g[0] == wnt[0] and abs(g[1][3] - wnt[1][3]) <= 1e-6
                assert abs(g[1][1] - wnt[1][1]) <= 1e-5 * max(1.0, abs(wnt[1][1])) and abs(g[1][2] - wnt[1][2]) <= 1e-5 * max(1.0, abs(wnt[1][2]))
