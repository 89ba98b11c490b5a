"""-m gpu: the drop-in decoders (public API) against the reference's golden outputs and the oracle,
plus size-independent properties at the benchmark's full map size."""
import numpy as np
import pytest
import torch

from oracle import sdnet_oracle as O
from oracle import torch_port as TP
from structuredetector_b200 import CoreMLDecoder, Decoder, KeypointDecoder, ImageAnnotation, ops
from structuredetector_b200.synth import CONFIGS, DecodeConfig, make_raw, split_outputs
from tests.helpers import (assert_objects_close, golden_args, golden_names, listify, load_golden, make_args, np_inputs,
                           packed_np, plain, plain_keypoints, torch_sigmoid_fn)

pytestmark = pytest.mark.gpu
TIE_FREE = [n for n in golden_names() if not n.startswith(("ties", "half"))]


def _golden_outputs(name, device):
    meta, arr = load_golden(name)
    b, m, n, h, w = meta["shape"]
    return meta, arr, split_outputs(torch.from_numpy(arr["raw"]).to(device), m, n)


@pytest.mark.parametrize("name", TIE_FREE)
def test_decoder_matches_reference_golden(cuda_device, name):
    """Structure, names, ordering, grouping and indices exact; scores within 1e-6 (CPU vs CUDA
    sigmoid differ in the last bit), coordinates within 1e-5 relative -- north_star's bar."""
    meta, arr, outs = _golden_outputs(name, cuda_device)
    dec = Decoder(golden_args(meta))
    anns = dec(outs)
    assert all(isinstance(a, ImageAnnotation) for a in anns) and len(anns) == meta["shape"][0]
    assert [str(a.image_path) for a in anns] == [f"batch_{i}" for i in range(len(anns))]
    assert_objects_close(listify(plain(anns)), meta["annotation"], score_atol=1e-6, coord_rtol=1e-5, what=name)
    out = dec(outs, return_metadata=True)
    assert list(out) == ["annotation", "anchor_hm_sig", "part_hm_sig", "embeddings", "topk_anchor", "topk_kp",
                         "raw_parts", "raw_embeddings", "raw_offsets"]
    np.testing.assert_array_equal(out["topk_anchor"][1].cpu().numpy(), arr["a_inds"])
    np.testing.assert_array_equal(out["topk_kp"][1].cpu().numpy(), arr["p_inds"])
    np.testing.assert_array_equal(out["topk_anchor"][2].cpu().numpy(), arr["a_labels"])
    np.testing.assert_allclose(out["topk_anchor"][0].cpu().numpy(), arr["a_scores_masked"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(out["topk_kp"][0].cpu().numpy(), arr["p_scores_masked"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(out["topk_anchor"][4].cpu().numpy(), arr["a_xs"], rtol=1e-6)
    np.testing.assert_allclose(out["topk_kp"][3].cpu().numpy(), arr["p_ys"], rtol=1e-6)
    np.testing.assert_array_equal(out["embeddings"].cpu().numpy(), arr["embeddings"])
    assert out["raw_embeddings"] is outs["embeddings"] and out["raw_offsets"] is outs["offsets"]
    got_rp = [[[k.kind, k.x, k.y, k.score] for k in img] for img in out["raw_parts"]]
    assert [len(i) for i in got_rp] == [len(i) for i in meta["raw_parts"]]
    if "anchor_sig" in arr:
        np.testing.assert_allclose(out["anchor_hm_sig"].cpu().numpy(), arr["anchor_sig"], rtol=0, atol=1e-6)
    # per-call overrides behave like the reference's (decoders.py:33-38)
    none = dec(outs, conf_thresh=0.9999999)
    assert all(a.is_empty for a in none)


@pytest.mark.parametrize("name", TIE_FREE)
def test_coreml_and_keypoint_decoders_match_reference_golden(cuda_device, name):
    meta, arr, outs = _golden_outputs(name, cuda_device)
    args = golden_args(meta)
    kps = KeypointDecoder(args)(outs)
    got = listify(plain_keypoints(kps))
    assert [len(i) for i in got] == [len(i) for i in meta["keypoints"]]
    for gi, wi in zip(got, meta["keypoints"]):
        for g, w in zip(gi, wi):
            assert g[0] == w[0] and abs(g[3] - w[3]) <= 1e-6 and abs(g[1] - w[1]) <= 1e-5 * max(1, abs(w[1]))
    # CoreMLDecoder: feed maps activated + suppressed by stock torch ops on the device
    pre = dict(outs)
    pre["anchor_hm"] = TP.suppress(TP.activate(outs["anchor_hm"]))
    pre["part_hm"] = TP.suppress(TP.activate(outs["part_hm"]))
    cm = CoreMLDecoder(args)(pre, return_metadata=True)
    assert "anchor_hm_sig" not in cm
    np.testing.assert_array_equal(cm["topk_anchor"][1].cpu().numpy(), arr["coreml_a_inds"])
    np.testing.assert_array_equal(cm["topk_kp"][1].cpu().numpy(), arr["coreml_p_inds"])
    assert_objects_close(listify(plain(cm["annotation"])), meta["coreml_annotation"], score_atol=1e-6, coord_rtol=1e-5,
                         what=f"coreml {name}")


@pytest.mark.parametrize("name,mode", [("cfg1", "noise"), ("cfg1", "ties"), ("cfg4", "noise")])
def test_pre_activated_and_keypoint_paths_match_oracle_bit_exact(cuda_device, name, mode):
    cfg = CONFIGS[name]
    raw = make_raw(cfg, mode, batch=2)
    outs_cpu = split_outputs(raw, cfg.labels, cfg.parts)
    outs = split_outputs(raw.to(cuda_device), cfg.labels, cfg.parts)
    pre = dict(outs)
    pre["anchor_hm"] = TP.suppress(TP.activate(outs["anchor_hm"]))
    pre["part_hm"] = TP.suppress(TP.activate(outs["part_hm"]))
    got = packed_np(ops.decode_packed(pre, cfg.max_objects, cfg.max_parts, cfg.conf_threshold, cfg.dist_thresh,
                                      pre_activated=True))
    want = O.decode_packed(pre["anchor_hm"].cpu().numpy(), pre["part_hm"].cpu().numpy(), *np_inputs(outs_cpu)[2:],
                           cfg.max_objects, cfg.max_parts, cfg.conf_threshold, cfg.dist_thresh, pre_activated=True)
    for key in ("anchor_inds", "part_inds", "assign", "counts", "anchor_out", "part_out"):
        np.testing.assert_array_equal(got[key], want[key], err_msg=f"pre-activated {key}")
    args = make_args(cfg)
    kps = KeypointDecoder(args)(outs)
    want_kp = O.keypoint_decode(*np_inputs(outs_cpu)[:3], cfg.max_objects, cfg.max_parts, cfg.conf_threshold, 4.0,
                                args._r_labels, args._r_parts, sigmoid_fn=torch_sigmoid_fn(cuda_device))
    assert plain_keypoints(kps) == want_kp


def test_full_size_properties(cuda_device):
    """cfg5-sized maps (512 x 612), a 48-image shard: bit-exact against the reference's own op
    sequence on the device, plus properties that do not need an oracle."""
    cfg = CONFIGS["cfg5"]
    raw = make_raw(cfg, "noise", batch=12).to(cuda_device)
    raw = raw[torch.arange(48, device=cuda_device) % 12].contiguous()
    outs = split_outputs(raw, cfg.labels, cfg.parts)
    pk = ops.decode_packed(outs, cfg.max_objects, cfg.max_parts, cfg.conf_threshold, cfg.dist_thresh)
    ref = TP.decode_tensors(outs, cfg.max_objects, cfg.max_parts, cfg.conf_threshold, cfg.dist_thresh)
    for key in ("anchor_inds", "part_inds", "anchor_out", "part_out", "assign", "counts"):
        assert torch.equal(getattr(pk, key), ref[key].to(getattr(pk, key).dtype)), key
    # sortedness: scores never increase along the slots
    assert bool((pk.anchor_out[:, 1:, 2] <= pk.anchor_out[:, :-1, 2]).all())
    assert bool((pk.part_out[:, 1:, 2] <= pk.part_out[:, :-1, 2]).all())
    # images are independent: identical inputs give identical outputs wherever they sit in the batch
    assert torch.equal(pk.anchor_inds[:12], pk.anchor_inds[36:48]) and torch.equal(pk.assign[:12], pk.assign[24:36])
    # batch permutation commutes with decoding
    perm = torch.randperm(48, device=cuda_device, generator=torch.Generator(device=cuda_device).manual_seed(3))
    pk2 = ops.decode_packed(split_outputs(raw[perm].contiguous(), cfg.labels, cfg.parts), cfg.max_objects, cfg.max_parts,
                            cfg.conf_threshold, cfg.dist_thresh)
    assert torch.equal(pk2.part_out, pk.part_out[perm]) and torch.equal(pk2.assign, pk.assign[perm])
    # the bounded-memory exact select and the any-alignment kernel agree with the TMA kernel
    for kw in ({"exact_select": True}, {"warp_kernel": True}):
        alt = ops.decode_packed(outs, cfg.max_objects, cfg.max_parts, cfg.conf_threshold, cfg.dist_thresh, **kw)
        assert torch.equal(alt.anchor_inds, pk.anchor_inds) and torch.equal(alt.part_out, pk.part_out)
    # every grouped part points at an anchor that is itself above the threshold
    slots = pk.assign.long().clamp(min=0)
    anchor_score = torch.gather(pk.anchor_out[..., 2], 1, slots)
    assert bool(((pk.assign < 0) | (anchor_score > np.float32(cfg.conf_threshold))).all())


def test_host_buffer_entry_point_matches_device_path(cuda_device):
    cfg = CONFIGS["cfg3"]
    raw = make_raw(cfg, "blobs", batch=3)
    host = raw.pin_memory()
    M, N, H, W, K, P = cfg.labels, cfg.parts, cfg.height, cfg.width, cfg.max_objects, cfg.max_parts
    dev_outs = split_outputs(raw.to(cuda_device), M, N)
    want = ops.decode_packed(dev_outs, K, P, cfg.conf_threshold, cfg.dist_thresh)
    plan = ops.DecodePlan(cuda_device, 3, M, N, H, W, K, P)
    staging = torch.empty(3 * (M + N) * H * W * 4, dtype=torch.uint8, device=cuda_device)
    h = split_outputs(host, M, N)
    got = plan.run_host(h["anchor_hm"], h["part_hm"], h["offsets"], h["embeddings"], ops._f32(cfg.conf_threshold),
                        ops._f32(cfg.dist_thresh * min(W, H)), staging)
    torch.cuda.synchronize()
    for key in ("anchor_inds", "part_inds", "anchor_out", "part_out", "assign", "counts"):
        assert torch.equal(getattr(got, key), getattr(want, key)), key


def test_concurrent_streams_do_not_interfere(cuda_device):
    cfg = CONFIGS["cfg2"]
    raws = [make_raw(cfg, "noise", batch=8, seed=100 + i).to(cuda_device) for i in range(4)]
    want = [ops.decode_packed(split_outputs(r, cfg.labels, cfg.parts), 100, 100, 0.4, 0.1) for r in raws]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(device=cuda_device) for _ in raws]
    got = []
    for r, s in zip(raws, streams):
        with torch.cuda.stream(s):
            got.append(ops.decode_packed(split_outputs(r, cfg.labels, cfg.parts), 100, 100, 0.4, 0.1))
    torch.cuda.synchronize()
    for g, w in zip(got, want):
        assert torch.equal(g.blob[: g.anchor_inds.numel() * 16], w.blob[: w.anchor_inds.numel() * 16])
        assert torch.equal(g.part_out, w.part_out) and torch.equal(g.assign, w.assign)
