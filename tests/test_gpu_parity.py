"""-m gpu: the CUDA path (through the C ABI) against the oracle on identical inputs.

Bar: bit-exact for indices, slot order, classes, part->anchor assignment and -- because the
oracle is fed the device's own sigmoid -- for scores and coordinates too.
"""
import numpy as np
import pytest
import torch

from oracle import sdnet_oracle as O
from oracle import torch_port as TP
from structuredetector_b200 import ops
from structuredetector_b200.synth import CONFIGS, DecodeConfig, make_raw, split_outputs
from tests.helpers import assert_packed_equal, np_inputs, packed_np, torch_sigmoid_fn

pytestmark = pytest.mark.gpu


def _decode_both(cfg, raw_cpu, device, *, exact_select=False, radius=2, conf=None, dist=None, warp_kernel=False):
    conf = cfg.conf_threshold if conf is None else conf
    dist = cfg.dist_thresh if dist is None else dist
    outs_cpu = split_outputs(raw_cpu, cfg.labels, cfg.parts)
    outs_gpu = split_outputs(raw_cpu.to(device), cfg.labels, cfg.parts)
    got = ops.decode_packed(outs_gpu, cfg.max_objects, cfg.max_parts, conf, dist, exact_select=exact_select,
                            radius=radius, warp_kernel=warp_kernel)
    torch.cuda.synchronize()
    want = O.decode_packed(*np_inputs(outs_cpu), cfg.max_objects, cfg.max_parts, conf, dist,
                           sigmoid_fn=torch_sigmoid_fn(device), radius=radius)
    return packed_np(got), want


def test_activate_bit_exact_vs_torch_cuda(cuda_device):
    g = torch.Generator().manual_seed(11)
    x = torch.cat([torch.randn(1 << 20, generator=g) * 6, torch.linspace(-20, 20, 1 << 16),
                   torch.tensor([0.0, -0.0, 13.8, 13.81, 13.82, 14.0, 30.0, 88.0, 100.0, -13.8, -13.82, -14.0, -30.0,
                                 -88.0, -100.0, 1e-30, -1e-30])])
    pad = (-x.numel()) % 64
    x = torch.cat([x, torch.zeros(pad)]).view(1, 1, -1, 64).to(cuda_device)
    got = ops.activate_maps(x)
    want = torch.clamp(torch.sigmoid(x), min=1e-6, max=1 - 1e-6)
    assert torch.equal(got.view(torch.int32), want.view(torch.int32))


def test_score_function_is_monotone_and_margins_hold(cuda_device):
    """The kernel runs its stencil on logits; that is exact iff S is monotone and the
    near-tie margins in sdnet_decode.cu (kNearTie, kHiZone, kLoZone, kSatX) are safe."""
    import struct
    f2b = lambda f: struct.unpack("<I", struct.pack("<f", f))[0]
    step = 1 << 24

    def scores(bits_lo, bits_hi):  # all floats with bit patterns in [lo, hi)
        bits = torch.arange(bits_lo, bits_hi, device=cuda_device, dtype=torch.int64)
        bits = torch.where(bits >= (1 << 31), bits - (1 << 32), bits).to(torch.int32)
        x = bits.view(torch.float32)
        n = x.numel()
        pad = (-n) % 1024
        xp = torch.cat([x, x[-1:].expand(pad)]) if pad else x
        return x, ops.activate_maps(xp.view(1, 1, -1, 1024)).view(-1)[:n]

    prev = None
    for s in range(0, f2b(20.0) + 1, step):  # x >= 0, increasing
        x, y = scores(s, min(s + step, f2b(20.0) + 1))
        assert bool((y[1:] >= y[:-1]).all())
        if prev is not None:
            assert float(y[0]) >= prev
        prev = float(y[-1])
        # margins on the positive side
        mid = (x <= 8.0) & (x >= 2e-3)
        lo = ops.activate_maps((x - 2e-3).view(1, 1, 1, -1)).view(-1)
        assert bool((lo[mid] < y[mid]).all())
    prev = None
    for s in range(0x80000000, f2b(-20.0) + 1, step):  # x <= 0, decreasing with the bit pattern
        x, y = scores(s, min(s + step, f2b(-20.0) + 1))
        assert bool((y[1:] <= y[:-1]).all())
        if prev is not None:
            assert float(y[0]) <= prev
        prev = float(y[-1])
        mid = x >= -13.0
        lo = ops.activate_maps((x - 2e-3).view(1, 1, 1, -1)).view(-1)
        assert bool((lo[mid] < y[mid]).all())
    pts = ops.activate_maps(torch.tensor([7.0, 8.0, 14.0, 1e9, -14.0, -1e9], device=cuda_device).view(1, 1, 1, -1)).view(-1)
    assert float(pts[0]) < float(pts[1])
    assert float(pts[2]) == float(pts[3]) and float(pts[4]) == float(pts[5])


CASES = [
    ("cfg1", "noise", 3), ("cfg1", "blobs", 3), ("cfg1", "ladder", 2), ("cfg1", "ties", 3),
    ("cfg2", "noise", 64), ("cfg2", "blobs", 64),
    ("cfg3", "noise", 2), ("cfg3", "blobs", 2), ("cfg3", "ties", 1),
    ("cfg4", "noise", 2), ("cfg4", "ties", 1),
]


@pytest.mark.parametrize("warp_kernel", [False, True], ids=["tma", "warp"])
@pytest.mark.parametrize("name,mode,batch", CASES)
def test_decode_matches_oracle(cuda_device, name, mode, batch, warp_kernel):
    cfg = CONFIGS[name]
    raw = make_raw(cfg, mode, batch=batch)
    got, want = _decode_both(cfg, raw, cuda_device, warp_kernel=warp_kernel)
    assert_packed_equal(got, want, what=f"{name}/{mode}")
    # the TMA kernel's lists must not overflow on tie-free inputs; the per-lane fallback cuts small batches into
    # many short strips that all warm up at once and may hand a plane to the exact select (same result, checked above)
    assert int(got["diag"][:, 1].sum()) == 0 or mode == "ties" or warp_kernel
    # every config's fp32 maps are 16-byte aligned: the default path is the TMA tile kernel
    outs = split_outputs(raw[:1].to(cuda_device), cfg.labels, cfg.parts)
    assert ops.peaks_path(outs, cfg.max_objects, cfg.max_parts, warp_kernel=warp_kernel) == ("warp" if warp_kernel else "tile")


@pytest.mark.parametrize("name,mode,batch", [("cfg1", "noise", 2), ("cfg1", "ties", 2), ("cfg3", "blobs", 1), ("cfg4", "noise", 1)])
def test_exact_select_path_matches_oracle(cuda_device, name, mode, batch):
    cfg = CONFIGS[name]
    raw = make_raw(cfg, mode, batch=batch)
    got, want = _decode_both(cfg, raw, cuda_device, exact_select=True)
    assert_packed_equal(got, want, what=f"exact {name}/{mode}")
    assert int(got["diag"][:, 1].min()) == 1


@pytest.mark.parametrize("h,w,k,p", [(5, 5, 25, 25), (7, 9, 10, 63), (37, 53, 40, 20), (64, 130, 100, 100),
                                     (33, 257, 64, 64), (200, 4, 50, 50), (1, 300, 30, 30), (300, 1, 30, 30),
                                     (3, 128, 20, 20), (41, 132, 60, 60), (70, 896, 100, 100), (40, 900, 100, 100),
                                     (35, 1028, 100, 100), (9, 8, 72, 10)])
@pytest.mark.parametrize("mode", ["noise", "ties"])
def test_odd_shapes(cuda_device, h, w, k, p, mode):
    cfg = DecodeConfig("odd", 2, 3, 2, h, w, k, p, cfg_id=9)
    raw = make_raw(cfg, mode)
    got, want = _decode_both(cfg, raw, cuda_device)
    assert_packed_equal(got, want, what=f"odd {h}x{w}/{mode}")


@pytest.mark.parametrize("warp_kernel", [False, True], ids=["tma", "warp"])
@pytest.mark.parametrize("name", ["cfg1", "cfg3"])
def test_radius_one(cuda_device, name, warp_kernel):
    cfg = CONFIGS[name]
    raw = make_raw(cfg, "noise", batch=2)
    got, want = _decode_both(cfg, raw, cuda_device, radius=1, warp_kernel=warp_kernel)
    assert_packed_equal(got, want, what="radius 1")


@pytest.mark.parametrize("name,mode,batch", [("cfg3", "noise", 4), ("cfg4", "noise", 2), ("cfg2", "blobs", 16)])
def test_decode_matches_torch_cuda_port(cuda_device, name, mode, batch):
    """The reference's own op sequence on CUDA tensors (k > 32: torch's topk is canonical)."""
    cfg = CONFIGS[name]
    raw = make_raw(cfg, mode, batch=batch).to(cuda_device)
    outs = split_outputs(raw, cfg.labels, cfg.parts)
    got = packed_np(ops.decode_packed(outs, cfg.max_objects, cfg.max_parts, cfg.conf_threshold, cfg.dist_thresh))
    want = {k: v.cpu().numpy() for k, v in
            TP.decode_tensors(outs, cfg.max_objects, cfg.max_parts, cfg.conf_threshold, cfg.dist_thresh).items()}
    assert_packed_equal(got, want, what=f"torch-cuda {name}/{mode}")


def test_saturated_planes_take_the_exact_path(cuda_device):
    cfg = DecodeConfig("sat", 2, 1, 1, 96, 160, 50, 50, cfg_id=10)
    raw = make_raw(cfg, "noise")
    raw[:, 0] = 3.0 + 0.0 * raw[:, 0]          # one mid-value plateau covering a whole plane
    raw[1, 1] = -20.0                           # a plane entirely at the low clamp
    raw[0, 1, 10:60, 20:140] = 25.0             # a big high-clamp plateau
    got, want = _decode_both(cfg, raw, cuda_device)
    assert_packed_equal(got, want, what="saturated")


def test_strided_channel_views_and_errors(cuda_device):
    cfg = CONFIGS["cfg1"]
    raw = make_raw(cfg, "blobs", batch=2).to(cuda_device)
    outs = split_outputs(raw, cfg.labels, cfg.parts)
    assert not outs["part_hm"].is_contiguous()
    with pytest.raises(RuntimeError):
        ops.decode_packed({k: v.cpu() for k, v in outs.items()}, 100, 100, 0.4, 0.1)
    with pytest.raises(TypeError):
        ops.decode_packed({k: v.double() for k, v in outs.items()}, 100, 100, 0.4, 0.1)
    with pytest.raises(RuntimeError):
        ops.decode_packed(outs, 128 * 128 + 1, 100, 0.4, 0.1)


def test_randomised_shapes_and_thresholds(cuda_device):
    """Seeded fuzz over shapes, class counts, K/P, thresholds, radius and input modes: every case
    bit-exact against the oracle fed the device's sigmoid."""
    rng = np.random.default_rng(20260118)
    modes = ("noise", "blobs", "ties", "ladder")
    for case in range(40):
        h = int(rng.integers(1, 90))
        w = int(rng.choice([rng.integers(1, 40), rng.integers(40, 300), 4 * rng.integers(1, 80)]))
        m, n = int(rng.integers(1, 5)), int(rng.integers(1, 4))
        k = int(rng.integers(1, min(h * w, 120) + 1))
        p = int(rng.integers(1, min(h * w, 120) + 1))
        cfg = DecodeConfig(f"fuzz{case}", int(rng.integers(1, 4)), m, n, h, w, k, p, cfg_id=200 + case)
        mode = modes[case % len(modes)]
        raw = make_raw(cfg, mode)
        conf = float(rng.choice([0.05, 0.3, 0.4, 0.5, 0.9]))
        dist = float(rng.choice([0.02, 0.1, 0.5, 2.0]))
        radius = 2 if case % 5 else 1
        got, want = _decode_both(cfg, raw, cuda_device, conf=conf, dist=dist, radius=radius,
                                 warp_kernel=bool(case % 7 == 3), exact_select=bool(case % 11 == 5))
        assert_packed_equal(got, want, what=f"fuzz {case}: {h}x{w} M{m} N{n} K{k} P{p} {mode} conf{conf} dist{dist} r{radius}")


def test_pipeline_of_decodes_matches_serial(cuda_device):
    """ops.DecodePipeline: three decodes in flight on three streams give what one-at-a-time decodes give."""
    cfg = CONFIGS["cfg2"]
    raws = [make_raw(cfg, mode, batch=4).to(cuda_device) for mode in ("noise", "blobs", "ties", "noise", "blobs")]
    conf32, dist32 = ops._f32(cfg.conf_threshold), ops._f32(cfg.dist_thresh * min(cfg.width, cfg.height))
    pipe = ops.DecodePipeline.for_shape(cuda_device, 3, 4, cfg.labels, cfg.parts, cfg.height, cfg.width, cfg.max_objects,
                                        cfg.max_parts)
    done = torch.cuda.Event()
    done.record()
    pipe.after(done)
    got = []
    for raw in raws:
        o = split_outputs(raw, cfg.labels, cfg.parts)
        out, ev = pipe.submit(o["anchor_hm"], o["part_hm"], o["offsets"], o["embeddings"], conf32, dist32)
        ev.synchronize()
        got.append(out.blob.clone())
    for raw, blob in zip(raws, got):
        want = ops.decode_packed(split_outputs(raw, cfg.labels, cfg.parts), cfg.max_objects, cfg.max_parts,
                                 cfg.conf_threshold, cfg.dist_thresh)
        torch.cuda.synchronize()
        n = blob.numel() - want.diag.numel() * 4  # everything but the trailing diagnostics is deterministic
        assert torch.equal(blob[:n], want.blob[:n])


def test_workspace_is_left_clean_between_decodes(cuda_device):
    """From its second run on a DecodePlan skips the per-call memset (SDNET_FLAG_WORKSPACE_CLEAN): every decode has to
    leave the workspace header as a memset would.  One plan whose workspace starts as garbage decodes different inputs
    run after run -- among them a batch whose planes overflow into the exact select, and forced exact selects -- and
    every run has to equal what a one-off decode of the same inputs gives."""
    cfg = DecodeConfig("clean", 3, 2, 1, 96, 160, 50, 50, cfg_id=11)
    sat = make_raw(cfg, "noise", seed=77)
    sat[:, 0] = 3.0 + 0.0 * sat[:, 0]
    sat[1, 2, 8:70, 16:150] = 25.0
    runs = [(make_raw(cfg, "noise", seed=71), 0), (make_raw(cfg, "blobs", seed=72), 0), (sat, 0),
            (make_raw(cfg, "ties", seed=73), 0), (make_raw(cfg, "noise", seed=74), ops.FLAG_EXACT_SELECT),
            (make_raw(cfg, "ladder", seed=75), 0), (make_raw(cfg, "noise", seed=71), 0)]
    conf32, dist32 = ops._f32(cfg.conf_threshold), ops._f32(cfg.dist_thresh * min(cfg.width, cfg.height))
    plan = ops.DecodePlan(cuda_device, cfg.batch, cfg.labels, cfg.parts, cfg.height, cfg.width, cfg.max_objects, cfg.max_parts)
    plan.workspace.fill_(0xA5)
    for it, (raw, flags) in enumerate(runs):
        o = split_outputs(raw.to(cuda_device), cfg.labels, cfg.parts)
        got = {k: v.clone() for k, v in plan.run(o["anchor_hm"], o["part_hm"], o["offsets"], o["embeddings"], conf32, dist32,
                                                 flags=flags).as_dict().items()}
        assert plan._clean and bool(plan.params.flags & ops.FLAG_WORKSPACE_CLEAN) == (it > 0)
        once = ops.DecodePlan(cuda_device, cfg.batch, cfg.labels, cfg.parts, cfg.height, cfg.width, cfg.max_objects, cfg.max_parts)
        want = once.run(o["anchor_hm"], o["part_hm"], o["offsets"], o["embeddings"], conf32, dist32, flags=flags)
        torch.cuda.synchronize()
        for key, val in want.as_dict().items():  # field by field: the blob has alignment padding nobody writes
            if key != "diag":  # candidates per plane depend on timing
                assert torch.equal(got[key], val), f"run {it}: {key}"
        assert torch.equal(got["diag"][:, 1], want.diag[:, 1]), f"run {it}: which planes went through the exact select"
    planes = cfg.batch * (cfg.labels + cfg.parts)
    assert int(plan.workspace.view(torch.int32)[:planes].abs().sum()) == 0, "candidate counts left behind"
