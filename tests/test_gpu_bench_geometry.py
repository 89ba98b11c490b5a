"""-m gpu: parity at the geometry bench.py actually runs.

BASELINE config 5 decodes 1024 / 512 / 256 / 128 images of (7, 512, 612) maps per GPU.  At those sizes the
peaks kernel's work line has more columns than the device has resident warps, so the schedule uses BOTH
tiers (whole-column units with long-lived pruning floors, then balanced chunks that straddle column
ends); below ~200 images everything is chunks.  The small parity cases never reach that, so here the
same launches the bench times are compared, every packed field bit for bit, with the reference's own op
sequence on the device (oracle/torch_port.py on CUDA tensors, K = 100 > 32: canonical topk order), and
the schedule that ran is asserted through ``sdnet_decode_schedule``."""
import pytest
import torch

from oracle import torch_port as TP
from structuredetector_b200 import ops
from structuredetector_b200.synth import CONFIGS, make_raw, split_outputs

pytestmark = pytest.mark.gpu
UNIQUE = 32  # distinct images, as in bench.py; tiled with a per-copy rotation so no two planes are queued alike
FIELDS = ("anchor_inds", "part_inds", "assign", "counts", "anchor_out", "part_out", "part_emb")


def shard(device, mode, images, dtype, seed=None):
    cfg = CONFIGS["cfg5"]
    uniq = make_raw(cfg, mode, batch=min(UNIQUE, images), seed=seed).to(device)
    idx = (torch.arange(images, device=device) * 7 + 3) % uniq.shape[0]  # 7 is coprime with 32: every image appears
    return cfg, uniq[idx].contiguous().to(dtype)


def port_in_chunks(outs, cfg, chunk=32):
    parts = []
    n = outs["anchor_hm"].shape[0]
    for lo in range(0, n, chunk):
        sub = {k: v[lo:lo + chunk] for k, v in outs.items()}
        ref = TP.decode_tensors(sub, cfg.max_objects, cfg.max_parts, cfg.conf_threshold, cfg.dist_thresh)
        parts.append({k: ref[k] for k in FIELDS})
    return {k: torch.cat([p[k] for p in parts]) for k in FIELDS}


def check(device, mode, images, dtype, want_tier1, want_chunks):
    cfg, raw = shard(device, mode, images, dtype)
    outs = split_outputs(raw, cfg.labels, cfg.parts)
    M, N, H, W, K, P = cfg.labels, cfg.parts, cfg.height, cfg.width, cfg.max_objects, cfg.max_parts
    plan = ops.DecodePlan(device, images, M, N, H, W, K, P, dtype)
    sched = plan.schedule(outs["anchor_hm"], outs["part_hm"], outs["offsets"], outs["embeddings"])
    assert sched["path"] == ("tile" if dtype == torch.float32 else "tile_row_pairs")
    assert (sched["tier1_units"] > 0) == want_tier1, sched
    assert (sched["chunk_units"] > 0) == want_chunks, sched
    if want_chunks:  # chunks are shorter than a column, i.e. units start and end inside columns
        assert sched["chunk_groups"] < sched["groups_per_column"], sched
    conf = float(torch.tensor(cfg.conf_threshold, dtype=dtype))
    got = plan.run(outs["anchor_hm"], outs["part_hm"], outs["offsets"], outs["embeddings"], conf,
                   ops._f32(cfg.dist_thresh * min(W, H)))
    torch.cuda.synchronize()
    assert int(got.diag[:, 1].sum()) == 0, "no plane should need the exact select on these inputs"
    want = port_in_chunks(outs, cfg)
    for key in FIELDS:
        g, w = getattr(got, key), want[key]
        assert torch.equal(g, w.to(g.dtype)), f"{mode} {images} {dtype}: {key}"
    return sched


@pytest.mark.parametrize("mode", ["noise", "blobs"])
@pytest.mark.parametrize("images", [256, 1024])
def test_two_tier_schedule_fp32(cuda_device, mode, images):
    check(cuda_device, mode, images, torch.float32, want_tier1=True, want_chunks=True)


@pytest.mark.parametrize("mode", ["noise", "blobs"])
def test_chunk_only_schedule_fp32(cuda_device, mode):
    """The 8-GPU shard of config 5 (128 images): fewer columns than resident warps, one wave of chunks."""
    sched = check(cuda_device, mode, 128, torch.float32, want_tier1=False, want_chunks=True)
    assert sched["units"] <= sched["ctas"] * sched["warps_per_cta"]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("mode,images", [("noise", 512), ("blobs", 128)])
def test_bench_shards_reduced_precision(cuda_device, dtype, mode, images):
    check(cuda_device, mode, images, dtype, want_tier1=images > 256, want_chunks=True)
