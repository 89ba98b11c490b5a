"""CPU, world_size 2 over gloo: the host logic of the multi-GPU path (contiguous sharding,
fixed-capacity packed blobs, one all-gather, stitching).  The per-rank "decode" is played by the
oracle here -- the CUDA path itself is covered by the -m gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sdnet_oracle as O
from structuredetector_b200 import ops, parallel
from structuredetector_b200.synth import DecodeConfig, make_raw, split_outputs
from tests.helpers import np_inputs

CFG = DecodeConfig("dist", 5, 2, 1, 24, 36, 12, 16, cfg_id=55)  # 5 images over 2 ranks: uneven split 3 + 2


def _pack(pk: dict, n: int) -> torch.Tensor:
    """Oracle result -> the same byte layout the C ABI writes (ops._carve)."""
    C = CFG.labels + CFG.parts
    blob = torch.zeros(ops.packed_nbytes(n, CFG.max_objects, CFG.max_parts, C), dtype=torch.uint8)
    view = ops._carve(blob, n, CFG.max_objects, CFG.max_parts, C)
    view.anchor_out.copy_(torch.from_numpy(pk["anchor_out"]))
    view.part_out.copy_(torch.from_numpy(pk["part_out"]))
    view.anchor_inds.copy_(torch.from_numpy(pk["anchor_inds"]))
    view.part_inds.copy_(torch.from_numpy(pk["part_inds"]))
    view.part_emb.copy_(torch.from_numpy(pk["part_emb"]))
    view.assign.copy_(torch.from_numpy(pk["assign"]))
    view.counts.copy_(torch.from_numpy(pk["counts"]))
    return blob


def _oracle(raw):
    outs = split_outputs(raw, CFG.labels, CFG.parts)
    return O.decode_packed(*np_inputs(outs), CFG.max_objects, CFG.max_parts, CFG.conf_threshold, CFG.dist_thresh)


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        raw = make_raw(CFG, "noise")
        lo, hi = parallel.shard_bounds(CFG.batch, world, rank)
        local = _pack(_oracle(raw[lo:hi]), hi - lo)
        sizes = parallel.shard_sizes(CFG.batch, world)
        merged = parallel.all_gather_packed(local, sizes, CFG.max_objects, CFG.max_parts, CFG.labels + CFG.parts)
        torch.save({k: getattr(merged, k) for k in ("anchor_out", "part_out", "anchor_inds", "part_inds", "assign", "counts")},
                   os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_shard_split_is_contiguous_and_complete():
    for total in (1, 5, 16, 1024):
        for world in (1, 2, 3, 4, 8):
            bounds = [parallel.shard_bounds(total, world, r) for r in range(world)]
            assert bounds[0][0] == 0 and bounds[-1][1] == total
            assert all(bounds[r][1] == bounds[r + 1][0] for r in range(world - 1))
            assert max(h - l for l, h in bounds) - min(h - l for l, h in bounds) <= 1


def test_two_rank_gather_equals_single_rank_result(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    want = _oracle(make_raw(CFG, "noise"))
    for rank in range(2):
        got = torch.load(os.path.join(tmp_path, f"rank{rank}.pt"))
        for key in ("anchor_out", "part_out", "anchor_inds", "part_inds", "assign", "counts"):
            np.testing.assert_array_equal(got[key].numpy(), want[key], err_msg=f"rank {rank}: {key}")
