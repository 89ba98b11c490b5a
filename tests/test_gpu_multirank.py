"""-m gpu, needs >= 2 GPUs (skipped otherwise): the sharded decode over NCCL returns, on every rank,
exactly what one GPU returns for the whole batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from structuredetector_b200.synth import CONFIGS, make_raw, split_outputs

pytestmark = pytest.mark.gpu
BATCH = 10  # uneven over 4 ranks, even over 2


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from structuredetector_b200 import parallel

        cfg = CONFIGS["cfg2"]
        raw = make_raw(cfg, "noise", batch=BATCH)
        lo, hi = parallel.shard_bounds(BATCH, world, rank)
        outs = split_outputs(raw[lo:hi].to(f"cuda:{rank}"), cfg.labels, cfg.parts)
        merged = parallel.ShardedDecoder(cfg.max_objects, cfg.max_parts, cfg.conf_threshold, cfg.dist_thresh)(outs, BATCH)
        torch.save({k: getattr(merged, k).cpu() for k in ("anchor_out", "part_out", "anchor_inds", "part_inds", "assign")},
                   os.path.join(out_dir, f"rank{rank}.pt"))
        # the fused path: tail kernel stores into every peer's copy + one barrier
        from structuredetector_b200 import ops

        fp = parallel.FusedGatherPlan(f"cuda:{rank}", BATCH, cfg.labels, cfg.parts, cfg.height, cfg.width,
                                      cfg.max_objects, cfg.max_parts)
        for _ in range(3):  # repeated runs reuse the symmetric blob
            res = fp.run(outs["anchor_hm"], outs["part_hm"], outs["offsets"], outs["embeddings"],
                         ops._f32(cfg.conf_threshold), ops._f32(cfg.dist_thresh * min(cfg.width, cfg.height)))
        torch.cuda.synchronize()
        torch.save({k: getattr(res, k).cpu() for k in ("anchor_out", "part_out", "anchor_inds", "part_inds", "assign")},
                   os.path.join(out_dir, f"fused{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_decode_equals_single_gpu(cuda_device, tmp_path, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from structuredetector_b200 import ops

    cfg = CONFIGS["cfg2"]
    raw = make_raw(cfg, "noise", batch=BATCH)
    want = ops.decode_packed(split_outputs(raw.to(cuda_device), cfg.labels, cfg.parts), cfg.max_objects, cfg.max_parts,
                             cfg.conf_threshold, cfg.dist_thresh)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for rank in range(world):
        for kind in ("rank", "fused"):
            got = torch.load(os.path.join(tmp_path, f"{kind}{rank}.pt"))
            for key, val in got.items():
                assert torch.equal(val, getattr(want, key).cpu()), f"{kind} {rank}: {key}"
