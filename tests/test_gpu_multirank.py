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
FUSED_RUNS = 5
# (SDNET_GATHER_SYNC, SDNET_GATHER_STORES, lazy arrival waits)
FUSED_VARIANTS = (('flags', 'auto', False), ('barrier', 'auto', False), ('flags', 'peer', False), ('flags', 'auto', True))


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

        # Repeated runs reuse the plan's two symmetric result buffers.  Every run decodes DIFFERENT inputs
        # and its result is read (copied on the run stream) before the next run is enqueued, with no host
        # synchronisation in between and rank 1 deliberately slowed down: a missing write-after-read guard
        # would let the fast rank's next run overwrite rows the slow rank has not copied yet.
        # Every way of getting the rows across (one multimem.st per value / one store per peer) and of knowing they
        # have arrived (completion flags / symmetric-memory barrier) has to give the same bits.
        keys = ("anchor_out", "part_out", "anchor_inds", "part_inds", "assign")
        for variant, (sync, stores, lazy) in enumerate(FUSED_VARIANTS):
            os.environ.update(SDNET_GATHER_SYNC=sync, SDNET_GATHER_STORES=stores)
            fp = parallel.FusedGatherPlan(f"cuda:{rank}", BATCH, cfg.labels, cfg.parts, cfg.height, cfg.width,
                                          cfg.max_objects, cfg.max_parts)
            assert fp.sync == sync and (stores == "auto" or fp.stores == stores)
            fp.lazy = lazy  # the reader then asks for arrival itself, on the run stream, before it copies
            copies = []
            for it in range(FUSED_RUNS):
                raw_it = make_raw(cfg, "noise", batch=BATCH, seed=500 + it)
                o = split_outputs(raw_it[lo:hi].to(f"cuda:{rank}"), cfg.labels, cfg.parts)
                res = fp.run(o["anchor_hm"], o["part_hm"], o["offsets"], o["embeddings"],
                             ops._f32(cfg.conf_threshold), ops._f32(cfg.dist_thresh * min(cfg.width, cfg.height)))
                if rank == 1:
                    torch.cuda._sleep(20_000_000)  # ~10 ms of device time before this rank reads its copy
                fp.wait_arrival()
                copies.append({k: getattr(res, k).clone() for k in keys})
            torch.cuda.synchronize()
            torch.save([{k: v.cpu() for k, v in c.items()} for c in copies], os.path.join(out_dir, f"fused{variant}_{rank}.pt"))
            dist.barrier()
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
    want_runs = []
    for it in range(FUSED_RUNS):
        raw_it = make_raw(cfg, "noise", batch=BATCH, seed=500 + it)
        w = ops.decode_packed(split_outputs(raw_it.to(cuda_device), cfg.labels, cfg.parts), cfg.max_objects,
                              cfg.max_parts, cfg.conf_threshold, cfg.dist_thresh)
        want_runs.append({k: getattr(w, k).cpu() for k in ("anchor_out", "part_out", "anchor_inds", "part_inds", "assign")})
    for rank in range(world):
        got = torch.load(os.path.join(tmp_path, f"rank{rank}.pt"))
        for key, val in got.items():
            assert torch.equal(val, getattr(want, key).cpu()), f"nccl gather, rank {rank}: {key}"
        for variant, names in enumerate(FUSED_VARIANTS):
            runs = torch.load(os.path.join(tmp_path, f"fused{variant}_{rank}.pt"))
            assert len(runs) == FUSED_RUNS
            for it, (g, w) in enumerate(zip(runs, want_runs)):
                for key, val in g.items():
                    assert torch.equal(val, w[key]), f"fused gather {names}, rank {rank}, run {it}: {key}"
