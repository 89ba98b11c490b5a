"""Batch sharding across GPUs and the one collective of the path: gathering the packed detections.

Images are independent (the reference never mixes batch entries: src/sdnet/data/decoders.py:44-100
is batched on dim 0 and the object loop iterates images, :104), so the only parallelism is a
contiguous split of the batch over ranks plus ONE all-gather of fixed-capacity packed records at
the end.  One process per GPU; ``torch.distributed`` (NCCL on GPUs, gloo in the CPU
tests) is plumbing only.
"""
from __future__ import annotations

import ctypes
import os

import torch
import torch.distributed as dist

from . import _native, ops

__all__ = ["shard_bounds", "shard_sizes", "all_gather_packed", "merge_packed", "ShardedDecoder", "FusedGatherPlan"]

_FIELDS = ("anchor_out", "part_out", "anchor_inds", "part_inds", "part_emb", "assign", "counts", "diag")


def shard_sizes(total: int, world: int) -> list[int]:
    """Contiguous split, the first ``total % world`` ranks take one extra image."""
    base, extra = divmod(total, world)
    return [base + (1 if r < extra else 0) for r in range(world)]


def shard_bounds(total: int, world: int, rank: int) -> tuple[int, int]:
    sizes = shard_sizes(total, world)
    lo = sum(sizes[:rank])
    return lo, lo + sizes[rank]


def merge_packed(blobs: list[torch.Tensor], sizes: list[int], K: int, P: int, C: int) -> ops.PackedDetections:
    """Per-rank packed blobs (rank order) -> one PackedDetections for the whole batch."""
    parts = [ops._carve(blob, n, K, P, C) for blob, n in zip(blobs, sizes) if n > 0]
    fields = {name: torch.cat([getattr(p, name) for p in parts], dim=0) for name in _FIELDS}
    return ops.PackedDetections(**fields, blob=None)


def all_gather_packed(local_blob: torch.Tensor, sizes: list[int], K: int, P: int, C: int, group=None) -> ops.PackedDetections:
    """All-gather every rank's packed blob and stitch the global result (identical on all ranks).

    Blobs have a fixed capacity per image, so no size exchange is needed; with an uneven split the
    shorter blobs are padded to the longest one for the collective.
    """
    world = dist.get_world_size(group)
    assert len(sizes) == world
    nbytes = [ops.packed_nbytes(n, K, P, C) for n in sizes]
    width = max(nbytes)
    send = local_blob
    if local_blob.numel() != width:
        send = torch.zeros(width, dtype=torch.uint8, device=local_blob.device)
        send[: local_blob.numel()] = local_blob
    recv = torch.empty(world * width, dtype=torch.uint8, device=local_blob.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    blobs = [recv[r * width : r * width + nbytes[r]] for r in range(world)]
    return merge_packed(blobs, sizes, K, P, C)


class ShardedDecoder:
    """Decode this rank's contiguous slice of a global batch on its GPU and all-gather the packed
    detections; ``rank``/``world`` come from the default process group."""

    def __init__(self, max_objects: int, max_parts: int, conf_thresh: float, dist_thresh: float, group=None):
        self.K, self.P, self.conf, self.dist, self.group = max_objects, max_parts, conf_thresh, dist_thresh, group

    def __call__(self, local_outputs: dict, global_batch: int) -> ops.PackedDetections:
        world = dist.get_world_size(self.group)
        sizes = shard_sizes(global_batch, world)
        rank = dist.get_rank(self.group)
        assert local_outputs["anchor_hm"].shape[0] == sizes[rank], "local shard does not match the contiguous split"
        packed = ops.decode_packed(local_outputs, self.K, self.P, self.conf, self.dist)
        C = local_outputs["anchor_hm"].shape[1] + local_outputs["part_hm"].shape[1]
        return all_gather_packed(packed.blob, sizes, self.K, self.P, C, self.group)


class FusedGatherPlan:
    """Decode + gather in one pass: the tail kernel stores every rank's packed detections straight
    into EVERY rank's copy of the global result (peer-mapped symmetric memory; one ``multimem.st`` per
    value through the NVSwitch multicast mapping when there is one, else a plain store per peer over
    NVLink), and per-rank completion flags replace the all-gather collective and its synchronisation.

    The global result is one packed blob for the whole batch (``ops._carve`` layout); rank r owns the
    image rows ``shard_bounds(B, world, r)`` of every field.  ``dest_delta[j] = peer_base[j] -
    local_base`` (``SdnetDecodeParams.dest_delta``) is all the kernel needs, because symmetric buffers
    share one layout.  Requires equal per-field offsets on all ranks, i.e. the same (B, K, P, C).

    Arrival: the last CTA of a rank's tail kernel releases that rank's completion flag (= its run count)
    into every copy, and ``run`` ends with a one-CTA wait kernel that spins on the local flags until all
    ranks have completed this run (``SDNET_GATHER_SYNC=barrier``: torch's symmetric-memory barrier instead).

    Write-after-read safety across ranks: the plan owns TWO result buffers and alternates between them.
    Run i stores into buffer i % 2 on every rank, then waits.  A rank can only start the remote stores of
    run i + 2 (the next writer of the same buffer) after it has passed the wait of run i + 1, i.e. after
    every peer has finished the tail kernel of ITS run i + 1 -- which that peer enqueued after whatever
    it did with result i.  So: consume (or copy) a result ON THE RUN STREAM, or make the run stream wait
    for your consumer, before calling ``run`` again; the tensors returned by run i stay valid until run
    i + 2 is enqueued.  No extra barrier is needed.  ``lazy = True`` moves the wait from the end of a run to the
    point where it is needed (see ``run`` / ``wait_arrival``).
    """

    def __init__(self, device, global_batch: int, M: int, N: int, H: int, W: int, K: int, P: int, group=None,
                 dtype: torch.dtype = torch.float32):
        import torch.distributed._symmetric_memory as symm

        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.sizes = shard_sizes(global_batch, self.world)
        self.lo, self.hi = shard_bounds(global_batch, self.world, self.rank)
        C = M + N
        nbytes = (ops.packed_nbytes(global_batch, K, P, C) + 255) // 256 * 256
        # two result buffers + one completion flag per rank (a 256-byte line of its own)
        self.blob = symm.empty(2 * nbytes + 256, dtype=torch.uint8, device=device)
        self.blob[2 * nbytes:].zero_()
        self.handle = symm.rendezvous(self.blob, self.group)
        self._flags = self.blob[2 * nbytes:2 * nbytes + 4 * self.world].view(torch.int32)  # flags[j] = decodes rank j has completed
        # the gathered detections of run i live in results[i % 2]
        self.results = [ops._carve(self.blob[k * nbytes:(k + 1) * nbytes], global_batch, K, P, C) for k in range(2)]
        self.result = self.results[0]  # what the most recent run() returned
        self._runs = 0
        self._waited = 0  # highest run number whose arrival has been waited for on the run stream
        self.lazy = False
        shard = self.hi - self.lo
        if min(self.sizes) == 0:
            raise ValueError(f"global batch {global_batch} leaves a rank of {self.world} without images: every rank "
                             "takes part in the completion flags, so every rank needs at least one image")
        self.plan = ops.DecodePlan(device, shard, M, N, H, W, K, P, dtype)
        # this rank's rows of each global buffer: where the plan's outputs point ...
        lo, hi = self.lo, self.hi
        self._mine = [ops.PackedDetections(
            r.anchor_out[lo:hi], r.part_out[lo:hi], r.anchor_inds[lo:hi], r.part_inds[lo:hi], r.part_emb[lo:hi],
            r.assign[lo:hi], r.counts[lo:hi], r.diag[lo * C:hi * C], None) for r in self.results]
        # ... and every store reaches every rank's copy: ONE multimem.st to the blob's NVSwitch multicast mapping when
        # the fabric offers one (the switch replicates it, the local copy included), else one st.global per peer
        ptrs = list(self.handle.buffer_ptrs)
        prm = self.plan.params
        mc = int(getattr(self.handle, "multicast_ptr", 0) or 0)
        want = os.environ.get("SDNET_GATHER_STORES", "auto")  # auto | multicast | peer
        if want not in ("auto", "multicast", "peer"):
            raise ValueError(f"SDNET_GATHER_STORES={want!r}: expected auto, multicast or peer")
        if want == "multicast" and mc == 0:
            raise RuntimeError("SDNET_GATHER_STORES=multicast but the symmetric allocation has no multicast mapping")
        self.stores = "multicast" if (mc != 0 and want != "peer") else "peer"
        if self.stores == "multicast":
            prm.n_dest, prm.dest_mode = 1, _native.DEST_MULTICAST
            prm.dest_delta[0] = mc - int(ptrs[self.rank])
        else:
            prm.n_dest, prm.dest_mode = self.world, _native.DEST_PEER_STORES
            for j in range(self.world):
                prm.dest_delta[j] = int(ptrs[j]) - int(ptrs[self.rank])
        # How a rank learns that every rank's rows have arrived: "flags" (default) = the last CTA of each rank's tail
        # kernel releases that rank's completion flag into every copy and a one-CTA wait kernel spins on the local flags
        # (one-sided: no handshake, no round trip over NVLink); "barrier" = torch's symmetric-memory barrier kernel.
        self.sync = os.environ.get("SDNET_GATHER_SYNC", "flags")
        if self.sync not in ("flags", "barrier"):
            raise ValueError(f"SDNET_GATHER_SYNC={self.sync!r}: expected flags or barrier")
        if self.sync == "flags":
            prm.done_flag = self._flags.data_ptr() + 4 * self.rank
        # timing diagnostics only (results are wrong or unsafe): "local" = no remote stores, "nobarrier" = no barrier / wait
        self._diag = os.environ.get("SDNET_GATHER_DIAG", "")
        if "local" in self._diag:
            prm.n_dest, prm.dest_mode, prm.done_flag = 0, _native.DEST_PEER_STORES, None
            self.sync = "barrier"
        torch.cuda.synchronize(self.plan.device)  # the flags are zero ...
        self.handle.barrier()  # ... and everyone is mapped before the first remote store

    def _enqueue_wait(self, value: int, stream) -> None:
        """Everything enqueued on ``stream`` after this sees every rank's rows of run number ``value`` (and earlier)."""
        if value <= self._waited or "nobarrier" in self._diag:
            return
        if self.sync == "flags":
            raw = getattr(stream, "cuda_stream", stream)
            rc = self.plan.lib.sdnet_gather_wait_launch(ctypes.c_void_p(self._flags.data_ptr()), self.world,
                                                        ctypes.c_uint32(value & 0xFFFFFFFF), ctypes.c_void_p(raw))
            _native.check(rc, "sdnet_gather_wait_launch")
        else:  # a barrier cannot wait for the past: lazy mode degenerates to one barrier per run
            with torch.cuda.stream(stream):
                self.handle.barrier()
        self._waited = value

    def run(self, anchor_hm, part_hm, offsets, embeddings, conf_f32, dist_abs_f32, radius=2, flags=0,
            stream: torch.cuda.Stream | None = None) -> ops.PackedDetections:
        """Enqueue decode + remote stores + the arrival wait on ``stream`` (default: the current one); returns
        the global result.  Every rank must call its plans in the same order, and a plan must stay on one stream.

        ``self.lazy = True`` (completion flags only) defers the arrival wait: ``run`` then only waits -- before its own
        stores -- for the run before the previous one's successor, which is what write-after-read safety of the two
        result buffers needs, and the caller asks for arrival with ``wait_arrival`` when it wants to read a result.
        Ranks then drift apart by up to two runs per plan instead of meeting after every run."""
        if stream is None:
            stream = torch.cuda.current_stream(self.plan.device)
        k = self._runs % 2
        self._runs += 1
        run_no = self._runs
        lazy = self.lazy and self.sync == "flags"
        if lazy and run_no >= 3:
            self._enqueue_wait(run_no - 1, stream)  # every peer is past its run (run_no - 1): it has consumed result (run_no - 2)
        self.plan._bind_outputs(self._mine[k])
        self.plan.params.done_value = run_no & 0xFFFFFFFF
        self.plan.run(anchor_hm, part_hm, offsets, embeddings, conf_f32, dist_abs_f32, radius, flags, stream=stream)
        if not lazy:
            self._enqueue_wait(run_no, stream)
        self.result = self.results[k]
        return self.result

    def wait_arrival(self, stream: torch.cuda.Stream | None = None) -> None:
        """Lazy mode: make ``stream`` (the plan's run stream) wait until every rank's rows of the most recent run
        have arrived.  A no-op when that has been waited for already."""
        if stream is None:
            stream = torch.cuda.current_stream(self.plan.device)
        self._enqueue_wait(self._runs, stream)
