#!/usr/bin/env python
"""bench.py -- decoded images/s of the SDNet decoding path on 1..8 B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode noise|blobs]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload: BASELINE.json configs[4] ("cfg5"): a global batch of 1024 network outputs of the
2448x2048-input configuration, i.e. raw (1024, 2+1+4, 512, 612) fp32, K = P = 100,
conf 0.4, dist 0.1, split evenly over the ranks (strong scaling).  One *step* = one decode of
the whole global batch: every rank decodes its shard (3 kernels through the C ABI), then
-- for N > 1 -- the packed detections are all-gathered over NCCL so every rank holds all
1024 results.  Inputs are resident in HBM and larger than L2 (8.98 GB / N per rank).

Prints ONE JSON line (see the fields below).  ``--impl reference`` times the reference
decoder's own algorithm on the host CPU instead (``oracle/torch_port.py``: the same stock
torch CPU ops the pure-Python reference calls, all host threads), on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOAD = "cfg5"
UNIQUE_IMAGES = 32  # distinct synthetic images generated on the CPU; tiled to fill the shard
CPU_BATCH = 16  # images per CPU-baseline step (the reference's cfg3 batch)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", choices=("ours", "reference"), default="ours")
    ap.add_argument("--mode", choices=("noise", "blobs"), default="noise")
    ap.add_argument("--dtype", choices=("f32", "f16", "bf16"), default="f32",
                    help="element type of the network outputs (f32 = the headline; f16/bf16 = the --amp validation path)")
    ap.add_argument("--global-batch", type=int, default=None, help="override cfg5's 1024 (debug only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-objects", action="store_true", help="skip the Decoder -> Python objects leg")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--pipeline", type=int, default=0,
                    help="decodes in flight: consecutive steps alternate between this many plans/streams; 0 = auto "
                         "(2 for shards of >= 512 images, up to 4 for smaller ones: measured at 8 GPUs x 128 images "
                         "6.55 / 6.90 / 7.42 M img/s with 2 / 3 / 4)")
    ap.add_argument("--gather", choices=("fused", "nccl"), default="fused",
                    help="N > 1: tail kernel stores into every peer (symmetric memory) + barrier, or ncclAllGather")
    return ap.parse_args()


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    REASONS = {  # nvmlClocksEventReasons bits
        0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
        0x80: "hw_power_brake_slowdown",
    }

    def __init__(self, index: int, period_s: float = 0.01):
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception as exc:  # noqa: BLE001
            self.nv, self.error = None, repr(exc)

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                for bit, name in self.REASONS.items():
                    if bits & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0,
                    "note": "NVML unavailable" if self.nv is None else "timed region shorter than one sample"}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def physical_gpu_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except (ValueError, IndexError):
            pass
    return local_rank


def workload_name(cfg) -> str:
    return (f"{WORKLOAD}: global batch {cfg.batch} of 2448x2048 inputs -> raw ({cfg.batch}, {cfg.channels}, {cfg.height}, "
            f"{cfg.width}) fp32, K=P={cfg.max_objects}, conf {cfg.conf_threshold}, dist {cfg.dist_thresh}")


# ----------------------------------------------------------------------------- CPU reference arm
def cpu_reference_rate(cfg, mode: str, seconds: float, steps: int | None = None, warmup: int = 1):
    """images/s of the reference algorithm on the host CPU (oracle port, all torch threads)."""
    import torch

    from oracle import torch_port
    from structuredetector_b200.synth import make_raw, split_outputs

    # all the host threads the process may use (torchrun exports OMP_NUM_THREADS=1, which would make
    # the CPU arm ten times slower than the same code started with plain `python`)
    try:
        usable = len(os.sched_getaffinity(0))
    except AttributeError:
        usable = os.cpu_count() or 1
    torch.set_num_threads(max(1, usable))
    raw = make_raw(cfg, mode, batch=CPU_BATCH)
    outs = split_outputs(raw, cfg.labels, cfg.parts)
    labels = {i: f"label{i}" for i in range(cfg.labels)}
    parts = {i: f"part{i}" for i in range(cfg.parts)}
    run = lambda: torch_port.decode(outs, labels, parts, "anchor", 4.0, cfg.max_objects, cfg.max_parts,
                                    cfg.conf_threshold, cfg.dist_thresh)
    for _ in range(max(1, warmup)):
        run()
    times = []
    t_end = time.perf_counter() + seconds
    while (steps is not None and len(times) < steps) or (steps is None and (time.perf_counter() < t_end or len(times) < 2)):
        t0 = time.perf_counter()
        run()
        times.append(time.perf_counter() - t0)
    total = sum(times)
    return {
        "value": CPU_BATCH * len(times) / total,
        "unit": "images/s",
        "cores": torch.get_num_threads(),
        "host_cpus": os.cpu_count(),
        "kind": "port",
        "sample": f"{len(times)} steps x {CPU_BATCH} images of {WORKLOAD} maps ({cfg.height}x{cfg.width}, mode {mode}), "
                  f"tensor ops + Python object assembly, torch {torch.__version__} CPU",
        "ms_per_step": 1e3 * total / len(times),
        "steps": len(times),
    }


def run_reference_arm(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # the CPU arm runs once, on rank 0
    res = cpu_reference_rate(cfg, args.mode, seconds=0.0, steps=max(1, args.steps), warmup=max(1, args.warmup))
    line = {
        "impl": "reference",
        "metric": "decoded images/s",
        "value": res["value"],
        "unit": "images/s",
        "n_gpus": args.gpus,
        "steps": res["steps"],
        "warmup": max(1, args.warmup),
        "ms_per_step": res["ms_per_step"],
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(cfg), "mode": args.mode, "images_per_step": CPU_BATCH,
                   "note": "CPU arm: each step decodes a bounded sample of the workload (16 images), all host threads"},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------- our arm
def emit(line: dict):
    """The ONE JSON line goes to the real stdout; everything else any library prints to fd 1 while
    the bench runs (e.g. NCCL's version banner) was diverted to stderr by main()."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    args = parse_args()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    from structuredetector_b200.synth import CONFIGS, DecodeConfig, make_raw, split_outputs

    cfg = CONFIGS[WORKLOAD]
    if args.global_batch:
        cfg = DecodeConfig(cfg.name, args.global_batch, cfg.labels, cfg.parts, cfg.height, cfg.width,
                           cfg.max_objects, cfg.max_parts, cfg.conf_threshold, cfg.dist_thresh, cfg.cfg_id)
    if args.impl == "reference":
        run_reference_arm(args, cfg)
        return

    import torch
    import torch.distributed as dist

    from structuredetector_b200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run for N > 1")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the decode path has no CPU fallback")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    if cfg.batch % world:
        raise SystemExit(f"global batch {cfg.batch} does not divide over {world} ranks")
    shard = cfg.batch // world
    M, N, H, W, K, P = cfg.labels, cfg.parts, cfg.height, cfg.width, cfg.max_objects, cfg.max_parts

    # ---- synthetic inputs: identical bits on every run; each rank takes its slice of the global batch
    uniq = make_raw(cfg, args.mode, batch=min(UNIQUE_IMAGES, cfg.batch))
    uniq_dev = uniq.to(device)
    idx = (torch.arange(shard, device=device) + rank * shard) % uniq_dev.shape[0]
    tdtype = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}[args.dtype]
    esize = 4 if args.dtype == "f32" else 2
    raw = uniq_dev[idx].contiguous().to(tdtype)  # (shard, M+N+4, H, W), resident in HBM
    del uniq_dev
    outs = split_outputs(raw, M, N)
    conf32 = float(torch.tensor(cfg.conf_threshold, dtype=tdtype))
    dist32 = float(torch.tensor(cfg.dist_thresh * min(W, H), dtype=torch.float32))
    plan = ops.DecodePlan(device, shard, M, N, H, W, K, P, tdtype)
    blob_bytes = plan.out.blob.numel()
    if args.dtype != "f32":
        args.no_e2e = True       # the host-buffer entry point is fp32-only
        args.gather = "nccl" if args.gather == "fused" else args.gather
    gathered = None
    fused = None
    gather_kind = "none"
    if world > 1 and args.gather == "fused":
        try:
            from structuredetector_b200.parallel import FusedGatherPlan

            fused = FusedGatherPlan(device, cfg.batch, M, N, H, W, K, P)
            gather_kind = ("fused: tail kernel stores each rank's packed detections into every peer's copy "
                           "(symmetric memory, st.global over NVLink) + one symmetric-memory barrier")
        except Exception as exc:  # noqa: BLE001
            print(f"[bench] symmetric memory unavailable ({exc!r}); falling back to ncclAllGather", file=sys.stderr)
            fused = None
    if world > 1 and fused is None:
        gathered = torch.empty(world * blob_bytes, dtype=torch.uint8, device=device)
        gather_kind = "ncclAllGather of packed detections (every rank holds all results)"

    # Software pipeline over batches: consecutive steps alternate between `depth` plans (own workspace,
    # own outputs, own gather buffer) on `depth` streams, so one batch's tail kernel and gather overlap
    # the next batch's peaks kernel.  --pipeline 1 = strictly one decode at a time.
    depth = args.pipeline if args.pipeline > 0 else (2 if shard >= 512 else 4)
    pipe = None
    if depth > 1:
        if fused is not None:
            extra = [FusedGatherPlan(device, cfg.batch, M, N, H, W, K, P) for _ in range(depth - 1)]
            pipe = ops.DecodePipeline(device, depth, lambda i: fused if i == 0 else extra[i - 1])
        elif world == 1:
            pipe = ops.DecodePipeline(device, depth, lambda i: plan if i == 0 else ops.DecodePlan(device, shard, M, N, H, W, K, P, tdtype))
        else:
            depth = 1  # the ncclAllGather baseline stays serial

    def step():
        if pipe is not None:
            pipe.submit(outs["anchor_hm"], outs["part_hm"], outs["offsets"], outs["embeddings"], conf32, dist32)
            return
        if fused is not None:
            fused.run(outs["anchor_hm"], outs["part_hm"], outs["offsets"], outs["embeddings"], conf32, dist32)
            return
        plan.run(outs["anchor_hm"], outs["part_hm"], outs["offsets"], outs["embeddings"], conf32, dist32)
        if world > 1:
            dist.all_gather_into_tensor(gathered, plan.out.blob)

    def fence():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    for _ in range(max(3, args.warmup)):
        step()
    fence()

    # ---- timed region: exactly K steps, CUDA events on the launching stream, max over ranks
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(physical_gpu_index(local_rank)) as clocks:
        fence()
        ev0.record()
        if pipe is not None:
            pipe.after(ev0)
        for _ in range(args.steps):
            step()
        if pipe is not None:
            pipe.drain()
        ev1.record()
        fence()
    ms_total = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = cfg.batch / (ms_per_step * 1e-3)

    # ---- sanity: the timed path produced real detections (and every rank agrees after the gather)
    res = fused.result if fused is not None else plan.out
    counts = res.counts.sum(dim=0).tolist()  # fused: the whole batch, gathered; else this rank's shard
    if fused is None and world > 1:
        counts = [c * world for c in counts]
    diag_overflow = int(res.diag[:, 1].sum())
    cand_mean = float(res.diag[:, 0].float().mean())

    # ---- roofline leg: device time of each kernel (events between the launches), averaged
    reps = 20
    kms = [plan.run_timed(outs["anchor_hm"], outs["part_hm"], outs["offsets"], outs["embeddings"], conf32, dist32)
           for _ in range(reps)]
    peaks_ms = statistics.mean(k[0] for k in kms)
    exact_ms = statistics.mean(k[1] for k in kms)
    tail_ms = statistics.mean(k[2] for k in kms)
    peaks_bytes = shard * (M + N) * H * W * esize  # algorithmic bytes of the dominant kernel: heat maps read once
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak_gbs, peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak_gbs, peak_src = 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"
    achieved = peaks_bytes / (peaks_ms * 1e-3) / 1e9
    traffic = None
    traffic_path = ROOT / "profiles" / "peaks_kernel_traffic.json"
    if traffic_path.exists() and args.dtype == "f32":
        try:
            traffic = json.loads(traffic_path.read_text()).get(f"n{world}", {}).get("dram_bytes_per_launch")
        except Exception:  # noqa: BLE001
            traffic = None
    step_s = ms_per_step * 1e-3
    roofline = {
        "bound": "hbm", "kernel": {"tile": "sdnet_peaks_tile_kernel (TMA tiles)", "tile_row_pairs": "sdnet_peaks_tile_kernel (TMA tiles over row pairs)",
                                  "warp": "sdnet_peaks_kernel (per-lane feed)"}[
                           plan.peaks_path(outs["anchor_hm"], outs["part_hm"], outs["offsets"], outs["embeddings"])], "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
        "frac": achieved / peak_gbs, "traffic": traffic, "peak_source": peak_src,
        "kernel_ms": {"peaks": peaks_ms, "exact_select": exact_ms, "tail": tail_ms},
        "kernel_share_of_step": peaks_ms / (peaks_ms + exact_ms + tail_ms),
        "algorithmic_bytes_per_launch": peaks_bytes,
        # whole step (all kernels + gather), per rank, under the two denominators of SURVEY 8(d)
        "step_achieved_min_gbs": shard * cfg.min_bytes_per_image * esize / 4 / step_s / 1e9,
        "step_achieved_contract_gbs": shard * cfg.contract_bytes_per_image * esize / 4 / step_s / 1e9,
        "frac_of_8tbs_min": shard * cfg.min_bytes_per_image * esize / 4 / step_s / 8e12,
        "frac_of_8tbs_contract": shard * cfg.contract_bytes_per_image * esize / 4 / step_s / 8e12,
    }

    # ---- the reference-facing Python call: device tensors -> list[ImageAnnotation] (decode + ONE D2H copy +
    # Python object assembly), reported separately; rank 0, bounded sample
    objects = None
    if rank == 0 and not args.no_objects:
        from types import SimpleNamespace

        from structuredetector_b200 import Decoder

        n_obj = min(shard, 128)
        dec = Decoder(SimpleNamespace(_r_labels={i: f"label{i}" for i in range(M)}, _r_parts={i: f"part{i}" for i in range(N)},
                                      anchor_name="anchor", down_ratio=4.0, max_objects=K, max_parts=P,
                                      conf_threshold=cfg.conf_threshold, decoder_dist_thresh=cfg.dist_thresh))
        sample = {k: v[:n_obj] for k, v in outs.items()}
        dec(sample)
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        reps_obj = 3
        for _ in range(reps_obj):
            anns = dec(sample)
        dt = (time.perf_counter() - t0) / reps_obj
        objects = {"value": n_obj / dt, "unit": "images/s", "ms_per_image": dt / n_obj * 1e3,
                   "objects_per_image": sum(len(a.objects) for a in anns) / n_obj,
                   "parts_per_image": sum(a.nb_parts for a in anns) / n_obj,
                   "sample": f"Decoder(args)(outputs) on {n_obj} images resident on the device -> list[ImageAnnotation], wall clock, 1 Python thread"}

    # ---- end to end through the C ABI with HOST buffers: pinned inputs -> packed results on the host
    e2e = None
    if not args.no_e2e:
        host_raw = torch.empty(raw.shape, dtype=raw.dtype, pin_memory=True)
        host_raw.copy_(raw)
        h_outs = split_outputs(host_raw, M, N)
        staging = torch.empty(shard * (M + N) * H * W * 4, dtype=torch.uint8, device=device)
        host_blob = torch.empty(blob_bytes, dtype=torch.uint8, pin_memory=True)

        def e2e_step():
            plan.run_host(h_outs["anchor_hm"], h_outs["part_hm"], h_outs["offsets"], h_outs["embeddings"],
                          conf32, dist32, staging)
            host_blob.copy_(plan.out.blob, non_blocking=True)

        e2e_step()
        fence()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        fence()
        dt = (time.perf_counter() - t0) / args.e2e_steps
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        same = bool(torch.equal(host_blob[: plan.out.anchor_inds.numel() * 8],
                                plan.out.blob.cpu()[: plan.out.anchor_inds.numel() * 8]))
        heat = shard * (M + N) * H * W * 4
        gathers = shard * (K + 2 * P) * 2 * 32  # zero-copy reads of offsets/embeddings, one 32 B sector each
        e2e = {"value": cfg.batch / dt, "unit": "images/s", "h2d_bytes_per_step": (heat + gathers) * world,
               "d2h_bytes_per_step": blob_bytes * world, "ms_per_step": dt * 1e3, "steps": args.e2e_steps,
               "path": "sdnet_decode_host_launch: heat planes cudaMemcpy2DAsync from pinned host, "
                       "offsets/embeddings read in place at the selected peaks, packed results copied back",
               "results_copied_back_match": same}
        del host_raw, staging

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.dtype == "f32":
        res = cpu_reference_rate(cfg, args.mode, seconds=12.0)
        cpu_baseline = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
        cpu_baseline["host_cpus"] = res["host_cpus"]

    if rank == 0:
        line = {
            "metric": "decoded images/s",
            "value": value,
            "unit": "images/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step,
            "higher_is_better": True,
            "scaling": "strong",
            "vs_baseline": None,
            "dtype": args.dtype,
            "data": "synthetic",
            "config": {
                "workload": workload_name(cfg) if args.dtype == "f32" else workload_name(cfg).replace("fp32", args.dtype),
                "mode": args.mode, "images_per_rank": shard, "parallelism": f"batch-shard x{world}",
                "gather": gather_kind,
                "l2": f"inputs larger than L2 ({raw.numel() * esize / 1e9:.2f} GB per rank, no flush needed)",
                "unique_images": int(min(UNIQUE_IMAGES, cfg.batch)),
                "pipeline": (f"{depth} decodes in flight: consecutive steps alternate between {depth} plans (own workspace, outputs"
                             f"{' and gather buffer' if world > 1 else ''}) on {depth} streams; every step is a full decode of the batch")
                            if depth > 1 else "1 (each step waits for the one before)",
            },
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "e2e": e2e,
            "python_objects": objects,
            "gpu_launches": ops.gpu_launches_per_decode() * args.steps,
            "clocks": clocks.summary(),
            "detections": {"anchors_above_conf": counts[0], "parts_above_conf": counts[1],
                           "planes_via_exact_select": diag_overflow,
                           "candidates_per_plane_mean": cand_mean},
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
