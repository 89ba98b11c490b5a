"""Turn the captures of tools/evidence.sh (gpurun_out/) into the tracked files under profiles/.
usage: python tools/make_profiles.py r02   -> copies the bench lines + launch list, writes profiles/<tag>_kernels.md (one
table per ncu --set full capture) and profiles/peaks_kernel_traffic.json"""
import csv, json, shutil, subprocess, sys
from pathlib import Path

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
out, prof = Path("gpurun_out"), Path("profiles")
BENCH = ("bench_n1", "bench_ref", "bench_n1_serial", "bench_n1_f16", "bench_n1_bf16", "bench_n1_gb128", "bench_cfg2", "bench_cfg3",
         "bench_cfg4")
CAPTURES = (  # (report suffix, what was captured, (images per rank, input mode, dtype) of a cfg5 peaks-kernel capture)
    ("kernels_n1", "cfg5, 1024 images, fp32, noise: `python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-objects --no-parity --pipeline 1 --mode noise`", (1024, "noise", "f32")),
    ("kernels_n1_blobs", "same, `--mode blobs`", (1024, "blobs", "f32")),
    ("kernels_n1_gb128", "the 8-GPU shard: `--mode blobs --global-batch 128`", (128, "blobs", "f32")),
    ("peaks_f16", "peaks kernel, `--mode noise --dtype f16` (row-pair tiles)", (1024, "noise", "f16")),
    ("peaks_bf16", "peaks kernel, `--mode noise --dtype bf16` (row-pair tiles)", (1024, "noise", "bf16")),
    ("suppress", "`python tools/time_suppress.py 256` (dense sigmoid + NMS maps, 256 cfg5 images, fp32)", None),
)
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3,
         "byte/block": 1, "Kbyte/block": 1e3, "Mbyte/block": 1e6}
HEAD = ("| kernel | grid × block | regs | dyn. smem | duration (share) | DRAM read | DRAM written | warp-instr. | issue active | warps active | L2 hit |\n"
        "|---|---|---|---|---|---|---|---|---|---|---|")


def table(rep: Path):
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(raw.splitlines()) if r]
    hdr, units, kern = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(hdr)}

    def num(r, name):
        return float(r[col[name]].replace(",", "")) * SCALE.get(units[col[name]], 1)

    durs = [num(r, "gpu__time_duration.sum") for r in kern]
    lines, traffic = [HEAD], {}
    for r, dur in zip(kern, durs):
        name = r[col["Kernel Name"]].replace("void <unnamed>::", "").split("(")[0]
        rd, wr = num(r, "dram__bytes_read.sum"), num(r, "dram__bytes_write.sum")
        first = lambda s: s.strip("()").split(",")[0]
        lines.append(
            f"| `{name}` | {first(r[col['Grid Size']])} × {first(r[col['Block Size']])} | {int(num(r, 'launch__registers_per_thread'))} | "
            f"{num(r, 'launch__shared_mem_per_block_dynamic') / 1e3:.1f} KB | {dur:.1f} µs ({100 * dur / sum(durs):.1f} %) | "
            f"{rd / 1e9:.3f} GB | {wr / 1e6:.1f} MB | {num(r, 'smsp__inst_executed.sum') / 1e6:.1f} M | "
            f"{num(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} % | "
            f"{num(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} % | {num(r, 'lts__t_sector_hit_rate.pct'):.1f} % |")
        if "peaks" in name or "suppress" in name:
            traffic = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr, "kernel": name}
            stalls = {n.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""): num(r, n) for n in hdr
                      if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio")}
            top = sorted(stalls.items(), key=lambda kv: -kv[1])[:8]
            lines.append("")
            lines.append(f"`{name}` stalls per issued instruction: " + ", ".join(f"{k} {v:.2f}" for k, v in top))
            lines.append("")
            lines.append(HEAD) if r is not kern[-1] else None
    while lines and lines[-1] in ("", HEAD):
        lines.pop()
    return lines, traffic


for name in BENCH:
    src = out / f"{tag}_{name}.json"
    if src.exists() and src.stat().st_size:
        shutil.copy(src, prof / src.name)
for extra in (f"{tag}_launches_n1.csv", f"{tag}_suppress.log", f"{tag}_pytest_gpu.log"):
    if (out / extra).exists():
        shutil.copy(out / extra, prof / extra)
md = [f"# {tag}: ncu `--set full --clock-control none --import-source on` captures, one B200", "",
      "Each capture was taken after the same command had run clean without ncu.  Durations under ncu are cold-cache and",
      "serialised: compare SHARES, not absolute times; every throughput number in the bench lines comes from CUDA events in a",
      "plain run.  Columns: registers per thread, dynamic shared memory per CTA, DRAM bytes per launch, executed warp",
      "instructions, issue-slot utilisation, achieved occupancy, L2 hit rate.", ""]
all_traffic = {}
for suffix, what, case in CAPTURES:
    rep = out / f"{tag}_{suffix}.ncu-rep"
    if not rep.exists():
        continue
    lines, traffic = table(rep)
    md += [f"## {suffix} — {what}", ""] + lines + [""]
    if traffic and case:
        traffic["source"] = f"profiles/{tag}_kernels.md section {suffix}"
        traffic["images"], traffic["mode"], traffic["dtype"] = case
        all_traffic[suffix] = traffic
(prof / f"{tag}_kernels.md").write_text("\n".join(md))
if all_traffic:  # bench.py looks its (images per rank, mode, dtype) up in this list
    (prof / "peaks_kernel_traffic.json").write_text(json.dumps({"captures": list(all_traffic.values())}, indent=1))
print("\n".join(md))
