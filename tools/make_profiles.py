"""Turn the captures of tools/evidence.sh (gpurun_out/) into the tracked files under profiles/.
usage: python tools/make_profiles.py r01 [report.ncu-rep]   -> copies bench lines + launch list, prints the kernel table,
writes profiles/peaks_kernel_traffic.json"""
import csv, json, shutil, subprocess, sys
from pathlib import Path

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out, prof = Path("gpurun_out"), Path("profiles")
rep = Path(sys.argv[2]) if len(sys.argv) > 2 else out / f"{tag}_kernels_n1.ncu-rep"
if len(sys.argv) <= 2:
    for name in ("bench_n1", "bench_ref", "bench_n1_blobs", "bench_n1_serial", "bench_n1_f16", "bench_n1_bf16"):
        src = out / f"{tag}_{name}.json"
        if src.exists() and src.stat().st_size:
            shutil.copy(src, prof / src.name)
    shutil.copy(out / f"{tag}_launches_n1.csv", prof / f"{tag}_launches_n1.csv")

raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = [r for r in csv.reader(raw.splitlines()) if r]
hdr, units, kern = rows[0], rows[1], rows[2:]
col = {n: i for i, n in enumerate(hdr)}
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3,
         "byte/block": 1, "Kbyte/block": 1e3, "Mbyte/block": 1e6}


def num(r, name):
    return float(r[col[name]].replace(",", "")) * SCALE.get(units[col[name]], 1)


durs = [num(r, "gpu__time_duration.sum") for r in kern]
traffic = {}
for r, dur in zip(kern, durs):
    name = r[col["Kernel Name"]].replace("void <unnamed>::", "").split("(")[0]
    rd, wr = num(r, "dram__bytes_read.sum"), num(r, "dram__bytes_write.sum")
    first = lambda s: s.strip("()").split(",")[0]
    print(f"| `{name}` | {first(r[col['Grid Size']])} × {first(r[col['Block Size']])} | {int(num(r, 'launch__registers_per_thread'))} | "
          f"{num(r, 'launch__shared_mem_per_block_dynamic') / 1e3:.1f} KB | {dur:.1f} µs ({100 * dur / sum(durs):.1f} %) | "
          f"{rd / 1e9:.3f} GB | {wr / 1e6:.1f} MB | {num(r, 'smsp__inst_executed.sum') / 1e6:.1f} M | "
          f"{num(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} % | "
          f"{num(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} % | {num(r, 'lts__t_sector_hit_rate.pct'):.1f} % |")
    if "peaks" in name:
        traffic = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr, "kernel": name,
                   "source": f"profiles/{tag}_kernels_n1.md (ncu --set full, cfg5, 1024 images)"}
        stalls = {n.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""): num(r, n) for n in hdr
                  if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio")}
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:9]
        print("peaks kernel, stalls per issued instruction:", ", ".join(f"{k} {v:.2f}" for k, v in top))
if traffic and len(sys.argv) <= 2:
    (prof / "peaks_kernel_traffic.json").write_text(json.dumps({"n1": traffic}, indent=1))
print(json.dumps(traffic))
