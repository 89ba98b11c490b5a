"""One-screen summary of an .ncu-rep: key metrics + stall reasons (+ optional per-line ranking).
usage: python tools/ncu_summary.py report.ncu-rep"""
import csv, subprocess, sys
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; units=rows[1]; data=rows[2:]
want=["gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","dram__throughput.avg.pct_of_peak_sustained_elapsed","sm__warps_active.avg.pct_of_peak_sustained_active","smsp__inst_executed.sum","lts__t_sector_hit_rate.pct","smsp__issue_active.avg.pct_of_peak_sustained_active","smsp__thread_inst_executed_per_inst_executed.ratio","launch__grid_size","launch__block_size","launch__occupancy_limit_registers","launch__occupancy_limit_shared_mem","launch__registers_per_thread","launch__shared_mem_per_block_dynamic"]
for w in want:
    if w in hdr:
        i=hdr.index(w); print(f"{w:72s} {units[i]:12s} {[r[i] for r in data]}")
for i,h in enumerate(hdr):
    if "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h:
        v=float(data[0][i])
        if v>0.2: print("  stall", h.replace("smsp__average_warps_issue_stalled_","").replace("_per_issue_active.ratio",""), round(v,2))
