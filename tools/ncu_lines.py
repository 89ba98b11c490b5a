"""Rank CUDA source lines of an .ncu-rep by executed instructions / stall samples.
usage: python tools/ncu_lines.py report.ncu-rep [topN]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
his = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
for k, hi in enumerate(his):
    hdr = rows[hi]
    iex, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    name = [r for r in rows[:hi] if r and r[0] == "Function Name"][-1][1]
    num = lambda v: int(v) if v.lstrip("-").isdigit() else 0
    lines = []
    end = his[k + 1] if k + 1 < len(his) else len(rows)
    for r in rows[hi + 1:end]:
        if r and r[0].isdigit():
            lines.append((int(r[0]), r[1].strip(), num(r[iex]), num(r[isamp])))
    tot = sum(l[2] for l in lines) or 1; ts = sum(l[3] for l in lines) or 1
    print(f"== {name}: {tot} warp-inst, {ts} samples")
    for l in sorted(lines, key=lambda x: -x[2])[:top]:
        print(f"{l[0]:5d} {100*l[2]/tot:5.1f}% inst {100*l[3]/ts:5.1f}% samp  {l[1][:120]}")
