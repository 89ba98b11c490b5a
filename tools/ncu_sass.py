"""Print hot SASS of the first kernel in an .ncu-rep with per-row execution counts.
usage: python tools/ncu_sass.py report.ncu-rep rows_total [min_per_row]"""
import csv, subprocess, sys
rep=sys.argv[1]; rows_total=float(sys.argv[2]); thr=float(sys.argv[3]) if len(sys.argv)>3 else 0.5
out=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source=cuda,sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
his=[i for i,r in enumerate(rows) if r and r[0]=="Line No"]
hi=his[0]; hdr=rows[hi]; iex=hdr.index("Instructions Executed")
cur=None; seen={}
for r in rows[hi+1:his[1] if len(his)>1 else len(rows)]:
    if r[0].isdigit(): cur=int(r[0]); continue
    if len(r)>3 and r[2].startswith("0x"):
        n=int(r[iex]) if r[iex].isdigit() else 0
        if r[2] not in seen: seen[r[2]]=(cur,r[3].strip(),n)
tot=0
for a in sorted(seen):
    l,sass,n=seen[a]; tot+=n
    if n/rows_total>=thr: print(f"{a[-5:]} L{l:4d} {n/rows_total:6.2f}  {sass}")
print("total per row", tot/rows_total)
