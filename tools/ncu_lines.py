"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line.
usage: python tools/ncu_lines.py export.csv [top_n]"""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
rows = list(csv.reader(open(path)))
cur_file = None
hdr = None
per = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        i_ins = hdr.index("Instructions Executed")
        i_smp = hdr.index("# Samples")
        continue
    if hdr is None or len(r) < len(hdr) - 2:
        continue
    if r[0] != "":  # a source line row (aggregated over its SASS)
        try:
            per.append((int(r[i_ins]), int(r[i_smp]), cur_file, r[0], r[1].strip()[:100]))
        except ValueError:
            pass
tot_i = sum(p[0] for p in per) or 1
tot_s = sum(p[1] for p in per) or 1
print(f"total warp-instructions {tot_i}  samples {tot_s}")
for p in sorted(per, key=lambda x: -x[1])[:top]:
    print(f"{p[0] / tot_i * 100:5.1f}% inst {p[1] / tot_s * 100:5.1f}% stall-samples  {p[2]}:{p[3]:>4}  {p[4]}")
