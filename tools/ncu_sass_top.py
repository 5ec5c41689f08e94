"""Top SASS instructions by stall samples from an `ncu --page source --csv --print-source cuda,sass` export,
with the dominant stall reasons and the preceding instructions. usage: python tools/ncu_sass_top.py export.csv [n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
hdr, sass = None, []
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        i_ins, i_smp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) < len(hdr) - 2 or r[0] != "" or r[2] == "...":
        continue
    try:
        sass.append((int(r[2], 16), r[3].strip(), int(r[i_ins]), int(r[i_smp]), r))
    except ValueError:
        pass
sass.sort()
tot = sum(x[3] for x in sass) or 1
st = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not" not in h]
order = sorted(range(len(sass)), key=lambda i: -sass[i][3])[:top]
for i in order:
    a, t, ins, smp, r = sass[i]
    why = sorted(((int(r[j]), hdr[j]) for j in st if r[j] not in ("", "0", "-")), reverse=True)[:3]
    print(f"{smp / tot * 100:5.1f}% samples  exec {ins:9d}  {t[:70]:70s} {why}")
    for j in range(max(0, i - 4), i):
        print(f"        prev   exec {sass[j][2]:9d}  {sass[j][1][:70]}")
