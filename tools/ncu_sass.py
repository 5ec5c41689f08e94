"""Print SASS (address order) with samples / executed counts / top stall reasons for source lines of a file.
usage: python tools/ncu_sass.py export.csv file.cu lo hi"""
import csv
import sys

path, fname, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
rows = list(csv.reader(open(path)))
hdr = None
cur = None
out = []
f = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        f = r[1].split('/')[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) - 2:
        continue
    if r[0] != "":
        cur = (f, int(r[0]))
        continue
    out.append((cur, r))
ia = hdr.index('Address')
ismp = hdr.index('# Samples')
iins = hdr.index('Instructions Executed')
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
sel = [(c, r) for c, r in out if c[0] == fname and lo <= c[1] <= hi]
seen = set()
for c, r in sorted(sel, key=lambda t: t[1][ia]):
    if r[ia] in seen:
        continue
    seen.add(r[ia])
    st = []
    for i, h in stall_cols:
        try:
            v = int(r[i])
        except ValueError:
            continue
        if v:
            st.append((h.replace('stall_', ''), v))
    st.sort(key=lambda t: -t[1])
    print(f"{c[1]:4d} {r[3][:72]:72s} smp {r[ismp]:>5} ins {r[iins]:>9} {st[:3]}")
