"""Per-function / per-region executed warp-instruction counts from an
`ncu --page source --csv --print-source cuda,sass` export (SASS rows only, so nothing is double counted).
usage: python tools/ncu_regions.py export.csv n_macroblocks"""
import csv
import re
import sys

path, nmb = sys.argv[1], float(sys.argv[2])
rows = list(csv.reader(open(path)))
cur_file, cur_line, hdr = None, None, None
per = {}
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
    if r[0] != "":
        cur_line = int(r[0])
        continue
    try:
        ins, smp = int(r[i_ins]), int(r[i_smp])
    except ValueError:
        continue
    k = (cur_file, cur_line)
    a = per.setdefault(k, [0, 0])
    a[0] += ins
    a[1] += smp
tot = sum(v[0] for v in per.values())
tots = sum(v[1] for v in per.values())
print(f"total {tot} warp-instr = {tot / nmb:.1f} per MB, {tots} samples")


def region_table(fname, marks):
    src = open(fname).read().split("\n")
    base = fname.split("/")[-1]
    pts = []
    for i, l in enumerate(src):
        for pat, name in marks:
            if re.search(pat, l):
                pts.append((i + 1, name))
    pts.sort()
    pts.append((len(src) + 1, "END"))
    for (a, name), (b, _) in zip(pts, pts[1:]):
        ins = sum(v[0] for (f, ln), v in per.items() if f == base and a <= ln < b)
        smp = sum(v[1] for (f, ln), v in per.items() if f == base and a <= ln < b)
        if ins:
            print(f"  {ins / nmb:8.1f} instr/MB {smp / tots * 100:5.1f}% samples  {base}:{a}-{b - 1} {name}")


region_table("dryv_b200/csrc/recon_kernels.cuh",
             [(r"^__device__ __forceinline__ \w[\w ]* (\w+)\(", "fn"), (r"^template", "tmpl")])
region_table("dryv_b200/csrc/recon.cu",
             [(r"^__device__ __forceinline__", "helper"), (r"^__global__", "kernel entry"), (r"front warp ====", "FRONT"),
              (r"// prefetch macroblock 0", "front: row setup"), (r"for \(int x = 0; x < W; x\+\+\)", "front: prefetch+header"),
              (r"// ring slot: wait", "front: wait slot"), (r"// 1\. residual", "front: residual call"),
              (r"// 2\. line x of the row above", "front: wait line"), (r"// 3\. prediction modes", "front: modes+handoff"),
              (r"// 4\. chroma prediction", "front: chroma+store+carry"), (r"// no more rows", "front: exit"),
              (r"luma warp ====", "LUMA"), (r"// row start$", "luma: row start"), (r"const int mbcls = slot.mbcls", "luma: wait line+predict"),
              (r"if \(lane < 16\) \{\s*$", "luma: store"), (r"// carry: right-most column -> left-neighbour column, top-row slots", "luma: carry"),
              (r"^// Residual only", "residual kernel")])
others = {}
for (f, ln), v in per.items():
    if f not in ("recon_kernels.cuh", "recon.cu"):
        others[f] = others.get(f, 0) + v[0]
for f, v in others.items():
    print(f"  {v / nmb:8.1f} instr/MB  {f}")
