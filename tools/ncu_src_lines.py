"""Read an .ncu-rep captured with --import-source on: stall samples aggregated per CUDA SOURCE LINE (the SASS view of
tools/ncu_src_top.py says which instruction waits, this one says which line of which warp role).
    python tools/ncu_src_lines.py gpurun_out/x.ncu-rep [n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 20
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(i for i, r in enumerate(rows) if "Line No" in r and "# Samples" in r)
h = rows[hdr]
si, li = h.index("# Samples"), h.index("Line No")
so = h.index("Source")
data = [(int(r[si]), r[li], r[so].strip()[:120]) for r in rows[hdr + 1:] if len(r) > si and r[li].strip().isdigit() and r[si].isdigit() and int(r[si]) > 0]
tot = sum(d[0] for d in data) or 1
print(f"{rows[0][1] if len(rows[0]) > 1 else ''}  {rows[1][1] if len(rows[1]) > 1 else ''}\ntotal samples {tot}")
for d in sorted(data, reverse=True)[:topn]:
    print(f"{d[0]:6d} {100 * d[0] / tot:5.1f}%  line {d[1]:>5s}  {d[2]}")
