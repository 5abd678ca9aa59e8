"""Turn the ncu launch list of one training step (tools/ncu_step.py under ncu, NVTX-renamed kernels) into
profiles/<tag>_launches.md (per plan-launch kind: launches, device time, share of the step, DRAM bytes) and
profiles/traffic.json ({kind: mean dram bytes per launch}) which bench.py reports as roofline.traffic.

    python tools/ncu_summarize.py gpurun_out/r01_launches.csv r01
"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, tag = sys.argv[1], sys.argv[2]
# the plain (un-profiled) run of tools/ncu_step.py that precedes the ncu pass records which build it was
_meta_path = src.replace("_launches.csv", "_step_launches.json")
build_digest = json.load(open(_meta_path)).get("build_digest") if os.path.exists(_meta_path) else None
rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
h = rows[0]
ki, mi, vi, ii, ui = (h.index(c) for c in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Metric Unit"))
per = collections.OrderedDict()
for r in rows[1:]:
    d = per.setdefault(r[ii], {"name": r[ki]})
    d[r[mi]] = float(r[vi].replace(",", ""))
    d["unit:" + r[mi]] = r[ui]
scale_t = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}
scale_b = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
agg = collections.OrderedDict()
labelled = 0
for d in per.values():
    name = d["name"]
    if "|" in name:       # NVTX label "idx|kind|desc" (possibly wrapped as range/kernel by ncu)
        parts = name.split("|")
        kind = parts[1]
        labelled += 1
    else:
        kind = name.split("(")[0].replace("void ", "").replace("mtbc::", "")
    us = d.get("gpu__time_duration.sum", 0.0) * scale_t.get(d.get("unit:gpu__time_duration.sum", "ns"), 1e-3)
    rd = d.get("dram__bytes_read.sum", 0.0) * scale_b.get(d.get("unit:dram__bytes_read.sum", "byte"), 1.0)
    wr = d.get("dram__bytes_write.sum", 0.0) * scale_b.get(d.get("unit:dram__bytes_write.sum", "byte"), 1.0)
    a = agg.setdefault(kind, {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
    a["n"] += 1; a["us"] += us; a["rd"] += rd; a["wr"] += wr
tot = sum(a["us"] for a in agg.values())
lines = [f"# ncu launch list of one training step ({tag})", "",
         f"Source: `{os.path.basename(src)}` — `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
         "--clock-control none` around one eager step of the bench workload (tools/ncu_step.py).  Times are cold-cache and "
         "serialised: compare SHARES with bench.py's CUDA-event numbers, not absolutes.", "",
         f"{len(per)} kernel launches ({labelled} NVTX-labelled), {tot / 1e3:.3f} ms summed device time.", "",
         "| plan launch kind | kernels | time (us) | share | DRAM read (MB) | DRAM write (MB) | DRAM GB/s |",
         "|---|---:|---:|---:|---:|---:|---:|"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    gbs = (a["rd"] + a["wr"]) / (a["us"] * 1e-6) / 1e9 if a["us"] else 0.0
    lines.append(f"| {k} | {a['n']} | {a['us']:.1f} | {100 * a['us'] / tot:.1f}% | {a['rd'] / 1e6:.1f} | {a['wr'] / 1e6:.1f} | {gbs:.0f} |")
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
open(os.path.join(ROOT, "profiles", f"{tag}_launches.md"), "w").write("\n".join(lines) + "\n")
traffic = {k: {"dram_bytes_per_step": a["rd"] + a["wr"], "kernels": a["n"], "us": a["us"]} for k, a in agg.items()}
json.dump({"tag": tag, "build_digest": build_digest, "by_kind": traffic}, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print("\n".join(lines))
