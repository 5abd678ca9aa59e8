"""Per-kernel SASS evidence that the tensor kernels are tcgen05 / TMA code (B200_PROFILING.md mnemonics): counts of
UTCHMMA / UTCQMMA (tcgen05.mma), UTMALDG (TMA tensor load), UBLKCP (bulk copy), LDTM (tcgen05.ld), SYNCS (mbarrier),
UTCBAR (tcgen05.commit) in every kernel of libmtbc.so.     python tools/sass_counts.py > profiles/<tag>_sass_counts.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

lib = Path(__file__).resolve().parent.parent / "multi_task_breast_cancer_b200" / "libmtbc.so"
out = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
demangle = lambda s: subprocess.run(["cu++filt", s], capture_output=True, text=True).stdout.strip() or s
keys = ["UTCHMMA", "UTMALDG", "UBLKCP", "LDTM", "UTCBAR", "SYNCS", "HMMA", "REDG", "ATOMG"]
cur, counts, arch = None, collections.OrderedDict(), set()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    if cur:
        for k in keys:
            if re.search(r"\b" + k + r"\b|\b" + k + r"\.", line):
                counts[cur][k] += 1
print(f"# SASS mnemonic counts per kernel of {lib.name} (arch {sorted(arch)}; {len(counts)} kernels)")
print("# tensor kernels = those with UTCHMMA (tcgen05.mma kind::f16 / tf32); UTMALDG = cp.async.bulk.tensor; LDTM = tcgen05.ld")
print("| kernel | " + " | ".join(keys) + " |")
print("|---|" + "---:|" * len(keys))
tot = collections.Counter()
for fn, c in counts.items():
    tot.update(c)
    if c["UTCHMMA"] or c["UTMALDG"] or c["UBLKCP"] or c["LDTM"]:
        name = demangle(fn)
        name = re.sub(r"\(mtbc::\w+\)$|\((?:const |unsigned |float|int|long|void|__nv_bfloat16|mtbc::)[^<>]*\)$", "", name)
        name = name.replace("mtbc::", "").replace("void ", "").replace("(int)", "")
        print(f"| {name} | " + " | ".join(str(c[k]) for k in keys) + " |")
print(f"| **all {len(counts)} kernels** | " + " | ".join(str(tot[k]) for k in keys) + " |")
