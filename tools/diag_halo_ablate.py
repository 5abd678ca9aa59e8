"""GPU tool: ablation matrix of the halo conv kernel on one layer shape (default the 24 -> 24 level-0 layer).
MTBC_HALO_DBG bits: 2 = no MMAs, 4 = no TMA traffic after the first tile, 8 = epilogue releases the accumulator
without reading it.  Each configuration runs in its own process (the flags are read at plan time).
    python tools/diag_halo_ablate.py "24" 24 256      |  python tools/diag_halo_ablate.py "24,24,24,24,48" 24 256"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, torch
sys.path.insert(0, %r)
from multi_task_breast_cancer_b200 import ops
from multi_task_breast_cancer_b200.ops import Feat
src_C = [int(v) for v in sys.argv[1].split(",")]; Cout = int(sys.argv[2]); S = int(sys.argv[3]); N = 32
srcs = [Feat.empty(N, S, S, c) for c in src_C]
for f in srcs: f.t.normal_()
out = Feat.empty(N, S, S, Cout)
offs, ktot = ops.k_offsets(srcs)
wf = torch.randn(9, out.Ck, ktot, device="cuda").to(torch.bfloat16)
ssum = torch.zeros(N, out.Cp, device="cuda"); ssq = torch.zeros(N, out.Cp, device="cuda")
op = ops.conv3x3_fwd_op(srcs, wf, out, stat_sum=ssum, stat_sq=ssq)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for it in range(8):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); op.launch(); e1.record(); torch.cuda.synchronize()
    if it >= 2: ts.append(e0.elapsed_time(e1) * 1e3)
print(f"{sorted(ts)[len(ts)//2]:.1f}")
''' % ROOT
args = sys.argv[1:4] if len(sys.argv) >= 4 else ["24", "24", "256"]
configs = [("baseline", {}), ("no MMA", {"MTBC_HALO_DBG": "2"}), ("no TMA", {"MTBC_HALO_DBG": "4"}),
           ("no epilogue", {"MTBC_HALO_DBG": "8"}), ("no MMA, no TMA", {"MTBC_HALO_DBG": "6"}),
           ("no MMA, no epilogue", {"MTBC_HALO_DBG": "10"}), ("no TMA, no epilogue", {"MTBC_HALO_DBG": "12"}),
           ("skeleton (none of the three)", {"MTBC_HALO_DBG": "14"}), ("one MMA lane", {"MTBC_HALO_LANES": "1"}),
           ("G = 1", {"MTBC_HALO_G": "1"}), ("2 CTAs/SM", {"MTBC_HALO_CTAS": "2"}), ("8 epilogue warps", {"MTBC_HALO_EPI": "2"}),
           ("skeleton, one lane", {"MTBC_HALO_DBG": "14", "MTBC_HALO_LANES": "1"})]
if os.environ.get("ABLATE_ONLY"):
    keep = os.environ["ABLATE_ONLY"].split(";")
    configs = [c for c in configs if c[0] in keep]
configs.append(("split lanes off", {"MTBC_HALO_SPLIT": "0"}))
for name, env in configs:
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, "-c", CHILD, *args], capture_output=True, text=True, env=e)
    print(f"{args[0]}->{args[1]} @{args[2]}  {name:32s} {r.stdout.strip() or r.stderr.strip()[-200:]} us", flush=True)
