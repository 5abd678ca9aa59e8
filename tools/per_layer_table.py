"""GPU tool: the per-layer dual-roof table SURVEY 8d asks for.  Every launch of one training step is timed with CUDA
events on the launching stream (eager replay, 3 repetitions after a warm-up) and set against BOTH roofs:

    t_tensor = true FLOPs / bf16 peak        t_hbm = algorithmic bytes / HBM peak       bound = max(t_tensor, t_hbm)
    fraction = bound / measured time

FLOPs: 2*M*N*K with true channel counts; bytes: every tensor read once and written once in its stored dtype (plan.py
`_annot` / `_mk_op(true_bytes=)`).  Peaks: MEASURED_PEAKS.json (sustained bf16, copy bandwidth), fallback from the
profiling guide if absent.  Output: markdown on stdout (tee it into profiles/<tag>_per_layer.md).

    python tools/per_layer_table.py [arch B S]
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import torch_oracle as O  # synthetic batch generator only
from multi_task_breast_cancer_b200 import models as M
from multi_task_breast_cancer_b200.ops import stream_ptr
from multi_task_breast_cancer_b200.train import TrainStep

arch = sys.argv[1] if len(sys.argv) > 1 else "unetpp"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
S = int(sys.argv[3]) if len(sys.argv) > 3 else 256
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pk):
    d = json.load(open(pk))
    TC, HBM, src = d.get("bf16_tflops_sustained", d["bf16_tflops"]) * 1e12, d["hbm_gbs"] * 1e9, "MEASURED_PEAKS.json"
else:
    TC, HBM, src = 1400e12, 6650e9, "fallback (B200_PROFILING.md)"
torch.manual_seed(1993)
model = {"unetpp": lambda: M.MTUNetPlusPlus(deep_supervision=True), "nnunet": lambda: M.MTnnUNet(1, 1, 3),
         "bts": lambda: M.Multi_BTS_UNet(1, 1, 3, 32, True)}[arch]().cuda()
ts = TrainStep(model, (B, 1, S, S), use_graph=False)
img, mask, onehot, _ = O.synthetic_batch(B, S, S, device="cuda")
ts.load_batch(img, mask, onehot)
launches = [l for l in ts.launches_fb + ts.launches_opt if l.kind != "bucket_ready"]
st = C.c_void_p(stream_ptr())
reps = 4
acc = [0.0] * len(launches)
for rep in range(reps):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(launches) + 1)]
    evs[0].record()
    for i, l in enumerate(launches):
        l(st); evs[i + 1].record()
    torch.cuda.synchronize()
    if rep:
        for i in range(len(launches)):
            acc[i] += evs[i].elapsed_time(evs[i + 1]) * 1e3 / (reps - 1)   # us
tot = sum(acc)
print(f"# Per-layer dual-roof table: {arch} B={B} {S}x{S}, one training step\n")
print(f"Peaks ({src}): bf16 {TC / 1e12:.1f} TFLOP/s (sustained), HBM {HBM / 1e9:.1f} GB/s.  "
      f"{len(launches)} launches, {tot / 1e3:.3f} ms summed per-launch CUDA-event time (eager, launch gaps included "
      "in each launch's interval; the graph replay of the same list is what bench.py times).\n")
print("`bound` = the larger of FLOPs/peak and bytes/peak; `frac` = bound / time.  Bytes are ALGORITHMIC (read once, "
      "write once, stored dtype); the two-pass InstanceNorm backward moves 10 B/elem against 6 B/elem algorithmic.\n")
kinds = {}
rows = []
for l, t in zip(launches, acc):
    fl, by = getattr(l, "true_flops", 0.0), getattr(l, "true_bytes", 0.0)
    t_tc, t_hbm = fl / TC * 1e6, by / HBM * 1e6
    bound = max(t_tc, t_hbm)
    which = "-" if bound == 0 else ("tensor" if t_tc >= t_hbm else "hbm")
    rows.append((l.kind, getattr(l, "desc", ""), fl, by, t_tc, t_hbm, which, t, bound / t if t > 0 else 0.0))
    k = kinds.setdefault(l.kind, [0, 0.0, 0.0, 0.0, 0.0])
    k[0] += 1; k[1] += t; k[2] += bound; k[3] += fl; k[4] += by
print("## By kind\n")
print("| kind | launches | time (us) | share | GFLOP | MB | sum of bounds (us) | frac |")
print("|---|---:|---:|---:|---:|---:|---:|---:|")
for k, (n, t, b, fl, by) in sorted(kinds.items(), key=lambda kv: -kv[1][1]):
    print(f"| {k} | {n} | {t:.1f} | {100 * t / tot:.1f}% | {fl / 1e9:.1f} | {by / 1e6:.1f} | {b:.1f} | {b / t if t else 0:.2f} |")
tb = sum(k[2] for k in kinds.values())
print(f"| **all** | {len(launches)} | {tot:.1f} | 100% | {sum(k[3] for k in kinds.values()) / 1e9:.1f} | "
      f"{sum(k[4] for k in kinds.values()) / 1e6:.1f} | {tb:.1f} | {tb / tot:.2f} |\n")
print("## Every launch (forward order, then backward)\n")
print("| # | kind | layer | GFLOP | MB | t_tensor (us) | t_hbm (us) | bound | time (us) | frac |")
print("|---:|---|---|---:|---:|---:|---:|---|---:|---:|")
for i, (kind, desc, fl, by, t_tc, t_hbm, which, t, fr) in enumerate(rows):
    if fl == 0 and by == 0 and t < 5:
        continue
    print(f"| {i} | {kind} | {desc} | {fl / 1e9:.2f} | {by / 1e6:.1f} | {t_tc:.1f} | {t_hbm:.1f} | {which} | {t:.1f} | {fr:.2f} |")
