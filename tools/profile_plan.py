"""GPU tool: CUDA-event time of every launch of one training step (eager replay of the launch list), top-N report.

    python tools/profile_plan.py [arch B S topN]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import torch
from oracle import torch_oracle as O
from multi_task_breast_cancer_b200 import models as M
from multi_task_breast_cancer_b200.ops import stream_ptr
from multi_task_breast_cancer_b200.train import TrainStep

arch = sys.argv[1] if len(sys.argv) > 1 else "unetpp"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
S = int(sys.argv[3]) if len(sys.argv) > 3 else 256
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 60
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
            acc[i] += evs[i].elapsed_time(evs[i + 1]) / (reps - 1)
tot = sum(acc)
print(f"{arch} B{B} {S}x{S}: {len(launches)} launches, sum of per-launch times {tot:.3f} ms")
kinds = {}
for l, t in zip(launches, acc):
    k = kinds.setdefault(l.kind, [0.0, 0, 0.0]); k[0] += t; k[1] += 1; k[2] += getattr(l, "true_flops", 0.0)
for k, (t, n, fl) in sorted(kinds.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:28s} {t:8.3f} ms  {n:4d} launches  {100 * t / tot:5.1f}%  {fl / t / 1e9 if fl else 0:8.1f} TF/s(true)")
print("top launches:")
order = sorted(range(len(launches)), key=lambda i: -acc[i])[:topn]
for i in order:
    l = launches[i]
    fl = getattr(l, "true_flops", 0.0)
    print(f"  {acc[i]:7.3f} ms  {fl / acc[i] / 1e9 if fl else 0:7.1f} TF/s  {l.kind:16s} {getattr(l, 'desc', '')}")
