"""GPU tool for ncu: runs ONE eager training step of the bench workload between cudaProfilerStart/Stop and writes the
ordered (kind, desc, true_flops) list of its launches so that tools/ncu_summarize.py can label ncu's launch list.

    python tools/ncu_step.py gpurun_out/r01_step_launches.json &&
    ncu --profile-from-start off --nvtx --print-nvtx-rename kernel --print-units base --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --csv --log-file gpurun_out/r01_launches.csv python tools/ncu_step.py /dev/null
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

out = sys.argv[1] if len(sys.argv) > 1 else "/dev/null"
arch = os.environ.get("ARCH", "unetpp")
B = int(os.environ.get("BATCH", "32"))
S = int(os.environ.get("SIZE", "256"))
torch.manual_seed(1993)
model = {"unetpp": lambda: M.MTUNetPlusPlus(deep_supervision=True), "nnunet": lambda: M.MTnnUNet(1, 1, 3),
         "bts": lambda: M.Multi_BTS_UNet(1, 1, 3, 32, True)}[arch]().cuda()
ts = TrainStep(model, (B, 1, S, S), use_graph=False)
img, mask, onehot, _ = O.synthetic_batch(B, S, S, device="cuda")
ts.load_batch(img, mask, onehot)
launches = [l for l in ts.launches_fb + ts.launches_opt if l.kind != "bucket_ready"]
st = C.c_void_p(stream_ptr())
for _ in range(3):
    for l in launches:
        l(st)
torch.cuda.synchronize()
rt = torch.cuda.cudart()
only = os.environ.get("ONLY")          # profile just the launches whose "kind|desc" label contains this substring
if only is None:
    rt.cudaProfilerStart()
for i, l in enumerate(launches):
    # NVTX range = label of the plan launch; `ncu --nvtx --print-nvtx-rename kernel` renames the kernels with it
    label = f"{i}|{l.kind}|{getattr(l, 'desc', '')}"
    hit = only is not None and only in label
    if hit:
        torch.cuda.synchronize()
        rt.cudaProfilerStart()
    torch.cuda.nvtx.range_push(label)
    l(st)
    torch.cuda.nvtx.range_pop()
    if hit:
        torch.cuda.synchronize()
        rt.cudaProfilerStop()
torch.cuda.synchronize()
if only is None:
    rt.cudaProfilerStop()
meta = [{"kind": l.kind, "desc": getattr(l, "desc", ""), "flops": getattr(l, "true_flops", 0.0),
         "nk": getattr(l, "n_kernels", 1)} for l in launches]
if out != "/dev/null":
    from multi_task_breast_cancer_b200 import build as _b
    json.dump({"arch": arch, "B": B, "S": S, "build_digest": _b.lib_digest(), "launches": meta}, open(out, "w"))
print(f"ncu_step: {len(launches)} plan launches")
