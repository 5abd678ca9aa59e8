"""GPU tool: where the end-to-end step (pinned host batch -> H2D -> step -> D2H of the loss) loses time against the
device-resident step.  Variants: device tensors (no H2D), host tensors without the D2H, full."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import torch_oracle as O  # noqa: E402  (synthetic data only)
from multi_task_breast_cancer_b200 import models as M  # noqa: E402
from multi_task_breast_cancer_b200.criterions import refine_predictions  # noqa: E402
from multi_task_breast_cancer_b200.train import TrainStep  # noqa: E402

B, S, steps = 32, 256, 20
torch.manual_seed(1993)
model = M.MTUNetPlusPlus(in_channels=1, out_channels=1, n_classes=3, deep_supervision=True).cuda()
ts = TrainStep(model, (B, 1, S, S))
img, mask, onehot, _ = O.synthetic_batch(B, S, S)
h = (img.pin_memory(), mask.pin_memory(), onehot.pin_memory())
d = (img.cuda(), mask.cuda(), onehot.cuda())
h_loss = torch.zeros(4).pin_memory()
ts.load_batch(*h)


def run(name, load, refine=True, d2h=True):
    for _ in range(3):
        if load is not None: ts.load_batch(*load)
        ts.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        if load is not None: ts.load_batch(*load)
        ts.step()
        if refine: refine_predictions(ts.plan.outputs_seg[-1], ts.plan.outputs_cls[0])
        if d2h: h_loss.copy_(ts.losses(), non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print(f"{name:44s} {e0.elapsed_time(e1) / steps:.3f} ms/step", flush=True)


run("step only (batch resident)", None, refine=False, d2h=False)
run("step + refine (= bench value)", None, d2h=False)
run("step + refine + D2H loss", None)
run("device-tensor load_batch + step + refine + D2H", d)
run("pinned-host load_batch + step + refine", h, d2h=False)
run("pinned-host load_batch + step + refine + D2H", h)


def run2(name, load):
    for _ in range(3):
        ts.load_batch(*load); ts.step(); ts.losses_to_host(h_loss)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        ts.load_batch(*load)
        ts.step()
        refine_predictions(ts.plan.outputs_seg[-1], ts.plan.outputs_cls[0])
        ts.losses_to_host(h_loss)
    e1.record(); torch.cuda.synchronize()
    print(f"{name:44s} {e0.elapsed_time(e1) / steps:.3f} ms/step   loss {h_loss.tolist()}", flush=True)


run2("pinned-host + step + refine + read-back stream", h)
