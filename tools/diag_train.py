"""GPU diagnostic: loss trajectory of the fused TrainStep vs the fp32 oracle loop on the same fixed batch, plus a quick
step-time measurement.

    python tools/diag_train.py arch B H W steps [--nograph]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

from oracle import torch_oracle as O
from multi_task_breast_cancer_b200 import models as M
from multi_task_breast_cancer_b200.train import TrainStep


def build(mod, arch):
    if arch == "unetpp":
        return mod.MTUNetPlusPlus(in_channels=1, out_channels=1, n_classes=3, deep_supervision=True)
    if arch == "nnunet":
        return mod.MTnnUNet(1, 1, 3)
    return mod.Multi_BTS_UNet(1, 1, 3, 32, True)


def main():
    arch = sys.argv[1]
    B, H, W, steps = (int(v) for v in sys.argv[2:6])
    use_graph = "--nograph" not in sys.argv
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(1993)
    ref = build(O, arch)
    new = build(M, arch)
    new.load_state_dict(ref.state_dict())
    ref, new = ref.cuda(), new.cuda()
    img, mask, onehot, _ = O.synthetic_batch(B, H, W, device="cuda")

    ts = TrainStep(new, (B, 1, H, W), lr=1e-4, eps=1e-4, alpha=0.35, inversely_weighted=True, use_graph=use_graph)
    ts.load_batch(img, mask, onehot)
    print(f"{arch} B{B} {H}x{W}: launches per step {ts.n_launches} (plan {ts.plan.launch_counts()})", flush=True)
    opt = O.make_optimizer(ref, 1e-4)
    traj_new, traj_ref = [], []
    for s in range(steps):
        ts.step()
        traj_new.append(ts.losses().clone())
        tot, seg, cls, _, _ = O.train_step(ref, opt, img, mask, onehot)
        traj_ref.append(torch.stack([tot, seg, cls]))
    torch.cuda.synchronize()
    worst = 0.0
    for s in range(steps):
        a, b = traj_new[s].tolist(), traj_ref[s].tolist()
        d = abs(a[0] - b[0]) / abs(b[0])
        worst = max(worst, d)
        if s < 5 or s % max(1, steps // 20) == 0 or s == steps - 1:
            print(f"  step {s:3d}: total {a[0]:.5f} vs {b[0]:.5f} ({100 * d:.3f}%)  seg {a[1]:.5f}/{b[1]:.5f}  cls {a[2]:.5f}/{b[2]:.5f}  nan {a[3]}")
    print(f"  worst relative deviation of the total loss over {steps} steps: {100 * worst:.3f}%")
    # timing
    for _ in range(3):
        ts.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        ts.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fl = ts.plan.tc_flops_fwd + ts.plan.tc_flops_bwd
    print(f"  step {ms:.3f} ms -> {B / ms * 1000:.1f} img/s; padded tensor-core flops/step {fl / 1e9:.1f} GF -> {fl / ms / 1e9:.1f} TF/s")
    # oracle eager timing on the same GPU (cuDNN fp32, TF32 off)
    for _ in range(2):
        O.train_step(ref, opt, img, mask, onehot)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        O.train_step(ref, opt, img, mask, onehot)
    torch.cuda.synchronize()
    print(f"  oracle torch-eager fp32 step on the same GPU: {(time.perf_counter() - t0) / 5 * 1000:.2f} ms")


if __name__ == "__main__":
    main()
