"""GPU diagnostic: run-to-run spread of the 200-step loss-trajectory deviation against the fp32 oracle loop
(tests/test_models_gpu.py::test_loss_trajectory_200_steps), default and deterministic mode.

    python tools/diag_traj.py arch B S steps reps [det]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from oracle import torch_oracle as O
from multi_task_breast_cancer_b200 import models as M
from multi_task_breast_cancer_b200.train import TrainStep

arch, B, S, steps, reps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
det = len(sys.argv) > 6 and sys.argv[6] == "det"


def build(mod):
    if arch == "unetpp":
        return mod.MTUNetPlusPlus(in_channels=1, out_channels=1, n_classes=3, deep_supervision=True)
    if arch == "nnunet":
        return mod.MTnnUNet(1, 1, 3)
    return mod.Multi_BTS_UNet(1, 1, 3, 32, True)


for rep in range(reps):
    torch.manual_seed(1993)
    ref = build(O); new = build(M); new.load_state_dict(ref.state_dict())
    ref, new = ref.cuda(), new.cuda()
    if det:
        new.set_precision("bf16", deterministic=True)
    img, mask, onehot, _ = O.synthetic_batch(B, S, S, device="cuda")
    ts = TrainStep(new, (B, 1, S, S), lr=1e-4, eps=1e-4, alpha=0.35, inversely_weighted=True)
    ts.load_batch(img, mask, onehot)
    opt = O.make_optimizer(ref, 1e-4)
    worst, at, dev = 0.0, -1, []
    for s in range(steps):
        ts.step()
        mine = ts.losses().clone()
        tot, *_ = O.train_step(ref, opt, img, mask, onehot)
        d = abs(mine[0].item() - tot.item()) / abs(tot.item())
        dev.append(d)
        if d > worst:
            worst, at = d, s
    dev.sort()
    print(f"{arch} B{B} {S}x{S} det={int(det)} rep {rep}: worst {100 * worst:.3f}% at step {at}; "
          f"median {100 * dev[len(dev) // 2]:.3f}%, p90 {100 * dev[int(.9 * len(dev))]:.3f}%, loss {tot.item():.4f}", flush=True)
    del ts, new, ref, opt
