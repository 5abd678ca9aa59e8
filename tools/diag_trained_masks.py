"""Thresholded-mask / class-argmax agreement with the fp32 oracle on TRAINED weights (north_star: >= 99.9 %).

At random init the mask logits hover around 0 and a bf16 rounding flips ~0.5 % of the pixels; after the oracle has
trained for a few hundred Adam steps the logits have moved away from the threshold.  Usage:
    python tools/diag_trained_masks.py [arch] [B] [S] [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import torch_oracle as O  # noqa: E402  (checker)
from multi_task_breast_cancer_b200 import models as M  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "unetpp"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
S = int(sys.argv[3]) if len(sys.argv) > 3 else 128
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 200
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def build(mod):
    if arch == "unetpp":
        return mod.MTUNetPlusPlus(in_channels=1, out_channels=1, n_classes=3, deep_supervision=True)
    if arch == "nnunet":
        return mod.MTnnUNet(1, 1, 3)
    return mod.Multi_BTS_UNet(1, 1, 3, 32, True)


torch.manual_seed(1993)
ref = build(O).cuda()
batches = [O.synthetic_batch(B, S, S, seed=1993 + i, device="cuda") for i in range(4)]
opt = O.make_optimizer(ref, 1e-4)
for s in range(steps + 1):
    if s % 50 == 0:
        new = build(M)
        new.load_state_dict(ref.state_dict())
        new = new.cuda()
        agree, n, cls_ok, relerr = 0.0, 0, True, 0.0
        for img, mask, onehot, _ in batches:
            with torch.no_grad():
                rl, ro = ref(img)
                nl, no = new(img)
            agree += ((no[-1] > 0) == (ro[-1] > 0)).float().mean().item()
            relerr = max(relerr, ((no[-1] - ro[-1]).norm() / ro[-1].norm()).item())
            cls_ok &= torch.equal(nl[-1].argmax(1), rl[-1].argmax(1))
            n += 1
        print(f"{arch} after {s:4d} oracle steps: masks identical {100 * agree / n:.4f} %  class argmax identical {cls_ok}"
              f"  full-decoder logits rel L2 {relerr:.4f}", flush=True)
    img, mask, onehot, _ = batches[s % 4]
    O.train_step(ref, opt, img, mask, onehot)
