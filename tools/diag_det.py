"""GPU tool: where does run-to-run non-determinism of the forward pass start?  Two independent builds of the same model
(deterministic mode unless DET=0) run the same input; the plan tensors are compared in creation order.
    python tools/diag_det.py [arch] [size] [batch]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import torch_oracle as O
from multi_task_breast_cancer_b200 import models as M
arch = sys.argv[1] if len(sys.argv) > 1 else "nnunet"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
B = int(sys.argv[3]) if len(sys.argv) > 3 else 2
det = os.environ.get("DET", "1") != "0"
mk = {"unetpp": lambda: M.MTUNetPlusPlus(deep_supervision=True), "nnunet": lambda: M.MTnnUNet(1, 1, 3)}[arch]
torch.manual_seed(1993)
sd = mk().state_dict()
img, *_ = O.synthetic_batch(B, S, S, device="cuda")
plans = []
for k in range(2):
    m = mk().cuda(); m.load_state_dict(sd); m.set_precision("bf16", deterministic=det)
    with torch.no_grad():
        for rep in range(2):
            m(img)
    torch.cuda.synchronize()
    plan = next(iter(m._plans.values()))
    plans.append((m, plan, {n: t.feat.t.clone() for n, t in plan.tensors.items()}))
    # same plan, run again: does one plan reproduce itself?
    with torch.no_grad():
        m(img)
    torch.cuda.synchronize()
    bad = [n for n, t in plan.tensors.items() if not torch.equal(t.feat.t, plans[-1][2][n])]
    print(f"build {k}: {len(plan.tensors)} tensors; same plan run twice differs in {len(bad)} tensors; first: {bad[:3]}")
a, b = plans[0][2], plans[1][2]
nbad = 0
for n in a:
    if not torch.equal(a[n], b[n]):
        d = (a[n].float() - b[n].float()).abs()
        nbad += 1
        if nbad <= 6:
            print(f"  DIFF {n:45s} shape {tuple(a[n].shape)} frac {(d > 0).float().mean().item():.2e} max {d.max().item():.3e}")
print(f"{nbad} of {len(a)} tensors differ between two builds (deterministic={det})")
