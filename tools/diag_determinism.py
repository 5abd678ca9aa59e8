"""GPU tool: run-to-run spread of the gradients (fp32 atomics reorder between runs).  Builds the same model twice with
identical weights / inputs and prints the cosine between the two flat gradient vectors, several trials.
    python tools/diag_determinism.py [arch] [size] [batch] [trials]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import torch_oracle as O
from multi_task_breast_cancer_b200 import models as M, criterions as Cr
arch = sys.argv[1] if len(sys.argv) > 1 else "nnunet"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
B = int(sys.argv[3]) if len(sys.argv) > 3 else 2
trials = int(sys.argv[4]) if len(sys.argv) > 4 else 6
mk = {"unetpp": lambda: M.MTUNetPlusPlus(deep_supervision=True), "nnunet": lambda: M.MTnnUNet(1, 1, 3)}[arch]
torch.manual_seed(1993)
ref = mk().cuda()
sd = ref.state_dict()
img, mask, onehot, _ = O.synthetic_batch(B, S, S, device="cuda")
def grads():
    m = mk().cuda(); m.load_state_dict(sd)
    logits, outs = m(img)
    seg, cls = Cr.apply_criterion_multitask_segmentation_classification(
        Cr.init_criterion_segmentation("DICE"), mask, outs, Cr.init_criterion_classification(3, None, "Focal"), onehot, logits, True)
    (0.35 * seg + 0.65 * cls).backward()
    torch.cuda.synchronize()
    return torch.cat([p.grad.flatten() for p in m.parameters() if p.grad is not None]), float(seg), float(cls)
g0, s0, c0 = grads()
for t in range(trials):
    g, s, c = grads()
    cos = torch.nn.functional.cosine_similarity(g0, g, dim=0).item()
    print(f"trial {t}: cos {cos:.5f}  rel {((g - g0).norm() / g0.norm()).item():.4f}  seg {s:.6f} ({s0:.6f}) cls {c:.6f} ({c0:.6f})", flush=True)
