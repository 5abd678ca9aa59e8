"""GPU diagnostic for the classification-only siblings: agreement with the oracle (forward, loss,
head gradients, all-parameter gradient cosine) and eager forward+backward time per batch.  Prints, never asserts."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

from multi_task_breast_cancer_b200 import criterions as Cr, models as M
from oracle import torch_oracle as O   # checker only

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
CASES = {"unetpp_cls": (lambda m: m.UNetPlusPlusClassifier(in_channels=1, n_classes=3), 128, 256),
         "nnunet_cls": (lambda m: m.nnUNetClassifier(1, 3), 128, 256),
         "nnunet_cls_binary": (lambda m: m.nnUNetClassifier(1, 2), 128, 256),
         "btsunet_cls": (lambda m: m.BTSUNetClassifier(1, 3, 16), 128, 128),
         "btsunet_cls_w32": (lambda m: m.BTSUNetClassifier(1, 3, 32), 128, 128)}


def objective(out, onehot, focal):
    if out.shape[1] == 1:
        return torch.nn.functional.binary_cross_entropy_with_logits(out, onehot[:, 1:2])
    return focal(out, onehot)


for name, (mk, S, Sb) in CASES.items():
    try:
        torch.manual_seed(1993)
        ref = mk(O).cuda()
        new = mk(M)
        new.load_state_dict(ref.state_dict())
        new = new.cuda()
        img, _, onehot, _ = O.synthetic_batch(4, S, S, device="cuda")
        ro, no = ref(img), new(img)
        lr_, ln_ = objective(ro, onehot, O.FocalLoss()), objective(no, onehot, Cr.init_criterion_classification(3, None, "Focal"))
        lr_.backward(); ln_.backward()
        pr, pn = dict(ref.named_parameters()), dict(new.named_parameters())
        ga = torch.cat([pn[n].grad.flatten() for n in pn if pn[n].grad is not None and pr[n].grad is not None])
        gb = torch.cat([pr[n].grad.flatten() for n in pn if pn[n].grad is not None and pr[n].grad is not None])
        cos = torch.nn.functional.cosine_similarity(ga, gb, dim=0).item()
        worst = max(((pn[n].grad - pr[n].grad).norm() / (pr[n].grad.norm() + 1e-12)).item()
                    for n in pn if pn[n].grad is not None and pr[n].grad is not None and pr[n].grad.norm() > 1e-6)
        print(f"{name}: out |ref|max {ro.abs().max().item():.4f} max abs err {(no - ro).abs().max().item():.5f} "
              f"loss {ln_.item():.6f} vs {lr_.item():.6f} grad cosine {cos:.5f} worst per-param rel {worst:.4f} "
              f"none-grad new {sum(p.grad is None for p in pn.values())} ref {sum(p.grad is None for p in pr.values())}",
              flush=True)
        del ref
        B = 32
        img, _, onehot, _ = O.synthetic_batch(B, Sb, Sb, device="cuda")
        focal = Cr.init_criterion_classification(3, None, "Focal")
        for it in range(8):
            if it == 3:
                torch.cuda.synchronize(); t0 = time.perf_counter()
            for p in new.parameters():
                p.grad = None
            objective(new(img), onehot, focal).backward()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 5 * 1e3
        print(f"{name}: eager forward+backward B={B} {Sb}x{Sb}: {ms:.2f} ms ({B / ms * 1e3:.0f} img/s)", flush=True)
    except Exception as e:  # keep going: this is a diagnostic
        print(f"{name}: FAILED {type(e).__name__}: {e}", flush=True)
    torch.cuda.empty_cache()
