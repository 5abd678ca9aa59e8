"""Pin the oracle's classification-only siblings against the UNMODIFIED reference modules
src/models/classification/{UnetPlusPlus_Classifier,nnUNet_classifier,BTS_UNET_classifier}.py (the U-Net++ one through
oracle/monai_standin, MONAI being absent): same seed -> bit-identical parameters, outputs, focal loss and gradient norm;
writes tests/golden/classifiers.pt.   Run in the build container only:   python tests/golden/make_golden_classifiers.py"""
import hashlib
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "monai_standin"))
sys.path.insert(0, "/root/reference")

from oracle import torch_oracle as O  # noqa: E402
from src.models.classification.UnetPlusPlus_Classifier import UNetPlusPlusClassifier as RefPP  # noqa: E402
from src.models.classification.nnUNet_classifier import nnUNetClassifier as RefNN  # noqa: E402
from src.models.classification.BTS_UNET_classifier import BTSUNetClassifier as RefBTS  # noqa: E402
from src.utils.criterions import FocalLoss as RefFocal  # noqa: E402

SEED = 1993
CASES = {
    "unetpp_cls": dict(B=2, S=64, ref=lambda: RefPP(spatial_dims=2, in_channels=1, n_classes=3),
                       ora=lambda: O.UNetPlusPlusClassifier(in_channels=1, n_classes=3)),
    "nnunet_cls": dict(B=2, S=64, ref=lambda: RefNN(sequences=1, n_classes=3), ora=lambda: O.nnUNetClassifier(1, 3)),
    "nnunet_cls_binary": dict(B=2, S=64, ref=lambda: RefNN(sequences=1, n_classes=2),
                              ora=lambda: O.nnUNetClassifier(1, 2)),
    "btsunet_cls": dict(B=2, S=128, ref=lambda: RefBTS(sequences=1, classes=3, width=16),
                        ora=lambda: O.BTSUNetClassifier(1, 3, 16)),
}


def digest(sd):
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode()); h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def objective(out, onehot, focal):
    if out.shape[1] == 1:   # binary models emit one logit: BCE-with-logits on the class-1 indicator
        return torch.nn.functional.binary_cross_entropy_with_logits(out, onehot[:, 1:2])
    return focal(out, onehot)


out = {}
for name, c in CASES.items():
    torch.manual_seed(SEED); ref = c["ref"]()
    torch.manual_seed(SEED); ora = c["ora"]()
    assert list(ref.state_dict()) == list(ora.state_dict()), name
    assert digest(ref.state_dict()) == digest(ora.state_dict()), name
    img, _, onehot, _ = O.synthetic_batch(c["B"], c["S"], c["S"], seed=SEED)
    res = []
    for m, focal in ((ref, RefFocal()), (ora, O.FocalLoss())):
        m.zero_grad(set_to_none=True)
        o = m(img)
        loss = objective(o, onehot, focal)
        loss.backward()
        no_grad = sorted(n for n, p in m.named_parameters() if p.grad is None)
        gn = torch.cat([p.grad.flatten() for p in m.parameters() if p.grad is not None]).norm().item()
        res.append((o, loss.item(), gn, no_grad))
    assert torch.equal(res[0][0], res[1][0]), name
    assert res[0][1:] == res[1][1:], name
    out[name] = {"seed": SEED, "B": c["B"], "S": c["S"], "state_digest": digest(ref.state_dict()),
                 "output": res[0][0].detach().clone(), "loss": res[0][1], "grad_norm": res[0][2],
                 "params_without_grad": res[0][3], "n_params": sum(p.numel() for p in ref.parameters())}
    print(name, "ok", out[name]["n_params"], out[name]["loss"], len(res[0][3]))
torch.save(out, os.path.join(HERE, "classifiers.pt"))
