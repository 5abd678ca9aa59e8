"""Pin the oracle's single-task siblings against the UNMODIFIED reference modules
src/models/segmentation/nnUNet.py (nnUNet2021) and BTS_UNet.py (BTSUNet): same seed -> bit-identical parameters,
outputs, Dice loss and gradients; writes tests/golden/single_task.pt (digests, outputs at a small size).
Run in the build container only:   python tests/golden/make_golden_single_task.py"""
import hashlib
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import torch_oracle as O  # noqa: E402
from src.models.segmentation.nnUNet import nnUNet2021 as RefNN  # noqa: E402
from src.models.segmentation.BTS_UNet import BTSUNet as RefBTS  # noqa: E402
from src.models.segmentation.ResidualUNet import ResidualUNet as RefRes  # noqa: E402

SEED = 1993
CASES = {
    "nnunet2021": dict(B=2, S=64, ref=lambda: RefNN(sequences=1, regions=1), ora=lambda: O.nnUNet2021(1, 1)),
    "btsunet_ds": dict(B=2, S=64, ref=lambda: RefBTS(sequences=1, regions=1, width=32, deep_supervision=True),
                       ora=lambda: O.BTSUNet(1, 1, 32, True)),
    "btsunet": dict(B=2, S=48, ref=lambda: RefBTS(sequences=1, regions=1, width=16, deep_supervision=False),
                    ora=lambda: O.BTSUNet(1, 1, 16, False)),
    # BatchNorm2d (training mode: batch statistics + running-statistics update) + F.dropout: both sides draw their masks
    # from torch's generator in the same call order, so one seed in front of each forward gives identical outputs
    "residualunet": dict(B=2, S=64, ref=lambda: RefRes(sequences=1, regions=1, width=24),
                         ora=lambda: O.ResidualUNet(1, 1, 24), reseed=True),
}


def digest(sd):
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode()); h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


out = {}
for name, c in CASES.items():
    torch.manual_seed(SEED); ref = c["ref"]()
    torch.manual_seed(SEED); ora = c["ora"]()
    assert list(ref.state_dict()) == list(ora.state_dict())
    assert digest(ref.state_dict()) == digest(ora.state_dict()), name
    img, mask, _, _ = O.synthetic_batch(c["B"], c["S"], c["S"], seed=SEED)
    dice = O.DiceLoss()
    res = []
    for m in (ref, ora):
        m.zero_grad(set_to_none=True)
        if c.get("reseed"):
            torch.manual_seed(SEED + 1)       # same dropout draws on both sides
        outs = m(img)
        lst = outs if isinstance(outs, list) else [outs]
        loss = sum(dice(o, mask) / (n + 1) for n, o in enumerate(reversed(lst)))
        loss.backward()
        res.append((lst, loss.item(), torch.cat([p.grad.flatten() for p in m.parameters() if p.grad is not None]).norm().item()))
    for a, b in zip(res[0][0], res[1][0]):
        assert torch.equal(a, b), name
    assert res[0][1] == res[1][1] and res[0][2] == res[1][2], name
    assert digest(ref.state_dict()) == digest(ora.state_dict()), name      # BatchNorm running statistics moved alike
    gr, go = dict(ref.named_parameters()), dict(ora.named_parameters())
    assert all((gr[k].grad is None) == (go[k].grad is None) for k in gr), name
    torch.manual_seed(SEED); fresh = c["ref"]()
    out[name] = {"seed": SEED, "B": c["B"], "S": c["S"], "state_digest": digest(fresh.state_dict()),
                 "state_digest_after_step": digest(ref.state_dict()),
                 "no_grad_params": sorted(k for k, p in ref.named_parameters() if p.grad is None),
                 "outputs": [o.detach().clone() for o in res[0][0]], "loss": res[0][1], "grad_norm": res[0][2],
                 "n_params": sum(p.numel() for p in ref.parameters()), "reseed": bool(c.get("reseed"))}
    print(name, "ok", out[name]["n_params"], out[name]["loss"])
torch.save(out, os.path.join(HERE, "single_task.pt"))
