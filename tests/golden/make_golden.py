"""Generate the golden fixtures from the UNMODIFIED reference modules (run in the build container only).

    python tests/golden/make_golden.py

Imports /root/reference/src/models/multitask/{MTUNetPlusPlus,MTnnUNet,Multi_BTS_UNet}.py and
/root/reference/src/utils/criterions.py as they are (MTUNetPlusPlus through oracle/monai_standin, because MONAI 1.3.0
is pinned by the reference but neither vendored nor installable offline -- parity is UNPINNED at that boundary).
For every architecture it
  1. builds the reference model and the oracle restatement under the same seed and asserts bit-identical parameters,
  2. runs the reference forward + Dice/focal criterion + backward and asserts the oracle reproduces it exactly,
  3. runs a short Adam(eps=1e-4) trajectory with the reference loop body (training_multitask.py:87-103),
  4. writes tests/golden/<arch>.pt with the inputs' seed, outputs, losses, gradient norms and the trajectory.
The fixtures travel with the repo; /root/reference does not exist on the GPU box.
"""
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
from monai.losses import DiceLoss as StandinDice  # noqa: E402  (stand-in restating MONAI 1.3.0)
from src.models.multitask.MTUNetPlusPlus import MTUNetPlusPlus as RefUNetPP  # noqa: E402
from src.models.multitask.MTnnUNet import MTnnUNet as RefNNUNet  # noqa: E402
from src.models.multitask.Multi_BTS_UNet import Multi_BTS_UNet as RefBTS  # noqa: E402
from src.utils.criterions import FocalLoss as RefFocal  # noqa: E402
from src.utils.criterions import apply_criterion_multitask_segmentation_classification as ref_apply  # noqa: E402

SEED = 1993  # the reference's seed (src/config.yaml:23)
CASES = {
    "unetpp": dict(B=2, H=64, W=64, steps=12,
                   ref=lambda: RefUNetPP(in_channels=1, out_channels=1, n_classes=3, deep_supervision=True),
                   ora=lambda: O.MTUNetPlusPlus(in_channels=1, out_channels=1, n_classes=3, deep_supervision=True)),
    "nnunet": dict(B=2, H=64, W=64, steps=12,
                   ref=lambda: RefNNUNet(sequences=1, regions=1, n_classes=3),
                   ora=lambda: O.MTnnUNet(sequences=1, regions=1, n_classes=3)),
    "bts": dict(B=2, H=128, W=128, steps=8,
                ref=lambda: RefBTS(sequences=1, regions=1, n_classes=3, width=32, deep_supervision=True),
                ora=lambda: O.Multi_BTS_UNet(sequences=1, regions=1, n_classes=3, width=32, deep_supervision=True)),
}


def state_digest(sd) -> str:
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def fwd_bwd(model, dice, focal, apply_fn, img, mask, onehot):
    model.zero_grad(set_to_none=True)
    logits, outs = model(img)
    seg, cls = apply_fn(dice, mask, outs, focal, onehot, logits, True)
    total = 0.35 * seg + 0.65 * cls
    total.backward()
    return logits, outs, seg.detach(), cls.detach(), total.detach()


def main():
    torch.set_num_threads(8)
    for name, c in CASES.items():
        torch.manual_seed(SEED)
        ref = c["ref"]()
        torch.manual_seed(SEED)
        ora = c["ora"]()
        sd_r, sd_o = ref.state_dict(), ora.state_dict()
        assert list(sd_r.keys()) == list(sd_o.keys()), name
        assert all(torch.equal(sd_r[k], sd_o[k]) for k in sd_r), f"{name}: oracle init differs from the reference"
        digest = state_digest(sd_r)
        img, mask, onehot, label = O.synthetic_batch(c["B"], c["H"], c["W"], seed=SEED)

        dice_ref = StandinDice(include_background=True, sigmoid=True, smooth_dr=1, smooth_nr=1, squared_pred=True)
        lr, or_, seg_r, cls_r, tot_r = fwd_bwd(ref, dice_ref, RefFocal(alpha=1, gamma=2, reduction="mean"), ref_apply,
                                               img, mask, onehot)
        lo, oo, seg_o, cls_o, tot_o = fwd_bwd(ora, O.DiceLoss(), O.FocalLoss(alpha=1, gamma=2), O.multitask_criterion,
                                              img, mask, onehot)
        for a, b in zip(lr + or_, lo + oo):
            assert torch.equal(a, b), f"{name}: oracle forward differs from the reference"
        assert torch.equal(tot_r, tot_o) and torch.equal(seg_r, seg_o) and torch.equal(cls_r, cls_o), name
        gnorm = {}
        pr, po = dict(ref.named_parameters()), dict(ora.named_parameters())
        for k in pr:
            assert (pr[k].grad is None) == (po[k].grad is None), (name, k)
            if pr[k].grad is not None:
                assert torch.allclose(pr[k].grad, po[k].grad, rtol=0, atol=0), (name, k)
                gnorm[k] = float(pr[k].grad.norm())
        # reference refinement inputs/outputs (batched restatement checked against a per-image numpy evaluation)
        rm, rc, cnt = O.refine_predictions(or_[-1], lr[0])
        for b in range(c["B"]):
            m = (torch.sigmoid(or_[-1][b:b + 1]) > .5).float().numpy()
            cls_b = int(lr[0][b].argmax())
            n_b = int((m == 1).sum())
            if cls_b == 2:
                m[m > 0] = 0
            assert (rm[b:b + 1].numpy() == m).all() and int(cnt[b]) == n_b
            assert int(rc[b]) == (2 if n_b == 0 else cls_b)

        # short training trajectory with the reference loop body
        opt = torch.optim.Adam(ref.parameters(), lr=1e-4, eps=1e-4)
        traj = []
        focal = RefFocal(alpha=1, gamma=2, reduction="mean")
        for step in range(c["steps"]):
            opt.zero_grad(set_to_none=True)
            logits, outs = ref(img)
            seg, cls = ref_apply(dice_ref, mask, outs, focal, onehot, logits, True)
            total = 0.35 * seg + 0.65 * cls
            total.backward()
            opt.step()
            traj.append(float(total))
        fixture = {
            "arch": name, "B": c["B"], "H": c["H"], "W": c["W"], "seed": SEED, "state_digest": digest,
            "n_params": sum(p.numel() for p in ref.parameters()), "n_state": len(sd_r),
            "class_logits": [t.detach().clone() for t in lr], "mask_logits": [t.detach().clone() for t in or_],
            "seg_loss": float(seg_r), "cls_loss": float(cls_r), "total_loss": float(tot_r),
            "grad_norms": gnorm, "trajectory": traj,
            "refined_mask_sum": int(rm.sum()), "refined_class": rc.tolist(), "pixel_count": cnt.tolist(),
            "torch": torch.__version__,
        }
        path = os.path.join(HERE, f"{name}.pt")
        torch.save(fixture, path)
        print(f"{name}: params {fixture['n_params']} total loss {fixture['total_loss']:.6f} traj {traj[0]:.5f}->"
              f"{traj[-1]:.5f} -> {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
