"""Pin the oracle restatement against the fixtures generated from the UNMODIFIED reference modules
(tests/golden/make_golden.py, run where /root/reference exists) and against hand-checkable anchors."""
import hashlib
import math
import os

import pytest
import torch

from oracle import torch_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


def _digest(sd):
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def _build(arch):
    torch.manual_seed(1993)
    if arch == "unetpp":
        return O.MTUNetPlusPlus(in_channels=1, out_channels=1, n_classes=3, deep_supervision=True)
    if arch == "nnunet":
        return O.MTnnUNet(sequences=1, regions=1, n_classes=3)
    return O.Multi_BTS_UNet(sequences=1, regions=1, n_classes=3, width=32, deep_supervision=True)


@pytest.mark.parametrize("arch", ["unetpp", "nnunet", "bts"])
def test_oracle_matches_reference_fixture(arch):
    fx = torch.load(os.path.join(HERE, "golden", f"{arch}.pt"), weights_only=False)
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    model = _build(arch)
    assert sum(p.numel() for p in model.parameters()) == fx["n_params"]
    assert len(model.state_dict()) == fx["n_state"]
    assert _digest(model.state_dict()) == fx["state_digest"], "seeded init differs from the reference's"
    img, mask, onehot, _ = O.synthetic_batch(fx["B"], fx["H"], fx["W"], seed=fx["seed"])
    logits, outs = model(img)
    for a, b in zip(logits, fx["class_logits"]):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-5)
    for a, b in zip(outs, fx["mask_logits"]):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-5)
    seg, cls = O.multitask_criterion(O.DiceLoss(), mask, outs, O.FocalLoss(), onehot, logits, True)
    total = 0.35 * seg + 0.65 * cls
    assert abs(float(seg) - fx["seg_loss"]) < 1e-5 and abs(float(cls) - fx["cls_loss"]) < 1e-5
    assert abs(float(total) - fx["total_loss"]) < 1e-5
    total.backward()
    for n, p in model.named_parameters():
        if n in fx["grad_norms"]:
            assert p.grad is not None, n
            assert abs(float(p.grad.norm()) - fx["grad_norms"][n]) <= 1e-3 * fx["grad_norms"][n] + 1e-7, n
        else:
            assert p.grad is None, n
    rm, rc, cnt = O.refine_predictions(outs[-1].detach(), logits[0].detach())
    assert int(rm.sum()) == fx["refined_mask_sum"] and rc.tolist() == fx["refined_class"]
    assert cnt.tolist() == fx["pixel_count"]


def test_oracle_trajectory_matches_reference_fixture():
    fx = torch.load(os.path.join(HERE, "golden", "nnunet.pt"), weights_only=False)
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    model = _build("nnunet")
    opt = O.make_optimizer(model, 1e-4)
    img, mask, onehot, _ = O.synthetic_batch(fx["B"], fx["H"], fx["W"], seed=fx["seed"])
    for step, want in enumerate(fx["trajectory"][:6]):
        tot, *_ = O.train_step(model, opt, img, mask, onehot)
        assert abs(float(tot) - want) < 2e-4 * abs(want), (step, float(tot), want)


def test_hand_anchors():
    # Dice with all-zero logits and an all-zero mask on N pixels = 1 - 1/(N/4 + 1); N = 16 -> 0.8
    d = O.DiceLoss()(torch.zeros(1, 1, 4, 4), torch.zeros(1, 1, 4, 4))
    assert abs(float(d) - 0.8) < 1e-6
    # focal with logits [0,0,0] and any one-hot target = (2/3)^2 * ln 3
    f = O.FocalLoss(alpha=1, gamma=2)(torch.zeros(2, 3), torch.tensor([[1., 0, 0], [0, 0, 1.]]))
    assert abs(float(f) - (2 / 3) ** 2 * math.log(3)) < 1e-6


def test_refinement_semantics():
    mask_logits = torch.full((3, 1, 4, 4), -1.0)
    mask_logits[0, 0, 1, 1] = 2.0      # one tumour pixel, class benign -> kept
    mask_logits[1, 0, :2, :2] = 3.0    # 4 pixels but predicted normal -> mask zeroed, class stays normal
    cls_logits = torch.tensor([[2.0, 0.1, 0.0], [0.0, 0.1, 3.0], [0.0, 2.0, 0.1]])  # sample 2: malignant, empty mask
    m, c, cnt = O.refine_predictions(mask_logits, cls_logits)
    assert cnt.tolist() == [1, 4, 0]
    assert m[0].sum() == 1 and m[1].sum() == 0 and m[2].sum() == 0
    assert c.tolist() == [0, 2, 2]  # empty mask forces "normal"
    m2, c2, _ = O.refine_predictions(mask_logits, cls_logits, seg_by_class=False, class_by_seg=False)
    assert m2[1].sum() == 4 and c2.tolist() == [0, 2, 1]
