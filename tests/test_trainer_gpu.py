"""GPU tests of the epoch drivers: sync-free metric accumulation, validation, checkpoints in
the reference's format, LR schedule reaching the captured Adam launch, batched test-time inference with refinement.

Integer / bookkeeping work is checked EXACTLY against the oracle's restatement of the reference loops applied to the
very logits the CUDA path produced; the losses are checked against the fp32 oracle model within the bf16 tolerance."""
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _cuda(lib):
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _pair(arch="nnunet"):
    from oracle import torch_oracle as O
    from multi_task_breast_cancer_b200 import models as M
    torch.manual_seed(1993)
    mk = {"nnunet": lambda m: m.MTnnUNet(1, 1, 3),
          "unetpp": lambda m: m.MTUNetPlusPlus(in_channels=1, out_channels=1, n_classes=3, deep_supervision=True)}[arch]
    ref, new = mk(O), mk(M)
    new.load_state_dict(ref.state_dict())
    return ref.cuda(), new.cuda()


def _batches(n, B, S, device="cpu", seed=1993):
    from oracle import torch_oracle as O
    return [O.synthetic_batch(B, S, S, seed=seed + i, device=device)[:3] for i in range(n)]


@pytest.mark.parametrize("use_graph", [True, False])
def test_train_one_epoch_bookkeeping_is_exact_and_losses_match_oracle(use_graph):
    from oracle import torch_oracle as O
    from multi_task_breast_cancer_b200.train import TrainStep
    from multi_task_breast_cancer_b200.trainer import EpochRunner
    ref, new = _pair("nnunet")
    B, S, n = 3, 128, 5
    host = _batches(n, B, S)
    ts = TrainStep(new, (B, 1, S, S), lr=1e-4, use_graph=use_graph)
    runner = EpochRunner(ts, 3)
    seen = []

    def loader():
        # the body after a `yield` runs when the runner asks for the next batch, i.e. right after the previous step's
        # bookkeeping launches: snapshot what that step produced
        for i, (img, mask, onehot) in enumerate(host):
            if i % 2:
                yield {"image": img, "mask": mask, "label": onehot.argmax(1, keepdim=True).float()}   # loader protocol
            else:
                yield (img, mask, onehot)
            seen.append((ts.loss_out.clone(), ts.plan.outputs_seg[-1].clone(), ts.plan.outputs_cls[0].clone()))

    loss, dice, acc, f1w = runner.train_one_epoch(loader())
    assert len(seen) == n
    # exact: same arithmetic as the reference loop on the same logits
    exp_loss = sum(float(s[0][0].item()) for s in seen) / n
    exp_dice = sum(O.hard_dice(host[i][1].cuda(), seen[i][1]) for i in range(n)) / n
    gt, pr = [], []
    for i in range(n):
        gt, pr = O.class_lists([seen[i][2]], host[i][2].cuda(), gt, pr)
    exp_acc, exp_f1 = O.classification_scores(gt, pr)
    assert abs(loss - exp_loss) < 1e-6 * abs(exp_loss)
    # hard Dice: the kernel thresholds logit > 0, the reference sigmoid(logit) > .5 -- the same set except for logits in
    # (0, ~1.2e-7], where fp32 sigmoid rounds to exactly 0.5 (seen once: one pixel of 245 760); integer-exactness of the
    # counts against `logit > 0` is asserted in test_models_gpu.py::test_prediction_refinement_bit_exact
    assert abs(dice - exp_dice) < 2e-5 and abs(acc - exp_acc) < 1e-12 and abs(f1w - exp_f1) < 1e-12
    # against the fp32 oracle running the reference loop on the same data
    opt = O.make_optimizer(ref, 1e-4)
    rl, rd, ra, rf = O.train_one_epoch(ref, opt, [tuple(t.cuda() for t in b) for b in host])
    assert abs(loss - rl) < 1e-2 * abs(rl), (loss, rl)
    assert abs(dice - rd) < 2e-2, (dice, rd)
    assert int(ts.step_dev.item()) == n


def test_validate_one_epoch_matches_oracle_and_leaves_weights_alone():
    """Validation on TRAINED weights: at random init the class logits are O(0.05) apart and a bf16-level difference
    flips argmax, which made the old accuracy / F1 tolerances (0.5 / 0.6) vacuous.  The oracle first trains 40 Adam
    steps; then accuracy and weighted F1 must be IDENTICAL unless the oracle's own top-1 margin on some sample is below
    the bf16 noise floor (then accuracy may move by exactly those samples)."""
    from oracle import torch_oracle as O
    from multi_task_breast_cancer_b200.train import TrainStep
    from multi_task_breast_cancer_b200.trainer import EpochRunner
    ref, new = _pair("unetpp")
    B, S, n = 2, 64, 3
    dev = _batches(n, B, S, device="cuda", seed=7)
    opt = O.make_optimizer(ref, 1e-3)
    for s in range(40):
        O.train_step(ref, opt, *dev[s % n])
    new.load_state_dict(ref.state_dict())
    ts = TrainStep(new, (B, 1, S, S))
    runner = EpochRunner(ts, 3)
    before = ts.flat_p.clone()
    out = runner.validate_one_epoch(iter(dev))
    out2 = runner.validate_one_epoch(iter(dev))          # graph replay path
    exp = O.validate_one_epoch(ref, dev)
    assert torch.equal(before, ts.flat_p) and int(ts.step_dev.item()) == 0
    with torch.no_grad():
        margins = []
        for img, _, _ in dev:
            top = torch.stack(ref(img)[0]).mean(0).topk(2, dim=1).values
            margins += (top[:, 0] - top[:, 1]).tolist()
    unstable = sum(m < 3e-2 for m in margins)
    names = ["val_loss", "val_dice", "val_acc", "val_f1", "seg_val_loss", "cls_val_loss"]
    for k, a, a2, b in zip(names, out, out2, exp):
        if k == "val_acc":
            assert abs(a - b) <= unstable / (n * B) + 1e-12, (k, a, b, margins)
        elif k == "val_f1":
            assert unstable > 0 or abs(a - b) < 1e-12, (k, a, b, margins)
        elif k == "val_dice":
            assert abs(a - b) < 2e-2, (k, a, b)
        else:
            assert abs(a - b) < 1e-2 * abs(b), (k, a, b)
        assert abs(a - a2) < 1e-3 * max(1.0, abs(a)), (k, a, a2)
    assert unstable <= 1, margins      # the trained case is meant to be decisive: at most one borderline sample
    assert abs(out[0] - (0.35 * out[4] + 0.65 * out[5])) < 1e-5
    with pytest.raises(ValueError):
        runner.validate_one_epoch(iter([tuple(t[:, :, :32] if t.dim() == 4 else t for t in dev[0])]))  # other H: error


def test_ragged_last_batch_trains_and_validates_like_the_reference():
    """The reference loaders have no drop_last (BUSI_dataloader.py:146-148) and train_one_epoch trains on the tail batch
    (training_multitask.py:79): batches of 3, 3, 2 go through two plans that share parameters and Adam state."""
    from oracle import torch_oracle as O
    from multi_task_breast_cancer_b200.train import TrainStep
    from multi_task_breast_cancer_b200.trainer import EpochRunner
    ref, new = _pair("nnunet")
    S = 128
    full = _batches(2, 3, S, device="cuda")
    tail = tuple(t[:2].contiguous() for t in _batches(1, 3, S, device="cuda", seed=77)[0])
    data = full + [tail]
    ts = TrainStep(new, (3, 1, S, S), lr=1e-4)
    runner = EpochRunner(ts, 3)
    p0 = ts.flat_p.clone()
    loss, dice, acc, f1w = runner.train_one_epoch(iter(data))
    assert int(ts.step_dev.item()) == 3 and 2 in runner._tail
    t2 = runner._tail[2]
    assert t2.flat_p.data_ptr() == ts.flat_p.data_ptr() and t2.exp_avg.data_ptr() == ts.exp_avg.data_ptr()
    assert t2.step_dev.data_ptr() == ts.step_dev.data_ptr() and t2.plan is not ts.plan
    # the tail step really updated the shared parameters: a third Adam step moves every weight again
    opt = O.make_optimizer(ref, 1e-4)
    rl, rd, ra, rf = O.train_one_epoch(ref, opt, data)
    assert abs(loss - rl) < 1e-2 * abs(rl), (loss, rl)
    assert abs(dice - rd) < 2e-2, (dice, rd)
    ref_flat = torch.cat([p.detach().flatten() for p in ref.parameters()])
    new_flat = torch.cat([p.detach().flatten() for p in new.parameters()])
    moved = (new_flat - ref_flat).abs().max().item()
    assert moved <= 6.2e-4, moved                 # three Adam steps of at most lr = 1e-4 each, on either side
    assert not torch.equal(p0, ts.flat_p)
    # second epoch: both graphs are replayed (no re-capture), still three more steps
    runner.train_one_epoch(iter(data))
    assert int(ts.step_dev.item()) == 6
    # validation on the ragged loader, against the oracle carrying OUR weights (six bf16 / fp32 Adam steps apart the
    # two trainings have drifted by a few percent in the class loss; that is not what is checked here)
    ref.load_state_dict(new.state_dict())
    v = runner.validate_one_epoch(iter(data))
    e = O.validate_one_epoch(ref, data)
    assert abs(v[0] - e[0]) < 1e-2 * abs(e[0]), (v, e)
    assert abs(v[4] - e[4]) < 5e-3 * abs(e[4]) and abs(v[5] - e[5]) < 2e-2 * abs(e[5]), (v, e)
    assert set(runner._eval) == {3, 2}


def test_checkpoint_round_trip_in_reference_format(tmp_path):
    from oracle import torch_oracle as O
    from multi_task_breast_cancer_b200 import models as M
    from multi_task_breast_cancer_b200.train import TrainStep
    from multi_task_breast_cancer_b200.trainer import EpochRunner, FlatAdam, load_pretrained_model, save_checkpoint
    ref, new = _pair("nnunet")
    B, S = 2, 128
    batches = _batches(4, B, S, device="cuda")
    ts = TrainStep(new, (B, 1, S, S), lr=2e-4)
    runner = EpochRunner(ts, 3)
    runner.train_one_epoch(iter(batches[:3]))
    path = str(tmp_path / "model_20260101_fold_0")
    save_checkpoint(path, 0, new, runner.optimizer, 0.5)
    ck = torch.load(path, weights_only=False)
    assert list(ck) == ["epoch", "model_state_dict", "optimizer_state_dict", "scheduler", "val_loss"]
    assert list(ck["model_state_dict"]) == list(ref.state_dict())
    # the reference side reads it: plain torch modules + torch.optim.Adam
    ref.load_state_dict(ck["model_state_dict"])
    opt = O.make_optimizer(ref, 2e-4)
    opt.load_state_dict(ck["optimizer_state_dict"])
    assert float(opt.state_dict()["state"][0]["step"]) == 3.0
    # resume on a fresh fused step: the continuation matches the uninterrupted run
    p_before = ts.flat_p.clone()
    ts.load_batch(*batches[3]); ts.step()
    want = ts.losses().clone()
    p_want = ts.flat_p.clone()
    torch.manual_seed(5)
    fresh = M.MTnnUNet(1, 1, 3).cuda()
    ts2 = TrainStep(fresh, (B, 1, S, S), lr=1e-3)
    load_pretrained_model(fresh, path)
    fa = FlatAdam(ts2)
    fa.load_state_dict(ck["optimizer_state_dict"])
    assert int(ts2.step_dev.item()) == 3 and ts2.lr == pytest.approx(2e-4)
    ts2.load_batch(*batches[3]); ts2.step()
    got = ts2.losses()
    assert abs(got[0].item() - want[0].item()) < 1e-2 * abs(want[0].item())
    # same weights AND same Adam moments: the 4th update has the same direction and size (a fresh Adam state would take
    # a sign-like first step instead); run-to-run atomics noise leaves cos > 0.97
    u1, u2 = (p_want - p_before), (ts2.flat_p - p_before)
    cos = torch.nn.functional.cosine_similarity(u1, u2, dim=0).item()
    assert cos > 0.97 and 0.9 < (u2.norm() / u1.norm()).item() < 1.1, (cos, u1.norm().item(), u2.norm().item())
    # and the oracle continues from the same checkpoint with the same loss (bf16 tolerance)
    tot, *_ = O.train_step(ref, opt, *batches[3])
    assert abs(tot.item() - want[0].item()) < 1e-2 * abs(tot.item())


def test_lr_schedule_reaches_the_captured_adam_kernel():
    from multi_task_breast_cancer_b200.train import TrainStep
    from multi_task_breast_cancer_b200.trainer import EpochRunner, init_lr_scheduler
    _, new = _pair("nnunet")
    B, S = 2, 64
    batches = _batches(1, B, S, device="cuda")
    ts = TrainStep(new, (B, 1, S, S), lr=1e-3)
    runner = EpochRunner(ts, 3)
    sched = init_lr_scheduler(runner.optimizer, "plateau", factor=0.5, min_lr=1e-6, patience=0)
    moved = []
    for epoch in range(3):
        before = ts.flat_p.clone()
        runner.train_one_epoch(iter(batches))
        moved.append((ts.flat_p - before).abs().max().item())
        sched.step(1.0)                        # never improves after the first epoch -> halves from epoch 2 on
    assert ts.lr == pytest.approx(5e-4)                      # what epoch 3 ran with
    runner.optimizer.sync_lr()                               # (pushed at the start of the next epoch otherwise)
    assert ts.lr == pytest.approx(2.5e-4) and ts.lr_dev.item() == pytest.approx(2.5e-4)
    # |Adam update| <= ~lr early on: the halved rate is visible in what the graph's Adam launch did to the weights
    assert moved[0] <= 1.01e-3 and moved[1] <= 1.01e-3 and 0 < moved[2] <= 0.51e-3 and moved[2] < 0.6 * moved[1]


@pytest.mark.parametrize("flags", [(False, False), (True, True)])
def test_batched_inference_with_refinement_is_exact(flags):
    from oracle import torch_oracle as O
    from multi_task_breast_cancer_b200.trainer import inference_multitask
    _, new = _pair("unetpp")
    B, S = 4, 64
    img, mask, onehot, label = O.synthetic_batch(B, S, S, device="cuda")
    with torch.no_grad():
        # push two samples into the refinement branches: a class head that says "normal", an empty prediction
        new.classifier[5].bias.data = torch.tensor([0.0, 0.0, 0.3], device="cuda")
    loader = [{"image": img[:2].cpu(), "mask": mask[:2].cpu(), "label": label[:2].view(-1, 1).float().cpu(),
               "patient_id": torch.tensor([11, 12])},
              {"image": img[2:].cpu(), "mask": mask[2:].cpu(), "label": label[2:].view(-1, 1).float().cpu(),
               "patient_id": torch.tensor([13, 14])}]
    keep = {}
    seg_rows, cls_rows = inference_multitask(new, loader, device="cuda", overlap_seg_based_on_class=flags[0],
                                             overlap_class_based_on_seg=flags[1], keep=keep)
    assert [r["patient_id"] for r in seg_rows] == [11, 12, 13, 14]
    e_seg, e_cls = [], []
    for d, ml, cl in zip(loader, keep["mask_logits"], keep["class_logits"]):
        # the reference's per-image numpy procedure applied to the very logits the batched path used
        s, c = O.inference_multitask(ml, cl, d["mask"].cuda(), d["label"].flatten().long(), flags[0], flags[1])
        e_seg += s; e_cls += c
    assert not torch.equal(keep["mask_logits"][0], keep["mask_logits"][1])   # outputs of different batches do not alias
    for a, b in zip(seg_rows, e_seg):
        for k, v in b.items():
            assert (math.isnan(a[k]) and math.isnan(v)) or a[k] == pytest.approx(v, abs=1e-12), (k, a[k], v)
    for a, b in zip(cls_rows, e_cls):
        assert a["ground_truth"] == b["ground_truth"] and a["predicted_label"] == b["predicted_label"]
    assert len(keep["masks"]) == 2 and keep["masks"][0].dtype == torch.uint8
    if flags[0]:
        assert all(int(m.sum()) == 0 for m in keep["masks"])        # every sample predicted "normal": masks cleared
    # close to the fp32 oracle model as well (bf16 tolerance: a few boundary pixels may differ)
    ref, _ = _pair("unetpp")
    with torch.no_grad():
        ref.classifier[5].bias.data = torch.tensor([0.0, 0.0, 0.3], device="cuda")
        rl, ro = ref(img)
    agree = ((ro[-1] > 0) == (torch.cat(keep["mask_logits"]) > 0)).float().mean().item()
    assert agree > 0.97, agree


def test_row_hausdorff_kernel_is_bit_exact():
    """mtbc_row_hausdorff + the host rule == the reference's haussdorf_distance (utils/metrics.py:236-252): on the
    fixture the reference itself produced (24x24: a partial last bit word) and, against the oracle restatement, on
    ellipse-like and random masks up to 512x512, empty / one-sided-empty cases included."""
    import json
    import numpy as np
    from oracle import torch_oracle as O
    from multi_task_breast_cancer_b200 import _lib, trainer as T
    from multi_task_breast_cancer_b200.ops import ptr

    def run(seg, gt):   # (B,1,H,W) uint8 / float32 on the device -> list of distances
        B, _, H, W = seg.shape
        out = torch.full((B, 4), -1, dtype=torch.int32, device="cuda")
        _lib.call("mtbc_row_hausdorff", ptr(seg), ptr(gt), B, H, W, ptr(out), None)
        torch.cuda.synchronize()
        rows = out.cpu().tolist()
        for r, s, g in zip(rows, seg, gt):
            assert r[2] == int((s != 0).sum()) and r[3] == int((g != 0).sum())
        return [T.hausdorff_from_row_distances(*r) for r in rows]

    fx = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics.json")))
    n = fx["H"] * fx["W"]
    gts = np.stack([np.unpackbits(np.array(c["gt"], dtype=np.uint8))[:n].reshape(1, fx["H"], fx["W"]) for c in fx["cases"]])
    segs = np.stack([np.unpackbits(np.array(c["seg"], dtype=np.uint8))[:n].reshape(1, fx["H"], fx["W"]) for c in fx["cases"]])
    got = run(torch.from_numpy(segs).cuda().contiguous(), torch.from_numpy(gts).float().cuda().contiguous())
    for g, c in zip(got, fx["cases"]):
        assert (c["hausdorff"] is None and math.isnan(g)) or g == c["hausdorff"], (g, c["hausdorff"])

    gen = torch.Generator().manual_seed(5)
    for H, W in [(256, 256), (64, 40), (512, 512), (32, 1000)]:
        B = 5
        _, mask, _, _ = O.synthetic_batch(B, H, W, seed=H + W) if H == W else (None, None, None, None)
        gt = mask if mask is not None else (torch.rand(B, 1, H, W, generator=gen) < 0.2).float()
        seg = ((gt + (torch.rand(B, 1, H, W, generator=gen) < 0.02).float()) > 0).float()
        seg = torch.roll(seg, shifts=(3, -5), dims=(2, 3))
        seg[1] = 0                      # empty prediction
        gt = gt.clone(); gt[2] = 0      # empty ground truth
        seg[3] = 0; gt[3] = 0           # both empty
        got = run(seg.to(torch.uint8).cuda().contiguous(), gt.cuda().contiguous())
        for b in range(B):
            want = O.hausdorff_rows(gt[b:b + 1].numpy(), seg[b:b + 1].numpy())
            assert (math.isnan(want) and math.isnan(got[b])) or got[b] == want, (H, W, b, got[b], want)
    assert math.isnan(got[1]) and math.isnan(got[2]) and got[3] == 0.0
    lib = _lib.load()
    assert lib.mtbc_row_hausdorff(None, None, 1, 8, 2048, None, None) != 0     # W > 1024: an error, not a crash
