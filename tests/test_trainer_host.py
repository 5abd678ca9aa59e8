"""CPU tests of the epoch-level host logic: metric helpers against the reference-generated
fixture and scikit-learn, the torch.optim.Adam state_dict <-> flat buffer mapping (the reference's
'optimizer_state_dict'), and LR schedulers driving the fused step's learning rate."""
import json
import math
import os

import numpy as np
import pytest
import torch

from oracle import torch_oracle as O
from multi_task_breast_cancer_b200 import trainer as T
from multi_task_breast_cancer_b200.plan import flat_layout

HERE = os.path.dirname(os.path.abspath(__file__))


def _fixture():
    fx = json.load(open(os.path.join(HERE, "golden", "metrics.json")))
    n = fx["H"] * fx["W"]
    for c in fx["cases"]:
        gt = np.unpackbits(np.array(c["gt"], dtype=np.uint8))[:n].reshape(1, 1, fx["H"], fx["W"])
        seg = np.unpackbits(np.array(c["seg"], dtype=np.uint8))[:n].reshape(1, 1, fx["H"], fx["W"])
        yield gt, seg, c


def _same(a, b):
    return (b is None and math.isnan(a)) or (b is not None and abs(a - b) < 1e-12)


def test_segmentation_metrics_match_reference_fixture():
    """Oracle restatement and the product's count-based helper against what src/utils/metrics.py itself returned."""
    n = 0
    for gt, seg, c in _fixture():
        ora = O.segmentation_metrics(gt, seg)
        tp = int((seg & gt).sum()); fp = int((seg & (1 - gt)).sum()); fn = int(((1 - seg) & gt).sum())
        tn = gt.size - tp - fp - fn
        mine = T.segmentation_metrics_from_counts(tp, fp, fn, tn)
        for k, v in c["metrics"].items():
            assert _same(float(ora[k]), v), (k, ora[k], v)
            assert _same(float(mine[k]), v), (k, mine[k], v)
        assert abs(O.hard_dice(torch.from_numpy(gt).float(), torch.from_numpy(seg).float() * 2 - 1) - c["hard_dice"]) < 1e-12
        # Hausdorff distance as the reference computes it (rows are the points): oracle bit-exact, and the product's
        # host rule fed with the integers the device kernel produces (computed here in numpy)
        ho = O.hausdorff_rows(gt, seg)
        assert (c["hausdorff"] is None and math.isnan(ho)) or ho == c["hausdorff"]
        d2 = (seg[0, 0][:, None, :] != gt[0, 0][None, :, :]).sum(-1)
        hm = T.hausdorff_from_row_distances(int(d2.min(1).max()), int(d2.min(0).max()), int(seg.sum()), int(gt.sum()))
        assert (c["hausdorff"] is None and math.isnan(hm)) or hm == c["hausdorff"]
        assert math.isnan(mine["Haussdorf distance"])    # not supplied -> NaN
        n += 1
    assert n >= 20


def test_classification_scores_match_sklearn():
    from sklearn.metrics import accuracy_score, f1_score
    rng = np.random.default_rng(7)
    for trial in range(20):
        n = int(rng.integers(1, 60))
        gt = rng.integers(0, 3, n)
        pr = rng.integers(0, 3 if trial % 3 else 2, n)   # some trials never predict class 2
        if trial == 5:
            gt[:] = 1                                     # a single true class
        conf = np.zeros((3, 3), dtype=np.int64)
        for g, p in zip(gt, pr):
            conf[g, p] += 1
        acc, f1w = T.classification_scores(conf.tolist())
        assert abs(acc - accuracy_score(gt, pr)) < 1e-12
        assert abs(f1w - f1_score(gt, pr, labels=[0, 1, 2], average="weighted", zero_division=0)) < 1e-12
        oa, of = O.classification_scores([float(v) for v in gt], [float(v) for v in pr])
        assert abs(oa - acc) < 1e-12 and abs(of - f1w) < 1e-12


def _small_model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Conv2d(1, 3, 3), torch.nn.Conv2d(3, 2, 1), torch.nn.Linear(5, 4))


def test_adam_state_dict_round_trip_through_flat_buffers():
    """torch.optim.Adam.state_dict() (what the reference stores as 'optimizer_state_dict') -> flat moment buffers
    -> back, bit-exact, in the layout TrainStep uses (reverse registration order, 64-float slots)."""
    m = _small_model()
    opt = torch.optim.Adam(m.parameters(), lr=3e-4, eps=1e-4)
    for _ in range(3):
        opt.zero_grad()
        sum((p ** 2).sum() for p in m.parameters()).backward()
        opt.step()
    sd = opt.state_dict()
    params = dict(m.named_parameters())
    names = list(params)
    ranges, total = flat_layout(params)
    ea, es = torch.full((total,), 7.0), torch.full((total,), 7.0)
    step, group = T.adam_state_dict_to_flat(sd, names, ranges, ea, es)
    assert step == 3 and group["lr"] == 3e-4 and group["eps"] == 1e-4
    for i, n in enumerate(names):
        a, b = ranges[n]
        assert torch.equal(ea[a:b].view_as(params[n]), sd["state"][i]["exp_avg"])
        assert torch.equal(es[a:b].view_as(params[n]), sd["state"][i]["exp_avg_sq"])
    used = sum(p.numel() for p in params.values())
    assert ea.abs().sum() > 0 and (ea == 0).sum() >= total - used      # padding between slots is zeroed
    back = T.adam_state_dict_from_flat(names, ranges, {n: p.shape for n, p in params.items()}, ea, es, step, group)
    assert back["param_groups"][0]["params"] == sd["param_groups"][0]["params"]
    for i in sd["state"]:
        assert float(back["state"][i]["step"]) == float(sd["state"][i]["step"])
        assert torch.equal(back["state"][i]["exp_avg"], sd["state"][i]["exp_avg"])
        assert torch.equal(back["state"][i]["exp_avg_sq"], sd["state"][i]["exp_avg_sq"])
    # and torch accepts it: a fresh Adam continues from the restored state exactly like the original
    m2 = _small_model()
    m2.load_state_dict(m.state_dict())
    opt2 = torch.optim.Adam(m2.parameters(), lr=3e-4, eps=1e-4)
    opt2.load_state_dict(back)
    for o, mm in ((opt, m), (opt2, m2)):
        o.zero_grad()
        sum((p ** 2).sum() for p in mm.parameters()).backward()
        o.step()
    assert all(torch.equal(a, b) for a, b in zip(m.parameters(), m2.parameters()))
    with pytest.raises(ValueError):
        T.adam_state_dict_to_flat({"state": {}, "param_groups": [{"params": [0]}]}, names, ranges, ea, es)


class _StubStep:
    """The attributes FlatAdam reads from a TrainStep, on the CPU."""

    def __init__(self, model, lr):
        self.model = model
        self.lr = lr
        self.lr_dev = torch.full((1,), lr)
        self.betas, self.eps = (0.9, 0.999), 1e-4
        params = dict(model.named_parameters())
        self.param_ranges, total = flat_layout(params)
        self.exp_avg, self.exp_avg_sq = torch.zeros(total), torch.zeros(total)
        self.step_dev = torch.zeros(1, dtype=torch.int32)
        self.pushed = []

    def set_lr(self, lr):
        self.lr = lr
        self.lr_dev.fill_(lr)
        self.pushed.append(lr)


@pytest.mark.parametrize("kind", ["plateau", "cosine"])
def test_lr_schedulers_drive_the_fused_step(kind):
    """init_lr_scheduler (experiment_init.py:266-283) on FlatAdam yields the same learning-rate sequence as on
    torch.optim.Adam, and every change reaches the device scalar the captured Adam kernel reads."""
    ts = _StubStep(_small_model(), 1e-3)
    fa = T.FlatAdam(ts)
    ref = torch.optim.Adam(_small_model().parameters(), lr=1e-3, eps=1e-4)
    sa = T.init_lr_scheduler(fa, kind, t_max=5, factor=0.5, min_lr=1e-6, patience=2)
    sb = T.init_lr_scheduler(ref, kind, t_max=5, factor=0.5, min_lr=1e-6, patience=2)
    val = [1.0, 0.9, 0.95, 0.96, 0.97, 0.98, 0.99, 1.0, 1.1, 1.2, 1.3, 0.5, 0.6, 0.7, 0.8, 0.9]
    for v in val:
        fa.step(); ref.step()
        if kind == "cosine":
            sa.step(); sb.step()
        else:
            sa.step(v); sb.step(v)
        fa.step()   # next epoch's first step pushes the new rate
        assert fa.param_groups[0]["lr"] == pytest.approx(ref.param_groups[0]["lr"], rel=1e-12)
        assert ts.lr_dev.item() == pytest.approx(ref.param_groups[0]["lr"], rel=1e-6)
    assert len(ts.pushed) >= 2
    with pytest.raises(SystemExit):
        T.init_lr_scheduler(fa, "step")


def test_flat_adam_state_dict_has_torch_adam_format():
    ts = _StubStep(_small_model(), 1e-4)
    fa = T.FlatAdam(ts)
    assert fa.state_dict()["state"] == {}                      # like torch before the first step
    ts.exp_avg.uniform_(); ts.exp_avg_sq.uniform_(); ts.step_dev.fill_(4)
    sd = fa.state_dict()
    ref = torch.optim.Adam(_small_model().parameters(), lr=1e-4, eps=1e-4)
    assert set(sd["param_groups"][0]) == set(ref.state_dict()["param_groups"][0])
    ref.load_state_dict(sd)                                    # torch accepts it
    assert float(ref.state_dict()["state"][0]["step"]) == 4.0
    ts2 = _StubStep(_small_model(), 5e-5)
    fb = T.FlatAdam(ts2)
    fb.load_state_dict(sd)
    for a, b in ts.param_ranges.values():                      # (the padding between slots is not state)
        assert torch.equal(ts2.exp_avg[a:b], ts.exp_avg[a:b]) and torch.equal(ts2.exp_avg_sq[a:b], ts.exp_avg_sq[a:b])
    assert int(ts2.step_dev.item()) == 4 and ts2.lr_dev.item() == pytest.approx(1e-4)


def test_checkpoint_has_the_reference_keys(tmp_path):
    m = _small_model()
    opt = T.init_optimizer(m, "Adam", 1e-4)
    assert isinstance(opt, torch.optim.Adam) and opt.defaults["eps"] == 1e-4
    assert isinstance(T.init_optimizer(m, "SGD", 1e-2), torch.optim.SGD)
    assert isinstance(T.init_optimizer(m, "AdamW", 1e-2), torch.optim.AdamW)
    p = str(tmp_path / "model_x_fold_0")
    T.save_checkpoint(p, 7, m, opt, 0.25)
    ck = torch.load(p, weights_only=False)
    assert list(ck) == ["epoch", "model_state_dict", "optimizer_state_dict", "scheduler", "val_loss"]
    assert ck["scheduler"] == "scheduler" and ck["epoch"] == 7
    m2 = _small_model()
    with torch.no_grad():
        for q in m2.parameters():
            q.add_(1.0)
    T.load_pretrained_model(m2, p)
    assert all(torch.equal(a, b) for a, b in zip(m.parameters(), m2.parameters()))
    with pytest.raises(ValueError):
        T.load_pretrained_model(m2, str(tmp_path / "missing"))
