"""CPU tests of the input-pipeline bookkeeping: oversampling indices, per-rank sharding (also under
gloo with 2 ranks in test_host_logic), and the RNG call order of the per-sample transform draws."""
import random

import pytest
import torch

from oracle import torch_oracle as O
from multi_task_breast_cancer_b200 import data as D


def test_deterministic_oversampling_matches_restatement_and_busi_counts():
    # the fold composition SURVEY 8d quotes: 133 benign / 98 malignant / 38 normal -> x2 / x3 / x7 = 826 rows
    classes = ["benign"] * 133 + ["malignant"] * 98 + ["normal"] * 38
    random.Random(3).shuffle(classes)
    idx = D.deterministic_oversampling_indices(classes)
    assert idx == O.deterministic_oversampling(classes)
    assert len(idx) == 133 * 2 + 98 * 3 + 38 * 7 == 826
    assert idx[:len(classes)] == list(range(len(classes)))
    per = {c: sum(classes[i] == c for i in idx) for c in set(classes)}
    assert per == {"benign": 266, "malignant": 294, "normal": 266}
    for trial in range(20):
        rng = random.Random(trial)
        cl = [rng.choice(["a", "b", "c", "d"][:rng.randint(1, 4)]) for _ in range(rng.randint(1, 60))]
        assert D.deterministic_oversampling_indices(cl) == O.deterministic_oversampling(cl)
    assert D.deterministic_oversampling_indices(["x"] * 5) == list(range(5)) * 2      # factor 1: the else-branch copy


def test_shard_indices_partition_the_epoch():
    idx = list(range(103))
    for world, batch in [(1, 8), (2, 8), (4, 5), (8, 3)]:
        per_rank = [D.shard_indices(idx, r, world, batch) for r in range(world)]
        n_global = 103 // (world * batch)
        assert all(len(p) == n_global for p in per_rank)
        for g in range(n_global):
            got = sum((per_rank[r][g] for r in range(world)), [])
            assert got == idx[g * world * batch:(g + 1) * world * batch]        # rank r = r-th slice of the global batch
        tail = [D.shard_indices(idx, r, world, batch, drop_last=False) for r in range(world)]
        # data-parallel safety: same number of steps on every rank (else the per-bucket all-reduces
        # deadlock) and -- for world > 1 -- only full batches (1/world averaging of per-rank means stays unbiased)
        assert len({len(t) for t in tail}) == 1
        seen = sum((sum(t, []) for t in tail), [])
        assert set(seen) == set(idx)                                             # nothing lost
        if world == 1:
            assert sorted(seen) == idx and len(tail[0][-1]) == 103 % batch       # ragged tail, as the reference
        else:
            assert all(len(b) == batch for t in tail for b in t)
            extra = len(seen) - len(idx)                                         # wrapped-around padding only
            assert 0 <= extra < world * batch
            last_global = sum((t[-1] for t in tail), [])
            assert last_global[:103 - n_global * world * batch] == idx[n_global * world * batch:]


def test_transform_draws_follow_torchvision_call_order():
    torch.manual_seed(1993)
    hf, vf, ang = D.draw_transform_params(50)
    torch.manual_seed(1993)
    rh, rv, ra = O.draw_reference_transform_params(50)
    assert hf == rh and vf == rv and ang == ra
    assert any(hf) and not all(hf) and any(vf) and not all(vf) and min(ang) < -100 and max(ang) > 100


def test_rotation_theta_matches_torchvision_matrix():
    from torchvision.transforms.functional import _get_inverse_affine_matrix
    for angle, H, W in [(0.0, 64, 64), (37.5, 128, 256), (-123.4, 256, 256), (90.0, 64, 32), (359.9, 96, 100)]:
        m = _get_inverse_affine_matrix([0.0, 0.0], -angle, [0.0, 0.0], 1.0, [0.0, 0.0])
        theta = torch.tensor(m, dtype=torch.float32).reshape(1, 2, 3)
        rescaled = theta.transpose(1, 2) / torch.tensor([0.5 * W, 0.5 * H])
        want = [rescaled[0, 0, 0], rescaled[0, 1, 0], rescaled[0, 2, 0], rescaled[0, 0, 1], rescaled[0, 1, 1], rescaled[0, 2, 1]]
        got = D.rotation_theta(angle, H, W)
        assert all(float(a) == b for a, b in zip(want, got)), (angle, want, got)


def test_device_dataset_rejects_cpu_and_bad_shapes():
    from multi_task_breast_cancer_b200._lib import MtbcError
    img = torch.zeros(2, 8, 8, dtype=torch.uint8)
    with pytest.raises(MtbcError):
        D.DeviceBUSI(img, img, torch.zeros(2), device="cpu")
