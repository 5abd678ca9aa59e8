"""GPU tests of the device-resident input pipeline against torchvision's own functional ops run on
the host the way the reference's Dataset does (oracle.augment_sample).

Gather, cast, flips and labels are byte work: bit-exact.  The rotation's source pixel is nearbyint() of an fp32
coordinate that torchvision computes with a bmm whose summation order is not specified; a pixel may therefore differ
only where that coordinate is within 1e-3 px of a rounding tie, which the test verifies pixel by pixel in float64."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _cuda(lib):
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")


def _dataset(n, H, W, seed=0):
    g = torch.Generator().manual_seed(seed)
    images = torch.randint(0, 256, (n, H, W), generator=g, dtype=torch.uint8)
    masks = (torch.rand(n, H, W, generator=g) < 0.3).to(torch.uint8)
    labels = torch.randint(0, 3, (n,), generator=g)
    return images, masks, labels


def test_plain_gather_and_flips_are_bit_exact():
    from oracle import torch_oracle as O
    from multi_task_breast_cancer_b200.data import DeviceBUSI
    images, masks, labels = _dataset(7, 32, 48)
    ds = DeviceBUSI(images, masks, labels)
    ids = [6, 0, 3, 3, 5]
    img, mask, onehot = ds.batch(ids)                                 # validation / test loader: no transforms
    for b, i in enumerate(ids):
        ri, rm = O.augment_sample(images[i], masks[i], False, False, None)
        assert torch.equal(img[b].cpu(), ri) and torch.equal(mask[b].cpu(), rm)
    assert torch.equal(onehot.cpu(), torch.nn.functional.one_hot(labels[ids], 3).float())
    hf, vf = [True, False, True, False, True], [True, True, False, False, True]
    img, mask, _ = ds.batch(ids, hf, vf)
    for b, i in enumerate(ids):
        ri, rm = O.augment_sample(images[i], masks[i], hf[b], vf[b], None)
        assert torch.equal(img[b].cpu(), ri) and torch.equal(mask[b].cpu(), rm)
    with pytest.raises(IndexError):
        ds.batch([7])


def _near_tie(angle, H, W, eps=1e-3):
    """float64 source coordinates of torchvision's rotate; True where either lies within eps of a rounding tie."""
    rot = math.radians(-angle)
    c, s = math.cos(rot), math.sin(rot)
    x = torch.arange(W, dtype=torch.float64) + 0.5 - W / 2
    y = (torch.arange(H, dtype=torch.float64) + 0.5 - H / 2).unsqueeze(1)
    gx = (c * x + s * y) / (0.5 * W)
    gy = (-s * x + c * y) / (0.5 * H)
    fx = ((gx + 1) * W - 1) / 2
    fy = ((gy + 1) * H - 1) / 2
    tie = lambda f: ((f - torch.floor(f)) - 0.5).abs() < eps
    return tie(fx) | tie(fy)


@pytest.mark.parametrize("H,W", [(64, 64), (128, 256)])
def test_flip_flip_rotate_matches_torchvision(H, W):
    from oracle import torch_oracle as O
    from multi_task_breast_cancer_b200.data import DeviceBUSI
    images, masks, labels = _dataset(9, H, W, seed=1)
    ds = DeviceBUSI(images, masks, labels)
    angles = [0.0, 90.0, 180.0, -45.0, 17.3, -201.75, 359.0, 123.456, -7.0]
    hf = [False, True, False, True, False, True, False, True, False]
    vf = [False, False, True, True, False, False, True, True, False]
    ids = list(range(9))
    img, mask, _ = ds.batch(ids, hf, vf, angles)
    total_bad = 0
    for b in ids:
        ri, rm = O.augment_sample(images[b], masks[b], hf[b], vf[b], angles[b])
        bad = (img[b].cpu() != ri)[0]
        bad_m = (mask[b].cpu() != rm)[0]
        tie = _near_tie(angles[b], H, W)
        assert not (bad & ~tie).any(), (angles[b], int((bad & ~tie).sum()))
        assert not (bad_m & ~tie).any()
        total_bad += int(bad.sum())
        # image and mask always move together (same draw for both channels, BUSI_dataset.py:152)
        moved = (img[b].cpu() != ri)[0] | bad_m
        assert not (moved & ~tie).any()
    assert total_bad <= 2e-3 * 9 * H * W, total_bad
    # rotation really happened and zero fill is in place: a 45-degree turn blanks the corners
    assert img[3, 0, 0, 0].item() == 0.0
    assert (img[3] == 0).float().mean().item() > (img[0] == 0).float().mean().item()


def test_epoch_of_augmented_batches_feeds_the_training_step():
    from multi_task_breast_cancer_b200 import models as M
    from multi_task_breast_cancer_b200.data import DeviceBUSI, deterministic_oversampling_indices, shard_indices
    from multi_task_breast_cancer_b200.train import TrainStep
    from multi_task_breast_cancer_b200.trainer import EpochRunner
    images, masks, labels = _dataset(10, 64, 64, seed=2)
    masks[labels == 2] = 0                                             # "normal" images have empty masks
    ds = DeviceBUSI(images, masks, labels)
    names = [["benign", "malignant", "normal"][int(l)] for l in labels]
    idx = deterministic_oversampling_indices(names)
    batches = shard_indices(idx, rank=0, world=1, batch=4)
    torch.manual_seed(1993)
    model = M.MTnnUNet(1, 1, 3).cuda()
    ts = TrainStep(model, (4, 1, 64, 64))
    runner = EpochRunner(ts, 3)
    g = torch.Generator().manual_seed(5)
    loss, dice, acc, f1w = runner.train_one_epoch(ds.epoch(batches, augment=True, generator=g))
    assert math.isfinite(loss) and 0.0 <= dice <= 1.0 and 0.0 <= acc <= 1.0 and 0.0 <= f1w <= 1.0
    assert int(ts.step_dev.item()) == len(batches) >= 3
    vloss, *_ = runner.validate_one_epoch(ds.epoch(batches[:2], augment=False))
    assert math.isfinite(vloss)
