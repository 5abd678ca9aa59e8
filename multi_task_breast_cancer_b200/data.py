"""GPU input pipeline.

The reference keeps the BUSI images as uint8 numpy arrays on the host, augments ONE sample at a time in
`BUSI.__getitem__` (src/dataset/BUSI_dataset.py:97-163: cat([mask, image]) -> RandomHorizontalFlip(0.5) ->
RandomVerticalFlip(0.5) -> RandomRotation(degrees=360), built at src/training_multitask.py:193-197), collates fp32
tensors and copies them to the device from pageable memory (training_multitask.py:82).  At the step rates of the CUDA
path (thousands of images/s) that host pipeline is the bottleneck, so here

  * the dataset lives on the device as uint8 (`DeviceBUSI`: images, masks [n][H][W], labels [n]),
  * one launch (`mtbc_augment_batch`, csrc/augment.cu) gathers a batch, applies the per-sample draw and writes the
    fp32 image / mask / one-hot label tensors the training step consumes,
  * the per-sample draws follow the reference's RNG call order (`draw_transform_params`), so a given torch seed yields
    the parameters the reference's transforms would have drawn,
  * `deterministic_oversampling_indices` / `shard_indices` restate the index bookkeeping of
    src/dataset/BUSI_dataloader.py:320-340 and split an epoch across data-parallel ranks.

CUDA only; there is no CPU fallback (the oracle's torchvision restatement lives in oracle/torch_oracle.py).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Iterator, List, Optional, Sequence, Tuple

import torch

from . import _lib
from .ops import ptr, stream_ptr

CLASS_IDS = {"benign": 0, "malignant": 1, "normal": 2}   # BUSI_dataset.py:63-71 (non-semantic branch)


# ======================================================================================================================
# host-side index / parameter bookkeeping (CPU-testable)
# ======================================================================================================================
def deterministic_oversampling_indices(classes: Sequence) -> List[int]:
    """Row indices after `deterministic_oversampling` (src/dataset/BUSI_dataloader.py:320-340), in the row order the
    reference's concatenated DataFrame has: all rows, then per class (most frequent first, `value_counts` order) the
    class's rows repeated `factor - 1` times, factor = round(1 / class frequency) (half-to-even, pandas `round`); a
    class whose factor is 1 is appended once more (the reference's else-branch)."""
    classes = list(classes)
    n = len(classes)
    first_seen, counts = {}, {}
    for i, c in enumerate(classes):
        first_seen.setdefault(c, i)
        counts[c] = counts.get(c, 0) + 1
    order = sorted(counts, key=lambda c: (-counts[c], first_seen[c]))
    out = list(range(n))
    for c in order:
        factor = int(round(1.0 / (counts[c] / n)))          # Python round == numpy rint: half to even
        rows = [i for i, k in enumerate(classes) if k == c]
        out += rows * (factor - 1 if factor > 1 else 1)
    return out


def shard_indices(indices: Sequence[int], rank: int, world: int, batch: int, drop_last: bool = True) -> List[List[int]]:
    """Per-rank batches of an epoch: global batch g takes indices[g*world*batch : (g+1)*world*batch] and rank r the
    r-th slice of `batch` of it (rank r gets samples [r*B, (r+1)*B) of the global batch).

    Every rank gets the SAME number of batches and every batch has exactly `batch` samples -- the data-parallel step
    issues one bucketed all-reduce sequence per step and averages per-rank means with 1/world, so unequal step counts
    would deadlock NCCL and unequal batch sizes would bias the mean.  With drop_last=False the tail (fewer than
    world*batch samples) is therefore completed to one more full global batch by wrapping around to the head of the
    epoch's index list (the convention of torch's DistributedSampler(drop_last=False)).  On ONE rank nothing has to line
    up, so the tail batch stays ragged: the reference trains on it (BUSI_dataloader.py:146-148 has no drop_last)."""
    indices = list(indices)
    per_global = world * batch
    n_full = len(indices) // per_global
    out = []
    for g in range(n_full):
        a = g * per_global + rank * batch
        out.append(indices[a:a + batch])
    rest = len(indices) - n_full * per_global
    if not drop_last and rest:
        tail = indices[n_full * per_global:]
        if world == 1:
            out.append(tail)                     # ragged last batch, exactly the reference's
        else:
            need = per_global - rest
            pad = (indices * (need // max(1, len(indices)) + 1))[:need]
            full = tail + pad
            out.append(full[rank * batch:(rank + 1) * batch])
    return out


def draw_transform_params(n: int, p_hflip: float = 0.5, p_vflip: float = 0.5, degrees: float = 360.0,
                          generator: Optional[torch.Generator] = None) -> Tuple[List[bool], List[bool], List[float]]:
    """Per-sample draws in the reference's order: for every sample RandomHorizontalFlip draws `torch.rand(1)`,
    RandomVerticalFlip draws `torch.rand(1)`, RandomRotation draws `torch.empty(1).uniform_(-degrees, degrees)`
    (torchvision transforms.py: forward / get_params).  With the same torch seed and a single-process loader these are
    the values the reference's `transforms(joined)` consumes for samples 0..n-1."""
    hf, vf, ang = [], [], []
    for _ in range(n):
        hf.append(bool(torch.rand(1, generator=generator) < p_hflip))
        vf.append(bool(torch.rand(1, generator=generator) < p_vflip))
        ang.append(float(torch.empty(1).uniform_(-float(degrees), float(degrees), generator=generator).item()))
    return hf, vf, ang


def rotation_theta(angle_deg: float, H: int, W: int) -> List[float]:
    """theta / (W/2, H/2) of torchvision.transforms.functional.rotate(img, angle) for tensors: the inverse affine matrix
    `_get_inverse_affine_matrix([0, 0], -angle, [0, 0], 1.0, [0.0, 0.0])` (python doubles), cast to fp32 and divided by
    the half extents in fp32 (`_gen_affine_grid`).  Row major: gx = x*t[0] + y*t[1] + t[2], gy = x*t[3] + y*t[4] + t[5]."""
    rot = math.radians(-angle_deg)
    a, b, c, d = math.cos(rot), -math.sin(rot), math.sin(rot), math.cos(rot)   # shear 0: RSS = rotation
    m = [d, -b, 0.0, -c, a, 0.0]
    m[2] += m[0] * 0.0 + m[1] * 0.0
    m[5] += m[3] * 0.0 + m[4] * 0.0
    theta = torch.tensor(m, dtype=torch.float32).reshape(2, 3)
    rescaled = theta.transpose(0, 1) / torch.tensor([0.5 * W, 0.5 * H], dtype=torch.float32)   # (3, 2) as in torchvision
    return [float(rescaled[0, 0]), float(rescaled[1, 0]), float(rescaled[2, 0]),
            float(rescaled[0, 1]), float(rescaled[1, 1]), float(rescaled[2, 1])]


# ======================================================================================================================
# device-resident dataset
# ======================================================================================================================
class DeviceBUSI:
    """uint8 images / masks and int32 labels resident in HBM; `batch()` is one launch.

    `images`, `masks`: uint8 (n, H, W) (masks in {0, 1}, as after BUSI_dataset.py:54-55); `labels`: (n,) integer class ids
    (benign 0, malignant 1, normal 2).  180 GB of HBM hold ~1.4 M 256x256 image/mask pairs; BUSI has 780."""

    def __init__(self, images: torch.Tensor, masks: torch.Tensor, labels: torch.Tensor, device="cuda", n_classes: int = 3):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.MtbcError("DeviceBUSI keeps the dataset on a CUDA device (no CPU fallback)")
        if images.dtype != torch.uint8 or masks.dtype != torch.uint8 or images.shape != masks.shape or images.dim() != 3:
            raise ValueError("images and masks must be uint8 tensors of the same (n, H, W) shape")
        if images.shape[2] % 4 != 0:
            raise ValueError("image width must be a multiple of 4")
        self.device = dev
        self.images = images.to(dev).contiguous()
        self.masks = masks.to(dev).contiguous()
        self.labels = labels.flatten().to(dev, torch.int32).contiguous()
        self.n, self.H, self.W = (int(v) for v in images.shape)
        self.K = int(n_classes)

    def __len__(self):
        return self.n

    def batch(self, indices: Sequence[int], hflip: Optional[Sequence[bool]] = None, vflip: Optional[Sequence[bool]] = None,
              angle: Optional[Sequence[float]] = None, out: Optional[Tuple[torch.Tensor, ...]] = None):
        """-> (image fp32 (B,1,H,W), mask fp32 (B,1,H,W), one-hot fp32 (B,K)).  No draws (all None) = the validation /
        test loaders of the reference (no transforms): an exact gather + cast."""
        B = len(indices)
        if B == 0:
            raise ValueError("empty batch")
        if min(indices) < 0 or max(indices) >= self.n:
            raise IndexError("sample index out of range")
        flips = [0] * B
        theta = [[0.0] * 6 for _ in range(B)]
        for b in range(B):
            f = (1 if hflip is not None and hflip[b] else 0) | (2 if vflip is not None and vflip[b] else 0)
            if angle is not None:
                f |= 4
                theta[b] = rotation_theta(float(angle[b]), self.H, self.W)
            flips[b] = f
        with torch.cuda.device(self.device):
            idx_d = torch.tensor(list(indices), dtype=torch.int32).to(self.device, non_blocking=True)
            fl_d = torch.tensor(flips, dtype=torch.uint8).to(self.device, non_blocking=True)
            th_d = torch.tensor(theta, dtype=torch.float32).to(self.device, non_blocking=True)
            if out is None:
                img = torch.empty(B, 1, self.H, self.W, dtype=torch.float32, device=self.device)
                mask = torch.empty(B, 1, self.H, self.W, dtype=torch.float32, device=self.device)
                onehot = torch.empty(B, self.K, dtype=torch.float32, device=self.device)
            else:
                img, mask, onehot = out
            _lib.call("mtbc_augment_batch", ptr(self.images), ptr(self.masks), ptr(self.labels), ptr(idx_d), ptr(fl_d),
                      ptr(th_d), B, self.H, self.W, self.K, ptr(img), ptr(mask), ptr(onehot), C.c_void_p(stream_ptr()))
        return img, mask, onehot

    def epoch(self, batches: Sequence[Sequence[int]], augment: bool = True,
              generator: Optional[torch.Generator] = None) -> Iterator[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
        """Iterate an epoch's batches (e.g. from `shard_indices`), drawing fresh transform parameters per sample when
        `augment` (training loader) and none otherwise (validation / test loaders)."""
        for ids in batches:
            if augment:
                hf, vf, ang = draw_transform_params(len(ids), generator=generator)
                yield self.batch(ids, hf, vf, ang)
            else:
                yield self.batch(ids)
