"""Loss criteria with the reference's call conventions, computed by the fused CUDA kernels of csrc/loss.cu.

    DiceLoss   == monai.losses.DiceLoss(include_background=True, sigmoid=True, smooth_nr=1, smooth_dr=1,
                                        squared_pred=True)      (built at src/utils/experiment_init.py:209-211)
    FocalLoss  == src/utils/criterions.py:6-24   (alpha, gamma=2, soft one-hot targets, mean reduction)
    apply_criterion_multitask_segmentation_classification == src/utils/criterions.py:52-76
    refine_predictions == prediction-refining module, src/utils/models.py:316-332 and :366-386 (batched)

Both criteria return differentiable fp32 scalars (0-dim CUDA tensors).  CUDA only -- no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import logging
import sys
from typing import List, Sequence, Union

import torch

from . import _lib
from .ops import ptr, stream_ptr


def _st():
    return C.c_void_p(stream_ptr())


def _require_cuda(*ts):
    for t in ts:
        if not t.is_cuda:
            raise _lib.MtbcError("criteria run on CUDA sm_100a only (no CPU fallback)")


def _on_device_of(fn):
    """Run `fn` with the first tensor argument's device current: allocations and `torch.cuda.current_stream()` (the
    launch stream, `_st()`) then belong to the device that holds the data, whichever device is current outside."""
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kw):
        dev = next((a.device for a in args if torch.is_tensor(a) and a.is_cuda), None)
        if dev is None:
            for a in args:   # lists of heads (deep supervision)
                if isinstance(a, (list, tuple)) and a and torch.is_tensor(a[-1]) and a[-1].is_cuda:
                    dev = a[-1].device
                    break
        if dev is None:
            return fn(*args, **kw)
        with torch.cuda.device(dev):
            return fn(*args, **kw)
    return wrapped


class _DiceFn(torch.autograd.Function):
    @staticmethod
    @_on_device_of
    def forward(ctx, logits, target):
        _require_cuda(logits, target)
        logits = logits.contiguous().float()
        target = target.contiguous().float()
        if logits.shape != target.shape:
            raise AssertionError(f"ground truth has different shape ({target.shape}) from input ({logits.shape})")
        B, Cc = logits.shape[0], logits.shape[1]
        HW = logits[0, 0].numel()
        N = B * Cc  # per-(b, c) dice terms, mean over all of them
        sums = torch.zeros(N, 3, dtype=torch.float32, device=logits.device)
        loss = torch.empty((), dtype=torch.float32, device=logits.device)
        _lib.call("mtbc_dice_sums", ptr(logits), ptr(target), N, HW, ptr(sums), _st())
        _lib.call("mtbc_dice_finalize", ptr(sums), N, ptr(loss), _st())
        ctx.save_for_backward(logits, target, sums)
        return loss

    @staticmethod
    @_on_device_of
    def backward(ctx, g):
        logits, target, sums = ctx.saved_tensors
        B, Cc = logits.shape[0], logits.shape[1]
        HW = logits[0, 0].numel()
        d = torch.empty_like(logits)
        g = g.contiguous().float()
        _lib.call("mtbc_dice_bwd", ptr(logits), ptr(target), B * Cc, HW, ptr(sums), ptr(g), C.c_float(1.0), ptr(d), _st())
        return d, None


class DiceLoss(torch.nn.Module):
    """Drop-in for the MONAI DiceLoss configuration the reference uses; other configurations are not implemented."""

    def __init__(self, include_background=True, to_onehot_y=False, sigmoid=False, softmax=False, other_act=None,
                 squared_pred=False, jaccard=False, reduction="mean", smooth_nr=1e-5, smooth_dr=1e-5, batch=False,
                 weight=None):
        super().__init__()
        ok = (include_background and not to_onehot_y and sigmoid and not softmax and other_act is None and squared_pred
              and not jaccard and reduction == "mean" and float(smooth_nr) == 1.0 and float(smooth_dr) == 1.0
              and not batch and weight is None)
        if not ok:
            raise NotImplementedError("only DiceLoss(include_background=True, sigmoid=True, smooth_nr=1, smooth_dr=1, "
                                      "squared_pred=True) -- the reference configuration -- is implemented")

    def forward(self, input, target):
        return _DiceFn.apply(input, target)


class _FocalFn(torch.autograd.Function):
    @staticmethod
    @_on_device_of
    def forward(ctx, logits, target, alpha, gamma):
        _require_cuda(logits, target)
        logits = logits.contiguous().float()
        target = target.contiguous().float()
        N, K = logits.shape
        loss = torch.empty((), dtype=torch.float32, device=logits.device)
        _lib.call("mtbc_focal_fwd", ptr(logits), ptr(target), N, K, C.c_float(alpha), C.c_float(gamma), ptr(loss), _st())
        ctx.save_for_backward(logits, target)
        ctx.alpha, ctx.gamma = alpha, gamma
        return loss

    @staticmethod
    @_on_device_of
    def backward(ctx, g):
        logits, target = ctx.saved_tensors
        N, K = logits.shape
        d = torch.empty_like(logits)
        g = g.contiguous().float()
        _lib.call("mtbc_focal_bwd", ptr(logits), ptr(target), N, K, C.c_float(ctx.alpha), C.c_float(ctx.gamma), ptr(g),
                  C.c_float(1.0), ptr(d), _st())
        return d, None, None, None


class FocalLoss(torch.nn.Module):
    """src/utils/criterions.py:6-24 (reduction='mean', no class weights: config classes_weighted is null)."""

    def __init__(self, alpha=1, gamma=2, reduction="mean", weight=None):
        super().__init__()
        if reduction != "mean" or weight is not None:
            raise NotImplementedError("only reduction='mean' without class weights (the shipped config) is implemented")
        self.alpha, self.gamma, self.reduction, self.weight = alpha, gamma, reduction, weight

    def forward(self, inputs, targets):
        return _FocalFn.apply(inputs, targets, float(self.alpha), float(self.gamma))


def apply_criterion_multitask_segmentation_classification(criterion_seg, ground_truth, segmentation, criterion_class,
                                                          label, predicted_class, inversely_weighted=False):
    """Host glue identical to src/utils/criterions.py:52-76 (including the NaN guard that exits the process)."""
    if isinstance(segmentation, list):
        if inversely_weighted:
            segmentation_loss = torch.sum(torch.stack(
                [criterion_seg(s, ground_truth) / (n + 1) for n, s in enumerate(reversed(segmentation))]))
        else:
            segmentation_loss = torch.sum(torch.stack([criterion_seg(s, ground_truth) for s in reversed(segmentation)]))
        classification_loss = torch.sum(torch.stack([criterion_class(c, label) for c in reversed(predicted_class)]))
    else:
        segmentation_loss = criterion_seg(segmentation, ground_truth)
        classification_loss = criterion_class(predicted_class, label)
    if not torch.isnan(segmentation_loss) and not torch.isnan(classification_loss):
        return segmentation_loss, classification_loss
    logging.info("NaN in model loss!!")
    sys.exit(1)


def init_criterion_segmentation(loss_function: str = "dice"):
    """src/utils/experiment_init.py:199-232 restricted to the multi-task path's shipped choice ('DICE')."""
    if loss_function == "DICE":
        return DiceLoss(include_background=True, sigmoid=True, smooth_dr=1, smooth_nr=1, squared_pred=True)
    raise NotImplementedError(f"segmentation criterion {loss_function!r} is outside the accelerated hot path")


def init_criterion_classification(n_classes: int = 2, classes_weighted=None, classification_criterion="CE"):
    """src/utils/experiment_init.py:235-263 restricted to the shipped choice (Focal, no class weights, 3 classes)."""
    if n_classes != 2 and not classes_weighted and classification_criterion == "Focal":
        return FocalLoss(alpha=1, gamma=2, reduction="mean")
    raise NotImplementedError("only the 3-class FocalLoss(alpha=1, gamma=2) criterion is on the accelerated hot path")


@_on_device_of
def refine_predictions(mask_logits: torch.Tensor, class_logits: torch.Tensor, normal_id: int = 2,
                       overlap_seg_based_on_class: bool = True, overlap_class_based_on_seg: bool = True,
                       threshold: int = 0):
    """Batched prediction-refining module: returns (uint8 mask (B,1,H,W), int32 class (B,), int32 pixel count (B,))."""
    # the models return lists under deep supervision (utils/models.py:313,360-361): last head, mean of the class logits
    if isinstance(mask_logits, (list, tuple)):
        mask_logits = mask_logits[-1]
    if isinstance(class_logits, (list, tuple)):
        class_logits = torch.mean(torch.stack(list(class_logits), dim=0), dim=0)
    _require_cuda(mask_logits, class_logits)
    mask_logits = mask_logits.contiguous().float()
    class_logits = class_logits.contiguous().float()
    B = mask_logits.shape[0]
    HW = mask_logits[0].numel()
    K = class_logits.shape[1]
    mask = torch.empty(mask_logits.shape, dtype=torch.uint8, device=mask_logits.device)
    cls = torch.empty(B, dtype=torch.int32, device=mask_logits.device)
    cnt = torch.empty(B, dtype=torch.int32, device=mask_logits.device)
    _lib.call("mtbc_refine_predictions", ptr(mask_logits), ptr(class_logits), B, HW, K, normal_id,
              int(overlap_seg_based_on_class), int(overlap_class_based_on_seg), int(threshold), ptr(mask), ptr(cls),
              ptr(cnt), _st())
    return mask, cls, cnt


@_on_device_of
def hard_dice_counts(mask_logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """tp/fp/fn of (logit > 0) vs target over the whole batch (training_multitask.py:65-71 without the host sync)."""
    _require_cuda(mask_logits, target)
    out = torch.zeros(3, dtype=torch.int64, device=mask_logits.device)
    _lib.call("mtbc_hard_dice_counts", ptr(mask_logits.contiguous().float()), ptr(target.contiguous().float()),
              mask_logits.numel(), ptr(out), _st())
    return out
