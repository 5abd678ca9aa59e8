"""monai.losses.DiceLoss (1.3.0) restated for the configuration the reference uses (experiment_init.py:209-211)."""
import torch
import torch.nn as nn


class DiceLoss(nn.Module):
    def __init__(self, include_background=True, to_onehot_y=False, sigmoid=False, softmax=False, other_act=None,
                 squared_pred=False, jaccard=False, reduction="mean", smooth_nr=1e-5, smooth_dr=1e-5, batch=False,
                 weight=None):
        super().__init__()
        assert include_background and not to_onehot_y and not softmax and other_act is None and not jaccard
        assert not batch and weight is None
        self.sigmoid, self.squared_pred, self.reduction = sigmoid, squared_pred, reduction
        self.smooth_nr, self.smooth_dr = float(smooth_nr), float(smooth_dr)

    def forward(self, input, target):
        if self.sigmoid:
            input = torch.sigmoid(input)
        if target.shape != input.shape:
            raise AssertionError(f"ground truth has different shape ({target.shape}) from input ({input.shape})")
        reduce_axis = list(range(2, input.dim()))
        intersection = torch.sum(target * input, dim=reduce_axis)
        if self.squared_pred:
            ground_o = torch.sum(target ** 2, dim=reduce_axis)
            pred_o = torch.sum(input ** 2, dim=reduce_axis)
        else:
            ground_o = torch.sum(target, dim=reduce_axis)
            pred_o = torch.sum(input, dim=reduce_axis)
        denominator = ground_o + pred_o
        f = 1.0 - (2.0 * intersection + self.smooth_nr) / (denominator + self.smooth_dr)
        if self.reduction == "mean":
            return torch.mean(f)
        if self.reduction == "sum":
            return torch.sum(f)
        return f
