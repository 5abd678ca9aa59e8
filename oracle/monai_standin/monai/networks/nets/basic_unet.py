"""monai.networks.nets.basic_unet.{TwoConv, Down, UpCat} (MONAI 1.3.0) restated for spatial_dims=2, upsample="deconv".

Module tree (hence state_dict keys), op order conv -> InstanceNorm -> Dropout -> LeakyReLU ("NDA"), ConvTranspose2d
k2 s2 with the `halves` channel rule, skip-first concatenation and the replicate-pad rule follow SURVEY.md Appendix B.
"""
import torch
import torch.nn as nn


def _make_act(act):
    name, kw = act if isinstance(act, (tuple, list)) else (act, {})
    name = name.lower()
    if name == "leakyrelu":
        return nn.LeakyReLU(**kw)
    if name == "relu":
        return nn.ReLU(**kw)
    raise NotImplementedError(name)


def _make_norm(norm, ch):
    name, kw = norm if isinstance(norm, (tuple, list)) else (norm, {})
    name = name.lower()
    if name == "instance":
        return nn.InstanceNorm2d(ch, **kw)
    if name == "batch":
        return nn.BatchNorm2d(ch, **kw)
    raise NotImplementedError(name)


class ADN(nn.Sequential):
    """Norm -> Dropout -> Act, children named N / D / A (monai.networks.blocks.ADN, ordering "NDA")."""

    def __init__(self, ch, act, norm, dropout):
        super().__init__()
        self.add_module("N", _make_norm(norm, ch))
        if dropout is not None:
            p = dropout[0] if isinstance(dropout, (tuple, list)) else dropout
            self.add_module("D", nn.Dropout(float(p)))
        self.add_module("A", _make_act(act))


class Convolution(nn.Sequential):
    """monai.networks.blocks.Convolution(2, in, out, act, norm, dropout, bias, padding=1): children conv, adn."""

    def __init__(self, spatial_dims, in_chns, out_chns, act, norm, dropout, bias, padding=1):
        super().__init__()
        assert spatial_dims == 2
        self.add_module("conv", nn.Conv2d(in_chns, out_chns, kernel_size=3, stride=1, padding=padding, bias=bias))
        self.add_module("adn", ADN(out_chns, act, norm, dropout))


class TwoConv(nn.Sequential):
    def __init__(self, spatial_dims, in_chns, out_chns, act, norm, bias, dropout=0.0):
        super().__init__()
        self.add_module("conv_0", Convolution(spatial_dims, in_chns, out_chns, act, norm, dropout, bias, padding=1))
        self.add_module("conv_1", Convolution(spatial_dims, out_chns, out_chns, act, norm, dropout, bias, padding=1))


class Down(nn.Sequential):
    def __init__(self, spatial_dims, in_chns, out_chns, act, norm, bias, dropout=0.0):
        super().__init__()
        self.add_module("max_pooling", nn.MaxPool2d(kernel_size=2))
        self.add_module("convs", TwoConv(spatial_dims, in_chns, out_chns, act, norm, bias, dropout))


class _UpSampleDeconv(nn.Sequential):
    """monai.networks.blocks.UpSample(mode="deconv"): single child `deconv`."""

    def __init__(self, in_chns, out_chns):
        super().__init__()
        self.add_module("deconv", nn.ConvTranspose2d(in_chns, out_chns, kernel_size=2, stride=2, bias=True))


class UpCat(nn.Module):
    def __init__(self, spatial_dims, in_chns, cat_chns, out_chns, act, norm, bias, dropout=0.0, upsample="deconv",
                 pre_conv="default", interp_mode="linear", align_corners=True, halves=True, is_pad=True):
        super().__init__()
        assert spatial_dims == 2 and upsample == "deconv"
        up_chns = in_chns // 2 if halves else in_chns
        self.upsample = _UpSampleDeconv(in_chns, up_chns)
        self.convs = TwoConv(spatial_dims, cat_chns + up_chns, out_chns, act, norm, bias, dropout)
        self.is_pad = is_pad

    def forward(self, x, x_e):
        x_0 = self.upsample(x)
        if x_e is not None:
            if self.is_pad:
                dimensions = len(x.shape) - 2
                sp = [0] * (dimensions * 2)
                for i in range(dimensions):
                    if x_e.shape[-i - 1] != x_0.shape[-i - 1]:
                        sp[i * 2 + 1] = 1
                x_0 = torch.nn.functional.pad(x_0, sp, "replicate")
            x = self.convs(torch.cat([x_e, x_0], dim=1))  # skip tensors first, upsampled tensor last
        else:
            x = self.convs(x_0)
        return x
