"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Plain-PyTorch fp32 restatement of the reference hot path of caumente/multi_task_breast_cancer, used only as the checker
by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg.  The product package
(multi_task_breast_cancer_b200/) must never import this file.

What is restated (reference file:line, relative to /root/reference):
  * MTUNetPlusPlus          src/models/multitask/MTUNetPlusPlus.py:11-136  (+ MONAI 1.3.0 TwoConv/Down/UpCat)
  * MTnnUNet                src/models/multitask/MTnnUNet.py:64-183
  * Multi_BTS_UNet          src/models/multitask/Multi_BTS_UNet.py:64-176
  * nnUNet2021, BTSUNet     src/models/segmentation/nnUNet.py:64-162, BTS_UNet.py:64-152 (single-task siblings, row f4)
  * UNetPlusPlusClassifier, nnUNetClassifier, BTSUNetClassifier
                            src/models/classification/UnetPlusPlus_Classifier.py:20-147, nnUNet_classifier.py:73-167,
                            BTS_UNET_classifier.py:55-116 (classification-only siblings, row f4)
  * FocalLoss               src/utils/criterions.py:6-24
  * DiceLoss                monai.losses.DiceLoss as configured at src/utils/experiment_init.py:209-211
  * multi-task criterion    src/utils/criterions.py:52-76
  * training step           src/training_multitask.py:79-103 (Adam eps=1e-4, src/utils/experiment_init.py:186-187)
  * prediction refinement   src/utils/models.py:316-332,366-386
  * epoch loops / metrics   src/training_multitask.py:33-159 (train_one_epoch, validate_one_epoch, class lists),
                            src/utils/metrics.py:26-76,173-267 (calculate_metrics and its scalar helpers; the
                            Hausdorff distance of :236-252 as hausdorff_rows), src/utils/models.py:273-397 (test-time inference, per image)
  * input pipeline          src/dataset/BUSI_dataset.py:97-163 (flip / flip / rotate on cat([mask, image]) through
                            torchvision itself), src/dataset/BUSI_dataloader.py:320-340 (deterministic oversampling;
                            restated for pandas 1.x semantics -- the function does not run under the pandas 3 of this
                            image, so this one is UNPINNED)

Pinning: the reference ships no tests or golden vectors for this path, and MONAI itself is absent, so parity is
UNPINNED at the MONAI boundary.  The restatement is pinned against the reference's own module files instead: in the
build container tests/golden/make_golden.py imports /root/reference (MTUNetPlusPlus through oracle/monai_standin),
checks that every model here has bit-identical parameters, outputs and losses for the same seed, and writes the
fixtures under tests/golden/ that tests/test_oracle_golden.py re-checks on any machine.

Module attribute names follow the reference exactly because they define the state_dict keys (checkpoints are raw
state_dicts: src/training_multitask.py:243-249).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import List, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


# ======================================================================================================================
# U-Net++ (MONAI-style blocks)
# ======================================================================================================================
class _ConvNormAct(nn.Sequential):
    """MONAI `Convolution`: children `conv` (3x3, pad 1, bias) and `adn` = (N: InstanceNorm affine, D: Dropout(0), A)."""

    def __init__(self, cin: int, cout: int, slope: float, bias: bool, dropout: float):
        super().__init__()
        self.add_module("conv", nn.Conv2d(cin, cout, 3, 1, 1, bias=bias))
        adn = nn.Sequential()
        adn.add_module("N", nn.InstanceNorm2d(cout, affine=True))
        adn.add_module("D", nn.Dropout(dropout))
        adn.add_module("A", nn.LeakyReLU(negative_slope=slope, inplace=True))
        self.add_module("adn", adn)


class _TwoConv(nn.Sequential):
    def __init__(self, cin, cout, slope, bias, dropout):
        super().__init__()
        self.add_module("conv_0", _ConvNormAct(cin, cout, slope, bias, dropout))
        self.add_module("conv_1", _ConvNormAct(cout, cout, slope, bias, dropout))


class _Down(nn.Sequential):
    def __init__(self, cin, cout, slope, bias, dropout):
        super().__init__()
        self.add_module("max_pooling", nn.MaxPool2d(kernel_size=2))
        self.add_module("convs", _TwoConv(cin, cout, slope, bias, dropout))


class _UpCat(nn.Module):
    def __init__(self, cin, ccat, cout, slope, bias, dropout, halves=True):
        super().__init__()
        cup = cin // 2 if halves else cin
        up = nn.Sequential()
        up.add_module("deconv", nn.ConvTranspose2d(cin, cup, kernel_size=2, stride=2, bias=True))
        self.upsample = up
        self.convs = _TwoConv(ccat + cup, cout, slope, bias, dropout)

    def forward(self, x, skip):
        u = self.upsample(x)
        pad = [0, 0, 0, 0]
        if skip.shape[-1] != u.shape[-1]:
            pad[1] = 1
        if skip.shape[-2] != u.shape[-2]:
            pad[3] = 1
        u = F.pad(u, pad, "replicate")
        return self.convs(torch.cat([skip, u], dim=1))


class MTUNetPlusPlus(nn.Module):
    """Restates src/models/multitask/MTUNetPlusPlus.py:11-136 (2-D, deconv upsampling only)."""

    def __init__(self, spatial_dims: int = 2, in_channels: int = 1, out_channels: int = 1, n_classes: int = 3,
                 features: Sequence[int] = (24, 48, 96, 192, 384, 24), deep_supervision: bool = False,
                 act=("LeakyReLU", {"negative_slope": 0.1, "inplace": True}), norm=("instance", {"affine": True}),
                 bias: bool = True, dropout: float = 0.0, upsample: str = "deconv"):
        super().__init__()
        assert spatial_dims == 2 and upsample == "deconv"
        slope = float(act[1].get("negative_slope", 0.01))
        self.deep_supervision = deep_supervision
        self.n_classes = 1 if n_classes == 2 else n_classes
        f = tuple(features)
        a = (slope, bias, dropout)
        self.conv_0_0 = _TwoConv(in_channels, f[0], *a)
        self.conv_1_0 = _Down(f[0], f[1], *a)
        self.conv_2_0 = _Down(f[1], f[2], *a)
        self.conv_3_0 = _Down(f[2], f[3], *a)
        self.conv_4_0 = _Down(f[3], f[4], *a)
        self.upcat_0_1 = _UpCat(f[1], f[0], f[0], *a, halves=False)
        self.upcat_1_1 = _UpCat(f[2], f[1], f[1], *a)
        self.upcat_2_1 = _UpCat(f[3], f[2], f[2], *a)
        self.upcat_3_1 = _UpCat(f[4], f[3], f[3], *a)
        self.upcat_0_2 = _UpCat(f[1], f[0] * 2, f[0], *a, halves=False)
        self.upcat_1_2 = _UpCat(f[2], f[1] * 2, f[1], *a)
        self.upcat_2_2 = _UpCat(f[3], f[2] * 2, f[2], *a)
        self.upcat_0_3 = _UpCat(f[1], f[0] * 3, f[0], *a, halves=False)
        self.upcat_1_3 = _UpCat(f[2], f[1] * 3, f[1], *a)
        self.upcat_0_4 = _UpCat(f[1], f[0] * 4, f[5], *a, halves=False)
        self.final_conv_0_1 = nn.Conv2d(f[0], out_channels, kernel_size=1)
        self.final_conv_0_2 = nn.Conv2d(f[0], out_channels, kernel_size=1)
        self.final_conv_0_3 = nn.Conv2d(f[0], out_channels, kernel_size=1)
        self.final_conv_0_4 = nn.Conv2d(f[5], out_channels, kernel_size=1)
        self.process_level_3 = _Down(f[3], f[4], *a)
        self.classifier = nn.Sequential(
            _TwoConv(f[4] * 3, 512, *a), nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(512, 256), nn.ReLU(),
            nn.Linear(256, self.n_classes))

    def forward(self, x):
        x00 = self.conv_0_0(x)
        x10 = self.conv_1_0(x00)
        x01 = self.upcat_0_1(x10, x00)
        x20 = self.conv_2_0(x10)
        x11 = self.upcat_1_1(x20, x10)
        x02 = self.upcat_0_2(x11, torch.cat([x00, x01], 1))
        x30 = self.conv_3_0(x20)
        x21 = self.upcat_2_1(x30, x20)
        x12 = self.upcat_1_2(x21, torch.cat([x10, x11], 1))
        x03 = self.upcat_0_3(x12, torch.cat([x00, x01, x02], 1))
        x40 = self.conv_4_0(x30)
        x31 = self.upcat_3_1(x40, x30)
        x22 = self.upcat_2_2(x31, torch.cat([x20, x21], 1))
        x13 = self.upcat_1_3(x22, torch.cat([x10, x11, x12], 1))
        x04 = self.upcat_0_4(x13, torch.cat([x00, x01, x02, x03], 1))
        o1, o2, o3, o4 = (self.final_conv_0_1(x01), self.final_conv_0_2(x02), self.final_conv_0_3(x03),
                          self.final_conv_0_4(x04))
        feat = torch.cat([self.process_level_3(x30), x40, self.process_level_3(x31)], 1)
        cls = self.classifier(feat)
        if self.deep_supervision:
            return [cls], [o1, o2, o3, o4]
        return cls, o4


# ======================================================================================================================
# nnU-Net style and BTS U-Net
# ======================================================================================================================
def _conv3x3(cin, cout):
    return nn.Conv2d(cin, cout, kernel_size=(3, 3), stride=1, padding=1, bias=False)


def _conv1x1(cin, cout):
    return nn.Conv2d(cin, cout, kernel_size=(1, 1))


class ConvInNormLeReLU(nn.Sequential):
    """conv3x3(bias=False) -> InstanceNorm2d (no affine) -> LeakyReLU(0.01) (MTnnUNet.py:19-39)."""

    def __init__(self, cin, cout):
        super().__init__(OrderedDict([("Conv", _conv3x3(cin, cout)), ("InNorm", nn.InstanceNorm2d(cout)),
                                      ("LeReLU", nn.LeakyReLU(inplace=True))]))


class LevelBlock(nn.Sequential):
    def __init__(self, cin, cmid, cout):
        super().__init__(OrderedDict([("ConvInNormLRelu1", ConvInNormLeReLU(cin, cmid)),
                                      ("ConvInNormLRelu2", ConvInNormLeReLU(cmid, cout))]))


def _kaiming_all_conv2d(module: nn.Module):
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, nonlinearity="leaky_relu")
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)


class MTnnUNet(nn.Module):
    """Restates src/models/multitask/MTnnUNet.py:64-183 (note: class-branch modules are created AFTER the kaiming
    re-initialisation, so they keep PyTorch's default init -- SURVEY Appendix C.2)."""

    def __init__(self, sequences, regions, n_classes=3):
        super().__init__()
        w = [32, 64, 128, 256, 320]
        self.n_classes = 1 if n_classes == 2 else n_classes
        self.encoder1 = LevelBlock(sequences, w[0], w[0])
        self.encoder2 = LevelBlock(w[0], w[1], w[1])
        self.encoder3 = LevelBlock(w[1], w[2], w[2])
        self.encoder4 = LevelBlock(w[2], w[3], w[3])
        self.encoder5 = LevelBlock(w[3], w[4], w[4])
        self.bottleneck = LevelBlock(w[4], w[4], w[4])
        self.decoder5 = LevelBlock(w[4] + w[4], w[3], w[3])
        self.decoder4 = LevelBlock(w[3] + w[3], w[2], w[2])
        self.decoder3 = LevelBlock(w[2] + w[2], w[1], w[1])
        self.decoder2 = LevelBlock(w[1] + w[1], w[0], w[0])
        self.decoder1 = LevelBlock(w[0] + w[0], w[0], w[0] // 2)
        self.upsample5 = nn.ConvTranspose2d(w[4], w[4], kernel_size=2, stride=2)
        self.upsample4 = nn.ConvTranspose2d(w[3], w[3], kernel_size=2, stride=2)
        self.upsample3 = nn.ConvTranspose2d(w[2], w[2], kernel_size=2, stride=2)
        self.upsample2 = nn.ConvTranspose2d(w[1], w[1], kernel_size=2, stride=2)
        self.upsample1 = nn.ConvTranspose2d(w[0], w[0], kernel_size=2, stride=2)
        self.downsample = nn.MaxPool2d(2, 2)
        self.output4 = nn.Sequential(nn.ConvTranspose2d(w[2], w[2], kernel_size=8, stride=8), _conv1x1(w[2], regions))
        self.output3 = nn.Sequential(nn.ConvTranspose2d(w[1], w[1], kernel_size=4, stride=4), _conv1x1(w[1], regions))
        self.output2 = nn.Sequential(nn.ConvTranspose2d(w[0], w[0], kernel_size=2, stride=2), _conv1x1(w[0], regions))
        self.output1 = _conv1x1(w[0] // 2, regions)
        _kaiming_all_conv2d(self)
        self.process_encoder_5 = ConvInNormLeReLU(w[4], w[4])
        self.process_decoder_5 = ConvInNormLeReLU(w[3], w[4])
        self.classifier = nn.Sequential(ConvInNormLeReLU(w[4] * 3, 512), nn.AdaptiveAvgPool2d(1), nn.Flatten(),
                                        nn.Linear(512, 256), nn.ReLU(), nn.Linear(256, self.n_classes))

    def forward(self, x):
        e1 = self.encoder1(x)
        e2 = self.encoder2(self.downsample(e1))
        e3 = self.encoder3(self.downsample(e2))
        e4 = self.encoder4(self.downsample(e3))
        e5 = self.encoder5(self.downsample(e4))
        bott = self.bottleneck(self.downsample(e5))
        d5 = self.decoder5(torch.cat([e5, self.upsample5(bott)], 1))
        d4 = self.decoder4(torch.cat([e4, self.upsample4(d5)], 1))
        d3 = self.decoder3(torch.cat([e3, self.upsample3(d4)], 1))
        d2 = self.decoder2(torch.cat([e2, self.upsample2(d3)], 1))
        d1 = self.decoder1(torch.cat([e1, self.upsample1(d2)], 1))
        feat = torch.cat([self.process_encoder_5(e5), self.upsample5(bott), self.process_decoder_5(d5)], 1)
        cls = self.classifier(feat)
        return [cls], [self.output4(d4), self.output3(d3), self.output2(d2), self.output1(d1)]


class Multi_BTS_UNet(nn.Module):
    """Restates src/models/multitask/Multi_BTS_UNet.py:64-176 (128x128 inputs only: Linear(widths[3]*16*16, 256))."""

    def __init__(self, sequences, regions, n_classes, width, deep_supervision):
        super().__init__()
        self.deep_supervision = deep_supervision
        w = [width * 2 ** i for i in range(4)]
        self.n_classes = 1 if n_classes == 2 else n_classes
        self.encoder1 = LevelBlock(sequences, w[0] // 2, w[0])
        self.encoder2 = LevelBlock(w[0], w[1] // 2, w[1])
        self.encoder3 = LevelBlock(w[1], w[2] // 2, w[2])
        self.encoder4 = LevelBlock(w[2], w[3] // 2, w[3])
        self.bottleneck = LevelBlock(w[3], w[3], w[3])
        self.bottleneck2 = ConvInNormLeReLU(w[3] * 2, w[2])
        self.decoder3 = LevelBlock(w[2] * 2, w[2], w[1])
        self.decoder2 = LevelBlock(w[1] * 2, w[1], w[0])
        self.decoder1 = LevelBlock(w[0] * 2, w[0], w[0] // 2)
        self.upsample = nn.Upsample(scale_factor=2, mode="nearest")
        self.downsample = nn.MaxPool2d(2, 2)
        self.softmax = nn.Softmax(dim=1)
        self.process_bottleneck2 = ConvInNormLeReLU(w[2], w[3])
        self.process_features_map = ConvInNormLeReLU(w[3] * 3, w[3])
        self.classifier = nn.Sequential(nn.Flatten(), nn.Linear(w[3] * 16 * 16, 256), nn.ReLU(),
                                        nn.Linear(256, self.n_classes))
        if self.deep_supervision:
            self.output3 = nn.Sequential(nn.ConvTranspose2d(w[1], w[1], kernel_size=4, stride=4), _conv1x1(w[1], regions))
            self.output2 = nn.Sequential(nn.ConvTranspose2d(w[0], w[0], kernel_size=2, stride=2), _conv1x1(w[0], regions))
        self.output1 = _conv1x1(w[0] // 2, regions)
        _kaiming_all_conv2d(self)

    def forward(self, x):
        e1 = self.encoder1(x)
        e2 = self.encoder2(self.downsample(e1))
        e3 = self.encoder3(self.downsample(e2))
        e4 = self.encoder4(self.downsample(e3))
        bott = self.bottleneck(e4)
        bott2 = self.bottleneck2(torch.cat([e4, bott], 1))
        d3 = self.decoder3(torch.cat([e3, self.upsample(bott2)], 1))
        d2 = self.decoder2(torch.cat([e2, self.upsample(d3)], 1))
        d1 = self.decoder1(torch.cat([e1, self.upsample(d2)], 1))
        feat = self.process_features_map(torch.cat([e4, bott, self.process_bottleneck2(bott2)], 1))
        cls = self.classifier(feat)
        if self.deep_supervision:
            return [cls], [self.output3(d3), self.output2(d2), self.output1(d1)]
        return cls, self.output1(d1)


class nnUNet2021(nn.Module):
    """Restates src/models/segmentation/nnUNet.py:64-162: MTnnUNet without the classification branch; returns the list
    [output4, output3, output2, output1]."""

    def __init__(self, sequences, regions):
        super().__init__()
        w = [32, 64, 128, 256, 320]
        self.encoder1 = LevelBlock(sequences, w[0], w[0])
        self.encoder2 = LevelBlock(w[0], w[1], w[1])
        self.encoder3 = LevelBlock(w[1], w[2], w[2])
        self.encoder4 = LevelBlock(w[2], w[3], w[3])
        self.encoder5 = LevelBlock(w[3], w[4], w[4])
        self.bottleneck = LevelBlock(w[4], w[4], w[4])
        self.decoder5 = LevelBlock(w[4] + w[4], w[3], w[3])
        self.decoder4 = LevelBlock(w[3] + w[3], w[2], w[2])
        self.decoder3 = LevelBlock(w[2] + w[2], w[1], w[1])
        self.decoder2 = LevelBlock(w[1] + w[1], w[0], w[0])
        self.decoder1 = LevelBlock(w[0] + w[0], w[0], w[0] // 2)
        self.upsample5 = nn.ConvTranspose2d(w[4], w[4], kernel_size=2, stride=2)
        self.upsample4 = nn.ConvTranspose2d(w[3], w[3], kernel_size=2, stride=2)
        self.upsample3 = nn.ConvTranspose2d(w[2], w[2], kernel_size=2, stride=2)
        self.upsample2 = nn.ConvTranspose2d(w[1], w[1], kernel_size=2, stride=2)
        self.upsample1 = nn.ConvTranspose2d(w[0], w[0], kernel_size=2, stride=2)
        self.downsample = nn.MaxPool2d(2, 2)
        self.output4 = nn.Sequential(nn.ConvTranspose2d(w[2], w[2], kernel_size=8, stride=8), _conv1x1(w[2], regions))
        self.output3 = nn.Sequential(nn.ConvTranspose2d(w[1], w[1], kernel_size=4, stride=4), _conv1x1(w[1], regions))
        self.output2 = nn.Sequential(nn.ConvTranspose2d(w[0], w[0], kernel_size=2, stride=2), _conv1x1(w[0], regions))
        self.output1 = _conv1x1(w[0] // 2, regions)
        _kaiming_all_conv2d(self)

    def forward(self, x):
        e1 = self.encoder1(x)
        e2 = self.encoder2(self.downsample(e1))
        e3 = self.encoder3(self.downsample(e2))
        e4 = self.encoder4(self.downsample(e3))
        e5 = self.encoder5(self.downsample(e4))
        bott = self.bottleneck(self.downsample(e5))
        d5 = self.decoder5(torch.cat([e5, self.upsample5(bott)], 1))
        d4 = self.decoder4(torch.cat([e4, self.upsample4(d5)], 1))
        d3 = self.decoder3(torch.cat([e3, self.upsample3(d4)], 1))
        d2 = self.decoder2(torch.cat([e2, self.upsample2(d3)], 1))
        d1 = self.decoder1(torch.cat([e1, self.upsample1(d2)], 1))
        return [self.output4(d4), self.output3(d3), self.output2(d2), self.output1(d1)]


class BTSUNet(nn.Module):
    """Restates src/models/segmentation/BTS_UNet.py:64-152: Multi_BTS_UNet without the classification branch."""

    def __init__(self, sequences, regions, width, deep_supervision):
        super().__init__()
        self.deep_supervision = deep_supervision
        w = [width * 2 ** i for i in range(4)]
        self.encoder1 = LevelBlock(sequences, w[0] // 2, w[0])
        self.encoder2 = LevelBlock(w[0], w[1] // 2, w[1])
        self.encoder3 = LevelBlock(w[1], w[2] // 2, w[2])
        self.encoder4 = LevelBlock(w[2], w[3] // 2, w[3])
        self.bottleneck = LevelBlock(w[3], w[3], w[3])
        self.bottleneck2 = ConvInNormLeReLU(w[3] * 2, w[2])
        self.decoder3 = LevelBlock(w[2] * 2, w[2], w[1])
        self.decoder2 = LevelBlock(w[1] * 2, w[1], w[0])
        self.decoder1 = LevelBlock(w[0] * 2, w[0], w[0] // 2)
        self.upsample = nn.Upsample(scale_factor=2, mode="nearest")
        self.downsample = nn.MaxPool2d(2, 2)
        if self.deep_supervision:
            self.output3 = nn.Sequential(nn.ConvTranspose2d(w[1], w[1], kernel_size=4, stride=4), _conv1x1(w[1], regions))
            self.output2 = nn.Sequential(nn.ConvTranspose2d(w[0], w[0], kernel_size=2, stride=2), _conv1x1(w[0], regions))
        self.output1 = _conv1x1(w[0] // 2, regions)
        _kaiming_all_conv2d(self)

    def forward(self, x):
        e1 = self.encoder1(x)
        e2 = self.encoder2(self.downsample(e1))
        e3 = self.encoder3(self.downsample(e2))
        e4 = self.encoder4(self.downsample(e3))
        bott = self.bottleneck(e4)
        bott2 = self.bottleneck2(torch.cat([e4, bott], 1))
        d3 = self.decoder3(torch.cat([e3, self.upsample(bott2)], 1))
        d2 = self.decoder2(torch.cat([e2, self.upsample(d3)], 1))
        d1 = self.decoder1(torch.cat([e1, self.upsample(d2)], 1))
        if self.deep_supervision:
            return [self.output3(d3), self.output2(d2), self.output1(d1)]
        return self.output1(d1)


class _RUInBlock(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, kernel_size=3, padding=1)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = nn.Conv2d(cout, cout, kernel_size=3, padding=1)
        self.conv3 = nn.Conv2d(cin, cout, kernel_size=3, padding=1)
        self.bn3 = nn.BatchNorm2d(cout)


class _RUResBlock(nn.Module):
    def __init__(self, cin, downsample=False):
        super().__init__()
        cout, st = (2 * cin, 2) if downsample else (cin, 1)
        self.bn1 = nn.BatchNorm2d(cin)
        self.conv1 = nn.Conv2d(cin, cout, kernel_size=3, stride=st, padding=1)
        self.bn2 = nn.BatchNorm2d(cout)
        self.conv2 = nn.Conv2d(cout, cout, kernel_size=3, padding=1)
        self.conv3 = nn.Conv2d(cin, cout, kernel_size=3, stride=st, padding=1)
        self.bn3 = nn.BatchNorm2d(cout)


class _RUEncoder(nn.Module):
    def __init__(self, bf):
        super().__init__()
        self.down_block2 = _RUResBlock(bf, True)
        self.down_block3 = _RUResBlock(bf * 2, True)
        self.down_block4 = _RUResBlock(bf * 4, True)


class _RUDecoder(nn.Module):
    def __init__(self, bf):
        super().__init__()
        self.upsample3 = nn.ConvTranspose2d(bf * 8, bf * 4, kernel_size=2, stride=2)
        self.conv3 = nn.Conv2d(bf * 8, bf * 4, kernel_size=1)      # skip-variant only: never called (ResidualUNet.py:356-362)
        self.up_block3 = _RUResBlock(bf * 4)
        self.upsample2 = nn.ConvTranspose2d(bf * 4, bf * 2, kernel_size=2, stride=2)
        self.conv2 = nn.Conv2d(bf * 4, bf * 2, kernel_size=1)
        self.up_block2 = _RUResBlock(bf * 2)
        self.upsample1 = nn.ConvTranspose2d(bf * 2, bf, kernel_size=2, stride=2)
        self.conv1 = nn.Conv2d(bf * 2, bf, kernel_size=1)
        self.up_block1 = _RUResBlock(bf)


class _RUOut(nn.Module):
    def __init__(self, bf, n_classes):
        super().__init__()
        self.conv = nn.Conv2d(bf, n_classes, kernel_size=1)


class ResidualUNet(nn.Module):
    """Restates src/models/segmentation/ResidualUNet.py:342-362 (ResidualUNet.forward: in_block -> encoder -> decoder
    -> out_block, i.e. the variant WITHOUT skip connections; `seg_path` :298-339 is not what the factory builds).
    BatchNorm2d with running statistics, stride-2 3x3 convs, F.leaky_relu (slope 0.01) and F.dropout(p=0.2) called
    with its default training=True (:61,139,145: active in eval mode too).  `dropout` is the hook parity tests use to
    substitute fixed masks for torch's generator; by default it IS F.dropout, drawn in the reference's call order."""

    def __init__(self, sequences=1, regions=1, width=24):
        super().__init__()
        self.in_block = _RUInBlock(sequences, width)
        self.encoder = _RUEncoder(width)
        self.decoder = _RUDecoder(width)
        self.out_block = _RUOut(width, regions)
        self.dropout = lambda t: F.dropout(t, p=0.2)

    def _res(self, blk, x):
        path = self.dropout(F.leaky_relu(blk.bn1(x)))
        path = self.dropout(F.leaky_relu(blk.bn2(blk.conv1(path))))
        path = blk.conv2(path)
        return path + blk.bn3(blk.conv3(x))

    def forward(self, x):
        ib = self.in_block
        path = ib.conv2(self.dropout(F.leaky_relu(ib.bn1(ib.conv1(x)))))
        x = path + ib.bn3(ib.conv3(x))
        for blk in (self.encoder.down_block2, self.encoder.down_block3, self.encoder.down_block4):
            x = self._res(blk, x)
        d = self.decoder
        x = self._res(d.up_block3, d.upsample3(x))
        x = self._res(d.up_block2, d.upsample2(x))
        x = self._res(d.up_block1, d.upsample1(x))
        return self.out_block.conv(x)


class UNetPlusPlusClassifier(nn.Module):
    """Restates src/models/classification/UnetPlusPlus_Classifier.py:20-147: encoder + upcat_3_1 + class branch of
    MTUNetPlusPlus; raw class logits (the softmax is commented out in the reference, :142-143)."""

    def __init__(self, spatial_dims: int = 2, in_channels: int = 1, n_classes: int = 3,
                 features: Sequence[int] = (24, 48, 96, 192, 384, 24),
                 act=("LeakyReLU", {"negative_slope": 0.1, "inplace": True}), norm=("instance", {"affine": True}),
                 bias: bool = True, dropout: float = 0.0, upsample: str = "deconv"):
        super().__init__()
        assert spatial_dims == 2 and upsample == "deconv"
        slope = float(act[1].get("negative_slope", 0.01))
        self.n_classes = 1 if n_classes == 2 else n_classes
        f = tuple(features)
        a = (slope, bias, dropout)
        self.conv_0_0 = _TwoConv(in_channels, f[0], *a)
        self.conv_1_0 = _Down(f[0], f[1], *a)
        self.conv_2_0 = _Down(f[1], f[2], *a)
        self.conv_3_0 = _Down(f[2], f[3], *a)
        self.conv_4_0 = _Down(f[3], f[4], *a)
        self.upcat_3_1 = _UpCat(f[4], f[3], f[3], *a)
        self.softmax = nn.Softmax(dim=1)
        self.process_level_3 = _Down(f[3], f[4], *a)
        self.classifier = nn.Sequential(
            _TwoConv(f[4] * 3, 512, *a), nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(512, 256), nn.ReLU(),
            nn.Linear(256, self.n_classes))

    def forward(self, x):
        x30 = self.conv_3_0(self.conv_2_0(self.conv_1_0(self.conv_0_0(x))))
        x40 = self.conv_4_0(x30)
        x31 = self.upcat_3_1(x40, x30)
        return self.classifier(torch.cat([self.process_level_3(x30), x40, self.process_level_3(x31)], 1))


class nnUNetClassifier(nn.Module):
    """Restates src/models/classification/nnUNet_classifier.py:73-167: encoder, bottleneck, decoder5 and the class branch
    of MTnnUNet (decoder4..1 are owned but never evaluated); softmax probabilities when n_classes > 2 (:165-166)."""

    def __init__(self, sequences, n_classes=3):
        super().__init__()
        w = [32, 64, 128, 256, 320]
        self.n_classes = 1 if n_classes == 2 else n_classes
        self.encoder1 = LevelBlock(sequences, w[0], w[0])
        self.encoder2 = LevelBlock(w[0], w[1], w[1])
        self.encoder3 = LevelBlock(w[1], w[2], w[2])
        self.encoder4 = LevelBlock(w[2], w[3], w[3])
        self.encoder5 = LevelBlock(w[3], w[4], w[4])
        self.bottleneck = LevelBlock(w[4], w[4], w[4])
        self.decoder5 = LevelBlock(w[4] + w[4], w[3], w[3])
        self.decoder4 = LevelBlock(w[3] + w[3], w[2], w[2])
        self.decoder3 = LevelBlock(w[2] + w[2], w[1], w[1])
        self.decoder2 = LevelBlock(w[1] + w[1], w[0], w[0])
        self.decoder1 = LevelBlock(w[0] + w[0], w[0], w[0] // 2)
        self.upsample5 = nn.ConvTranspose2d(w[4], w[4], kernel_size=2, stride=2)
        self.downsample = nn.MaxPool2d(2, 2)
        _kaiming_all_conv2d(self)      # the class branch below is created afterwards and keeps the default init
        self.softmax = nn.Softmax(dim=1)
        self.process_encoder_5 = ConvInNormLeReLU(w[4], w[4])
        self.process_decoder_5 = ConvInNormLeReLU(w[3], w[4])
        self.classifier = nn.Sequential(ConvInNormLeReLU(w[4] * 3, 512), nn.AdaptiveAvgPool2d(1), nn.Flatten(),
                                        nn.Linear(512, 256), nn.ReLU(), nn.Linear(256, self.n_classes))

    def forward(self, x):
        e1 = self.encoder1(x)
        e2 = self.encoder2(self.downsample(e1))
        e3 = self.encoder3(self.downsample(e2))
        e4 = self.encoder4(self.downsample(e3))
        e5 = self.encoder5(self.downsample(e4))
        bott = self.bottleneck(self.downsample(e5))
        d5 = self.decoder5(torch.cat([e5, self.upsample5(bott)], 1))
        feat = torch.cat([self.process_encoder_5(e5), self.upsample5(bott), self.process_decoder_5(d5)], 1)
        out = self.classifier(feat)
        return self.softmax(out) if self.n_classes > 2 else out


class BTSUNetClassifier(nn.Module):
    """Restates src/models/classification/BTS_UNET_classifier.py:55-116 (128x128 inputs: Linear(8*width*8*8, 256))."""

    def __init__(self, sequences, classes, width, deep_supervision=False):
        super().__init__()
        self.deep_supervision = deep_supervision
        w = [width * 2 ** i for i in range(4)]
        self.classes = 1 if classes == 2 else classes
        self.encoder = nn.Sequential(
            LevelBlock(sequences, w[0] // 2, w[0]), nn.MaxPool2d(2, 2), LevelBlock(w[0], w[1] // 2, w[1]),
            nn.MaxPool2d(2, 2), LevelBlock(w[1], w[2] // 2, w[2]), nn.MaxPool2d(2, 2),
            LevelBlock(w[2], w[3] // 2, w[3]), nn.MaxPool2d(2, 2), LevelBlock(w[3], w[3], w[3]))
        self.classifier = nn.Sequential(nn.Flatten(), nn.Linear(w[3] * 8 * 8, 256), nn.ReLU(),
                                        nn.Linear(256, self.classes))
        _kaiming_all_conv2d(self)

    def forward(self, x):
        return self.classifier(self.encoder(x))


def build_model(architecture: str, sequences: int = 1, regions: int = 1, n_classes: int = 3, width: int = 32,
                deep_supervision: bool = True) -> nn.Module:
    """Restates init_multitask_model (src/utils/experiment_init.py:130-174)."""
    if architecture == "Multi_BTSUNet":
        return Multi_BTS_UNet(sequences=sequences, regions=regions, n_classes=n_classes, width=width,
                              deep_supervision=deep_supervision)
    if architecture == "MTUNetPlusPlus":
        return MTUNetPlusPlus(in_channels=sequences, out_channels=regions, n_classes=n_classes,
                              deep_supervision=deep_supervision)
    if architecture == "MTnnUNet":
        return MTnnUNet(sequences=sequences, regions=regions, n_classes=n_classes)
    return nn.Module()  # the reference silently returns an empty module (experiment_init.py:160-163)


# ======================================================================================================================
# Losses, criterion glue, training step, refinement
# ======================================================================================================================
class FocalLoss(nn.Module):
    """src/utils/criterions.py:6-24: soft-target cross entropy -> (1 - exp(-ce))^gamma * ce."""

    def __init__(self, alpha=1, gamma=2, reduction="mean", weight=None):
        super().__init__()
        self.alpha, self.gamma, self.reduction, self.weight = alpha, gamma, reduction, weight

    def forward(self, inputs, targets):
        ce = F.cross_entropy(inputs, targets, reduction="none", weight=self.weight)
        fl = self.alpha * (1 - torch.exp(-ce)) ** self.gamma * ce
        if self.reduction == "mean":
            return fl.mean()
        if self.reduction == "sum":
            return fl.sum()
        return fl


class DiceLoss(nn.Module):
    """monai DiceLoss(include_background=True, sigmoid=True, smooth_nr=1, smooth_dr=1, squared_pred=True)."""

    def __init__(self, smooth_nr: float = 1.0, smooth_dr: float = 1.0):
        super().__init__()
        self.smooth_nr, self.smooth_dr = smooth_nr, smooth_dr

    def forward(self, logits, target):
        p = torch.sigmoid(logits)
        axes = list(range(2, p.dim()))
        inter = (target * p).sum(axes)
        den = (target ** 2).sum(axes) + (p ** 2).sum(axes)
        return (1.0 - (2.0 * inter + self.smooth_nr) / (den + self.smooth_dr)).mean()


def multitask_criterion(criterion_seg, ground_truth, segmentation, criterion_class, label, predicted_class,
                        inversely_weighted=False):
    """src/utils/criterions.py:52-76 without the host-side NaN exit (callers check)."""
    if isinstance(segmentation, list):
        terms = []
        for n, s in enumerate(reversed(segmentation)):
            t = criterion_seg(s, ground_truth)
            terms.append(t / (n + 1) if inversely_weighted else t)
        seg = torch.stack(terms).sum()
        cls = torch.stack([criterion_class(c, label) for c in reversed(predicted_class)]).sum()
    else:
        seg = criterion_seg(segmentation, ground_truth)
        cls = criterion_class(predicted_class, label)
    return seg, cls


def train_step(model, optimizer, inputs, masks, onehot, alpha=0.35, inversely_weighted=True,
               seg_criterion=None, cls_criterion=None):
    """src/training_multitask.py:87-103: zero_grad -> forward -> losses -> alpha mix -> backward -> step."""
    seg_criterion = seg_criterion or DiceLoss()
    cls_criterion = cls_criterion or FocalLoss(alpha=1, gamma=2)
    optimizer.zero_grad(set_to_none=True)
    logits, outputs = model(inputs)
    seg, cls = multitask_criterion(seg_criterion, masks, outputs, cls_criterion, onehot, logits, inversely_weighted)
    total = alpha * seg + (1 - alpha) * cls
    total.backward()
    optimizer.step()
    return total.detach(), seg.detach(), cls.detach(), logits, outputs


def make_optimizer(model, lr=1e-4):
    return torch.optim.Adam(model.parameters(), lr=lr, eps=1e-4)  # experiment_init.py:186-187


def refine_predictions(mask_logits, class_logits, normal_id=2, seg_by_class=True, class_by_seg=True, threshold=0):
    """Batched restatement of src/utils/models.py:316-332 and :366-386 (the reference runs it per image on numpy).
    Returns (uint8 mask (B,1,H,W), int64 class (B,), int64 pixel count (B,)); both refinements read the initial
    predictions."""
    m = (torch.sigmoid(mask_logits) > 0.5)
    cnt = m.flatten(1).sum(1)
    cls = class_logits.argmax(1)
    keep = torch.ones_like(cnt, dtype=torch.bool)
    if threshold > 0:
        keep &= cnt > threshold
    if seg_by_class:
        keep &= cls != normal_id
    refined_mask = (m & keep.view(-1, 1, 1, 1)).to(torch.uint8)
    refined_cls = torch.where((cnt == 0) & bool(class_by_seg), torch.full_like(cls, normal_id), cls)
    return refined_mask, refined_cls, cnt


def hard_dice(masks, outputs) -> float:
    """src/training_multitask.py:65-71 + src/utils/metrics.py:255-267 (whole-batch hard Dice)."""
    if isinstance(outputs, list):
        outputs = outputs[-1]
    seg = (torch.sigmoid(outputs) > 0.5)
    gt = masks.bool()
    tp = (seg & gt).sum().double()
    fp = (seg & ~gt).sum().double()
    fn = (~seg & gt).sum().double()
    if gt.sum() == 0:
        return 1.0 if seg.sum() == 0 else 0.0
    return float(2 * tp / (2 * tp + fp + fn))


def synthetic_batch(B, H, W, n_classes=3, seed=1993, device="cpu"):
    """Synthetic inputs of SURVEY 8(d): integer-valued 0..255 image, filled-ellipse masks (class 2 = empty mask),
    balanced labels."""
    g = torch.Generator().manual_seed(seed)
    img = torch.randint(0, 256, (B, 1, H, W), generator=g).float()
    label = torch.arange(B) % n_classes
    yy, xx = torch.meshgrid(torch.arange(H).float(), torch.arange(W).float(), indexing="ij")
    masks = torch.zeros(B, 1, H, W)
    for b in range(B):
        if int(label[b]) == 2:
            continue
        cy, cx = torch.rand(2, generator=g) * torch.tensor([H, W]) * 0.6 + torch.tensor([H, W]) * 0.2
        ry, rx = torch.rand(2, generator=g) * torch.tensor([H, W]) * 0.2 + torch.tensor([H, W]) * 0.05
        masks[b, 0] = (((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1).float()
    onehot = F.one_hot(label, n_classes).float()
    return img.to(device), masks.to(device), onehot.to(device), label.to(device)


# ======================================================================================================================
# epoch loops and metrics (checker for multi_task_breast_cancer_b200/trainer.py)
# ======================================================================================================================
def class_lists(pred_logits, onehot, gt_list, pred_list):
    """src/training_multitask.py:33-62, num_classes > 2 branch: mean over the logits list, softmax, argmax."""
    if isinstance(pred_logits, list):
        pred_logits = torch.mean(torch.stack(pred_logits, dim=0), dim=0)
    prob = F.softmax(pred_logits, dim=1)
    for la, p in zip(onehot, prob):
        gt_list.append(float(torch.argmax(la).item()))
        pred_list.append(float(torch.argmax(p).item()))
    return gt_list, pred_list


def classification_scores(gt_list, pred_list):
    """accuracy_score + f1(labels=[0, 1, 2], average='weighted') as at src/training_multitask.py:112-113."""
    from sklearn.metrics import accuracy_score, f1_score
    return (float(accuracy_score(gt_list, pred_list)),
            float(f1_score(y_true=gt_list, y_pred=pred_list, labels=[0, 1, 2], average="weighted", zero_division=0)))


def train_one_epoch(model, optimizer, batches, alpha=0.35, inversely_weighted=True):
    """src/training_multitask.py:74-116 over a list of (image, mask, one-hot) batches."""
    loss, dice, gt, pr = 0.0, 0.0, [], []
    for img, mask, onehot in batches:
        tot, _, _, logits, outs = train_step(model, optimizer, img, mask, onehot, alpha, inversely_weighted)
        loss += tot.item()
        dice += hard_dice(mask, outs)
        gt, pr = class_lists([l.detach() for l in logits] if isinstance(logits, list) else logits.detach(), onehot, gt, pr)
    acc, f1w = classification_scores(gt, pr)
    return loss / len(batches), dice / len(batches), acc, f1w


@torch.inference_mode()
def validate_one_epoch(model, batches, alpha=0.35, inversely_weighted=True):
    """src/training_multitask.py:119-159."""
    loss, segl, clsl, dice, gt, pr = 0.0, 0.0, 0.0, 0.0, [], []
    dl, fl = DiceLoss(), FocalLoss(alpha=1, gamma=2)
    for img, mask, onehot in batches:
        logits, outs = model(img)
        seg, cls = multitask_criterion(dl, mask, outs, fl, onehot, logits, inversely_weighted)
        loss += (alpha * seg + (1 - alpha) * cls).item()
        segl += seg.item()
        clsl += cls.item()
        dice += hard_dice(mask, outs)
        gt, pr = class_lists(logits, onehot, gt, pr)
    acc, f1w = classification_scores(gt, pr)
    n = len(batches)
    return loss / n, dice / n, acc, f1w, segl / n, clsl / n


def segmentation_metrics(gt, seg):
    """calculate_metrics (src/utils/metrics.py:26-76) without the Hausdorff distance, from boolean numpy-like arrays."""
    import numpy as np
    gt = np.asarray(gt).astype(float)
    seg = np.asarray(seg).astype(float)
    tp = float(np.sum(np.logical_and(seg, gt)))
    tn = float(np.sum(np.logical_and(np.logical_not(seg), np.logical_not(gt))))
    fp = float(np.sum(np.logical_and(seg, np.logical_not(gt))))
    fn = float(np.sum(np.logical_and(np.logical_not(seg), gt)))
    empty_gt, empty_seg = np.sum(gt) == 0, np.sum(seg) == 0
    return {
        "DICE": (1 if empty_seg else 0) if empty_gt else 2 * tp / (2 * tp + fp + fn),
        "Sensitivity": np.nan if tp == 0 else tp / (tp + fn),
        "Specificity": tn / (tn + fp),
        "Accuracy": (tp + tn) / (tp + tn + fp + fn),
        "Jaccard index": (1 if empty_seg else 0) if empty_gt else tp / (tp + fp + fn),
        "Precision": np.nan if tp == 0 else tp / (tp + fp),
    }


def hausdorff_rows(gt, seg) -> float:
    """haussdorf_distance (src/utils/metrics.py:236-252) restated in numpy: scipy's directed_hausdorff is handed the two
    (H, W) boolean images, so every image ROW is one point of {0,1}^W and the distance between rows is
    sqrt(Hamming).  NaN when exactly one of the masks is empty (the reference's second `if`), 0 when both are."""
    import numpy as np
    gt = np.asarray(gt, dtype=bool)
    seg = np.asarray(seg, dtype=bool)
    if gt.ndim == 4:
        gt, seg = gt[0, 0], seg[0, 0]
    if (gt.sum() == 0) != (seg.sum() == 0):
        return float("nan")
    d2 = (seg[:, None, :] != gt[None, :, :]).sum(-1)          # [i, j] = Hamming(seg row i, gt row j)
    return float(np.sqrt(float(max(d2.min(1).max(), d2.min(0).max()))))


@torch.no_grad()
def inference_multitask(mask_logits, class_logits, masks, labels, overlap_seg_based_on_class=False,
                        overlap_class_based_on_seg=False, normal_id=2):
    """src/utils/models.py:299-386 image by image, GIVEN the model outputs (full-decoder mask logits (B,1,H,W), class
    logits (B,K)): -> (segmentation rows, classification rows) like the two CSV files (no postprocess)."""
    seg_rows, cls_rows = [], []
    for b in range(mask_logits.shape[0]):
        out = (torch.sigmoid(mask_logits[b:b + 1]) > 0.5).float().cpu().numpy()
        pred = int(class_logits[b].argmax().item())
        seg_out = out.copy()
        if overlap_seg_based_on_class and pred == normal_id:
            seg_out[seg_out > 0] = 0
        row = segmentation_metrics(masks[b:b + 1].cpu().numpy(), seg_out)
        row["Haussdorf distance"] = hausdorff_rows(masks[b:b + 1].cpu().numpy(), seg_out)
        row["class"] = int(labels[b])
        seg_rows.append(row)
        tumor_pixels = int((out == 1).sum())
        final = normal_id if (overlap_class_based_on_seg and tumor_pixels == 0) else pred
        cls_rows.append({"ground_truth": int(labels[b]), "predicted_label": final})
    return seg_rows, cls_rows


# ======================================================================================================================
# input pipeline (checker for multi_task_breast_cancer_b200/data.py)
# ======================================================================================================================
def augment_sample(image_u8, mask_u8, hflip: bool, vflip: bool, angle):
    """What BUSI.__getitem__ (src/dataset/BUSI_dataset.py:100-163) + the transforms of training_multitask.py:193-197 do
    to one sample for a GIVEN draw, using torchvision's own functional ops on the host: joined = cat([mask, image]);
    hflip; vflip; rotate(angle) (nearest, zero fill).  `angle=None` = no transforms (validation / test loaders)."""
    import torchvision.transforms.functional as TF
    image = torch.unsqueeze(torch.as_tensor(image_u8, dtype=torch.float32), 0)
    mask = torch.unsqueeze(torch.as_tensor(mask_u8, dtype=torch.float32), 0)
    joined = torch.cat([mask, image], dim=0)
    if hflip:
        joined = TF.hflip(joined)
    if vflip:
        joined = TF.vflip(joined)
    if angle is not None:
        joined = TF.rotate(joined, float(angle))
    return joined[1:2], joined[0:1]


def draw_reference_transform_params(n: int):
    """Run torchvision's RandomHorizontalFlip / RandomVerticalFlip / RandomRotation(360) MODULES sample by sample on a
    probe image and read back what they drew (flip decisions from the probe, the angle from get_params' RNG call)."""
    import torchvision.transforms as T
    hf, vf, ang = [], [], []
    probe = torch.arange(16, dtype=torch.float32).reshape(1, 4, 4)
    for _ in range(n):
        h = T.RandomHorizontalFlip(p=0.5)(probe)
        hf.append(not torch.equal(h, probe))
        v = T.RandomVerticalFlip(p=0.5)(probe)
        vf.append(not torch.equal(v, probe))
        ang.append(float(T.RandomRotation.get_params([-360.0, 360.0])))
    return hf, vf, ang


def deterministic_oversampling(classes):
    """src/dataset/BUSI_dataloader.py:320-340 with pandas-1.x `value_counts().reset_index()` semantics, on a list of
    class names; returns the row indices of the concatenated frame."""
    import numpy as np
    classes = list(classes)
    names, first, counts = [], {}, {}
    for i, c in enumerate(classes):
        if c not in counts:
            names.append(c); first[c] = i; counts[c] = 0
        counts[c] += 1
    n = len(classes)
    order = sorted(names, key=lambda c: (-counts[c], first[c]))
    rows = list(range(n))
    for c in order:
        factor = int(np.round(1.0 / (counts[c] / n), 0))
        block = [i for i, k in enumerate(classes) if k == c]
        if factor > 1:
            for _ in range(factor - 1):
                rows += block
        else:
            rows += block
    return rows
