"""Drop-in nn.Modules for the reference's multi-task models.

Same constructors, same `forward(x) -> (class logits, mask logits)` (lists under deep supervision), same parameter
names / shapes / dtypes / initialisation as

    MTUNetPlusPlus   /root/reference/src/models/multitask/MTUNetPlusPlus.py:11-136   (MONAI TwoConv/Down/UpCat blocks)
    MTnnUNet         /root/reference/src/models/multitask/MTnnUNet.py:64-183
    Multi_BTS_UNet   /root/reference/src/models/multitask/Multi_BTS_UNet.py:64-176

The torch.nn sub-modules below are only PARAMETER CONTAINERS (they give the reference's state_dict keys, init and
`print(model)` output); they are never called.  `forward` runs a static `Plan` of hand-written sm_100a kernels
(plan.py) through one torch.autograd.Function; gradients land in `param.grad` (fp32, parameter shaped) for
torch.optim.  There is no CPU / cuDNN fallback: a non-CUDA input raises.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch
import torch.nn as nn

from . import _lib
from .plan import Plan, PTensor


# ======================================================================================================================
# autograd boundary
# ======================================================================================================================
class _PlanFunction(torch.autograd.Function):
    """image -> (class logits..., mask logits...) through the plan; backward writes param.grad directly."""

    @staticmethod
    def forward(ctx, module, plan, x, *params):
        if x.data_ptr() != plan.x_in.data_ptr():
            plan.x_in.copy_(x)
        plan.run_pack()
        plan.run_forward()
        ctx.module, ctx.plan = module, plan
        # The plan owns ONE set of activation buffers: a later forward through the same plan overwrites what this
        # forward's backward needs.  Remember which forward this is so that backward can refuse a stale one.
        plan.generation = getattr(plan, "generation", 0) + 1
        ctx.generation = plan.generation
        # fresh tensor objects aliasing the plan's static output buffers (no copy)
        return tuple(t.detach() for t in plan.outputs_cls) + tuple(t.detach() for t in plan.outputs_seg)

    @staticmethod
    def backward(ctx, *grads):
        plan: Plan = ctx.plan
        if ctx.generation != plan.generation:
            raise RuntimeError(
                "backward() through a forward whose activations are gone: this module keeps one static activation set "
                f"per input shape and forward #{plan.generation} has overwritten those of forward #{ctx.generation}. "
                "Call backward() before the next forward of the same shape (gradient accumulation as "
                "forward; backward; forward; backward works), or run the extra forward under torch.no_grad().")
        # Gradient accumulation (backward without zero_grad in between): a `.grad` delivered by an earlier backward IS a
        # view of the plan's gradient buffer, which this backward is about to overwrite.  Move such gradients to memory
        # of their own first; _deliver_grads then adds the new gradient to them (torch's `.grad +=` semantics).
        for name, p in plan.params.items():
            if p.grad is not None and p.grad.data_ptr() == plan.grad_view[name].data_ptr():
                p.grad = p.grad.clone()
        ncls = len(plan.outputs_cls)
        for buf, g in zip(plan.g_cls, grads[:ncls]):
            if g is None:
                buf.zero_()
            elif g.data_ptr() != buf.data_ptr():
                buf.copy_(g)
        for i, (buf, g) in enumerate(zip(plan.g_seg, grads[ncls:])):
            if not plan.seg_grad_active[i]:
                continue
            if g is None:
                buf.zero_()
            elif g.data_ptr() != buf.data_ptr():
                buf.copy_(g)
        plan.run_backward()
        ctx.module._deliver_grads(plan)
        return (None, None, None) + (None,) * len(plan.params)


class _PlanModule(nn.Module):
    """Shared machinery: plan cache keyed on (shape, device, training-needed, parameter storage)."""

    _slope: float = 0.01

    def _init_runtime(self):
        import os
        self._plans: Dict[Tuple, Plan] = {}
        # Arithmetic mode (not part of the reference's constructor: set the attribute, call set_precision(), or export
        # MTBC_PRECISION / MTBC_DETERMINISTIC before building the model).  "bf16" is the product path; "tf32" and
        # "tf32x3" are the parity modes of north_star ("1e-3 (TF32 mode)"): correct forward + backward, not fast.
        self.precision = os.environ.get("MTBC_PRECISION", "bf16")
        self.deterministic = os.environ.get("MTBC_DETERMINISTIC", "0") not in ("", "0")

    def set_precision(self, precision: str = "bf16", deterministic: Optional[bool] = None):
        """"bf16" (default: the product path), "tf32" (fp32 storage, tcgen05 kind::tf32 convolutions) or "tf32x3" (same
        with the three-product split: fp32-grade results; the mode the parity tests separate wiring from precision with).  deterministic=True makes the forward pass
        bit-reproducible (order-independent InstanceNorm statistics) at the price of one more read of every conv
        output."""
        from .plan import PRECISIONS
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {PRECISIONS}")
        self.precision = precision
        if deterministic is not None:
            self.deterministic = bool(deterministic)
        return self

    # nn.Module.__setstate__/deepcopy safety
    def __getstate__(self):
        d = self.__dict__.copy()
        d["_plans"] = {}
        return d

    def _build_graph(self, plan: Plan, x_in: torch.Tensor):  # pragma: no cover - overridden
        raise NotImplementedError

    def _named_params(self) -> "OrderedDict[str, nn.Parameter]":
        return OrderedDict(self.named_parameters())

    def _seg_heads_active(self) -> Optional[List[bool]]:
        return None

    def _plan_key_extra(self) -> Tuple:
        """Anything else a plan bakes in (ResidualUNet: BatchNorm2d train / eval mode)."""
        return ()

    def _get_plan(self, x: torch.Tensor, need_grad: bool) -> Plan:
        if not x.is_cuda:
            raise _lib.MtbcError("multi_task_breast_cancer_b200 runs on CUDA sm_100a only: move the model and the input to "
                                 "'cuda' (there is no CPU fallback)")
        if x.dtype != torch.float32 or x.dim() != 4:
            raise ValueError("expected a float32 (B, C, H, W) image batch")
        params = self._named_params()
        sig = tuple(p.data_ptr() for p in params.values())
        key = (tuple(x.shape), x.device.index, need_grad, sig, self.precision, self.deterministic) + self._plan_key_extra()
        plan = self._plans.get(key)
        if plan is None:
            # drop plans built for stale parameter storage (e.g. after .to())
            for k in [k for k in self._plans if k[3] != sig]:
                del self._plans[k]
            lib = _lib.load()
            _lib.check(lib.mtbc_device_check(), "device check")
            B, Cin, H, W = x.shape
            from . import plan as _plan_mod
            with torch.cuda.device(x.device):
                try:
                    plan = Plan(B, H, W, x.device, params, training=need_grad, precision=self.precision,
                                deterministic=self.deterministic)
                    # op CREATION consults the mode too (the halo conv drops its two-lane accumulation when the
                    # result has to be order independent), so it is set for the whole build, not just per launch
                    lib.mtbc_set_mode(plan.mode_flags)
                    plan.x_in = torch.zeros(B, Cin, H, W, dtype=torch.float32, device=x.device)
                    self._build_graph(plan, plan.x_in)
                    plan.finalize(self._seg_heads_active())
                finally:
                    lib.mtbc_set_mode(0)
                    _plan_mod._build_mode = 0
            self._plans[key] = plan
        return plan

    def _deliver_grads(self, plan: Plan):
        for name, p in plan.params.items():
            if not plan.has_grad[name] or not p.requires_grad:
                continue
            g = plan.grad_view[name]
            if p.grad is None:
                p.grad = g          # no copy: the optimizer reads the plan's buffer (valid until the next backward)
            elif p.grad.data_ptr() != g.data_ptr():
                p.grad.add_(g)

    def _run(self, x: torch.Tensor):
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        plan = self._get_plan(x, need_grad)
        with torch.cuda.device(x.device):
            if need_grad:
                outs = _PlanFunction.apply(self, plan, x, *plan.params.values())
            else:
                if x.data_ptr() != plan.x_in.data_ptr():
                    plan.x_in.copy_(x)
                plan.run_pack()
                plan.run_forward()
                # copies: the plan's output buffers are static and the next forward overwrites them; an inference
                # loop that keeps the logits of several batches (utils/models.py:309-386) must not see them change
                outs = tuple(t.clone() for t in plan.outputs_cls) + tuple(t.clone() for t in plan.outputs_seg)
        ncls = len(plan.outputs_cls)
        return list(outs[:ncls]), list(outs[ncls:])


# ======================================================================================================================
# parameter containers (names == reference state_dict keys)
# ======================================================================================================================
class _MonaiConvolution(nn.Sequential):
    def __init__(self, cin, cout, slope, bias, dropout):
        super().__init__()
        self.add_module("conv", nn.Conv2d(cin, cout, 3, 1, 1, bias=bias))
        adn = nn.Sequential()
        adn.add_module("N", nn.InstanceNorm2d(cout, affine=True))
        adn.add_module("D", nn.Dropout(dropout))
        adn.add_module("A", nn.LeakyReLU(negative_slope=slope, inplace=True))
        self.add_module("adn", adn)


class _MonaiTwoConv(nn.Sequential):
    def __init__(self, cin, cout, slope, bias, dropout):
        super().__init__()
        self.add_module("conv_0", _MonaiConvolution(cin, cout, slope, bias, dropout))
        self.add_module("conv_1", _MonaiConvolution(cout, cout, slope, bias, dropout))


class _MonaiDown(nn.Sequential):
    def __init__(self, cin, cout, slope, bias, dropout):
        super().__init__()
        self.add_module("max_pooling", nn.MaxPool2d(kernel_size=2))
        self.add_module("convs", _MonaiTwoConv(cin, cout, slope, bias, dropout))


class _MonaiUpCat(nn.Module):
    def __init__(self, cin, ccat, cout, slope, bias, dropout, halves=True):
        super().__init__()
        cup = cin // 2 if halves else cin
        up = nn.Sequential()
        up.add_module("deconv", nn.ConvTranspose2d(cin, cup, kernel_size=2, stride=2, bias=True))
        self.upsample = up
        self.convs = _MonaiTwoConv(ccat + cup, cout, slope, bias, dropout)


class MTUNetPlusPlus(_PlanModule):
    """Multi-task U-Net++ (reference MTUNetPlusPlus.py:11-136)."""

    def __init__(self, spatial_dims: int = 2, in_channels: int = 1, out_channels: int = 1, n_classes: int = 3,
                 features: Sequence[int] = (24, 48, 96, 192, 384, 24), deep_supervision: bool = False,
                 act: Union[str, tuple] = ("LeakyReLU", {"negative_slope": 0.1, "inplace": True}),
                 norm: Union[str, tuple] = ("instance", {"affine": True}), bias: bool = True,
                 dropout: Union[float, tuple] = 0.0, upsample: str = "deconv"):
        super().__init__()
        if spatial_dims != 2 or upsample != "deconv":
            raise NotImplementedError("only the reference configuration (2-D, deconv upsampling) is implemented")
        norm_name, norm_kw = norm if isinstance(norm, (tuple, list)) else (norm, {})
        act_name, act_kw = act if isinstance(act, (tuple, list)) else (act, {})
        if norm_name.lower() != "instance" or not norm_kw.get("affine", False) or act_name.lower() != "leakyrelu":
            raise NotImplementedError("only InstanceNorm(affine=True) + LeakyReLU blocks are implemented")
        if float(dropout if not isinstance(dropout, (tuple, list)) else dropout[0]) != 0.0:
            raise NotImplementedError("dropout must be 0.0 (the reference value)")
        if out_channels != 1:
            raise NotImplementedError("mask heads project to one region (reference: regions=1)")
        self._slope = float(act_kw.get("negative_slope", 0.01))
        self.deep_supervision = deep_supervision
        self.n_classes = 1 if n_classes == 2 else n_classes
        fea = tuple(features)
        assert len(fea) == 6
        a = (self._slope, bias, 0.0)
        self.conv_0_0 = _MonaiTwoConv(in_channels, fea[0], *a)
        self.conv_1_0 = _MonaiDown(fea[0], fea[1], *a)
        self.conv_2_0 = _MonaiDown(fea[1], fea[2], *a)
        self.conv_3_0 = _MonaiDown(fea[2], fea[3], *a)
        self.conv_4_0 = _MonaiDown(fea[3], fea[4], *a)
        self.upcat_0_1 = _MonaiUpCat(fea[1], fea[0], fea[0], *a, halves=False)
        self.upcat_1_1 = _MonaiUpCat(fea[2], fea[1], fea[1], *a)
        self.upcat_2_1 = _MonaiUpCat(fea[3], fea[2], fea[2], *a)
        self.upcat_3_1 = _MonaiUpCat(fea[4], fea[3], fea[3], *a)
        self.upcat_0_2 = _MonaiUpCat(fea[1], fea[0] * 2, fea[0], *a, halves=False)
        self.upcat_1_2 = _MonaiUpCat(fea[2], fea[1] * 2, fea[1], *a)
        self.upcat_2_2 = _MonaiUpCat(fea[3], fea[2] * 2, fea[2], *a)
        self.upcat_0_3 = _MonaiUpCat(fea[1], fea[0] * 3, fea[0], *a, halves=False)
        self.upcat_1_3 = _MonaiUpCat(fea[2], fea[1] * 3, fea[1], *a)
        self.upcat_0_4 = _MonaiUpCat(fea[1], fea[0] * 4, fea[5], *a, halves=False)
        self.final_conv_0_1 = nn.Conv2d(fea[0], out_channels, kernel_size=1)
        self.final_conv_0_2 = nn.Conv2d(fea[0], out_channels, kernel_size=1)
        self.final_conv_0_3 = nn.Conv2d(fea[0], out_channels, kernel_size=1)
        self.final_conv_0_4 = nn.Conv2d(fea[5], out_channels, kernel_size=1)
        self.process_level_3 = _MonaiDown(fea[3], fea[4], *a)
        self.classifier = nn.Sequential(
            _MonaiTwoConv(fea[4] * 3, 512, *a), nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(512, 256), nn.ReLU(),
            nn.Linear(in_features=256, out_features=self.n_classes))
        self._has_bias = bias
        self._init_runtime()

    # ---- graph -------------------------------------------------------------------------------------------------
    def _conv(self, plan, srcs, prefix, pool, first_input=None):
        w = f"{prefix}.conv.weight"
        b = f"{prefix}.conv.bias" if self._has_bias else None
        g, be = f"{prefix}.adn.N.weight", f"{prefix}.adn.N.bias"
        if first_input is not None:
            return plan.input_conv_in_act(first_input, w, b, g, be, self._slope, pool, prefix)
        return plan.conv_in_act(srcs, w, b, g, be, self._slope, pool, prefix)

    def _two(self, plan, srcs, prefix, pool=False, first_input=None):
        a1, _ = self._conv(plan, srcs, prefix + ".conv_0", False, first_input)
        return self._conv(plan, [a1], prefix + ".conv_1", pool)

    def _upcat(self, plan, low: PTensor, skips: List[PTensor], prefix, pool=False):
        up = plan.convT(low, f"{prefix}.upsample.deconv.weight", f"{prefix}.upsample.deconv.bias", 2, prefix + ".up")
        return self._two(plan, list(skips) + [up], prefix + ".convs", pool)

    def _build_graph(self, plan: Plan, x_in: torch.Tensor):
        B, Cin, H, W = x_in.shape
        if H % 16 or W % 16:
            raise ValueError("MTUNetPlusPlus needs H and W divisible by 16 (MTUNetPlusPlus.py:94-95)")
        if Cin > 4:
            raise NotImplementedError("first layer supports up to 4 input channels")
        x00, p00 = self._two(plan, None, "conv_0_0", pool=True, first_input=x_in)
        x10, p10 = self._two(plan, [p00], "conv_1_0.convs", pool=True)
        x01, _ = self._upcat(plan, x10, [x00], "upcat_0_1")
        x20, p20 = self._two(plan, [p10], "conv_2_0.convs", pool=True)
        x11, _ = self._upcat(plan, x20, [x10], "upcat_1_1")
        x02, _ = self._upcat(plan, x11, [x00, x01], "upcat_0_2")
        x30, p30 = self._two(plan, [p20], "conv_3_0.convs", pool=True)
        x21, _ = self._upcat(plan, x30, [x20], "upcat_2_1")
        x12, _ = self._upcat(plan, x21, [x10, x11], "upcat_1_2")
        x03, _ = self._upcat(plan, x12, [x00, x01, x02], "upcat_0_3")
        x40, _ = self._two(plan, [p30], "conv_4_0.convs")
        x31, p31 = self._upcat(plan, x40, [x30], "upcat_3_1", pool=True)
        x22, _ = self._upcat(plan, x31, [x20, x21], "upcat_2_2")
        x13, _ = self._upcat(plan, x22, [x10, x11, x12], "upcat_1_3")
        x04, _ = self._upcat(plan, x13, [x00, x01, x02, x03], "upcat_0_4")
        ds = self.deep_supervision
        # final_conv_0_1..3 are evaluated but dropped by the reference when deep_supervision=False
        # (MTUNetPlusPlus.py:120-123,131-134): skip the work, their grads stay None.
        plan.head1x1(x01, "final_conv_0_1.weight", "final_conv_0_1.bias", active=ds)
        plan.head1x1(x02, "final_conv_0_2.weight", "final_conv_0_2.bias", active=ds)
        plan.head1x1(x03, "final_conv_0_3.weight", "final_conv_0_3.bias", active=ds)
        plan.head1x1(x04, "final_conv_0_4.weight", "final_conv_0_4.bias", active=True)
        f0, _ = self._two(plan, [p30], "process_level_3.convs")
        f2, _ = self._two(plan, [p31], "process_level_3.convs")
        feat, _ = self._two(plan, [f0, x40, f2], "classifier.0")
        plan.gap_fc(feat, "classifier.3.weight", "classifier.3.bias", "classifier.5.weight", "classifier.5.bias")

    def forward(self, x: torch.Tensor):
        cls, seg = self._run(x)
        if self.deep_supervision:
            return cls, seg
        return cls[0], seg[-1]


# ----------------------------------------------------------------------------------------------------------------------
def conv1x1(in_channels, out_channels):
    return nn.Conv2d(in_channels=in_channels, out_channels=out_channels, kernel_size=(1, 1))


def conv3x3(in_channels, out_channels, stride=1, groups=1, dilation=1, bias=False):
    return nn.Conv2d(in_channels, out_channels, kernel_size=(3, 3), stride=stride, padding=dilation, groups=groups,
                     bias=bias, dilation=dilation)


class ConvInNormLeReLU(nn.Sequential):
    """Parameter container for conv3x3 -> InstanceNorm2d -> LeakyReLU (MTnnUNet.py:19-39)."""

    def __init__(self, in_channels, out_channels):
        super().__init__(OrderedDict([("Conv", conv3x3(in_channels, out_channels)),
                                      ("InNorm", nn.InstanceNorm2d(out_channels)),
                                      ("LeReLU", nn.LeakyReLU(inplace=True))]))


class LevelBlock(nn.Sequential):
    def __init__(self, in_channels, mid_channels, out_channels):
        super().__init__(OrderedDict([("ConvInNormLRelu1", ConvInNormLeReLU(in_channels, mid_channels)),
                                      ("ConvInNormLRelu2", ConvInNormLeReLU(mid_channels, out_channels))]))


def _kaiming_conv2d(module: nn.Module):
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, nonlinearity="leaky_relu")
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)


class _PlainUNetBase(_PlanModule):
    _slope = 0.01

    def _cil(self, plan, srcs, prefix, pool=False, first_input=None):
        w = f"{prefix}.Conv.weight"
        if first_input is not None:
            return plan.input_conv_in_act(first_input, w, None, None, None, self._slope, pool, prefix)
        return plan.conv_in_act(srcs, w, None, None, None, self._slope, pool, prefix)

    def _level(self, plan, srcs, prefix, pool=False, first_input=None):
        a1, _ = self._cil(plan, srcs, prefix + ".ConvInNormLRelu1", False, first_input)
        return self._cil(plan, [a1], prefix + ".ConvInNormLRelu2", pool)


class MTnnUNet(_PlainUNetBase):
    """Multi-task nnU-Net style network (reference MTnnUNet.py:64-183); always returns lists."""

    def __init__(self, sequences, regions, n_classes=3):
        super().__init__()
        if regions != 1:
            raise NotImplementedError("mask heads project to one region (reference: regions=1)")
        widths = [32, 64, 128, 256, 320]
        self.n_classes = 1 if n_classes == 2 else n_classes
        self.encoder1 = LevelBlock(sequences, widths[0], widths[0])
        self.encoder2 = LevelBlock(widths[0], widths[1], widths[1])
        self.encoder3 = LevelBlock(widths[1], widths[2], widths[2])
        self.encoder4 = LevelBlock(widths[2], widths[3], widths[3])
        self.encoder5 = LevelBlock(widths[3], widths[4], widths[4])
        self.bottleneck = LevelBlock(widths[4], widths[4], widths[4])
        self.decoder5 = LevelBlock(widths[4] + widths[4], widths[3], widths[3])
        self.decoder4 = LevelBlock(widths[3] + widths[3], widths[2], widths[2])
        self.decoder3 = LevelBlock(widths[2] + widths[2], widths[1], widths[1])
        self.decoder2 = LevelBlock(widths[1] + widths[1], widths[0], widths[0])
        self.decoder1 = LevelBlock(widths[0] + widths[0], widths[0], widths[0] // 2)
        self.upsample5 = nn.ConvTranspose2d(widths[4], widths[4], kernel_size=2, stride=2)
        self.upsample4 = nn.ConvTranspose2d(widths[3], widths[3], kernel_size=2, stride=2)
        self.upsample3 = nn.ConvTranspose2d(widths[2], widths[2], kernel_size=2, stride=2)
        self.upsample2 = nn.ConvTranspose2d(widths[1], widths[1], kernel_size=2, stride=2)
        self.upsample1 = nn.ConvTranspose2d(widths[0], widths[0], kernel_size=2, stride=2)
        self.downsample = nn.MaxPool2d(2, 2)
        self.output4 = nn.Sequential(nn.ConvTranspose2d(widths[2], widths[2], kernel_size=8, stride=8),
                                     conv1x1(widths[2], regions))
        self.output3 = nn.Sequential(nn.ConvTranspose2d(widths[1], widths[1], kernel_size=4, stride=4),
                                     conv1x1(widths[1], regions))
        self.output2 = nn.Sequential(nn.ConvTranspose2d(widths[0], widths[0], kernel_size=2, stride=2),
                                     conv1x1(widths[0], regions))
        self.output1 = conv1x1(widths[0] // 2, regions)
        self.weights_initialization()
        # created after the kaiming pass, exactly like the reference (MTnnUNet.py:120-132): default PyTorch init
        self.process_encoder_5 = ConvInNormLeReLU(widths[4], widths[4])
        self.process_decoder_5 = ConvInNormLeReLU(widths[3], widths[4])
        self.classifier = nn.Sequential(ConvInNormLeReLU(widths[4] * 3, 512), nn.AdaptiveAvgPool2d(1), nn.Flatten(),
                                        nn.Linear(512, 256), nn.ReLU(),
                                        nn.Linear(in_features=256, out_features=self.n_classes))
        self._init_runtime()

    def weights_initialization(self):
        _kaiming_conv2d(self)

    def _build_graph(self, plan: Plan, x_in: torch.Tensor):
        B, Cin, H, W = x_in.shape
        if H % 32 or W % 32:
            raise ValueError("MTnnUNet needs H and W divisible by 32 (five 2x2 poolings)")
        e1, p1 = self._level(plan, None, "encoder1", pool=True, first_input=x_in)
        e2, p2 = self._level(plan, [p1], "encoder2", pool=True)
        e3, p3 = self._level(plan, [p2], "encoder3", pool=True)
        e4, p4 = self._level(plan, [p3], "encoder4", pool=True)
        e5, p5 = self._level(plan, [p4], "encoder5", pool=True)
        bott, _ = self._level(plan, [p5], "bottleneck")
        # upsample5(bottleneck) is evaluated twice by the reference (MTnnUNet.py:160,174); both calls see the same
        # input and weights, so it is computed once and its two consumers' gradients are summed.
        up5 = plan.convT(bott, "upsample5.weight", "upsample5.bias", 2, "up5")
        d5, _ = self._level(plan, [e5, up5], "decoder5")
        up4 = plan.convT(d5, "upsample4.weight", "upsample4.bias", 2, "up4")
        d4, _ = self._level(plan, [e4, up4], "decoder4")
        up3 = plan.convT(d4, "upsample3.weight", "upsample3.bias", 2, "up3")
        d3, _ = self._level(plan, [e3, up3], "decoder3")
        up2 = plan.convT(d3, "upsample2.weight", "upsample2.bias", 2, "up2")
        d2, _ = self._level(plan, [e2, up2], "decoder2")
        up1 = plan.convT(d2, "upsample1.weight", "upsample1.bias", 2, "up1")
        d1, _ = self._level(plan, [e1, up1], "decoder1")
        pe5, _ = self._cil(plan, [e5], "process_encoder_5")
        pd5, _ = self._cil(plan, [d5], "process_decoder_5")
        feat, _ = self._cil(plan, [pe5, up5, pd5], "classifier.0")
        plan.gap_fc(feat, "classifier.3.weight", "classifier.3.bias", "classifier.5.weight", "classifier.5.bias")
        plan.dshead(d4, "output4.0.weight", "output4.0.bias", "output4.1.weight", "output4.1.bias", 8)
        plan.dshead(d3, "output3.0.weight", "output3.0.bias", "output3.1.weight", "output3.1.bias", 4)
        plan.dshead(d2, "output2.0.weight", "output2.0.bias", "output2.1.weight", "output2.1.bias", 2)
        plan.head1x1(d1, "output1.weight", "output1.bias")

    def forward(self, x):
        cls, seg = self._run(x)
        return cls, seg


class Multi_BTS_UNet(_PlainUNetBase):
    """Multi-task BTS U-Net (reference Multi_BTS_UNet.py:64-176); 128x128 inputs only, like the reference."""

    name = "Full-Scale-Bridge BTS U-Net"

    def __init__(self, sequences, regions, n_classes, width, deep_supervision):
        super().__init__()
        if regions != 1:
            raise NotImplementedError("mask heads project to one region (reference: regions=1)")
        self.deep_supervision = deep_supervision
        widths = [width * 2 ** i for i in range(4)]
        self.n_classes = 1 if n_classes == 2 else n_classes
        self.encoder1 = LevelBlock(sequences, widths[0] // 2, widths[0])
        self.encoder2 = LevelBlock(widths[0], widths[1] // 2, widths[1])
        self.encoder3 = LevelBlock(widths[1], widths[2] // 2, widths[2])
        self.encoder4 = LevelBlock(widths[2], widths[3] // 2, widths[3])
        self.bottleneck = LevelBlock(widths[3], widths[3], widths[3])
        self.bottleneck2 = ConvInNormLeReLU(widths[3] * 2, widths[2])
        self.decoder3 = LevelBlock(widths[2] * 2, widths[2], widths[1])
        self.decoder2 = LevelBlock(widths[1] * 2, widths[1], widths[0])
        self.decoder1 = LevelBlock(widths[0] * 2, widths[0], widths[0] // 2)
        self.upsample = nn.Upsample(scale_factor=2, mode="nearest")
        self.downsample = nn.MaxPool2d(2, 2)
        self.softmax = nn.Softmax(dim=1)
        self.process_bottleneck2 = ConvInNormLeReLU(widths[2], widths[3])
        self.process_features_map = ConvInNormLeReLU(widths[3] * 3, widths[3])
        self.classifier = nn.Sequential(nn.Flatten(), nn.Linear(widths[3] * 16 * 16, 256), nn.ReLU(),
                                        nn.Linear(256, self.n_classes))
        if self.deep_supervision:
            self.output3 = nn.Sequential(nn.ConvTranspose2d(widths[1], widths[1], kernel_size=4, stride=4),
                                         conv1x1(widths[1], regions))
            self.output2 = nn.Sequential(nn.ConvTranspose2d(widths[0], widths[0], kernel_size=2, stride=2),
                                         conv1x1(widths[0], regions))
        self.output1 = conv1x1(widths[0] // 2, regions)
        self.weights_initialization()
        self._init_runtime()

    def weights_initialization(self):
        _kaiming_conv2d(self)

    def _build_graph(self, plan: Plan, x_in: torch.Tensor):
        B, Cin, H, W = x_in.shape
        if (H, W) != (128, 128):
            raise ValueError("Multi_BTS_UNet only accepts 128x128 inputs: its classifier is Linear(widths[3]*16*16, 256) "
                             "(Multi_BTS_UNet.py:110)")
        e1, p1 = self._level(plan, None, "encoder1", pool=True, first_input=x_in)
        e2, p2 = self._level(plan, [p1], "encoder2", pool=True)
        e3, p3 = self._level(plan, [p2], "encoder3", pool=True)
        e4, _ = self._level(plan, [p3], "encoder4")
        bott, _ = self._level(plan, [e4], "bottleneck")
        bott2, _ = self._cil(plan, [e4, bott], "bottleneck2")
        up3 = plan.upsample2(bott2, "up3")
        d3, _ = self._level(plan, [e3, up3], "decoder3")
        up2 = plan.upsample2(d3, "up2")
        d2, _ = self._level(plan, [e2, up2], "decoder2")
        up1 = plan.upsample2(d2, "up1")
        d1, _ = self._level(plan, [e1, up1], "decoder1")
        pb2, _ = self._cil(plan, [bott2], "process_bottleneck2")
        feat, _ = self._cil(plan, [e4, bott, pb2], "process_features_map")
        plan.flat_fc(feat, "classifier.1.weight", "classifier.1.bias", "classifier.3.weight", "classifier.3.bias")
        if self.deep_supervision:
            plan.dshead(d3, "output3.0.weight", "output3.0.bias", "output3.1.weight", "output3.1.bias", 4)
            plan.dshead(d2, "output2.0.weight", "output2.0.bias", "output2.1.weight", "output2.1.bias", 2)
        plan.head1x1(d1, "output1.weight", "output1.bias")

    def forward(self, x):
        cls, seg = self._run(x)
        if self.deep_supervision:
            return cls, seg
        return cls[0], seg[-1]


# ======================================================================================================================
# single-task siblings sharing the same kernels: segmentation-only graphs
# ======================================================================================================================
class nnUNet2021(_PlainUNetBase):
    """Segmentation nnU-Net (reference src/models/segmentation/nnUNet.py:64-162): MTnnUNet's encoder / decoder / four
    mask heads without the classification branch; returns [output4, output3, output2, output1]."""

    name = "nn-UNet2021"

    def __init__(self, sequences, regions):
        super().__init__()
        if regions != 1:
            raise NotImplementedError("mask heads project to one region (reference: regions=1)")
        widths = [32, 64, 128, 256, 320]
        self.encoder1 = LevelBlock(sequences, widths[0], widths[0])
        self.encoder2 = LevelBlock(widths[0], widths[1], widths[1])
        self.encoder3 = LevelBlock(widths[1], widths[2], widths[2])
        self.encoder4 = LevelBlock(widths[2], widths[3], widths[3])
        self.encoder5 = LevelBlock(widths[3], widths[4], widths[4])
        self.bottleneck = LevelBlock(widths[4], widths[4], widths[4])
        self.decoder5 = LevelBlock(widths[4] + widths[4], widths[3], widths[3])
        self.decoder4 = LevelBlock(widths[3] + widths[3], widths[2], widths[2])
        self.decoder3 = LevelBlock(widths[2] + widths[2], widths[1], widths[1])
        self.decoder2 = LevelBlock(widths[1] + widths[1], widths[0], widths[0])
        self.decoder1 = LevelBlock(widths[0] + widths[0], widths[0], widths[0] // 2)
        self.upsample5 = nn.ConvTranspose2d(widths[4], widths[4], kernel_size=2, stride=2)
        self.upsample4 = nn.ConvTranspose2d(widths[3], widths[3], kernel_size=2, stride=2)
        self.upsample3 = nn.ConvTranspose2d(widths[2], widths[2], kernel_size=2, stride=2)
        self.upsample2 = nn.ConvTranspose2d(widths[1], widths[1], kernel_size=2, stride=2)
        self.upsample1 = nn.ConvTranspose2d(widths[0], widths[0], kernel_size=2, stride=2)
        self.downsample = nn.MaxPool2d(2, 2)
        self.output4 = nn.Sequential(nn.ConvTranspose2d(widths[2], widths[2], kernel_size=8, stride=8),
                                     conv1x1(widths[2], regions))
        self.output3 = nn.Sequential(nn.ConvTranspose2d(widths[1], widths[1], kernel_size=4, stride=4),
                                     conv1x1(widths[1], regions))
        self.output2 = nn.Sequential(nn.ConvTranspose2d(widths[0], widths[0], kernel_size=2, stride=2),
                                     conv1x1(widths[0], regions))
        self.output1 = conv1x1(widths[0] // 2, regions)
        self.weights_initialization()
        self._init_runtime()

    def weights_initialization(self):
        _kaiming_conv2d(self)

    def _build_graph(self, plan: Plan, x_in: torch.Tensor):
        B, Cin, H, W = x_in.shape
        if H % 32 or W % 32:
            raise ValueError("nnUNet2021 needs H and W divisible by 32 (five 2x2 poolings)")
        e1, p1 = self._level(plan, None, "encoder1", pool=True, first_input=x_in)
        e2, p2 = self._level(plan, [p1], "encoder2", pool=True)
        e3, p3 = self._level(plan, [p2], "encoder3", pool=True)
        e4, p4 = self._level(plan, [p3], "encoder4", pool=True)
        e5, p5 = self._level(plan, [p4], "encoder5", pool=True)
        bott, _ = self._level(plan, [p5], "bottleneck")
        up5 = plan.convT(bott, "upsample5.weight", "upsample5.bias", 2, "up5")
        d5, _ = self._level(plan, [e5, up5], "decoder5")
        up4 = plan.convT(d5, "upsample4.weight", "upsample4.bias", 2, "up4")
        d4, _ = self._level(plan, [e4, up4], "decoder4")
        up3 = plan.convT(d4, "upsample3.weight", "upsample3.bias", 2, "up3")
        d3, _ = self._level(plan, [e3, up3], "decoder3")
        up2 = plan.convT(d3, "upsample2.weight", "upsample2.bias", 2, "up2")
        d2, _ = self._level(plan, [e2, up2], "decoder2")
        up1 = plan.convT(d2, "upsample1.weight", "upsample1.bias", 2, "up1")
        d1, _ = self._level(plan, [e1, up1], "decoder1")
        plan.dshead(d4, "output4.0.weight", "output4.0.bias", "output4.1.weight", "output4.1.bias", 8)
        plan.dshead(d3, "output3.0.weight", "output3.0.bias", "output3.1.weight", "output3.1.bias", 4)
        plan.dshead(d2, "output2.0.weight", "output2.0.bias", "output2.1.weight", "output2.1.bias", 2)
        plan.head1x1(d1, "output1.weight", "output1.bias")

    def forward(self, x):
        _, seg = self._run(x)
        return seg


class BTSUNet(_PlainUNetBase):
    """Segmentation BTS U-Net (reference src/models/segmentation/BTS_UNet.py:64-152): Multi_BTS_UNet without the
    classification branch (so any H, W divisible by 8); a list under deep supervision, else the full-decoder logits."""

    name = "BTS U-Net"

    def __init__(self, sequences, regions, width, deep_supervision):
        super().__init__()
        if regions != 1:
            raise NotImplementedError("mask heads project to one region (reference: regions=1)")
        self.deep_supervision = deep_supervision
        widths = [width * 2 ** i for i in range(4)]
        self.encoder1 = LevelBlock(sequences, widths[0] // 2, widths[0])
        self.encoder2 = LevelBlock(widths[0], widths[1] // 2, widths[1])
        self.encoder3 = LevelBlock(widths[1], widths[2] // 2, widths[2])
        self.encoder4 = LevelBlock(widths[2], widths[3] // 2, widths[3])
        self.bottleneck = LevelBlock(widths[3], widths[3], widths[3])
        self.bottleneck2 = ConvInNormLeReLU(widths[3] * 2, widths[2])
        self.decoder3 = LevelBlock(widths[2] * 2, widths[2], widths[1])
        self.decoder2 = LevelBlock(widths[1] * 2, widths[1], widths[0])
        self.decoder1 = LevelBlock(widths[0] * 2, widths[0], widths[0] // 2)
        self.upsample = nn.Upsample(scale_factor=2, mode="nearest")
        self.downsample = nn.MaxPool2d(2, 2)
        if self.deep_supervision:
            self.output3 = nn.Sequential(nn.ConvTranspose2d(widths[1], widths[1], kernel_size=4, stride=4),
                                         conv1x1(widths[1], regions))
            self.output2 = nn.Sequential(nn.ConvTranspose2d(widths[0], widths[0], kernel_size=2, stride=2),
                                         conv1x1(widths[0], regions))
        self.output1 = conv1x1(widths[0] // 2, regions)
        self.weights_initialization()
        self._init_runtime()

    def weights_initialization(self):
        _kaiming_conv2d(self)

    def _build_graph(self, plan: Plan, x_in: torch.Tensor):
        B, Cin, H, W = x_in.shape
        if H % 8 or W % 8:
            raise ValueError("BTSUNet needs H and W divisible by 8 (three 2x2 poolings)")
        for k in range(4):   # plane extents the conv kernels tile: multiples of 16, or 1 / 2 / 4 / 8
            for v in (H >> k, W >> k):
                if not (v % 16 == 0 or v in (1, 2, 4, 8)):
                    raise ValueError(f"BTSUNet on sm_100a: the {v}-pixel planes of a {H}x{W} input are not tileable "
                                     "(every level's extent must be a multiple of 16, or 8 / 4 / 2 / 1)")
        e1, p1 = self._level(plan, None, "encoder1", pool=True, first_input=x_in)
        e2, p2 = self._level(plan, [p1], "encoder2", pool=True)
        e3, p3 = self._level(plan, [p2], "encoder3", pool=True)
        e4, _ = self._level(plan, [p3], "encoder4")
        bott, _ = self._level(plan, [e4], "bottleneck")
        bott2, _ = self._cil(plan, [e4, bott], "bottleneck2")
        up3 = plan.upsample2(bott2, "up3")
        d3, _ = self._level(plan, [e3, up3], "decoder3")
        up2 = plan.upsample2(d3, "up2")
        d2, _ = self._level(plan, [e2, up2], "decoder2")
        up1 = plan.upsample2(d2, "up1")
        d1, _ = self._level(plan, [e1, up1], "decoder1")
        if self.deep_supervision:
            plan.dshead(d3, "output3.0.weight", "output3.0.bias", "output3.1.weight", "output3.1.bias", 4)
            plan.dshead(d2, "output2.0.weight", "output2.0.bias", "output2.1.weight", "output2.1.bias", 2)
        plan.head1x1(d1, "output1.weight", "output1.bias")

    def forward(self, x):
        _, seg = self._run(x)
        return seg if self.deep_supervision else seg[-1]


# ----------------------------------------------------------------------------------------------------------------------
class _RUInBlock(nn.Module):
    """Parameter container of ResidualUNet.py:14-71 (registration order = the reference's, so that one seed gives one
    initialisation)."""

    def __init__(self, channel_in, channel_out):
        super().__init__()
        self.conv1 = nn.Conv2d(channel_in, channel_out, kernel_size=3, padding=1)
        self.bn1 = nn.BatchNorm2d(channel_out)
        self.conv2 = nn.Conv2d(channel_out, channel_out, kernel_size=3, padding=1)
        self.conv3 = nn.Conv2d(channel_in, channel_out, kernel_size=3, padding=1)
        self.bn3 = nn.BatchNorm2d(channel_out)


class _RUResBlock(nn.Module):
    """ResidualUNet.py:74-157: bn1 -> lrelu -> dropout -> conv1 (stride 2 when downsampling) -> bn2 -> lrelu -> dropout
    -> conv2, plus the residual conv3 -> bn3."""

    def __init__(self, channel_in, downsample=False):
        super().__init__()
        cout, st = (2 * channel_in, 2) if downsample else (channel_in, 1)
        self.stride = st
        self.bn1 = nn.BatchNorm2d(channel_in)
        self.conv1 = nn.Conv2d(channel_in, cout, kernel_size=3, stride=st, padding=1)
        self.bn2 = nn.BatchNorm2d(cout)
        self.conv2 = nn.Conv2d(cout, cout, kernel_size=3, padding=1)
        self.conv3 = nn.Conv2d(channel_in, cout, kernel_size=3, stride=st, padding=1)
        self.bn3 = nn.BatchNorm2d(cout)


class _RUEncoder(nn.Module):
    def __init__(self, bf):
        super().__init__()
        self.down_block2 = _RUResBlock(bf, True)
        self.down_block3 = _RUResBlock(bf * 2, True)
        self.down_block4 = _RUResBlock(bf * 4, True)


class _RUDecoder(nn.Module):
    """ResidualUNet.py:195-269.  The 1x1 convs conv3 / conv2 / conv1 belong to the skip-connected variant (`seg_path`,
    :298-339); ResidualUNet.forward (:356-362) never calls them, so they exist in the state_dict and get no gradient."""

    def __init__(self, bf):
        super().__init__()
        self.upsample3 = nn.ConvTranspose2d(bf * 8, bf * 4, kernel_size=2, stride=2)
        self.conv3 = nn.Conv2d(bf * 8, bf * 4, kernel_size=1)
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


class ResidualUNet(_PlanModule):
    """Residual U-Net (reference src/models/segmentation/ResidualUNet.py:342-362): in_block -> three stride-2 residual
    blocks -> three (ConvTranspose2d k2 s2 -> residual block) stages WITHOUT skip connections -> 1x1 conv.  The only
    model of the repo with BatchNorm2d (batch statistics + running statistics in training, running statistics in eval),
    stride-2 3x3 convs and F.dropout(p=0.2) -- which the reference calls with its default training=True, so the masks
    are drawn in eval mode as well.  Returns the mask logits (B, regions, H, W)."""

    name = "Residual UNet"
    _slope = 0.01          # F.leaky_relu default
    _p_drop = 0.2

    def __init__(self, sequences=1, regions=1, width=24):
        super().__init__()
        if regions != 1:
            raise NotImplementedError("mask heads project to one region (reference: regions=1)")
        if width % 8:
            raise NotImplementedError("width must be a multiple of 8 (16-byte channel vectors)")
        self.in_block = _RUInBlock(sequences, width)
        self.encoder = _RUEncoder(width)
        self.decoder = _RUDecoder(width)
        self.out_block = _RUOut(width, regions)
        self._external_dropout = False      # parity tests: masks are handed in (plan.dropout_masks) instead of drawn
        self._init_runtime()

    def _plan_key_extra(self):
        return (bool(self.training), bool(self._external_dropout))

    def _bn(self, plan, x, mod, prefix, slope, p, name, stats=None):
        return plan.bn_act(x, mod, prefix, slope, p, name, stats=stats, training=self.training)

    def _res_block(self, plan, x, blk: _RUResBlock, prefix):
        a = self._bn(plan, x, blk.bn1, prefix + ".bn1", self._slope, self._p_drop, prefix + ".a1")
        y, s, q = plan.conv_plain([a], prefix + ".conv1.weight", prefix + ".conv1.bias", prefix + ".conv1",
                                  stride=blk.stride, stats=True)
        b = self._bn(plan, y, blk.bn2, prefix + ".bn2", self._slope, self._p_drop, prefix + ".a2", stats=(s, q))
        path, _, _ = plan.conv_plain([b], prefix + ".conv2.weight", prefix + ".conv2.bias", prefix + ".conv2")
        y3, s3, q3 = plan.conv_plain([x], prefix + ".conv3.weight", prefix + ".conv3.bias", prefix + ".conv3",
                                     stride=blk.stride, stats=True)
        res = self._bn(plan, y3, blk.bn3, prefix + ".bn3", 1.0, 0.0, prefix + ".res", stats=(s3, q3))
        return plan.add(path, res, prefix + ".out")

    def _build_graph(self, plan: Plan, x_in: torch.Tensor):
        B, Cin, H, W = x_in.shape
        if H % 8 or W % 8 or (H * W) % 128 or Cin > 4:
            raise ValueError("ResidualUNet needs H and W divisible by 8 (three stride-2 convs), H * W a multiple of 128 "
                             "and at most 4 input channels")
        for k in range(4):   # plane extents the conv kernels tile: multiples of 16, or 1 / 2 / 4 / 8
            for v in (H >> k, W >> k):
                if not (v % 16 == 0 or v in (1, 2, 4, 8)):
                    raise ValueError(f"ResidualUNet on sm_100a: the {v}-pixel planes of a {H}x{W} input are not tileable "
                                     "(every level's extent must be a multiple of 16, or 8 / 4 / 2 / 1)")
        plan.dropout_external = bool(self._external_dropout)
        ib = self.in_block
        y1, s1, q1 = plan.conv_plain(None, "in_block.conv1.weight", "in_block.conv1.bias", "in_block.conv1", stats=True,
                                     first_input=x_in)
        a1 = self._bn(plan, y1, ib.bn1, "in_block.bn1", self._slope, self._p_drop, "in_block.a1", stats=(s1, q1))
        path, _, _ = plan.conv_plain([a1], "in_block.conv2.weight", "in_block.conv2.bias", "in_block.conv2")
        y3, s3, q3 = plan.conv_plain(None, "in_block.conv3.weight", "in_block.conv3.bias", "in_block.conv3", stats=True,
                                     first_input=x_in)
        res = self._bn(plan, y3, ib.bn3, "in_block.bn3", 1.0, 0.0, "in_block.res", stats=(s3, q3))
        l1 = plan.add(path, res, "down_level1")
        l2 = self._res_block(plan, l1, self.encoder.down_block2, "encoder.down_block2")
        l3 = self._res_block(plan, l2, self.encoder.down_block3, "encoder.down_block3")
        codes = self._res_block(plan, l3, self.encoder.down_block4, "encoder.down_block4")
        up3 = plan.convT(codes, "decoder.upsample3.weight", "decoder.upsample3.bias", 2, "up3")
        u3 = self._res_block(plan, up3, self.decoder.up_block3, "decoder.up_block3")
        up2 = plan.convT(u3, "decoder.upsample2.weight", "decoder.upsample2.bias", 2, "up2")
        u2 = self._res_block(plan, up2, self.decoder.up_block2, "decoder.up_block2")
        up1 = plan.convT(u2, "decoder.upsample1.weight", "decoder.upsample1.bias", 2, "up1")
        u1 = self._res_block(plan, up1, self.decoder.up_block1, "decoder.up_block1")
        plan.head1x1(u1, "out_block.conv.weight", "out_block.conv.bias")

    def forward(self, x):
        _, seg = self._run(x)
        return seg[-1]


# ======================================================================================================================
# single-task siblings: classification-only graphs (encoder + class branch of the multi-task parents)
# ======================================================================================================================
class UNetPlusPlusClassifier(MTUNetPlusPlus):
    """Classification U-Net++ (reference src/models/classification/UnetPlusPlus_Classifier.py:20-147): the encoder,
    upcat_3_1 and the class branch of MTUNetPlusPlus; returns the raw class logits (the reference's softmax is
    commented out, :142-143)."""

    def __init__(self, spatial_dims: int = 2, in_channels: int = 1, n_classes: int = 3,
                 features: Sequence[int] = (24, 48, 96, 192, 384, 24),
                 act: Union[str, tuple] = ("LeakyReLU", {"negative_slope": 0.1, "inplace": True}),
                 norm: Union[str, tuple] = ("instance", {"affine": True}), bias: bool = True,
                 dropout: Union[float, tuple] = 0.0, upsample: str = "deconv"):
        nn.Module.__init__(self)
        if spatial_dims != 2 or upsample != "deconv":
            raise NotImplementedError("only the reference configuration (2-D, deconv upsampling) is implemented")
        norm_name, norm_kw = norm if isinstance(norm, (tuple, list)) else (norm, {})
        act_name, act_kw = act if isinstance(act, (tuple, list)) else (act, {})
        if norm_name.lower() != "instance" or not norm_kw.get("affine", False) or act_name.lower() != "leakyrelu":
            raise NotImplementedError("only InstanceNorm(affine=True) + LeakyReLU blocks are implemented")
        if float(dropout if not isinstance(dropout, (tuple, list)) else dropout[0]) != 0.0:
            raise NotImplementedError("dropout must be 0.0 (the reference value)")
        self._slope = float(act_kw.get("negative_slope", 0.01))
        self.n_classes = 1 if n_classes == 2 else n_classes
        fea = tuple(features)
        assert len(fea) == 6
        a = (self._slope, bias, 0.0)
        self.conv_0_0 = _MonaiTwoConv(in_channels, fea[0], *a)
        self.conv_1_0 = _MonaiDown(fea[0], fea[1], *a)
        self.conv_2_0 = _MonaiDown(fea[1], fea[2], *a)
        self.conv_3_0 = _MonaiDown(fea[2], fea[3], *a)
        self.conv_4_0 = _MonaiDown(fea[3], fea[4], *a)
        self.upcat_3_1 = _MonaiUpCat(fea[4], fea[3], fea[3], *a)
        self.softmax = nn.Softmax(dim=1)
        self.process_level_3 = _MonaiDown(fea[3], fea[4], *a)
        self.classifier = nn.Sequential(
            _MonaiTwoConv(fea[4] * 3, 512, *a), nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(512, 256), nn.ReLU(),
            nn.Linear(in_features=256, out_features=self.n_classes))
        self._has_bias = bias
        self._init_runtime()

    def _build_graph(self, plan: Plan, x_in: torch.Tensor):
        B, Cin, H, W = x_in.shape
        if H % 16 or W % 16:
            raise ValueError("UNetPlusPlusClassifier needs H and W divisible by 16 (four 2x2 poolings)")
        if Cin > 4:
            raise NotImplementedError("first layer supports up to 4 input channels")
        _, p00 = self._two(plan, None, "conv_0_0", pool=True, first_input=x_in)
        _, p10 = self._two(plan, [p00], "conv_1_0.convs", pool=True)
        _, p20 = self._two(plan, [p10], "conv_2_0.convs", pool=True)
        x30, p30 = self._two(plan, [p20], "conv_3_0.convs", pool=True)
        x40, _ = self._two(plan, [p30], "conv_4_0.convs")
        _, p31 = self._upcat(plan, x40, [x30], "upcat_3_1", pool=True)
        f0, _ = self._two(plan, [p30], "process_level_3.convs")
        f2, _ = self._two(plan, [p31], "process_level_3.convs")
        feat, _ = self._two(plan, [f0, x40, f2], "classifier.0")
        plan.gap_fc(feat, "classifier.3.weight", "classifier.3.bias", "classifier.5.weight", "classifier.5.bias")

    def forward(self, x: torch.Tensor):
        cls, _ = self._run(x)
        return cls[0]


class nnUNetClassifier(_PlainUNetBase):
    """Classification nnU-Net (reference src/models/classification/nnUNet_classifier.py:73-167): encoder, bottleneck,
    decoder5 and the class branch of MTnnUNet.  Like the reference it still OWNS decoder4..decoder1 (never evaluated;
    their gradients stay None) and returns softmax probabilities when n_classes > 2 (:165-166)."""

    name = "nn-UNet2021"

    def __init__(self, sequences, n_classes=3):
        super().__init__()
        widths = [32, 64, 128, 256, 320]
        self.n_classes = 1 if n_classes == 2 else n_classes
        self.encoder1 = LevelBlock(sequences, widths[0], widths[0])
        self.encoder2 = LevelBlock(widths[0], widths[1], widths[1])
        self.encoder3 = LevelBlock(widths[1], widths[2], widths[2])
        self.encoder4 = LevelBlock(widths[2], widths[3], widths[3])
        self.encoder5 = LevelBlock(widths[3], widths[4], widths[4])
        self.bottleneck = LevelBlock(widths[4], widths[4], widths[4])
        self.decoder5 = LevelBlock(widths[4] + widths[4], widths[3], widths[3])
        self.decoder4 = LevelBlock(widths[3] + widths[3], widths[2], widths[2])
        self.decoder3 = LevelBlock(widths[2] + widths[2], widths[1], widths[1])
        self.decoder2 = LevelBlock(widths[1] + widths[1], widths[0], widths[0])
        self.decoder1 = LevelBlock(widths[0] + widths[0], widths[0], widths[0] // 2)
        self.upsample5 = nn.ConvTranspose2d(widths[4], widths[4], kernel_size=2, stride=2)
        self.downsample = nn.MaxPool2d(2, 2)
        self.weights_initialization()
        # created after the kaiming pass, as in the reference (:107-120): PyTorch default init
        self.softmax = nn.Softmax(dim=1)
        self.process_encoder_5 = ConvInNormLeReLU(widths[4], widths[4])
        self.process_decoder_5 = ConvInNormLeReLU(widths[3], widths[4])
        self.classifier = nn.Sequential(ConvInNormLeReLU(widths[4] * 3, 512), nn.AdaptiveAvgPool2d(1), nn.Flatten(),
                                        nn.Linear(512, 256), nn.ReLU(),
                                        nn.Linear(in_features=256, out_features=self.n_classes))
        self._init_runtime()

    def weights_initialization(self):
        _kaiming_conv2d(self)

    def _build_graph(self, plan: Plan, x_in: torch.Tensor):
        B, Cin, H, W = x_in.shape
        if H % 32 or W % 32:
            raise ValueError("nnUNetClassifier needs H and W divisible by 32 (five 2x2 poolings)")
        _, p1 = self._level(plan, None, "encoder1", pool=True, first_input=x_in)
        _, p2 = self._level(plan, [p1], "encoder2", pool=True)
        _, p3 = self._level(plan, [p2], "encoder3", pool=True)
        _, p4 = self._level(plan, [p3], "encoder4", pool=True)
        e5, p5 = self._level(plan, [p4], "encoder5", pool=True)
        bott, _ = self._level(plan, [p5], "bottleneck")
        # upsample5(bottleneck) appears twice (:155,163): one evaluation, two consumers (as in MTnnUNet)
        up5 = plan.convT(bott, "upsample5.weight", "upsample5.bias", 2, "up5")
        d5, _ = self._level(plan, [e5, up5], "decoder5")
        pe5, _ = self._cil(plan, [e5], "process_encoder_5")
        pd5, _ = self._cil(plan, [d5], "process_decoder_5")
        feat, _ = self._cil(plan, [pe5, up5, pd5], "classifier.0")
        plan.gap_fc(feat, "classifier.3.weight", "classifier.3.bias", "classifier.5.weight", "classifier.5.bias")
        if self.n_classes > 2:
            plan.softmax_cls()

    def forward(self, x):
        cls, _ = self._run(x)
        return cls[0]


class BTSUNetClassifier(_PlainUNetBase):
    """Classification BTS U-Net (reference src/models/classification/BTS_UNET_classifier.py:55-116): four pooled
    LevelBlocks + a bottleneck LevelBlock in one nn.Sequential `encoder`, then Flatten -> Linear(8*width*8*8, 256) ->
    ReLU -> Linear; 128x128 inputs only (the Linear hard-codes the 8x8 bottleneck, :92)."""

    name = "BTS U-Net Classifier"

    def __init__(self, sequences, classes, width, deep_supervision=False):
        super().__init__()
        self.deep_supervision = deep_supervision
        widths = [width * 2 ** i for i in range(4)]
        self.classes = 1 if classes == 2 else classes
        self.encoder = nn.Sequential(
            LevelBlock(sequences, widths[0] // 2, widths[0]), nn.MaxPool2d(2, 2),
            LevelBlock(widths[0], widths[1] // 2, widths[1]), nn.MaxPool2d(2, 2),
            LevelBlock(widths[1], widths[2] // 2, widths[2]), nn.MaxPool2d(2, 2),
            LevelBlock(widths[2], widths[3] // 2, widths[3]), nn.MaxPool2d(2, 2),
            LevelBlock(widths[3], widths[3], widths[3]))
        self.classifier = nn.Sequential(nn.Flatten(), nn.Linear(widths[3] * 8 * 8, 256), nn.ReLU(),
                                        nn.Linear(256, self.classes))
        self.weights_initialization()
        self._init_runtime()

    def weights_initialization(self):
        _kaiming_conv2d(self)

    def _build_graph(self, plan: Plan, x_in: torch.Tensor):
        B, Cin, H, W = x_in.shape
        if (H, W) != (128, 128):
            raise ValueError("BTSUNetClassifier only accepts 128x128 inputs: its classifier is Linear(widths[3]*8*8, "
                             "256) (BTS_UNET_classifier.py:92)")
        _, p = self._level(plan, None, "encoder.0", pool=True, first_input=x_in)
        _, p = self._level(plan, [p], "encoder.2", pool=True)
        _, p = self._level(plan, [p], "encoder.4", pool=True)
        _, p = self._level(plan, [p], "encoder.6", pool=True)
        feat, _ = self._level(plan, [p], "encoder.8")
        plan.flat_fc(feat, "classifier.1.weight", "classifier.1.bias", "classifier.3.weight", "classifier.3.bias")

    def forward(self, x):
        cls, _ = self._run(x)
        return cls[0]


def init_classification_model(architecture: str, sequences: int = 1, n_classes: int = 1, width: int = 48,
                              save_folder=None) -> nn.Module:
    """String-keyed factory with the reference's signature (src/utils/experiment_init.py:86-127)."""
    if architecture == "BTSUNetClassifier":
        model = BTSUNetClassifier(sequences=sequences, classes=n_classes, width=width)
    elif architecture == "UNetPlusPlusClassifier":
        model = UNetPlusPlusClassifier(spatial_dims=2, in_channels=sequences, n_classes=n_classes)
    elif architecture == "nnUNetClassifier":
        model = nnUNetClassifier(sequences=sequences, n_classes=n_classes)
    else:
        model = torch.nn.Module()  # the reference does not raise for unknown names (experiment_init.py:113-116)
    if save_folder is not None:
        with (save_folder / "model.txt").open("w") as f:
            print(model, file=f)
    return model


def init_segmentation_model(architecture: str, sequences: int = 1, regions: int = 1, width: int = 48, save_folder=None,
                            deep_supervision: bool = False) -> nn.Module:
    """The reference-local branches of src/utils/experiment_init.py:26-92 (the MONAI networks of that factory are
    third-party models outside this repo's path)."""
    if architecture == "BTSUNet":
        return BTSUNet(sequences=sequences, regions=regions, width=width, deep_supervision=deep_supervision)
    if architecture == "nnUNet":
        return nnUNet2021(sequences=sequences, regions=regions)
    if architecture == "ResidualUNet":
        return ResidualUNet(sequences=sequences, regions=regions, width=width)
    raise NotImplementedError(f"segmentation architecture {architecture!r} is outside the accelerated path "
                              "(MONAI UNet / AttentionUnet / SwinUNETR / SegResNet)")


def init_multitask_model(architecture: str, sequences: int = 1, regions: int = 1, n_classes: int = 2, width: int = 48,
                         save_folder=None, deep_supervision: bool = False) -> nn.Module:
    """String-keyed factory with the reference's signature (src/utils/experiment_init.py:130-174)."""
    if architecture == "Multi_BTSUNet":
        model = Multi_BTS_UNet(sequences=sequences, regions=regions, n_classes=n_classes, width=width,
                               deep_supervision=deep_supervision)
    elif architecture == "MTUNetPlusPlus":
        model = MTUNetPlusPlus(in_channels=sequences, out_channels=regions, n_classes=n_classes,
                               deep_supervision=deep_supervision)
    elif architecture == "MTnnUNet":
        model = MTnnUNet(sequences=sequences, regions=regions, n_classes=n_classes)
    else:
        model = torch.nn.Module()  # the reference does not raise for unknown names (experiment_init.py:160-163)
    if save_folder is not None:
        with (save_folder / "model.txt").open("w") as f:
            print(model, file=f)
    return model
