"""TEST INFRASTRUCTURE (never imported by the product package): the fp32 oracle with the B200 path's STORAGE precision
switched on, tensor by tensor.

The CUDA path stores four kinds of tensors in bf16: the raw conv / transposed-conv output `y`, the
normalised activation `a`, their gradients (`ga`, `gy`) and the packed conv weights `w`; everything else (statistics,
accumulators, heads, losses, parameters, optimizer) is fp32.  `with_bf16_storage(model)` returns a deep copy of an
oracle model whose forward / backward round exactly those tensors to bf16 (round to nearest even, what
`__float2bfloat16_rn` does) and nothing else.  Two uses:

* tools/emulate_bf16.py: how much of a measured difference is the price of the storage format (no kernel involved);
* tests/test_grad_wiring_gpu.py: the reverse-mode WIRING check.  Against the plain fp32 oracle the per-parameter
  gradient error of any bf16 implementation is dominated by LeakyReLU units whose pre-activation is within the forward
  rounding error of 0 and therefore take the other slope (a fraction ~0.8*eps of the units, i.e. a relative L2 error of
  ~0.9*sqrt(0.8*eps) ~ 8 % PER LAYER at eps = 1e-2, accumulating with depth; measured on the oracle itself: 10 % one
  layer below the head, 20-55 % in the encoder / class branch).  Against the oracle WITH the same storage rounding that
  error is common mode and what remains is summation order: a missing contribution, a wrong concat offset or a shared
  module counted once shows up as O(1).
"""
from __future__ import annotations

import copy

import torch
import torch.nn as nn


def _round_fn(fwd: bool, bwd: bool):
    class R(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x):
            return x.to(torch.bfloat16).float() if fwd else x.clone()

        @staticmethod
        def backward(ctx, g):
            return g.to(torch.bfloat16).float() if bwd else g
    return R


def is_fp32_head(name: str) -> bool:
    """Mask heads (1x1 and composed deep-supervision heads) run in fp32 on the bf16 activations: no rounding inside."""
    return name.startswith("output") or name.startswith("final_conv")


def with_bf16_storage(model: nn.Module, what=("y", "a", "w", "gy", "ga")) -> nn.Module:
    """Deep copy of `model` with bf16 rounding of: conv/convT outputs ('y'), LeakyReLU outputs ('a'), conv/convT weights
    ('w'), and the gradients arriving at those two kinds of tensors ('gy', 'ga')."""
    what = set(what)
    emu = copy.deepcopy(model)
    RoundY = _round_fn("y" in what, "gy" in what)
    RoundA = _round_fn("a" in what, "ga" in what)
    first = True
    for n, m in emu.named_modules():
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)) and not is_fp32_head(n):
            if "w" in what:
                # the master weights stay fp32 (the optimizer updates them); the conv consumes their bf16 rounding and
                # the weight gradient is taken w.r.t. the value used, exactly like the packed operands of plan.py
                _install_rounded_weight(m)
            is3x3 = isinstance(m, nn.Conv2d) and tuple(m.kernel_size) == (3, 3)
            m.register_forward_hook(_y_hook(RoundY, drop_bias=is3x3 and m.bias is not None, centre=is3x3 and first))
            first = first and not is3x3
        if isinstance(m, nn.LeakyReLU):
            m.register_forward_hook(lambda mod, i, o: RoundA.apply(o))
    return emu


def _y_hook(RoundY, drop_bias: bool, centre: bool):
    """What the CUDA path STORES for a conv output.  Every 3x3 conv of these models is followed by
    an InstanceNorm, which is invariant to a per-(sample, channel) shift, so the stored tensor may differ from the
    reference's by such a shift without changing anything downstream in exact arithmetic -- but the bf16 rounding grid
    moves with it, so the emulation has to round the very same numbers: the conv bias is not added, and the first layer
    (raw 0..255 image, DC level of many sigma) stores y - mean(y)."""
    def hook(mod, inputs, out):
        if drop_bias:
            out = out - mod.bias.view(1, -1, 1, 1)
        if centre:
            out = out - out.mean(dim=(2, 3), keepdim=True)
        return RoundY.apply(out)
    return hook


class _RoundSTE(torch.autograd.Function):
    """bf16 rounding with a straight-through gradient: d(round(w))/dw := 1 (what a packed bf16 copy of an fp32 master
    weight means for the optimizer)."""

    @staticmethod
    def forward(ctx, w):
        return w.to(torch.bfloat16).float()

    @staticmethod
    def backward(ctx, g):
        return g


def _install_rounded_weight(m: nn.Module):
    import torch.nn.utils.parametrize as P

    class Rounded(nn.Module):
        def forward(self, w):
            return _RoundSTE.apply(w)
    P.register_parametrization(m, "weight", Rounded(), unsafe=True)


def named_grads(model: nn.Module) -> dict:
    """{reference parameter name: gradient} of a (possibly weight-parametrized) model."""
    out = {}
    for n, p in model.named_parameters():
        out[n.replace(".parametrizations.weight.original", ".weight")] = p.grad
    return out
