"""CPU experiment: how much gradient error does bf16 STORAGE of activations / activation-gradients / conv weights
cause by itself?  The fp32 oracle is run twice on the same weights and inputs: once plainly, once with every conv
output, block output and their gradients rounded to bf16 (what the B200 path stores).  Kernel bugs cannot explain any
difference reported here; this is the precision floor the GPU parity tests are judged against.

    python tools/emulate_bf16.py {unetpp|nnunet|bts} B H W
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import copy

import torch
import torch.nn as nn

from oracle import torch_oracle as O


MODE = set(os.environ.get("EMU", "y,a,gy,ga,w").split(","))


def rel(a, b):
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


def build(arch):
    if arch == "unetpp":
        return O.MTUNetPlusPlus(deep_supervision=True)
    if arch == "nnunet":
        return O.MTnnUNet(1, 1, 3)
    return O.Multi_BTS_UNet(1, 1, 3, 32, True)


def run(model, img, mask, onehot):
    model.zero_grad(set_to_none=True)
    logits, outs = model(img)
    if os.environ.get("OBJ") == "random_seg":
        g = torch.Generator().manual_seed(11)
        obj = sum((t * torch.randn(t.shape, generator=g)).sum() for t in list(outs)) + 0.0 * sum(t.sum() for t in logits)
        obj.backward()
        return logits, outs
    if os.environ.get("OBJ") == "random":
        # random cotangent on every output (tests/test_grad_wiring_gpu.py): no Dice / InstanceNorm cancellation
        g = torch.Generator().manual_seed(11)
        obj = sum((t * torch.randn(t.shape, generator=g)).sum() for t in list(logits) + list(outs))
        obj.backward()
        return logits, outs
    seg, cls = O.multitask_criterion(O.DiceLoss(), mask, outs, O.FocalLoss(), onehot, logits, True)
    (0.35 * seg + 0.65 * cls).backward()
    return logits, outs


def main():
    arch = sys.argv[1]
    B, H, W = (int(v) for v in sys.argv[2:5])
    torch.manual_seed(1993)
    ref = build(arch)
    from oracle.emulation import named_grads, with_bf16_storage
    emu = with_bf16_storage(ref, MODE)
    img, mask, onehot, _ = O.synthetic_batch(B, H, W)
    rl, ro = run(ref, img, mask, onehot)
    el, eo = run(emu, img, mask, onehot)
    print(f"== {arch} B{B} {H}x{W}: bf16-storage emulation {sorted(MODE)} vs fp32")
    for a, b in zip(eo, ro):
        print(f"  mask threshold agreement {((a > 0) == (b > 0)).float().mean().item():.5f}")
    for a, b in zip(el, rl):
        print(f"  class logits rel {rel(a, b):.4g}")
    for a, b in zip(eo, ro):
        print(f"  mask logits rel {rel(a, b):.4g}")
    pr, pe = named_grads(ref), named_grads(emu)
    rows = []
    for n, g in pe.items():
        if g is None or pr[n] is None:
            continue
        rows.append((rel(g, pr[n]), n, pr[n].norm().item()))
    med = sorted(r[0] for r in rows)[len(rows) // 2]
    print(f"  grads: median rel {med:.4g}; max {max(r[0] for r in rows if r[2] > 1e-6):.4g}")
    for e, n, s in rows[::-1]:
        if s > 1e-6 and n.endswith("weight") and not os.environ.get("EMU_BRIEF"):
            print(f"    {n:55s} rel {e:.4g} |ref| {s:.4g}")


if __name__ == "__main__":
    main()
