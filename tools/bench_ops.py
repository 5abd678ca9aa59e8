"""GPU micro-benchmark of individual tensor-core ops at benchmark shapes (also the ncu target).

    python tools/bench_ops.py [case ...]      cases: l0_single l0_multi l1_multi l3 cls wgrad_l0 wgrad_l3 convT
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

from multi_task_breast_cancer_b200 import _lib, ops
from multi_task_breast_cancer_b200.ops import Feat

dev = "cuda"


def feat(N, H, W, C):
    f = Feat.empty(N, H, W, C)
    f.t[..., :C] = torch.randn(N, H, W, C, device=dev).to(torch.bfloat16)
    return f


def timeit(op, flops_true, name, reps=5):
    for _ in range(2):
        op.launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        op.launch()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:46s} {ms * 1e3:8.1f} us   {flops_true / ms / 1e9:7.1f} TF/s (true)   {op.flops / ms / 1e9:7.1f} TF/s (padded)", flush=True)


def conv_fwd(N, H, W, src_C, Cout, name, stats=True):
    srcs = [feat(N, H, W, c) for c in src_C]
    out = Feat.empty(N, H, W, Cout)
    offs, ktot = ops.k_offsets(srcs)
    wf = (torch.randn(9, out.Cp, ktot, device=dev) * 0.05).to(torch.bfloat16)
    ssum = torch.zeros(N, out.Cp, device=dev) if stats else None
    ssq = torch.zeros(N, out.Cp, device=dev) if stats else None
    bias = torch.zeros(out.Cp, device=dev)
    op = ops.conv3x3_fwd_op(srcs, wf, out, bias=bias, stat_sum=ssum, stat_sq=ssq)
    timeit(op, 2.0 * N * H * W * Cout * sum(src_C) * 9, name)


def conv_wgrad(N, H, W, Cin, Cout, name):
    x, dy = feat(N, H, W, Cin), feat(N, H, W, Cout)
    acc = torch.zeros(9, dy.Cp, x.Cp, device=dev)
    op = ops.conv3x3_wgrad_op(x, dy, acc, 0)
    timeit(op, 2.0 * N * H * W * Cout * Cin * 9, name)


def convT(N, H, W, Cin, Cout, name):
    x = feat(N, H, W, Cin)
    out = Feat.empty(N, 2 * H, 2 * W, Cout)
    wf = (torch.randn(1, 4 * out.Cp, x.Cp, device=dev) * 0.05).to(torch.bfloat16)
    op = ops.convT_fwd_op(x, wf, out, 2, torch.zeros(out.Cp, device=dev))
    timeit(op, 2.0 * N * H * W * Cin * Cout * 4, name)


def conv_dgrad_multi(N, H, W, src_C, Cout, name, acc):
    dy = feat(N, H, W, Cout)
    dxs = [feat(N, H, W, c) for c in src_C]
    rows = sum(d.Cp for d in dxs)
    wd = (torch.randn(9, rows, dy.Cp, device=dev) * 0.05).to(torch.bfloat16)
    op = ops.conv3x3_dgrad_multi_op(dy, wd, dxs, [acc] * len(dxs))
    timeit(op, 2.0 * N * H * W * Cout * sum(src_C) * 9, name)


CASES = {
    "dgrad_l0": lambda: conv_dgrad_multi(32, 256, 256, [24, 24, 24, 24, 48], 24, "dgrad L0 fused [24x4,48]<-24 store", False),
    "dgrad_l0_acc": lambda: conv_dgrad_multi(32, 256, 256, [24, 24, 24, 24, 48], 24, "dgrad L0 fused [24x4,48]<-24 accumulate", True),
    "dgrad_l1": lambda: conv_dgrad_multi(32, 128, 128, [48, 48, 48, 96], 48, "dgrad L1 fused [48x3,96]<-48 store", False),
    "l0_single": lambda: conv_fwd(32, 256, 256, [24], 24, "fwd L0 [24]->24 @256"),
    "l0_multi": lambda: conv_fwd(32, 256, 256, [24, 24, 24, 24, 48], 24, "fwd L0 [24x4,48]->24 @256"),
    "l1_multi": lambda: conv_fwd(32, 128, 128, [48, 48, 48, 48], 48, "fwd L1 [48x4]->48 @128"),
    "l2": lambda: conv_fwd(32, 64, 64, [96, 96], 96, "fwd L2 [96,96]->96 @64"),
    "l3": lambda: conv_fwd(32, 32, 32, [192], 192, "fwd L3 [192]->192 @32"),
    "l4": lambda: conv_fwd(32, 16, 16, [384], 384, "fwd L4 [384]->384 @16"),
    "cls": lambda: conv_fwd(32, 16, 16, [384, 384, 384], 512, "fwd cls [384x3]->512 @16"),
    "wgrad_l0": lambda: conv_wgrad(32, 256, 256, 24, 24, "wgrad L0 24x24 @256"),
    "wgrad_l1": lambda: conv_wgrad(32, 128, 128, 48, 48, "wgrad L1 48x48 @128"),
    "wgrad_l3": lambda: conv_wgrad(32, 32, 32, 192, 192, "wgrad L3 192x192 @32"),
    "wgrad_cls": lambda: conv_wgrad(32, 16, 16, 384, 512, "wgrad cls 384x512 @16"),
    "convT": lambda: convT(32, 128, 128, 48, 48, "convT 48->48 @128"),
}

if __name__ == "__main__":
    _lib.check(_lib.load().mtbc_device_check(), "device")
    torch.manual_seed(0)
    for c in (sys.argv[1:] or list(CASES)):
        CASES[c]()
