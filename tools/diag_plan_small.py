"""GPU diagnostic: tiny plans (2-3 conv blocks + head) vs torch autograd, to bisect plan-level backward bugs."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.nn as nn
import torch.nn.functional as F

from multi_task_breast_cancer_b200.plan import Plan


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-20)).item()


class Tiny(nn.Module):
    def __init__(self, c=24, depth=3, affine=True):
        super().__init__()
        self.convs = nn.ModuleList([nn.Conv2d(1 if i == 0 else c, c, 3, padding=1) for i in range(depth)])
        self.norms = nn.ModuleList([nn.InstanceNorm2d(c, affine=affine) for _ in range(depth)])
        self.head = nn.Conv2d(c, 1, 1)

    def forward(self, x):
        acts = []
        for cv, nm in zip(self.convs, self.norms):
            x = F.leaky_relu(nm(cv(x)), 0.1)
            x.retain_grad()
            acts.append(x)
        return self.head(x), acts


def main():
    torch.manual_seed(0)
    B, H, W, c, depth = 4, 64, 64, 24, 3
    gscale = float(sys.argv[1]) if len(sys.argv) > 1 else 1e-5
    m = Tiny(c, depth).cuda()
    x = torch.randint(0, 256, (B, 1, H, W), device="cuda").float()
    dl = (torch.randn(B, 1, H, W, device="cuda") * gscale)
    out, acts = m(x)
    out.backward(dl)

    params = dict(m.named_parameters())
    plan = Plan(B, H, W, x.device, params, training=True)
    plan.x_in = x.clone()
    a, _ = plan.input_conv_in_act(plan.x_in, "convs.0.weight", "convs.0.bias", "norms.0.weight", "norms.0.bias", 0.1,
                                  False, "b0")
    for i in range(1, depth):
        a, _ = plan.conv_in_act([a], f"convs.{i}.weight", f"convs.{i}.bias", f"norms.{i}.weight", f"norms.{i}.bias",
                                0.1, False, f"b{i}")
    logits = plan.head1x1(a, "head.weight", "head.bias")
    plan.finalize()
    plan.run_pack()
    plan.run_forward()
    plan.g_seg[0].copy_(dl)
    plan.run_backward()
    torch.cuda.synchronize()
    print("logits rel", rel(logits, out))
    for i in range(depth):
        t = plan.tensors[f"b{i}"]
        print(f"  b{i}: fwd rel {rel(t.feat.to_nchw(), acts[i]):.4g}  grad rel "
              f"{rel(t.g.to_nchw(), acts[i].grad) if t.g is not None else float('nan'):.4g}  max|new g| "
              f"{t.g.t.float().abs().max().item() if t.g is not None else 0:.4g} max|ref g| {acts[i].grad.abs().max().item():.4g}")
    for n, p in params.items():
        print(f"  {n:20s} rel {rel(plan.grad_view[n], p.grad):.4g} |ref| {p.grad.norm().item():.4g} |new| {plan.grad_view[n].norm().item():.4g}")


if __name__ == "__main__":
    main()
