"""GPU diagnostic: compare every intermediate activation and its gradient (product plan vs oracle autograd).

    python tools/diag_tensors.py {unetpp|nnunet|bts} [B H W]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

from oracle import torch_oracle as O
from multi_task_breast_cancer_b200 import models as M, criterions as Cr


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


def main():
    arch = sys.argv[1] if len(sys.argv) > 1 else "unetpp"
    B, H, W = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (4, 64 if arch != "bts" else 128, 64 if arch != "bts" else 128)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(1993)
    if arch == "unetpp":
        ref = O.MTUNetPlusPlus(deep_supervision=True)
        new = M.MTUNetPlusPlus(deep_supervision=True)
        block_types = (O._ConvNormAct,)
    elif arch == "nnunet":
        ref, new = O.MTnnUNet(1, 1, 3), M.MTnnUNet(1, 1, 3)
        block_types = (O.ConvInNormLeReLU,)
    else:
        ref, new = O.Multi_BTS_UNet(1, 1, 3, 32, True), M.Multi_BTS_UNet(1, 1, 3, 32, True)
        block_types = (O.ConvInNormLeReLU,)
    new.load_state_dict(ref.state_dict())
    ref, new = ref.cuda(), new.cuda()
    img, mask, onehot, label = O.synthetic_batch(B, H, W, device="cuda")

    captured = []  # (name, [outputs per call])

    def mk_hook(name):
        def hook(mod, inp, out):
            out.retain_grad()
            captured.append((name, out))
        return hook

    for n, m in ref.named_modules():
        if isinstance(m, block_types) or isinstance(m, torch.nn.ConvTranspose2d):
            m.register_forward_hook(mk_hook(n))
    rl, ro = ref(img)
    seg_r, cls_r = O.multitask_criterion(O.DiceLoss(), mask, ro, O.FocalLoss(), onehot, rl, True)
    (0.35 * seg_r + 0.65 * cls_r).backward()

    nl, no = new(img)
    seg_n, cls_n = Cr.apply_criterion_multitask_segmentation_classification(
        Cr.DiceLoss(sigmoid=True, squared_pred=True, smooth_nr=1, smooth_dr=1), mask, no, Cr.FocalLoss(), onehot, nl, True)
    (0.35 * seg_n + 0.65 * cls_n).backward()
    torch.cuda.synchronize()
    plan = next(iter(new._plans.values()))
    if os.environ.get("DIAG_STEP"):
        import ctypes as C
        from multi_task_breast_cancer_b200 import ops
        watch = plan.debug[os.environ["DIAG_STEP"]]
        am = lambda: ["%.3g" % watch.t[n].float().abs().max().item() for n in range(B)]
        print("after autograd backward:", am())
        plan.run_backward(); torch.cuda.synchronize()
        print("after plain re-run:", am())
        st = C.c_void_p(ops.stream_ptr())
        prev = None
        watch.t.zero_()
        for i, l in enumerate(plan.bwd):
            l(st)
            torch.cuda.synchronize()
            v = am()
            if v != prev:
                print("  step", i, l.kind, v)
                prev = v
        plan.run_backward(); torch.cuda.synchronize()
        print("after second plain re-run:", am())
    if os.environ.get("DIAG_AUX"):
        nm = os.environ["DIAG_AUX"]
        mean, rstd, gv, bv, s1, s2 = plan.debug[nm + ".aux"]
        t = plan.tensors[nm]
        y = plan.tensors[nm + ".y"].feat.t.float()
        g = t.g.t.float()
        Cp = y.shape[-1]
        HW = y.shape[1] * y.shape[2]
        print("aux", nm, "Cp", Cp)
        for n in range(B):
            print(f"  n={n}: mean absmax {mean[n].abs().max().item():.4g} rstd max {rstd[n].max().item():.4g} "
                  f"s1 absmax {s1[n].abs().max().item():.4g} s2 absmax {s2[n].abs().max().item():.4g}")
        print("  gamma", None if gv is None else gv.tolist()[:8], " beta", None if bv is None else bv.tolist()[:8])
        ym = y.mean((1, 2)); yv = y.var((1, 2), unbiased=False)
        print("  mean err", (ym - mean).abs().max(1).values.tolist(), " rstd err", ((yv + 1e-5).rsqrt() - rstd).abs().max(1).values.tolist())
        xh = (y - mean[:, None, None, :]) * rstd[:, None, None, :]
        gam = gv if gv is not None else torch.ones(Cp, device=y.device)
        bet = bv if bv is not None else torch.zeros(Cp, device=y.device)
        z = xh * gam + bet
        gg = torch.where(z > 0, g, g * 0.1)
        s1r = gg.sum((1, 2)); s2r = (gg * xh).sum((1, 2))
        print("  s1 err", (s1r - s1).abs().max(1).values.tolist(), " s2 err", (s2r - s2).abs().max(1).values.tolist())
        dyr = rstd[:, None, None, :] * gam * (gg - s1r[:, None, None, :] / HW - xh * s2r[:, None, None, :] / HW)
        dyp = plan.debug[nm + ".dy"].t.float()
        print("  recomputed dy absmax/sample", [dyr[n].abs().max().item() for n in range(B)], " plan dy", [dyp[n].abs().max().item() for n in range(B)])
    print("plan tensors:", len(plan.tensors), "launches", plan.launch_counts())
    seen = {}
    for name, out in captured:
        k = seen.get(name, 0)
        seen[name] = k + 1
        pname = name
        if pname.endswith(".upsample.deconv"):
            pname = pname[: -len(".upsample.deconv")] + ".up"
        if pname.startswith("upsample") and arch == "nnunet":
            pname = "up" + pname[len("upsample"):]
        t = plan.tensors.get(pname)
        if t is None or k > 0:
            print(f"  {name:45s} call {k}: (no 1:1 plan tensor)")
            continue
        fa = rel(t.feat.to_nchw(), out)
        if t.g is not None and out.grad is not None:
            ga = rel(t.g.to_nchw(), out.grad)
            gmax = t.g.t.float().abs().max().item()
            print(f"  {name:45s} fwd rel {fa:.4g}   grad rel {ga:.4g}  |ref g| {out.grad.norm().item():.4g} max|new g| {gmax:.4g}")
            if ga > 1 and os.environ.get("DIAG_WHERE"):
                d = (t.g.to_nchw() - out.grad).abs()
                thr = d.max().item() * 0.25
                idx = torch.nonzero(d > thr)
                print(f"      {idx.shape[0]} elements with |diff| > {thr:.3g}; per-dim unique counts: "
                      f"n {idx[:, 0].unique().numel()} c {idx[:, 1].unique().numel()} h {idx[:, 2].unique().numel()} "
                      f"w {idx[:, 3].unique().numel()}")
                print("      n:", idx[:, 0].unique().tolist()[:16], " c:", idx[:, 1].unique().tolist()[:32])
                print("      h:", idx[:, 2].unique().tolist()[:40])
                print("      w:", idx[:, 3].unique().tolist()[:40])
                for i in idx[:6]:
                    tt = tuple(i.tolist())
                    print("      at", tt, "new", t.g.to_nchw()[tt].item(), "ref", out.grad[tt].item())
                if arch == "unetpp" and name == "upcat_0_4.convs.conv_0":
                    import torch.nn.functional as Fn
                    dyf = plan.debug["upcat_0_4.convs.conv_1.dy"]
                    dyt = dyf.to_nchw()
                    wt = dict(ref.named_parameters())["upcat_0_4.convs.conv_1.conv.weight"]
                    mine = Fn.conv_transpose2d(dyt, wt.to(torch.bfloat16).float(), padding=1)
                    print("      torch dgrad from plan dy vs plan a.g:", rel(t.g.to_nchw(), mine), " vs ref:", rel(mine, out.grad))
                    print("      dy per-sample absmax:", [dyt[i].abs().max().item() for i in range(dyt.shape[0])])
                    print("      dy pad lanes absmax:", dyf.t[..., dyf.C:].float().abs().max().item())
                    wd = plan._packed["upcat_0_4.convs.conv_1.conv.weight"]["wd"][0].float()
                    wexp = wt.to(torch.bfloat16).float().flip(2, 3).permute(2, 3, 1, 0).reshape(9, 24, 24)
                    print("      wd check:", rel(wd[:, :24, :24], wexp), "wd pad absmax", wd[:, 24:, :].abs().max().item(), wd[:, :, 24:].abs().max().item())
                pad = t.g.t[..., t.g.C:]
                print("      pad lanes max:", pad.float().abs().max().item() if pad.numel() else None)
        else:
            print(f"  {name:45s} fwd rel {fa:.4g}   grad: new {t.g is not None} ref {out.grad is not None}")


if __name__ == "__main__":
    main()
