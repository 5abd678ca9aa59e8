"""GPU diagnostic: product nn.Modules vs the fp32 oracle restatement on the same weights and inputs.

    python tools/diag_models.py [arch ...]      arch in {unetpp, nnunet, bts}
Prints forward relative errors, loss values and per-parameter gradient errors.  Each arch runs in a subprocess.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def run(arch, B, H, W):
    import torch
    from oracle import torch_oracle as O
    from multi_task_breast_cancer_b200 import models as M, criterions as Cr

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(1993)
    if arch == "unetpp":
        ref = O.MTUNetPlusPlus(in_channels=1, out_channels=1, n_classes=3, deep_supervision=True)
        new = M.MTUNetPlusPlus(in_channels=1, out_channels=1, n_classes=3, deep_supervision=True)
    elif arch == "nnunet":
        ref = O.MTnnUNet(1, 1, 3)
        new = M.MTnnUNet(1, 1, 3)
    else:
        ref = O.Multi_BTS_UNet(1, 1, 3, 32, True)
        new = M.Multi_BTS_UNet(1, 1, 3, 32, True)
    assert list(ref.state_dict().keys()) == list(new.state_dict().keys())
    new.load_state_dict(ref.state_dict())
    ref, new = ref.cuda(), new.cuda()
    img, mask, onehot, label = O.synthetic_batch(B, H, W, device="cuda")

    # oracle
    rl, ro = ref(img)
    seg_r, cls_r = O.multitask_criterion(O.DiceLoss(), mask, ro, O.FocalLoss(), onehot, rl, True)
    tot_r = 0.35 * seg_r + 0.65 * cls_r
    tot_r.backward()
    # product
    nl, no = new(img)
    seg_n, cls_n = Cr.apply_criterion_multitask_segmentation_classification(
        Cr.DiceLoss(sigmoid=True, squared_pred=True, smooth_nr=1, smooth_dr=1), mask, no, Cr.FocalLoss(), onehot, nl, True)
    tot_n = 0.35 * seg_n + 0.65 * cls_n
    tot_n.backward()
    torch.cuda.synchronize()
    print(f"== {arch} B{B} {H}x{W}")
    for i, (a, b) in enumerate(zip(nl, rl)):
        print(f"  class logits[{i}] rel {rel(a, b):.4g}   max|ref| {b.abs().max().item():.3g}  argmax agree "
              f"{(a.argmax(1) == b.argmax(1)).float().mean().item():.3f}")
    for i, (a, b) in enumerate(zip(no, ro)):
        print(f"  mask logits[{i}]  rel {rel(a, b):.4g}   max|ref| {b.abs().max().item():.3g}  thresh agree "
              f"{((a > 0) == (b > 0)).float().mean().item():.5f}")
    print(f"  loss seg {seg_n.item():.6f} vs {seg_r.item():.6f} | cls {cls_n.item():.6f} vs {cls_r.item():.6f} | total "
          f"{tot_n.item():.6f} vs {tot_r.item():.6f}")
    worst = []
    pr = dict(ref.named_parameters())
    for n, p in new.named_parameters():
        gr = pr[n].grad
        if gr is None and p.grad is None:
            continue
        if (gr is None) != (p.grad is None):
            print(f"  GRAD PRESENCE MISMATCH {n}: ref {gr is not None} new {p.grad is not None}")
            continue
        e = rel(p.grad, gr)
        # bias of a conv followed by InstanceNorm has a mathematically zero gradient (fp noise in the reference)
        scale = gr.norm().item()
        worst.append((e, n, scale))
    worst.sort(reverse=True)
    big = [w for w in worst if w[0] > 0.05 and w[2] > 1e-6]
    print(f"  grads: {len(worst)} params, median rel {sorted(w[0] for w in worst)[len(worst) // 2]:.4g}, "
          f">5% (non-negligible norm): {len(big)}")
    if os.environ.get("DIAG_ALL"):
        order = {n: i for i, (n, _) in enumerate(new.named_parameters())}
        for e, n, s in sorted(worst, key=lambda w: -order[w[1]]):
            print(f"    {n:55s} rel {e:.4g} |ref| {s:.4g} |new| {dict(new.named_parameters())[n].grad.norm().item():.4g}")
    else:
        for e, n, s in worst[:12]:
            print(f"    {n:55s} rel {e:.4g} |ref| {s:.4g}")
    return 0


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        return run(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]))
    archs = sys.argv[1:] or ["unetpp", "nnunet", "bts"]
    sizes = {"unetpp": (4, 64, 64), "nnunet": (4, 64, 64), "bts": (4, 128, 128)}
    rc = 0
    for a in archs:
        B, H, W = sizes[a]
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", a, str(B), str(H), str(W)], timeout=600)
            rc |= p.returncode
        except subprocess.TimeoutExpired:
            print(a, "TIMED OUT")
            rc = 1
    return rc


if __name__ == "__main__":
    sys.exit(main())
