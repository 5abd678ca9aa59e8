"""ResidualUNet (last item; reference src/models/segmentation/ResidualUNet.py): the only backbone of
the reference with BatchNorm2d (running statistics), stride-2 3x3 convolutions and F.dropout(p=0.2).

CPU: the oracle restatement against the fixture the reference's own module produced (same seed -> bit-identical
parameters, outputs INCLUDING the dropout draws, loss, gradient norm, running statistics after the step, the set of
parameters that get no gradient); the drop-in's state_dict is the reference's.
GPU: forward / backward parity with the dropout masks handed to both sides (the device generator cannot reproduce torch's
stream), BatchNorm running-statistics update, eval mode, the stride-2 kernels, and the device generator's statistics."""
import hashlib
import os

import pytest
import torch

from oracle import torch_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


def _digest(sd):
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode()); h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-30)).item()


def test_oracle_and_dropin_match_reference_fixture():
    from multi_task_breast_cancer_b200 import models as M
    fx = torch.load(os.path.join(HERE, "golden", "single_task.pt"), weights_only=False)["residualunet"]
    torch.manual_seed(fx["seed"]); ora = O.ResidualUNet(1, 1, 24)
    torch.manual_seed(fx["seed"]); new = M.ResidualUNet(1, 1, 24)
    assert _digest(ora.state_dict()) == fx["state_digest"] == _digest(new.state_dict())
    assert list(new.state_dict()) == list(ora.state_dict())
    assert sum(p.numel() for p in new.parameters()) == fx["n_params"] == 1328809
    img, mask, _, _ = O.synthetic_batch(fx["B"], fx["S"], fx["S"], seed=fx["seed"])
    torch.manual_seed(fx["seed"] + 1)                    # the dropout draws of the fixture
    out = ora(img)
    assert torch.allclose(out, fx["outputs"][0], atol=1e-5)
    loss = O.DiceLoss()(out, mask)
    assert abs(loss.item() - fx["loss"]) < 1e-5
    loss.backward()
    assert sorted(k for k, p in ora.named_parameters() if p.grad is None) == fx["no_grad_params"]
    assert fx["no_grad_params"] == sorted(f"decoder.conv{i}.{w}" for i in (1, 2, 3) for w in ("weight", "bias"))
    assert _digest(ora.state_dict()) == fx["state_digest_after_step"]     # running statistics moved as the reference's
    assert isinstance(M.init_segmentation_model("ResidualUNet", width=24), M.ResidualUNet)


# ------------------------------------------------------------------------------------------------------------- GPU
def _pair(precision="bf16", train=True):
    from multi_task_breast_cancer_b200 import models as M
    torch.manual_seed(1993)
    ref = O.ResidualUNet(1, 1, 24)
    new = M.ResidualUNet(1, 1, 24)
    new.load_state_dict(ref.state_dict())
    ref, new = ref.cuda(), new.cuda()
    ref.train(train); new.train(train)
    new.set_precision(precision)
    new._external_dropout = True
    return ref, new


def _share_masks(ref, new, img, need_grad, seed=5):
    """Draw one Bernoulli(0.8) mask per dropout site (the oracle's call order = the plan's creation order) and hand it
    to both sides: NCHW float to the oracle's hook, NHWC uint8 into the plan's mask buffers."""
    shapes = []
    ref.dropout = lambda t: (shapes.append(tuple(t.shape)), t)[1]
    with torch.no_grad():
        ref.train(ref.training)
        saved = {k: v.clone() for k, v in ref.state_dict().items()}
        ref(img)
        ref.load_state_dict(saved)                       # the dry run moved the running statistics: put them back
    g = torch.Generator(device="cuda").manual_seed(seed)
    masks = [(torch.rand(s, device="cuda", generator=g) < 0.8) for s in shapes]
    it = iter(masks)
    ref.dropout = lambda t: t * next(it).float() / 0.8
    plan = new._get_plan(img, need_grad)
    assert len(plan.dropout_masks) == len(masks) == 13
    for buf, m in zip(plan.dropout_masks, masks):
        buf.copy_(m.permute(0, 2, 3, 1).contiguous().flatten().to(torch.uint8))
    return masks


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("bf16", 5e-2), ("tf32x3", 1e-3)])
def test_forward_parity_train_mode_and_running_statistics(lib, precision, tol):
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ref, new = _pair(precision)
    img, mask, *_ = O.synthetic_batch(4, 128, 128, device="cuda")
    _share_masks(ref, new, img, need_grad=False)
    with torch.no_grad():
        r = ref(img)
        n = new(img)
    e = rel(n, r)
    print(f"\nResidualUNet [{precision}] train-mode forward rel L2 {e:.2e}")
    assert n.shape == r.shape == (4, 1, 128, 128) and e < tol, e
    rs, ns = ref.state_dict(), new.state_dict()
    worst = 0.0
    for k in rs:
        if k.endswith("num_batches_tracked"):
            assert int(ns[k]) == int(rs[k]) == 1, k
        elif "running_" in k:
            worst = max(worst, rel(ns[k], rs[k]))
    print(f"   running statistics after one step: worst rel {worst:.2e}")
    assert worst < (3e-2 if precision == "bf16" else 1e-4), worst
    if precision == "tf32x3":
        assert ((n > 0) == (r > 0)).float().mean().item() >= 0.999


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,W", [(2, 64, 64), (3, 128, 256), (2, 256, 256)])
def test_other_input_extents(lib, B, H, W):
    """The reference accepts any extent divisible by 8; the kernels need the planes tileable (H, W % 16 == 0): small
    planes (8 x 8 at the bottom of a 64 x 64 input), a non-square input, and 256 x 256."""
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    ref, new = _pair("tf32x3")
    img = O.synthetic_batch(B, max(H, W), max(H, W), device="cuda")[0][:, :, :H, :W].contiguous()
    _share_masks(ref, new, img, need_grad=False)
    with torch.no_grad():
        r = ref(img)
        n = new(img)
    assert n.shape == r.shape == (B, 1, H, W) and rel(n, r) < 1e-3, rel(n, r)
    if (H, W) == (64, 64):
        with pytest.raises(ValueError, match="not tileable"):      # 192 / 8 = 24-pixel rows at the bottom level
            new(torch.zeros(1, 1, 128, 192, device="cuda"))


@pytest.mark.gpu
def test_eval_mode_uses_running_statistics(lib):
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    ref, new = _pair("tf32x3", train=True)
    img, *_ = O.synthetic_batch(2, 128, 128, device="cuda")
    # two training forwards on the oracle move the running statistics away from (0, 1); copy them over
    torch.manual_seed(0)
    ref.dropout = lambda t: torch.nn.functional.dropout(t, p=0.2)
    with torch.no_grad():
        ref(img); ref(img)
    new.load_state_dict(ref.state_dict())
    ref.eval(); new.eval()
    before = {k: v.clone() for k, v in new.state_dict().items()}
    _share_masks(ref, new, img, need_grad=False)          # F.dropout(training=True): drawn in eval mode as well
    with torch.no_grad():
        r = ref(img)
        n = new(img)
    assert rel(n, r) < 1e-3, rel(n, r)
    after = new.state_dict()
    assert all(torch.equal(before[k], after[k]) for k in before), "eval mode must not touch the running statistics"


@pytest.mark.gpu
def test_gradients_match_oracle_per_parameter(lib):
    """Wiring of the new ops (BatchNorm2d backward through batch-pooled sums, stride-2 conv data / weight gradients,
    residual add, dropout) in the fp32-grade mode, random cotangent, every parameter."""
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ref, new = _pair("tf32x3")
    img, *_ = O.synthetic_batch(2, 128, 128, device="cuda")
    _share_masks(ref, new, img, need_grad=True)
    g = torch.randn(2, 1, 128, 128, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    (ref(img) * g).sum().backward()
    (new(img) * g).sum().backward()
    torch.cuda.synchronize()
    pr, rows, cancelled = dict(ref.named_parameters()), [], []
    rms = lambda t: t.norm().item() / t.numel() ** 0.5
    scale = max(rms(p.grad) for p in pr.values() if p.grad is not None)
    for k, p in new.named_parameters():
        assert (p.grad is None) == (pr[k].grad is None), k
        if p.grad is None:
            continue
        if rms(pr[k].grad) < 1e-5 * scale:
            # bias of a conv in front of a BatchNorm2d: the batch-mean subtraction cancels it, its true gradient is 0 and
            # both sides hold rounding noise (reference ~1e-7 of the scale)
            assert k.endswith((".conv1.bias", ".conv3.bias")), k
            assert rms(p.grad) < 1e-4 * scale, (k, rms(p.grad), scale)
            cancelled.append(k)
            continue
        rows.append((rel(p.grad, pr[k].grad), k))
    rows.sort(reverse=True)
    print("\nResidualUNet [tf32x3] per-parameter gradient rel-L2, worst:", [(k, f"{e:.2e}") for e, k in rows[:5]],
          f"; {len(cancelled)} BatchNorm-cancelled conv biases ~0 on both sides")
    assert len(rows) + len(cancelled) == len(pr) - 6     # every parameter but the six of the never-called 1x1 convs
    assert rows[0][0] < 5e-2, rows[:5]
    unused = [k for k, p in new.named_parameters() if p.grad is None]
    assert sorted(unused) == sorted(f"decoder.conv{i}.{w}" for i in (1, 2, 3) for w in ("weight", "bias"))


@pytest.mark.gpu
def test_training_step_on_the_product_path(lib):
    """bf16, device-drawn dropout masks: loss decreases over a few Adam steps, masks change from step to step, keep
    rate is 0.8, and backward uses the very mask forward drew (the gradient of sum(out) w.r.t. a dropped unit is 0)."""
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    from multi_task_breast_cancer_b200 import criterions as Cr, models as M
    torch.manual_seed(7)
    net = M.ResidualUNet(1, 1, 24).cuda()
    img, mask, *_ = O.synthetic_batch(4, 128, 128, device="cuda")
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, eps=1e-4)
    dice = Cr.init_criterion_segmentation("DICE")
    losses, prev = [], None
    for step in range(6):
        opt.zero_grad(set_to_none=True)
        loss = dice(net(img), mask)
        loss.backward()
        opt.step()
        losses.append(loss.item())
        plan = next(iter(net._plans.values()))
        m = plan.dropout_masks[0].clone()
        keep = torch.cat([b.float() for b in plan.dropout_masks]).mean().item()
        assert abs(keep - 0.8) < 5e-3, keep
        if prev is not None:
            assert not torch.equal(m, prev), "dropout masks must be redrawn every forward"
        prev = m
    print("\nResidualUNet bf16 Dice over 6 Adam steps:", [f"{v:.4f}" for v in losses])
    assert all(torch.isfinite(torch.tensor(losses))) and losses[-1] < losses[0]
    assert int(net.in_block.bn1.num_batches_tracked) == 6


@pytest.mark.gpu
@pytest.mark.parametrize("N,H,W,Cin,Cout", [(2, 32, 32, 24, 48), (2, 64, 64, 48, 96), (3, 16, 16, 96, 192)])
def test_stride2_conv_kernels(lib, N, H, W, Cin, Cout):
    """Forward (four strided views through the generic implicit GEMM), weight gradient, and the data gradient as the
    stride-1 gradient of the zero-stuffed dy, against torch fp32 on bf16-rounded inputs."""
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    import ctypes as C
    import torch.nn.functional as F
    from multi_task_breast_cancer_b200 import _lib, ops
    from multi_task_breast_cancer_b200.ops import Feat, ptr
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    r16 = lambda *s, sc=1.0: (torch.randn(*s, device="cuda") * sc).to(torch.bfloat16).float()
    x, w, b = r16(N, Cin, H, W), r16(Cout, Cin, 3, 3, sc=0.1), r16(Cout)
    xf = Feat.from_nchw(x)
    out = Feat.empty(N, H // 2, W // 2, Cout)
    wf = torch.zeros(9, out.Ck, xf.Ck, dtype=torch.bfloat16, device="cuda")
    wd = torch.zeros(9, xf.Ck, out.Ck, dtype=torch.bfloat16, device="cuda")
    ops.pack_conv_weight(w, [Cin], [0], wf, [wd])
    bp = torch.zeros(out.Ck, device="cuda"); bp[:Cout] = b
    ops.conv3x3_s2_fwd_op(xf, wf, out, bias=bp).launch()
    ref = F.conv2d(x, w, b, stride=2, padding=1)
    assert rel(out.to_nchw(), ref) < 6e-3
    dy = r16(N, Cout, H // 2, W // 2)
    dyf = Feat.from_nchw(dy)
    acc = torch.zeros(9, out.Ck, xf.Ck, device="cuda")
    ops.conv3x3_s2_wgrad_op(xf, dyf, acc).launch()
    gw = torch.nn.grad.conv2d_weight(x, w.shape, dy, stride=2, padding=1)
    got = acc[:, :Cout, :Cin].permute(1, 2, 0).reshape(Cout, Cin, 3, 3)
    assert rel(got, gw) < 1e-3
    up = Feat.empty(N, H, W, Cout)
    _lib.call("mtbc_zero_stuff2", ptr(dyf.t), N, H // 2, W // 2, dyf.Cp, ptr(up.t), C.c_void_p(ops.stream_ptr()))
    dx = Feat.empty(N, H, W, Cin)
    ops.conv3x3_dgrad_op(up, wd, dx, accumulate=False).launch()
    torch.cuda.synchronize()
    gx = torch.nn.grad.conv2d_input(x.shape, w, dy, stride=2, padding=1)
    assert rel(dx.to_nchw(), gx) < 6e-3
