"""Reverse-mode WIRING of the static plans (plan.py) against the fp32 oracle, parameter by parameter.

Why a random cotangent.  With the real objective at random init the Dice gradient is nearly uniform over the plane, each
InstanceNorm backward subtracts the plane mean from it, and what is left is a small difference of large numbers: the
fp32 oracle's own gradient moves by 20 % (median over parameters, up to 60 %) when nothing but its WEIGHTS are rounded to
bf16 (tools/emulate_bf16.py, EMU=w; rounding the gradient tensors instead, EMU=ga,gy, moves it by 0.4-0.6 %).  A
whole-network check under that objective therefore cannot tell a missing contribution from rounding (round 1 asserted
cos > 0.6).  Filling every output's gradient with randn removes the cancellation: a bf16 forward perturbation of 1e-2
then moves a parameter gradient by O(1e-2), while a dropped consumer of a multi-consumer tensor, a shared module that
only counts one of its two applications, or a wrong K offset of one concat source moves it by O(1)
(`test_tolerance_would_catch_a_dropped_consumer` shows the margin on the oracle itself).

Asserted per parameter (relative L2 against the oracle's fp32 gradient): conv / InstanceNorm-affine / transposed-conv
parameters and the class FC layers <= 5e-2, mask-head parameters <= 1e-2 (bf16 activations, fp32 head arithmetic);
parameters the reference leaves without gradient must have none here either."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _cuda(lib):
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-30)).item()


def build(mod, arch, ds):
    if arch == "unetpp":
        return mod.MTUNetPlusPlus(in_channels=1, out_channels=1, n_classes=3, deep_supervision=ds)
    if arch == "nnunet":
        return mod.MTnnUNet(1, 1, 3)
    return mod.Multi_BTS_UNet(1, 1, 3, 32, ds)


def pair(arch, ds):
    from oracle import torch_oracle as O
    from multi_task_breast_cancer_b200 import models as M
    torch.manual_seed(1993)
    ref = build(O, arch, ds)
    new = build(M, arch, ds)
    new.load_state_dict(ref.state_dict())
    return ref.cuda(), new.cuda()


def as_lists(logits, outs):
    return (logits if isinstance(logits, list) else [logits]), (outs if isinstance(outs, list) else [outs])


def cotangents(logits, outs, seed=11):
    g = torch.Generator(device="cuda").manual_seed(seed)
    gc = [torch.randn(t.shape, device="cuda", generator=g) for t in logits]
    gs = [torch.randn(t.shape, device="cuda", generator=g) for t in outs]
    return gc, gs


def backward_random(model, img, gc=None, gs=None):
    model.zero_grad(set_to_none=True)
    logits, outs = as_lists(*model(img))
    if gc is None:
        gc, gs = cotangents(logits, outs)
    obj = sum((t * g).sum() for t, g in zip(logits, gc)) + sum((t * g).sum() for t, g in zip(outs, gs))
    obj.backward()
    return gc, gs


MASK_HEADS = ("final_conv_", "output")                              # 1x1 / composed deep-supervision mask heads
CLASS_FC = ("classifier.1.", "classifier.3.", "classifier.5.")        # Linear layers behind GAP / Flatten
# parameters whose gradient only comes out right if the explicit reverse-mode bookkeeping is right
MUST_CHECK = {
    "unetpp": [
        "conv_0_0.conv_1.conv.weight",           # x_0_0: five consumers (MTUNetPlusPlus.py:103-118) + max-pool
        "conv_1_0.convs.conv_1.conv.weight",     # x_1_0: four consumers + max-pool
        "upcat_0_1.convs.conv_1.conv.weight",    # x_0_1: three later concat consumers + (deep supervision) a head
        "process_level_3.convs.conv_0.conv.weight",   # shared Down(192->384) applied to x_3_0 AND x_3_1 (:128)
        "process_level_3.convs.conv_1.conv.weight",
        "upcat_3_1.convs.conv_1.conv.weight",    # x_3_1: decoder consumer + pooled class-branch consumer
        "conv_4_0.convs.conv_1.conv.weight",     # x_4_0: decoder + class-branch concat
        "upcat_0_4.convs.conv_0.conv.weight",    # five concat sources, K offsets per source
        "upcat_0_4.upsample.deconv.weight",
        "conv_0_0.conv_0.conv.weight",           # first layer (CUDA-core kernel), end of the chain
    ],
    "nnunet": [
        "upsample5.weight",                      # the same ConvTranspose2d applied twice (MTnnUNet.py:160,174)
        "bottleneck.ConvInNormLRelu2.Conv.weight",
        "encoder5.ConvInNormLRelu2.Conv.weight",  # e5: skip + pool + class branch
        "decoder5.ConvInNormLRelu2.Conv.weight",  # d5: next decoder + class branch + (no head)
        "encoder1.ConvInNormLRelu1.Conv.weight",
        "output4.0.weight", "output4.1.weight",   # composed deep-supervision head: gradients to BOTH original params
    ],
    "bts": [
        "encoder1.ConvInNormLRelu1.Conv.weight",
    ],
}


@pytest.mark.parametrize("arch,ds,B,S", [("unetpp", True, 2, 128), ("unetpp", False, 2, 128), ("nnunet", True, 2, 128),
                                         ("bts", True, 2, 128), ("bts", False, 2, 128), ("unetpp", True, 3, 64)])
def test_random_cotangent_gradients_match_oracle_per_parameter(arch, ds, B, S):
    from oracle import torch_oracle as O
    ref, new = pair(arch, ds)
    img, *_ = O.synthetic_batch(B, S, S, device="cuda")
    gc, gs = backward_random(ref, img)
    backward_random(new, img, gc, gs)
    torch.cuda.synchronize()
    pr, pn = dict(ref.named_parameters()), dict(new.named_parameters())
    scale = max(p.grad.norm().item() / p.numel() ** 0.5 for p in pr.values() if p.grad is not None)
    rows, bad = [], []
    for n, p in pn.items():
        r = pr[n]
        assert (p.grad is None) == (r.grad is None), f"{n}: gradient presence differs from the reference"
        if r.grad is None:
            continue
        assert p.grad.shape == p.shape and p.grad.dtype == torch.float32 and torch.isfinite(p.grad).all(), n
        rms = r.grad.norm().item() / r.grad.numel() ** 0.5
        if rms < 1e-6 * scale:
            # conv bias in front of an InstanceNorm: identically zero here, ~1e-9 noise in the reference
            assert p.grad.abs().max().item() <= 1e-5 * scale, (n, p.grad.abs().max().item())
            continue
        e = rel(p.grad, r.grad)
        # mask heads: fp32 arithmetic on bf16 activations -> 1e-2; the class FC layers sit behind the whole encoder (their
        # input carries its accumulated bf16 error, and hidden ReLU units near 0 may flip) -> the conv-stack bound
        tol = 1e-2 if n.startswith(MASK_HEADS) else 5e-2
        rows.append((e, n, tol))
        if e > tol:
            bad.append((n, e, tol))
    rows.sort(reverse=True)
    print(f"\n{arch} ds={ds} B={B} {S}x{S}: {len(rows)} parameters checked, worst:")
    for e, n, tol in rows[:8]:
        print(f"   {n:55s} rel {e:.4f} (tol {tol})")
    checked = {n for _, n, _ in rows}
    for n in MUST_CHECK[arch]:
        if n in pn and pr[n].grad is not None:
            assert n in checked, f"{n} was not compared"
    assert not bad, bad
    # whole-vector agreement follows from the per-parameter bound; assert it anyway (round 1 asserted > 0.6)
    a = torch.cat([pn[n].grad.flatten() for _, n, _ in rows])
    b = torch.cat([pr[n].grad.flatten() for _, n, _ in rows])
    cos = torch.nn.functional.cosine_similarity(a, b, dim=0).item()
    assert cos > 0.999, cos


def test_tolerance_would_catch_a_dropped_consumer():
    """The oracle against a deliberately mis-wired copy of itself: the skip tensors entering upcat_0_3 are detached, i.e.
    x_0_0, x_0_1 and x_0_2 each lose ONE of their consumers' contributions.  The per-parameter error this causes on the
    producers of those tensors is far above the 5e-2 the test above allows, so such a bug cannot hide in the tolerance."""
    from oracle import torch_oracle as O
    torch.manual_seed(1993)
    ref = build(O, "unetpp", True).cuda()
    img, *_ = O.synthetic_batch(2, 128, 128, device="cuda")
    gc, gs = backward_random(ref, img)
    good = {n: p.grad.clone() for n, p in ref.named_parameters() if p.grad is not None}
    h = ref.upcat_0_3.register_forward_pre_hook(lambda mod, args: (args[0], args[1].detach()))
    backward_random(ref, img, gc, gs)
    h.remove()
    for n in ["conv_0_0.conv_1.conv.weight", "upcat_0_1.convs.conv_1.conv.weight", "upcat_0_2.convs.conv_1.conv.weight"]:
        e = rel(dict(ref.named_parameters())[n].grad, good[n])
        print(f"dropped consumer: {n} rel {e:.3f}")
        assert e > 0.15, (n, e)


def test_second_forward_invalidates_the_first_backward():
    """forward; forward; backward; backward through one plan: the first graph's activations are gone -> clear error
    (round 1 silently back-propagated through the second forward's activations).  forward; backward twice (gradient
    accumulation) keeps working and adds."""
    from oracle import torch_oracle as O
    _, new = pair("nnunet", True)
    img, *_ = O.synthetic_batch(2, 64, 64, device="cuda")
    l1, o1 = new(img)
    obj1 = o1[-1].sum()
    l2, o2 = new(img)
    obj2 = o2[-1].sum()
    with pytest.raises(RuntimeError, match="overwritten"):
        obj1.backward()
    obj2.backward()                                    # the latest forward is still valid
    g1 = new.output1.weight.grad.clone()
    l3, o3 = new(img)
    o3[-1].sum().backward()                            # accumulation: p.grad += g
    assert rel(new.output1.weight.grad, 2 * g1) < 1e-3
    with torch.no_grad():                              # a no-grad forward (other plan) does not disturb a pending one
        l4, o4 = new(img)
        o4[-1].sum()
    l5, o5 = new(img)
    with torch.no_grad():
        new(img)
    o5[-1].sum().backward()
