"""GPU parity of every C-ABI kernel against torch fp32 on the same (bf16-rounded) inputs.  Tolerances: 2e-2 of the
reference's max magnitude for bf16 outputs (one bf16 rounding of the result), 1e-4..1e-2 for fp32 outputs."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.fixture(scope="module")
def D(lib):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    import diag_kernels as d
    from multi_task_breast_cancer_b200 import _lib
    _lib.check(lib.mtbc_device_check(), "device check")
    torch.manual_seed(0)
    return d


def _run(D, fn, *a, **k):
    D.RESULTS.clear()
    fn(*a, **k)
    bad = [r for r in D.RESULTS if not r[1]]
    assert D.RESULTS and not bad, bad


@pytest.mark.parametrize("N,H,W,src,cout,kw", [
    (2, 16, 16, [64], 64, dict(bias=False, stats=False, identity=True)),
    (2, 16, 16, [32], 32, dict(bias=False, stats=False, identity=True)),
    (2, 32, 32, [24, 24, 48], 24, {}),
    (2, 32, 32, [48, 48, 48], 48, {}),
    (2, 16, 16, [96, 96], 96, {}),
    (3, 16, 16, [128], 512, {}),
    (2, 16, 16, [384, 384, 384], 512, {}),
    (4, 8, 8, [320], 320, dict(stats=False)),      # small planes: several samples per 128-row tile
    (9, 4, 4, [64], 64, dict(stats=False)),        # ragged batch (9 is not a multiple of the 8 samples per tile)
    (2, 64, 64, [24, 24, 24, 24, 48], 24, {}),     # x_0_4 of U-Net++: five concat sources folded into the K loop
])
def test_conv3x3_forward(D, N, H, W, src, cout, kw):
    _run(D, D.conv_fwd_case, N, H, W, src, cout, **kw)


# the same convolution over pixel pairs (96-byte TMA rows for dense 24-channel tensors): paired operand, result,
# per-channel statistics folded over the two pixels of a pair; several samples per CTA range, bias, the benchmark shape
@pytest.mark.parametrize("args", [(2, 16, 32, 24, 24), (3, 48, 64, 24, 24, True), (2, 64, 64, 24, 24),
                                  (32, 256, 256, 24, 24)])
def test_conv3x3_forward_through_the_pixel_pair_view(D, args):
    _run(D, D.conv_fwd_pair_case, *args)


# ... and the data gradient with the fused InstanceNorm backward sums through the pair view (affine on / off, a sample
# change inside a CTA's tile range, the benchmark shape)
@pytest.mark.parametrize("args", [(2, 32, 32, 24, 24, True), (3, 48, 64, 24, 24, False), (32, 256, 256, 24, 24, True)])
def test_conv3x3_data_gradient_with_fused_norm_backward_sums_through_the_pixel_pair_view(D, args):
    _run(D, D.conv_dgrad_inbwd_case, *args, pair=True)


def test_pixel_pair_view_refuses_what_it_cannot_serve(D):
    """The generic (small-plane) kernel knows nothing about folded statistics: creation must fail, not mis-index."""
    import torch
    from multi_task_breast_cancer_b200 import _lib, ops
    from multi_task_breast_cancer_b200.ops import Feat
    x = Feat.empty(2, 8, 16, 24); y = Feat.empty(2, 8, 16, 24)     # H = 8: not a halo plane
    xp, yp = ops.pair_view(x), ops.pair_view(y)
    w = torch.zeros(9, yp.Ck, xp.Ck, dtype=torch.bfloat16, device="cuda")
    st = [torch.zeros(2, 24, device="cuda") for _ in range(2)]
    with pytest.raises(_lib.MtbcError):
        ops.conv3x3_fwd_op([xp], w, yp, stat_sum=st[0], stat_sq=st[1], stat_fold=24)


# fused InstanceNorm + LeakyReLU backward statistics in the data-gradient epilogue: G = 2 (24 / 32 channels on planes with
# H % 32 == 0), G = 1 with 32 columns (H % 32 != 0), 64 columns, two samples per CTA range (sample change mid-range), and
# the benchmark's level-0 / level-1 shapes
@pytest.mark.parametrize("args", [(2, 32, 32, 24, 24, False), (2, 64, 64, 24, 24, True), (3, 48, 40, 24, 24, True),
                                  (2, 32, 32, 40, 40, True), (2, 64, 64, 40, 40, False), (5, 32, 64, 56, 56, True),
                                  (32, 256, 256, 24, 24, True), (8, 128, 128, 56, 56, True)])
def test_conv3x3_data_gradient_with_fused_norm_backward_sums(D, args):
    _run(D, D.conv_dgrad_inbwd_case, *args)


# fused transposed-conv backward (data + weight + bias gradient in one pass over dout): one tile, several tiles per CTA
# with accumulation into dx, narrow / wide channel counts, the benchmark's level-0 shape
@pytest.mark.parametrize("args", [(1, 8, 16, 48, 48, False), (2, 16, 16, 48, 48, False), (2, 32, 16, 24, 48, True),
                                  (3, 24, 48, 40, 64, True), (2, 16, 32, 56, 24, False, False),
                                  (32, 128, 128, 48, 48, True)])
def test_transposed_conv_fused_backward(D, args):
    _run(D, D.convT_bwd_case, *args)


def test_transposed_conv_fused_backward_refuses_unserved_shapes(D):
    import torch
    from multi_task_breast_cancer_b200 import _lib, ops
    from multi_task_breast_cancer_b200.ops import Feat
    for (N, H, W, Cin, Cout) in [(2, 16, 16, 64, 48), (2, 16, 16, 96, 48), (2, 8, 8, 48, 96), (2, 4, 16, 48, 48)]:
        x = Feat.empty(N, H, W, Cin); dout = Feat.empty(N, 2 * H, 2 * W, Cout); dx = Feat.empty(N, H, W, Cin)
        wd = torch.zeros(4, x.Ck, dout.Ck, dtype=torch.bfloat16, device="cuda")
        acc = torch.zeros(4, dout.Ck, x.Ck, device="cuda")
        with pytest.raises(_lib.MtbcError):
            ops.convT_bwd_op(x, dout, wd, acc, None, dx, False)


def test_fused_norm_backward_sums_refuse_unserved_shapes(D):
    """Channel pitches whose dense y tile would be read with shared-memory bank conflicts (32 / 48 / 64 channels) and planes
    the halo kernel does not take are refused at creation: plan.py then keeps the two-pass InstanceNorm backward."""
    import torch
    from multi_task_breast_cancer_b200 import _lib, ops
    from multi_task_breast_cancer_b200.ops import Feat
    for (N, H, W, Cc) in [(2, 32, 32, 32), (2, 32, 32, 48), (2, 32, 32, 64), (2, 8, 8, 24)]:
        y = Feat.empty(N, H, W, Cc); dy = Feat.empty(N, H, W, Cc); ga = Feat.empty(N, H, W, Cc)
        wd = torch.zeros(9, ga.Ck, dy.Ck, dtype=torch.bfloat16, device="cuda")
        st = [torch.zeros(N, y.Cp, device="cuda") for _ in range(4)]
        with pytest.raises(_lib.MtbcError):
            ops.conv3x3_dgrad_op(dy, wd, ga, False, bwd_fuse=(y, st[0], st[1], None, None, 0.1), s1=st[2], s2=st[3])


@pytest.mark.parametrize("args", [(2, 16, 16, 64, 64, False), (2, 32, 32, 24, 48, True), (2, 16, 16, 192, 96, False),
                                  (4, 64, 64, 24, 24, False, 1e-5), (4, 4, 4, 512, 512, False, 1e-3)])
def test_conv3x3_dgrad(D, args):
    _run(D, D.conv_dgrad_case, *args)


@pytest.mark.parametrize("args", [(2, 32, 32, [24, 48], 24), (2, 64, 64, [24, 24, 24, 24, 48], 24),
                                  (2, 32, 32, [48, 48, 96], 48), (2, 16, 16, [96, 96, 192], 96)])
def test_conv3x3_dgrad_fused_over_concat_sources(D, args):
    """One launch reads dy once and writes every concat source's gradient (some stored, some accumulated)."""
    _run(D, D.conv_dgrad_multi_case, *args)


def test_conv3x3_dgrad_fused_refuses_what_it_cannot_serve(D):
    """Weights too wide to stay resident next to the halo ring: creation fails cleanly (the plan then emits one launch
    per source); nothing silently falls back inside the library."""
    import torch
    from multi_task_breast_cancer_b200 import _lib, ops
    dy = ops.Feat.empty(2, 16, 16, 512)
    dxs = [ops.Feat.empty(2, 16, 16, 384) for _ in range(3)]
    wd = torch.zeros(9, 3 * 384, 512, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(_lib.MtbcError):
        ops.conv3x3_dgrad_multi_op(dy, wd, dxs, [False] * 3)


@pytest.mark.parametrize("args", [(2, 16, 16, 64, 64), (2, 16, 16, 32, 32), (2, 32, 32, 24, 24), (2, 32, 32, 48, 24),
                                  (2, 16, 16, 192, 96), (2, 16, 16, 128, 256), (2, 16, 16, 384, 512), (4, 8, 8, 320, 320),
                                  (2, 64, 64, 24, 24, 7)])
def test_conv3x3_wgrad(D, args):
    _run(D, D.conv_wgrad_case, *args)


@pytest.mark.parametrize("args", [(2, 32, 32, [24], 24), (2, 32, 32, [24, 48], 24), (2, 64, 64, [24, 24, 24, 24, 48], 24),
                                  (2, 32, 32, [48, 48, 48, 48], 48), (2, 16, 16, [96, 96, 192], 96),
                                  (2, 16, 16, [384, 384, 384], 512), (3, 16, 24, [70, 33], 100)])
def test_conv3x3_wgrad_fused_over_concat_sources(D, args):
    """One launch reads dy once per pixel tile and accumulates every concat source's weight gradient."""
    _run(D, D.conv_wgrad_multi_case, *args)


@pytest.mark.parametrize("which", ["fwd", "dgrad", "wgrad", "fwd48", "wgrad48"])
def test_conv3x3_at_the_benchmark_shape(D, which):
    """The level-0 / level-1 multi-source layers of U-Net++ at BASELINE configs[1]'s real extent (B = 32, 256 x 256 /
    128 x 128): many tiles per persistent CTA, G = 2 row stacking, two MMA lanes, the fused multi-source data / weight
    gradients with their splits, two CTAs per SM -- the paths the small cases above only touch with a few tiles."""
    if which == "fwd":
        _run(D, D.conv_fwd_case, 32, 256, 256, [24, 24, 24, 24, 48], 24)
    elif which == "dgrad":
        _run(D, D.conv_dgrad_multi_case, 32, 256, 256, [24, 24, 24, 24, 48], 24)
    elif which == "wgrad":
        _run(D, D.conv_wgrad_multi_case, 32, 256, 256, [24, 24, 24, 24, 48], 24)
    elif which == "fwd48":
        _run(D, D.conv_fwd_case, 32, 128, 128, [48, 48, 48, 48], 48)
    else:
        _run(D, D.conv_wgrad_multi_case, 32, 128, 128, [48, 48, 48, 48], 48)


@pytest.mark.parametrize("args", [(2, 16, 16, 64, 32), (2, 16, 16, 48, 48), (2, 8, 8, 384, 192), (4, 8, 8, 320, 320),
                                  # several pixel tiles per CTA: the weight-gradient operand rings wrap around
                                  (8, 128, 128, 48, 48), (6, 64, 64, 96, 48)])
def test_conv_transpose_k2(D, args):
    _run(D, D.convT_case, *args)


@pytest.mark.parametrize("cout", [24, 32, 16])
def test_first_layer(D, cout):
    _run(D, D.first_conv_case, 2, 32, 32, cout)


@pytest.mark.parametrize("args", [(3, 64, 64, 24), (2, 128, 128, 32), (2, 16, 24, 24), (5, 32, 32, 16), (2, 64, 192, 24),
                                  (32, 256, 256, 24), (3, 72, 128, 16)])
def test_first_layer_shapes(D, args):
    """Whole 1024-pixel blocks (8 pixels per thread, register-resident statistics) and planes that are not (one pixel
    per thread); the weight gradient's pixel loop wraps around the grid; W % 64 == 0 and H % 8 == 0 take the row-strip
    weight gradient (one, two and four 64-pixel passes per row; the benchmark shape), the others the per-pixel one."""
    _run(D, D.first_conv_case, *args)


@pytest.mark.parametrize("args", [(2, 32, 32, 24, True, True), (2, 16, 16, 96, True, False),
                                  (2, 16, 16, 320, False, True, 0.01), (3, 8, 8, 512, False, False, 0.01),
                                  # >= 1 Mi elements: the bulk-copy pipelined kernels (stream_pipe.cu)
                                  (3, 128, 128, 24, True, False),          # cvec 4, whole tiles
                                  (3, 64, 64, 96, True, False),            # cvec 12: 252 threads, ragged last tile
                                  (5, 32, 32, 320, False, False, 0.01),    # cvec 40: 240 threads, several samples / CTA
                                  (40, 16, 16, 384, True, False),          # more samples than tiles per CTA
                                  # dA + y > 96 MB: the fused cooperative backward (group barriers, L2 second pass)
                                  (13, 256, 256, 24, True, False),
                                  (25, 128, 128, 48, False, False, 0.01)])
def test_instance_norm_lrelu_pool(D, args, monkeypatch):
    monkeypatch.setenv("MTBC_FUSED_INBWD", "1")   # also exercise the opt-in cooperative backward where it is eligible
    _run(D, D.norm_case, *args)


@pytest.mark.parametrize("args", [(24, [24]), (24, [24, 24, 48]), (48, [48, 96]), (100, [70, 33]), (512, [384, 384, 384])])
def test_param_jobs_tiled_pack_unpack(D, args):
    _run(D, D.param_jobs_case, *args)


@pytest.mark.parametrize("args", [(2, 16, 16, 48), (3, 128, 128, 48), (3, 64, 64, 96)])
def test_channel_sum(D, args):
    _run(D, D.chansum_case, *args)


@pytest.mark.parametrize("args", [(2, 32, 32, 24), (2, 16, 16, 16)])
def test_head1x1(D, args):
    _run(D, D.head1x1_case, *args)


@pytest.mark.parametrize("args", [(2, 8, 8, 128, 8), (2, 16, 16, 64, 4), (3, 16, 16, 32, 2), (2, 24, 40, 24, 2), (3, 32, 32, 48, 4),
                                  (32, 32, 32, 128, 8)])
def test_composed_deep_supervision_head(D, args):
    _run(D, D.dshead_case, *args)


@pytest.mark.parametrize("args", [(3, 4, 4, 512), (5, 16, 16, 512), (2, 8, 8, 320)])
def test_gap_fc_head(D, args):
    _run(D, D.gap_fc_case, *args)


def test_flatten_fc_head(D):
    _run(D, D.flat_fc_case, 3, 16, 16, 256)


def test_losses_and_adam(D):
    _run(D, D.loss_cases)
