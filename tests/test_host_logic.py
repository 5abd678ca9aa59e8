"""CPU tests of the host-side logic: drop-in module surface, flat layouts, data-parallel averaging (gloo, 2 ranks)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from oracle import torch_oracle as O
from multi_task_breast_cancer_b200 import models as M, ops
from multi_task_breast_cancer_b200.plan import flat_layout
from multi_task_breast_cancer_b200._lib import MtbcError


def _pairs():
    return [
        (lambda m: m.MTUNetPlusPlus(in_channels=1, out_channels=1, n_classes=3, deep_supervision=True), 14927455, 168),
        (lambda m: m.MTnnUNet(1, 1, 3), 15819799, 53),
        (lambda m: m.Multi_BTS_UNet(1, 1, 3, 32, True), 21751142, 33),
    ]


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_state_dict_and_init_identical_to_oracle(idx):
    mk, n_params, n_state = _pairs()[idx]
    torch.manual_seed(1993)
    a = mk(O)
    torch.manual_seed(1993)
    b = mk(M)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    assert all(torch.equal(sa[k], sb[k]) for k in sa)
    assert sum(p.numel() for p in b.parameters()) == n_params and len(sb) == n_state
    b.load_state_dict(sa)  # reference checkpoints load (training_multitask.py:243-249, utils/models.py:28-32)


def test_n_classes_two_means_one_logit():
    assert M.MTnnUNet(1, 1, n_classes=2).classifier[5].out_features == 1
    assert M.MTUNetPlusPlus(n_classes=2).classifier[5].out_features == 1


def test_factory_mirrors_reference():
    assert isinstance(M.init_multitask_model("MTnnUNet", 1, 1, 3), M.MTnnUNet)
    assert isinstance(M.init_multitask_model("MTUNetPlusPlus", 1, 1, 3, deep_supervision=True), M.MTUNetPlusPlus)
    assert isinstance(M.init_multitask_model("Multi_BTSUNet", 1, 1, 3, width=32), M.Multi_BTS_UNet)
    unknown = M.init_multitask_model("does-not-exist")
    assert type(unknown) is torch.nn.Module  # the reference silently returns an empty module


def test_cpu_input_fails_loudly():
    m = M.MTnnUNet(1, 1, 3)
    with pytest.raises(MtbcError):
        m(torch.zeros(1, 1, 64, 64))


def test_flat_layout_is_reverse_registration_order_and_padded():
    m = M.MTnnUNet(1, 1, 3)
    params = dict(m.named_parameters())
    ranges, total = flat_layout(params)
    names = list(params)
    assert ranges[names[-1]][0] == 0
    prev_end = 0
    for n in reversed(names):
        a, b = ranges[n]
        assert a % 64 == 0 and a >= prev_end and b - a == params[n].numel()
        prev_end = b
    assert total % 64 == 0 and total >= sum(p.numel() for p in params.values())


def test_k_offsets_alignment():
    class F:  # minimal stand-in for ops.Feat: Cp = channel pitch of the tensor, Ck = its GEMM extent (padded to 32)
        def __init__(self, cp): self.Cp, self.Ck = cp, ops.pad32(cp)
    offs, ktot = ops.k_offsets([F(24), F(48)])     # dense 24 / 48-channel tensors occupy 32 / 64 weight columns
    assert offs == [0, 64] and ktot == 128
    offs, ktot = ops.k_offsets([F(32), F(64)])
    assert offs == [0, 64] and ktot == 128          # 64-wide sources start on a 64 boundary
    offs, ktot = ops.k_offsets([F(32), F(32), F(96)])
    assert offs == [0, 32, 64] and ktot == 160
    assert ops.pad32(24) == 32 and ops.pad32(320) == 320 and ops.pad32(1) == 32


def test_gradient_bucket_plan():
    """Bucket arithmetic of the overlapped gradient all-reduce (plan.plan_buckets)."""
    from multi_task_breast_cancer_b200.plan import plan_buckets
    # flat layout = reverse registration order: the head (registered last) sits at offset 0 and finishes first
    ranges = {"enc.w": (600, 1000), "mid.w": (300, 600), "dec.w": (64, 300), "head.w": (0, 64)}
    done = {"head.w": 3, "dec.w": 10, "mid.w": 25, "enc.w": 40}
    buckets, idx, ready = plan_buckets(ranges, done, 1000, 4)
    assert buckets == [(0, 250), (250, 500), (500, 750), (750, 1000)]
    assert idx == {"head.w": 0, "dec.w": 0, "mid.w": 1, "enc.w": 2}
    assert ready == [10, 25, 40, 40]            # monotone; the empty last bucket follows the one before it
    # a parameter without gradient (unused head) does not hold its bucket back
    _, _, ready = plan_buckets(ranges, {k: v for k, v in done.items() if k != "dec.w"}, 1000, 4)
    assert ready[0] == 3
    # one bucket = the whole buffer, ready at the very end
    b1, _, r1 = plan_buckets(ranges, done, 1000, 1)
    assert b1 == [(0, 1000)] and r1 == [40]


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _dp_worker(rank, world, port, out_path):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    torch.manual_seed(1993)
    model = O.MTnnUNet(1, 1, 3)
    B = 4
    img, mask, onehot, _ = O.synthetic_batch(B, 64, 64)
    sl = slice(rank * B // world, (rank + 1) * B // world)
    logits, outs = model(img[sl])
    seg, cls = O.multitask_criterion(O.DiceLoss(), mask[sl], outs, O.FocalLoss(), onehot[sl], logits, True)
    (0.35 * seg + 0.65 * cls).backward()
    params = dict(model.named_parameters())
    ranges, total = flat_layout(params)
    flat = torch.zeros(total)
    for n, p in params.items():
        a, b = ranges[n]
        flat[a:b] = p.grad.reshape(-1)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)   # what TrainStep does over NCCL ...
    flat *= 1.0 / world                           # ... with the 1/world folded into the Adam kernel
    if rank == 0:
        torch.save(flat, out_path)
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_mean_equals_full_batch_gloo():
    """2 ranks x B/2 samples == 1 rank x B samples (no op couples samples): the DP averaging rule of train.py."""
    world, port = 2, _free_port()
    import tempfile
    ctx = mp.get_context("spawn")
    out_path = os.path.join(tempfile.mkdtemp(), "flat.pt")
    procs = [ctx.Process(target=_dp_worker, args=(r, world, port, out_path)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    flat = torch.load(out_path)
    torch.manual_seed(1993)
    model = O.MTnnUNet(1, 1, 3)
    img, mask, onehot, _ = O.synthetic_batch(4, 64, 64)
    logits, outs = model(img)
    seg, cls = O.multitask_criterion(O.DiceLoss(), mask, outs, O.FocalLoss(), onehot, logits, True)
    (0.35 * seg + 0.65 * cls).backward()
    params = dict(model.named_parameters())
    ranges, total = flat_layout(params)
    want = torch.zeros(total)
    for n, p in params.items():
        a, b = ranges[n]
        want[a:b] = p.grad.reshape(-1)
    torch.testing.assert_close(flat, want, rtol=2e-3, atol=1e-6)


def test_step_segments_split_at_bucket_markers():
    """TrainStep._segments: the launch list is cut at every bucket marker, in order, each cut carrying the bucket whose
    all-reduce is forked there; markers that coincide leave an EMPTY segment (not captured, nothing to replay) and the
    tail after the last marker has no bucket."""
    from types import SimpleNamespace
    from multi_task_breast_cancer_b200.train import TrainStep

    def L(kind, bucket=None):
        return SimpleNamespace(kind=kind, bucket=bucket)
    fb = [L("a"), L("b"), L("bucket_ready", 0), L("c"), L("bucket_ready", 1), L("bucket_ready", 2), L("d"), L("e"),
          L("bucket_ready", 3)]
    segs = TrainStep._segments(SimpleNamespace(launches_fb=fb))
    assert [[l.kind for l in ls] for ls, _ in segs] == [["a", "b"], ["c"], [], ["d", "e"], []]
    assert [k for _, k in segs] == [0, 1, 2, 3, None]
    assert sum(len(ls) for ls, _ in segs) == sum(l.kind != "bucket_ready" for l in fb)


def test_bf16_storage_alone_moves_gradients():
    """No kernel involved: the fp32 oracle against itself with nothing but the B200 path's STORAGE rounding switched on
    (oracle/emulation.py).  With a random cotangent on every output (no Dice / InstanceNorm cancellation) the forward
    moves by ~1e-2, yet per-parameter gradients move by 10 % and more: LeakyReLU units whose pre-activation lies within
    the rounding error of 0 take the other slope.  This is why tests/test_grad_wiring_gpu.py checks the CUDA path
    against the storage-emulating oracle (and why round 1's check against the plain oracle could only assert a
    direction)."""
    from oracle.emulation import named_grads, with_bf16_storage
    torch.manual_seed(1993)
    ref = O.MTUNetPlusPlus(in_channels=1, out_channels=1, n_classes=3, deep_supervision=True)
    emu = with_bf16_storage(ref)
    img, *_ = O.synthetic_batch(1, 64, 64)
    outs = []
    for m in (ref, emu):
        g = torch.Generator().manual_seed(11)
        logits, segs = m(img)
        sum((t * torch.randn(t.shape, generator=g)).sum() for t in list(logits) + list(segs)).backward()
        outs.append(segs[-1].detach())
    fwd = ((outs[1] - outs[0]).norm() / outs[0].norm()).item()
    gr, ge = named_grads(ref), named_grads(emu)

    def rel(n):
        return ((ge[n] - gr[n]).norm() / gr[n].norm()).item()
    assert 2e-3 < fwd < 5e-2, fwd                                   # the forward moves by ~1e-2 ...
    assert rel("final_conv_0_4.weight") < 5e-2                      # ... the head's own gradient likewise ...
    assert rel("upcat_0_4.convs.conv_0.conv.weight") > 5e-2         # ... one TwoConv below it: already beyond 5e-2 ...
    assert rel("conv_2_0.convs.conv_0.conv.weight") > 0.1           # ... and the encoder far beyond
    # rounding only the gradient tensors (forward untouched) costs two orders of magnitude less
    emu_g = with_bf16_storage(ref, what=("gy", "ga"))
    ref.zero_grad(set_to_none=True)
    for m in (ref, emu_g):
        g = torch.Generator().manual_seed(11)
        logits, segs = m(img)
        sum((t * torch.randn(t.shape, generator=g)).sum() for t in list(logits) + list(segs)).backward()
    gr, gg = named_grads(ref), named_grads(emu_g)
    e = ((gg["conv_2_0.convs.conv_0.conv.weight"] - gr["conv_2_0.convs.conv_0.conv.weight"]).norm()
         / gr["conv_2_0.convs.conv_0.conv.weight"].norm()).item()
    assert e < 2e-2, e


def test_pixel_pair_operand_is_the_same_convolution():
    """The operand MTBC_JOB_PACK_CONV_PAIR is defined to write (include/mtbc.h; element-wise restatement in
    tools/diag_kernels.py::pair_operand_ref, which the GPU test compares the pack job with) turns the 3x3 conv over pixels
    into a 3x3 conv over pixel PAIRS of the tensors viewed as (N, H, W/2, 2C): checked here in fp32 on the CPU, forward
    and data gradient, for a concat source that starts at channel 8 of a wider weight."""
    import sys
    import torch.nn.functional as F
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from diag_kernels import pair_operand_ref
    torch.manual_seed(0)
    N, C, Co, H, W, c0 = 2, 24, 24, 8, 12, 8
    wfull = torch.randn(Co, c0 + C + 5, 3, 3, dtype=torch.float64)
    w = wfull[:, c0:c0 + C]
    x = torch.randn(N, C, H, W, dtype=torch.float64)

    def pairs(t):      # NCHW -> NCHW of the pair view: channel = parity * C + c, width W/2
        n, c, h, ww = t.shape
        return t.permute(0, 2, 3, 1).reshape(n, h, ww // 2, 2 * c).permute(0, 3, 1, 2)

    def unpairs(t, c):
        n, _, h, w2 = t.shape
        return t.permute(0, 2, 3, 1).reshape(n, h, 2 * w2, c).permute(0, 3, 1, 2)

    wp = pair_operand_ref(wfull, c0, C, 64, 64, 0, 0, C, Co, 0).double()
    assert (wp != 0).sum().item() == (wfull[:, c0:c0 + C] != 0).sum().item() * 2      # every tap appears for both output parities
    y = F.conv2d(pairs(x), wp[:, :2 * Co, :2 * C].reshape(3, 3, 2 * Co, 2 * C).permute(2, 3, 0, 1), padding=1)
    assert torch.allclose(unpairs(y, Co), F.conv2d(x, w, padding=1), atol=1e-6)
    dy = torch.randn(N, Co, H, W, dtype=torch.float64)
    wd = pair_operand_ref(wfull, c0, C, 64, 64, 0, 0, Co, C, 1).double()
    dx = F.conv2d(pairs(dy), wd[:, :2 * C, :2 * Co].reshape(3, 3, 2 * C, 2 * Co).permute(2, 3, 0, 1), padding=1)
    assert torch.allclose(unpairs(dx, C), F.conv_transpose2d(dy, w, padding=1), atol=1e-6)
    # structure the kernel's issue sequence relies on (halo_issue_chunk_pair24): outside dq = 0 only one corner is non-zero
    for dh in range(3):
        left, right = wp[dh * 3 + 0], wp[dh * 3 + 2]
        assert left[Co:, :].abs().sum() == 0 and left[:, :C].abs().sum() == 0      # rows 0..23 x columns 24..47 only
        assert right[:Co, :].abs().sum() == 0 and right[:, C:].abs().sum() == 0     # rows 24..47 x columns 0..23 only
