"""GPU diagnostic: run every tcgen05 GEMM variant (and the streaming kernels) against torch fp32 on the same bf16-rounded
inputs and print an error summary per case.  Never stops at the first failure; meant for `gpurun`.

    python tools/diag_kernels.py [--quick]
"""
import os
import sys
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.nn.functional as F

from multi_task_breast_cancer_b200 import _lib, ops
from multi_task_breast_cancer_b200.ops import Feat, pad32

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
RESULTS = []


def report(name, got, ref, tol=2e-2):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    scale = ref.abs().max().item() + 1e-12
    rel = err.max().item() / scale
    ok = bool(torch.isfinite(got).all().item()) and rel < tol
    RESULTS.append((name, ok, rel))
    print(f"[{'OK ' if ok else 'BAD'}] {name}: max|err|={err.max().item():.4g} ref max={scale:.4g} rel={rel:.3g}", flush=True)
    if not ok:
        idx = torch.nonzero(err > tol * scale)[:6]
        for i in idx:
            t = tuple(i.tolist())
            print("      at", t, "got", got[t].item(), "ref", ref[t].item())
        print("      fraction bad:", (err > tol * scale).float().mean().item())
    return ok


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).to(torch.bfloat16).float()


def run_case(fn, name):
    try:
        fn()
        torch.cuda.synchronize()
    except Exception as e:  # noqa
        RESULTS.append((name, False, float("nan")))
        print(f"[EXC] {name}: {e}")
        traceback.print_exc()


def conv_fwd_case(N, H, W, src_C, Cout, bias=True, stats=True, identity=False):
    name = f"conv3x3_fwd N{N} {H}x{W} src{src_C}->{Cout} bias={bias} stats={stats} id={identity}"

    def fn():
        xs = [rnd(N, c, H, W) for c in src_C]
        Cin = sum(src_C)
        w = rnd(Cout, Cin, 3, 3, scale=0.1)
        if identity:
            w.zero_()
            for i in range(min(Cout, Cin)):
                w[i, i, 1, 1] = 1.0
        b = rnd(Cout) if bias else None
        srcs = [Feat.from_nchw(x) for x in xs]
        out = Feat.empty(N, H, W, Cout)
        offs, ktot = ops.k_offsets(srcs)
        wf = torch.zeros(9, out.Ck, ktot, dtype=torch.bfloat16, device=dev)
        ops.pack_conv_weight(w, src_C, offs, wf, [None] * len(srcs))
        bp = None
        if bias:
            bp = torch.zeros(out.Ck, device=dev)
            bp[:Cout] = b
        ssum = ssq = None
        if stats:
            ssum = torch.zeros(N, out.Cp, device=dev)
            ssq = torch.zeros(N, out.Cp, device=dev)
        op = ops.conv3x3_fwd_op(srcs, wf, out, bias=bp, stat_sum=ssum, stat_sq=ssq)
        op.launch()
        torch.cuda.synchronize()
        ref = F.conv2d(torch.cat(xs, 1), w, b, padding=1)
        report(name, out.to_nchw(), ref)
        if out.Cp > Cout:
            report(name + " [pad lanes zero]", out.t[..., Cout:].float(), torch.zeros_like(out.t[..., Cout:]).float(), tol=1e-6)
        if stats:
            report(name + " [sum]", ssum[:, :Cout], ref.sum((2, 3)), tol=1e-2)
            report(name + " [sumsq]", ssq[:, :Cout], (ref * ref).sum((2, 3)), tol=1e-2)

    run_case(fn, name)


def pair_operand_ref(w, c_begin, c_count, rows, ld, k0, n0, Ks, Np, dgrad):
    """Reference of MTBC_JOB_PACK_CONV_PAIR (include/mtbc.h): the 3x3 conv over pixel pairs, element by element."""
    Cout = w.shape[0]
    wp = torch.zeros(9, rows, ld, device=w.device)
    for dh in (-1, 0, 1):
        for dq in (-1, 0, 1):
            for par in (0, 1):
                for op in (0, 1):
                    dw = 2 * dq + par - op
                    if abs(dw) > 1:
                        continue
                    kh, kw = (1 - dh, 1 - dw) if dgrad else (dh + 1, dw + 1)
                    blk = w[:, c_begin:c_begin + c_count, kh, kw]   # [co][ci]
                    t = (dh + 1) * 3 + dq + 1
                    if dgrad:   # rows = ci, columns = co
                        wp[t, n0 + op * Np:n0 + op * Np + c_count, k0 + par * Ks:k0 + par * Ks + Cout] = blk.t()
                    else:
                        wp[t, n0 + op * Np:n0 + op * Np + Cout, k0 + par * Ks:k0 + par * Ks + c_count] = blk
    return wp


def conv_fwd_pair_case(N, H, W, C, Cout, bias=False):
    """3x3 conv of a dense C-channel tensor through the pixel-pair view (mtbc_conv_gemm_desc.stat_fold): the paired
    operand from the batched pack job against its element-wise definition, the result and the per-channel statistics
    against torch fp32, and against the same conv through the per-pixel view of the same kernel."""
    name = f"conv3x3_fwd PAIR N{N} {H}x{W} {C}->{Cout} bias={bias}"

    def fn():
        import ctypes as Ct
        from multi_task_breast_cancer_b200 import plan as P
        x = rnd(N, C, H, W)
        w = rnd(Cout, C, 3, 3, scale=0.1)
        b = rnd(Cout) if bias else None
        src = Feat.from_nchw(x)
        out = Feat.empty(N, H, W, Cout)
        assert src.Cp == C and out.Cp == Cout, "pair view needs dense tensors"
        sp, outp = ops.pair_view(src), ops.pair_view(out)
        wfp = torch.zeros(9, outp.Ck, sp.Ck, dtype=torch.bfloat16, device=dev)
        jt = P.JobTable()
        jt.add(_lib.JOB_PACK_CONV_PAIR, [Cout, C, 0, C, wfp.shape[1], wfp.shape[2], 0, 0, src.Cp, out.Cp, 0], w, wfp)
        st = Ct.c_void_p(ops.stream_ptr())
        for l in jt.launch():
            l(st)
        torch.cuda.synchronize()
        report(name + " [operand]", wfp, pair_operand_ref(w, 0, C, wfp.shape[1], wfp.shape[2], 0, 0, C, Cout, 0), tol=1e-6)
        bp = None
        if bias:
            bp = torch.zeros(out.Ck, device=dev)
            bp[:Cout] = b
        ssum = torch.zeros(N, out.Cp, device=dev)
        ssq = torch.zeros(N, out.Cp, device=dev)
        op = ops.conv3x3_fwd_op([sp], wfp, outp, bias=bp, stat_sum=ssum, stat_sq=ssq, stat_fold=out.Cp)
        op.launch()
        torch.cuda.synchronize()
        ref = F.conv2d(x, w, b, padding=1)
        report(name, out.to_nchw(), ref)
        report(name + " [sum]", ssum[:, :Cout], ref.sum((2, 3)), tol=1e-2)
        report(name + " [sumsq]", ssq[:, :Cout], (ref * ref).sum((2, 3)), tol=1e-2)
        # the per-pixel view of the same kernel: same products, other summation order inside the fp32 accumulator
        out1 = Feat.empty(N, H, W, Cout)
        offs, ktot = ops.k_offsets([src])
        wf = torch.zeros(9, out1.Ck, ktot, dtype=torch.bfloat16, device=dev)
        ops.pack_conv_weight(w, [C], offs, wf, [None])
        s1 = torch.zeros(N, out1.Cp, device=dev); q1 = torch.zeros(N, out1.Cp, device=dev)
        ops.conv3x3_fwd_op([src], wf, out1, bias=bp, stat_sum=s1, stat_sq=q1).launch()
        torch.cuda.synchronize()
        report(name + " [vs per-pixel view]", out.t, out1.t, tol=8e-3)
        report(name + " [sumsq vs per-pixel view]", ssq, q1, tol=1e-4)

    run_case(fn, name)


def conv_dgrad_case(N, H, W, Cin, Cout, accumulate, dyscale=1.0):
    name = f"conv3x3_dgrad N{N} {H}x{W} {Cin}<-{Cout} acc={accumulate} dyscale={dyscale}"

    def fn():
        w = rnd(Cout, Cin, 3, 3, scale=0.1)
        dy = rnd(N, Cout, H, W, scale=dyscale)
        dyf = Feat.from_nchw(dy)
        dx = Feat.empty(N, H, W, Cin)
        base = rnd(N, Cin, H, W, scale=dyscale)
        if accumulate:
            dx = Feat.from_nchw(base)
        wd = torch.zeros(9, dx.Ck, dyf.Ck, dtype=torch.bfloat16, device=dev)
        wf = torch.zeros(9, dyf.Ck, dx.Ck, dtype=torch.bfloat16, device=dev)
        ops.pack_conv_weight(w, [Cin], [0], wf, [wd])
        op = ops.conv3x3_dgrad_op(dyf, wd, dx, accumulate)
        op.launch()
        torch.cuda.synchronize()
        ref = F.conv_transpose2d(dy, w, padding=1)
        if accumulate:
            ref = ref + base
        report(name, dx.to_nchw(), ref)

    run_case(fn, name)


def conv_dgrad_inbwd_case(N, H, W, C, Cout, affine, slope=0.1, pair=False):
    """Data gradient of conv_1 with the InstanceNorm + LeakyReLU backward sums of its input's producer fused into the
    epilogue (mtbc_conv_gemm_desc.bwd_y), followed by the apply pass with slope = 1: dy of the whole
    conv -> IN -> LeakyReLU chain against torch autograd, and s1 / s2 against their definitions."""
    name = f"conv3x3_dgrad+in_bwd N{N} {H}x{W} {C}<-{Cout} affine={affine}" + (" PAIR" if pair else "")

    def fn():
        y = (rnd(N, C, H, W, scale=2.0) + 0.7).to(torch.bfloat16).float()     # stored pre-norm conv output of layer 0
        yf = Feat.from_nchw(y)
        Cp = yf.Cp
        g = b = gp = bp = None
        if affine:
            g = torch.rand(C, device=dev) + 0.5
            b = torch.randn(C, device=dev) * 0.3
            gp = torch.zeros(yf.Ck, device=dev); gp[:C] = g
            bp = torch.zeros(yf.Ck, device=dev); bp[:C] = b
        ssum = torch.zeros(N, Cp, device=dev); ssq = torch.zeros(N, Cp, device=dev)
        _lib.call("mtbc_in_stats", ops.ptr(yf.t), N, H * W, Cp, ops.ptr(ssum), ops.ptr(ssq), None)
        a = Feat.empty(N, H, W, C)
        mean = torch.zeros(N, Cp, device=dev); rstd = torch.zeros(N, Cp, device=dev)
        _lib.call("mtbc_in_apply", ops.ptr(yf.t), N, H, W, Cp, ops.ptr(ssum), ops.ptr(ssq), ops.ptr(gp), ops.ptr(bp), C,
                  1e-5, slope, ops.ptr(a.t), None, ops.ptr(mean), ops.ptr(rstd), None)
        w = rnd(Cout, C, 3, 3, scale=0.1)
        dy1 = rnd(N, Cout, H, W)
        dyf = Feat.from_nchw(dy1)
        ga = Feat.empty(N, H, W, C)
        wd = torch.zeros(9, ga.Ck, dyf.Ck, dtype=torch.bfloat16, device=dev)
        wf = torch.zeros(9, dyf.Ck, ga.Ck, dtype=torch.bfloat16, device=dev)
        ops.pack_conv_weight(w, [C], [0], wf, [wd])
        s1 = torch.zeros(N, Cp, device=dev); s2 = torch.zeros(N, Cp, device=dev)
        if pair:   # the same launch through the pixel-pair view (dy, the gradient and y as (N, H, W/2, 2C) tensors)
            import ctypes as Ct
            from multi_task_breast_cancer_b200 import plan as P
            dyp, gap, yp = ops.pair_view(dyf), ops.pair_view(ga), ops.pair_view(yf)
            wdp = torch.zeros(9, gap.Ck, dyp.Ck, dtype=torch.bfloat16, device=dev)
            jt = P.JobTable()
            jt.add(_lib.JOB_PACK_CONV_PAIR, [Cout, C, 0, C, wdp.shape[1], wdp.shape[2], 0, 0, dyf.Cp, ga.Cp, 1], w, wdp)
            for l in jt.launch():
                l(Ct.c_void_p(ops.stream_ptr()))
            torch.cuda.synchronize()
            report(name + " [operand]", wdp, pair_operand_ref(w, 0, C, wdp.shape[1], wdp.shape[2], 0, 0, dyf.Cp, ga.Cp, 1), tol=1e-6)
            op = ops.conv3x3_dgrad_op(dyp, wdp, gap, False, bwd_fuse=(yp, mean, rstd, gp, bp, slope), s1=s1, s2=s2,
                                      stat_fold=Cp)
        else:
            op = ops.conv3x3_dgrad_op(dyf, wd, ga, False, bwd_fuse=(yf, mean, rstd, gp, bp, slope), s1=s1, s2=s2)
        op.launch()
        dy0 = Feat.empty(N, H, W, C)
        dg = torch.zeros(C, device=dev) if affine else None
        db = torch.zeros(C, device=dev) if affine else None
        _lib.call("mtbc_in_bwd_apply", ops.ptr(ga.t), ops.ptr(yf.t), N, H * W, Cp, ops.ptr(mean), ops.ptr(rstd),
                  ops.ptr(gp), ops.ptr(bp), 1.0, ops.ptr(s1), ops.ptr(s2), ops.ptr(dy0.t), ops.ptr(dg), ops.ptr(db), C, None)
        torch.cuda.synchronize()
        # reference: autograd through IN -> LeakyReLU -> conv with the cotangent dy1
        yr = y.clone().requires_grad_(True)
        ar = F.leaky_relu(F.instance_norm(yr, weight=g, bias=b, eps=1e-5), slope)
        out = F.conv2d(ar, w, padding=1)
        out.backward(dy1)
        D = F.conv_transpose2d(dy1, w, padding=1)
        xh = (y - y.mean((2, 3), keepdim=True)) * torch.rsqrt(y.var((2, 3), unbiased=False, keepdim=True) + 1e-5)
        z = xh if not affine else xh * g.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)
        gg = torch.where(z > 0, D, slope * D)
        report(name + " gg", ga.to_nchw(), gg, tol=2e-2)
        report(name + " s1", s1[:, :C], gg.sum((2, 3)), tol=2e-3)
        report(name + " s2", s2[:, :C], (gg * xh).sum((2, 3)), tol=2e-3)
        report(name + " dy", dy0.to_nchw(), yr.grad, tol=3e-2)
        if affine:
            report(name + " dgamma", dg, (gg * xh).sum((0, 2, 3)), tol=2e-3)
            report(name + " dbeta", db, gg.sum((0, 2, 3)), tol=2e-3)

    run_case(fn, name)


def conv_dgrad_multi_case(N, H, W, src_C, Cout, acc_flags=None):
    """Fused data gradient: one launch writes the gradient of every concat source (some accumulating)."""
    name = f"conv3x3_dgrad_multi N{N} {H}x{W} {src_C}<-{Cout}"

    def fn():
        Cin = sum(src_C)
        w = rnd(Cout, Cin, 3, 3, scale=0.1)
        dy = rnd(N, Cout, H, W)
        dyf = Feat.from_nchw(dy)
        flags = list(acc_flags) if acc_flags is not None else [i % 2 == 1 for i in range(len(src_C))]
        bases = [rnd(N, c, H, W) for c in src_C]
        dxs = [Feat.from_nchw(b) if f else Feat.empty(N, H, W, c) for b, f, c in zip(bases, flags, src_C)]
        for d, f in zip(dxs, flags):
            if not f:
                d.t.fill_(7.0)  # must be overwritten, pad lanes included
        rows = sum(d.Ck for d in dxs)
        wd_all = torch.zeros(9, rows, dyf.Ck, dtype=torch.bfloat16, device=dev)
        c0 = r0 = 0
        for c, d in zip(src_C, dxs):
            _lib.call("mtbc_pack_conv_weight", ops.ptr(w), Cout, Cin, 3, c0, c, None, 0, 0, 0, ops.ptr(wd_all[0, r0:]),
                      rows, dyf.Ck, None)
            c0 += c
            r0 += d.Ck
        op = ops.conv3x3_dgrad_multi_op(dyf, wd_all, dxs, flags)
        op.launch()
        torch.cuda.synchronize()
        ref = F.conv_transpose2d(dy, w, padding=1)
        c0 = 0
        for i, (c, d, b, f) in enumerate(zip(src_C, dxs, bases, flags)):
            r = ref[:, c0:c0 + c] + (b if f else 0)
            report(f"{name} [src {i} acc={f}]", d.to_nchw(), r)
            if d.Cp > c and not f:
                report(f"{name} [src {i} pad lanes zero]", d.t[..., c:].float(), torch.zeros_like(d.t[..., c:].float()) + 0, tol=1e-6)
            c0 += c

    run_case(fn, name)


def conv_wgrad_case(N, H, W, Cin, Cout, splits=0):
    name = f"conv3x3_wgrad N{N} {H}x{W} {Cin}->{Cout} splits={splits}"

    def fn():
        x = rnd(N, Cin, H, W)
        dy = rnd(N, Cout, H, W)
        xf, dyf = Feat.from_nchw(x), Feat.from_nchw(dy)
        acc = torch.zeros(9, dyf.Ck, xf.Ck, device=dev)
        op = ops.conv3x3_wgrad_op(xf, dyf, acc, 0, splits=splits)
        op.launch()
        torch.cuda.synchronize()
        grad = torch.zeros(Cout, Cin, 3, 3, device=dev)
        _lib.call("mtbc_unpack_conv_wgrad", ops.ptr(acc), acc.shape[1], acc.shape[2], 0, ops.ptr(grad), Cout, Cin, 3, 0,
                  Cin, 0, None)
        torch.cuda.synchronize()
        ref = torch.nn.grad.conv2d_weight(x, (Cout, Cin, 3, 3), dy, padding=1)
        report(name, grad, ref)

    run_case(fn, name)


def conv_wgrad_multi_case(N, H, W, src_C, Cout):
    """Weight gradient over a folded concat, all sources in one launch."""
    name = f"conv3x3_wgrad_multi N{N} {H}x{W} {src_C}->{Cout}"

    def fn():
        xs = [rnd(N, c, H, W) for c in src_C]
        dy = rnd(N, Cout, H, W)
        xfs, dyf = [Feat.from_nchw(x) for x in xs], Feat.from_nchw(dy)
        offs, ktot = ops.k_offsets(xfs)
        acc = torch.zeros(9, dyf.Ck, ktot, device=dev)
        ops.conv3x3_wgrad_multi_op(xfs, dyf, acc, offs).launch()
        torch.cuda.synchronize()
        Cin = sum(src_C)
        grad = torch.zeros(Cout, Cin, 3, 3, device=dev)
        c0 = 0
        for c, off in zip(src_C, offs):
            _lib.call("mtbc_unpack_conv_wgrad", ops.ptr(acc), acc.shape[1], acc.shape[2], off, ops.ptr(grad), Cout, Cin, 3,
                      c0, c, 0, None)
            c0 += c
        torch.cuda.synchronize()
        ref = torch.nn.grad.conv2d_weight(torch.cat(xs, 1), (Cout, Cin, 3, 3), dy, padding=1)
        report(name, grad, ref)
        # nothing may leak into the pad columns / rows of the accumulator
        mask = torch.ones_like(acc, dtype=torch.bool)
        for c, off in zip(src_C, offs):
            mask[:, :Cout, off:off + c] = False
        if mask.any():
            report(name + " [pad entries zero]", acc[mask], torch.zeros_like(acc[mask]), tol=1e-6)

    run_case(fn, name)


def convT_case(N, H, W, Cin, Cout, k=2):
    name = f"convT k{k} N{N} {H}x{W} {Cin}->{Cout}"

    def fn():
        x = rnd(N, Cin, H, W)
        w = rnd(Cin, Cout, k, k, scale=0.1)
        b = rnd(Cout)
        xf = Feat.from_nchw(x)
        out = Feat.empty(N, H * k, W * k, Cout)
        wf = torch.zeros(1, k * k * out.Ck, xf.Ck, dtype=torch.bfloat16, device=dev)
        wd = torch.zeros(k * k, xf.Ck, out.Ck, dtype=torch.bfloat16, device=dev)
        ops.pack_convT_weight(w, out.Ck, wf, wd)
        bp = torch.zeros(out.Ck, device=dev)
        bp[:Cout] = b
        ops.convT_fwd_op(xf, wf, out, k, bp).launch()
        torch.cuda.synchronize()
        ref = F.conv_transpose2d(x, w, b, stride=k)
        report(name + " fwd", out.to_nchw(), ref)
        # dgrad
        dout = rnd(N, Cout, H * k, W * k)
        df = Feat.from_nchw(dout)
        dx = Feat.empty(N, H, W, Cin)
        ops.convT_dgrad_op(df, wd, dx, k, False).launch()
        torch.cuda.synchronize()
        refdx = F.conv2d(dout, w, stride=k)
        report(name + " dgrad", dx.to_nchw(), refdx)
        # wgrad
        acc = torch.zeros(k * k, out.Ck, xf.Ck, device=dev)
        ops.convT_wgrad_op(xf, df, acc, k).launch()
        grad = torch.zeros(Cin, Cout, k, k, device=dev)
        _lib.call("mtbc_unpack_convT_wgrad", ops.ptr(acc), acc.shape[0] * acc.shape[1], acc.shape[2], ops.ptr(grad), Cin,
                  Cout, k, 0, None)
        torch.cuda.synchronize()
        xr = x.clone().requires_grad_(False)
        wr = w.clone().requires_grad_(True)
        F.conv_transpose2d(xr, wr, None, stride=k).backward(dout)
        report(name + " wgrad", grad, wr.grad)

    run_case(fn, name)


def convT_bwd_case(N, H, W, Cin, Cout, accumulate, bias=True):
    """Fused backward of ConvTranspose2d(k = s = 2): data, weight and bias gradient from one pass over dout."""
    name = f"convT_bwd N{N} {H}x{W} {Cin}->{Cout} acc={accumulate} bias={bias}"

    def fn():
        k = 2
        x = rnd(N, Cin, H, W)
        w = rnd(Cin, Cout, k, k, scale=0.1)
        xf = Feat.from_nchw(x)
        dout = rnd(N, Cout, H * k, W * k)
        df = Feat.from_nchw(dout)
        wf = torch.zeros(1, k * k * df.Ck, xf.Ck, dtype=torch.bfloat16, device=dev)
        wd = torch.zeros(k * k, xf.Ck, df.Ck, dtype=torch.bfloat16, device=dev)
        ops.pack_convT_weight(w, df.Ck, wf, wd)
        base = rnd(N, Cin, H, W)
        dx = Feat.from_nchw(base) if accumulate else Feat.empty(N, H, W, Cin)
        acc = torch.zeros(k * k, df.Ck, xf.Ck, device=dev)
        dbias = torch.full((Cout,), 0.5, device=dev) if bias else None     # accumulated into
        op = ops.convT_bwd_op(xf, df, wd, acc, dbias, dx, accumulate)
        op.launch()
        grad = torch.zeros(Cin, Cout, k, k, device=dev)
        _lib.call("mtbc_unpack_convT_wgrad", ops.ptr(acc), acc.shape[0] * acc.shape[1], acc.shape[2], ops.ptr(grad), Cin,
                  Cout, k, 0, None)
        torch.cuda.synchronize()
        refdx = F.conv2d(dout, w, stride=k)
        if accumulate:
            refdx = refdx + base
        report(name + " dx", dx.to_nchw(), refdx)
        wr = w.clone().requires_grad_(True)
        F.conv_transpose2d(x, wr, None, stride=k).backward(dout)
        report(name + " dw", grad, wr.grad, tol=2e-3)
        if bias:
            report(name + " dbias", dbias, dout.sum((0, 2, 3)) + 0.5, tol=1e-4)
        # a second launch accumulates the parameter gradients (the accumulators are only zeroed by the caller)
        op.launch()
        torch.cuda.synchronize()
        if bias:
            report(name + " dbias twice", dbias, 2 * dout.sum((0, 2, 3)) + 0.5, tol=1e-4)

    run_case(fn, name)


def first_conv_case(N, H, W, Cout):
    name = f"conv_first N{N} {H}x{W} 1->{Cout}"

    def fn():
        x = torch.randint(0, 256, (N, 1, H, W), device=dev).float()
        w = rnd(Cout, 1, 3, 3, scale=0.1)
        b = rnd(Cout)
        out = Feat.empty(N, H, W, Cout)
        ssum = torch.zeros(N, out.Cp, device=dev)
        ssq = torch.zeros(N, out.Cp, device=dev)
        _lib.call("mtbc_conv_first_fwd", ops.ptr(x), N, 1, H, W, ops.ptr(w), ops.ptr(b), Cout, ops.ptr(out.t), out.Cp,
                  ops.ptr(ssum), ops.ptr(ssq), None, None)
        torch.cuda.synchronize()
        ref = F.conv2d(x, w, b, padding=1)
        report(name + " fwd", out.to_nchw(), ref)
        report(name + " sum", ssum[:, :Cout], ref.sum((2, 3)), tol=1e-3)
        report(name + " sumsq", ssq[:, :Cout], (ref * ref).sum((2, 3)), tol=1e-3)
        # centred storage: y - mean_(h,w)(y), statistics of the centred values
        ssum.zero_(); ssq.zero_()
        xs = torch.zeros(N, 9, device=dev)
        _lib.call("mtbc_conv_first_fwd", ops.ptr(x), N, 1, H, W, ops.ptr(w), ops.ptr(b), Cout, ops.ptr(out.t), out.Cp,
                  ops.ptr(ssum), ops.ptr(ssq), ops.ptr(xs), None)
        torch.cuda.synchronize()
        refc = ref - ref.mean((2, 3), keepdim=True)
        report(name + " fwd centred", out.to_nchw(), refc)
        report(name + " sumsq centred", ssq[:, :Cout], (refc * refc).sum((2, 3)), tol=1e-3)
        report(name + " mean centred (+1)", ssum[:, :Cout] / (H * W) + 1.0, torch.ones(N, Cout, device=dev), tol=1e-2)
        dy = rnd(N, Cout, H, W)
        dyf = Feat.from_nchw(dy)
        dw = torch.zeros(Cout, 1, 3, 3, device=dev)
        _lib.call("mtbc_conv_first_wgrad", ops.ptr(x), N, 1, H, W, ops.ptr(dyf.t), dyf.Cp, Cout, ops.ptr(dw), None)
        torch.cuda.synchronize()
        refw = torch.nn.grad.conv2d_weight(x, (Cout, 1, 3, 3), dy, padding=1)
        report(name + " wgrad", dw, refw)

    run_case(fn, name)


def norm_case(N, H, W, C, affine, pool, slope=0.1):
    name = f"in_lrelu N{N} {H}x{W} C{C} affine={affine} pool={pool}"

    def fn():
        y = rnd(N, C, H, W, scale=3.0) + 1.5
        y = y.to(torch.bfloat16).float()
        yf = Feat.from_nchw(y)
        Cp = yf.Cp
        g = b = gp = bp = None
        if affine:
            g = torch.rand(C, device=dev) + 0.5
            b = torch.randn(C, device=dev) * 0.3
            gp = torch.zeros(Cp, device=dev); gp[:C] = g
            bp = torch.zeros(Cp, device=dev); bp[:C] = b
        ssum = torch.zeros(N, Cp, device=dev); ssq = torch.zeros(N, Cp, device=dev)
        _lib.call("mtbc_in_stats", ops.ptr(yf.t), N, H * W, Cp, ops.ptr(ssum), ops.ptr(ssq), None)
        a = Feat.empty(N, H, W, C)
        pooled = Feat.empty(N, H // 2, W // 2, C) if pool else None
        mean = torch.zeros(N, Cp, device=dev); rstd = torch.zeros(N, Cp, device=dev)
        _lib.call("mtbc_in_apply", ops.ptr(yf.t), N, H, W, Cp, ops.ptr(ssum), ops.ptr(ssq), ops.ptr(gp), ops.ptr(bp), C,
                  1e-5, slope, ops.ptr(a.t), None if pooled is None else ops.ptr(pooled.t), ops.ptr(mean), ops.ptr(rstd),
                  None)
        torch.cuda.synchronize()
        yr = y.clone().requires_grad_(True)
        ref = F.leaky_relu(F.instance_norm(yr, weight=g, bias=b, eps=1e-5), slope)
        report(name + " fwd", a.to_nchw(), ref.detach())
        if pool:
            report(name + " pooled", pooled.to_nchw(), F.max_pool2d(a.to_nchw(), 2))
        # backward
        dA = rnd(N, C, H, W)
        ref.backward(dA)
        dAf = Feat.from_nchw(dA)
        s1 = torch.zeros(N, Cp, device=dev); s2 = torch.zeros(N, Cp, device=dev)
        _lib.call("mtbc_in_bwd_reduce", ops.ptr(dAf.t), ops.ptr(yf.t), N, H * W, Cp, ops.ptr(mean), ops.ptr(rstd),
                  ops.ptr(gp), ops.ptr(bp), slope, ops.ptr(s1), ops.ptr(s2), None)
        dy = Feat.empty(N, H, W, C)
        dg = torch.zeros(C, device=dev) if affine else None
        db = torch.zeros(C, device=dev) if affine else None
        _lib.call("mtbc_in_bwd_apply", ops.ptr(dAf.t), ops.ptr(yf.t), N, H * W, Cp, ops.ptr(mean), ops.ptr(rstd),
                  ops.ptr(gp), ops.ptr(bp), slope, ops.ptr(s1), ops.ptr(s2), ops.ptr(dy.t), ops.ptr(dg), ops.ptr(db), C,
                  None)
        torch.cuda.synchronize()
        report(name + " bwd dy", dy.to_nchw(), yr.grad, tol=3e-2)
        # the single-call backward (one cooperative launch with an L2-resident second pass when the planes exceed L2)
        s1b = torch.zeros(N, Cp, device=dev); s2b = torch.zeros(N, Cp, device=dev)
        cnt = torch.zeros(N, dtype=torch.int32, device=dev)
        dy2 = Feat.empty(N, H, W, C)
        dg2 = torch.zeros(C, device=dev) if affine else None
        db2 = torch.zeros(C, device=dev) if affine else None
        _lib.call("mtbc_in_bwd", ops.ptr(dAf.t), ops.ptr(yf.t), N, H * W, Cp, ops.ptr(mean), ops.ptr(rstd), ops.ptr(gp),
                  ops.ptr(bp), slope, ops.ptr(s1b), ops.ptr(s2b), ops.ptr(dy2.t), ops.ptr(dg2), ops.ptr(db2), C,
                  ops.ptr(cnt), None)
        torch.cuda.synchronize()
        report(name + " bwd dy (mtbc_in_bwd)", dy2.to_nchw(), yr.grad, tol=3e-2)
        report(name + " s1 (mtbc_in_bwd)", s1b, s1, tol=1e-4)
        report(name + " s2 (mtbc_in_bwd)", s2b, s2, tol=1e-4)
        if affine:
            report(name + " dgamma (mtbc_in_bwd)", dg2, dg, tol=1e-4)
            report(name + " dbeta (mtbc_in_bwd)", db2, db, tol=1e-4)
        if affine:
            # reference affine grads
            yr2 = y.clone()
            gg = g.clone().requires_grad_(True); bb = b.clone().requires_grad_(True)
            F.leaky_relu(F.instance_norm(yr2, weight=gg, bias=bb, eps=1e-5), slope).backward(dA)
            report(name + " dgamma", dg, gg.grad, tol=2e-2)
            report(name + " dbeta", db, bb.grad, tol=2e-2)
        if pool:
            dP = rnd(N, C, H // 2, W // 2)
            ar = a.to_nchw().requires_grad_(True)
            F.max_pool2d(ar, 2).backward(dP)
            dPf = Feat.from_nchw(dP)
            dA2 = Feat.empty(N, H, W, C)
            _lib.call("mtbc_maxpool2_bwd", ops.ptr(a.t), ops.ptr(dPf.t), N, H, W, Cp, ops.ptr(dA2.t), 0, None)
            torch.cuda.synchronize()
            report(name + " pool bwd", dA2.to_nchw(), ar.grad)

    run_case(fn, name)


def param_jobs_case(Cout, src_C):
    """Batched parameter jobs (one launch: tiled weight pack + weight-gradient unpack) vs the stand-alone kernels."""
    name = f"param_jobs Cout{Cout} src{src_C}"

    def fn():
        import ctypes as C
        from multi_task_breast_cancer_b200 import plan as P
        Cin = sum(src_C)
        w = rnd(Cout, Cin, 3, 3, scale=0.5)
        cks = [ops.pad32(c) for c in src_C]
        offs, o = [], 0
        for ck in cks:
            a = 64 if ck % 64 == 0 else 32
            o = (o + a - 1) // a * a
            offs.append(o); o += ck
        ktot, Ck = o, ops.pad32(Cout)
        wf_ref = torch.zeros(9, Ck, ktot, dtype=torch.bfloat16, device=dev)
        wd_ref = [torch.zeros(9, ck, Ck, dtype=torch.bfloat16, device=dev) for ck in cks]
        ops.pack_conv_weight(w, src_C, offs, wf_ref, wd_ref)
        wf = torch.zeros_like(wf_ref)
        wd = [torch.zeros_like(t) for t in wd_ref]
        jt = P.JobTable()
        c0 = 0
        for cs, off, d in zip(src_C, offs, wd):
            jt.add(_lib.JOB_PACK_CONV, [Cout, Cin, 3, c0, cs, wf.shape[1], wf.shape[2], off, d.shape[1], d.shape[2]], w, wf, d)
            c0 += cs
        acc = rnd(9, Ck, ktot)
        grad_ref = torch.full((Cout, Cin, 3, 3), 0.25, device=dev)
        grad = grad_ref.clone()
        c0 = 0
        for cs, off in zip(src_C, offs):
            _lib.call("mtbc_unpack_conv_wgrad", ops.ptr(acc), Ck, ktot, off, ops.ptr(grad_ref), Cout, Cin, 3, c0, cs, 1, None)
            jt.add(_lib.JOB_UNPACK_CONV, [Ck, ktot, off, Cout, Cin, 3, c0, cs, 1], acc, grad)
            c0 += cs
        st = C.c_void_p(ops.stream_ptr())
        for l in jt.launch():
            l(st)
        torch.cuda.synchronize()
        report(name + " wf", wf, wf_ref, tol=1e-6)
        for i, (a, b) in enumerate(zip(wd, wd_ref)):
            report(name + f" wd[{i}]", a, b, tol=1e-6)
        report(name + " unpack(add)", grad, grad_ref, tol=1e-6)

    run_case(fn, name)


def chansum_case(N, H, W, C):
    """Per-channel sum over all pixels (bias gradient of the transposed convolutions)."""
    name = f"channel_sum N{N} {H}x{W} C{C}"

    def fn():
        t = rnd(N, C, H, W).to(torch.bfloat16).float()
        tf = Feat.from_nchw(t)
        out = torch.full((C,), 0.5, device=dev)
        _lib.call("mtbc_channel_sum", ops.ptr(tf.t), N * H * W, tf.Cp, C, ops.ptr(out), 1, None)
        torch.cuda.synchronize()
        report(name, out, t.sum(dim=(0, 2, 3)) + 0.5, tol=1e-4)

    run_case(fn, name)


def head1x1_case(N, H, W, C):
    name = f"head1x1 N{N} {H}x{W} C{C}"

    def fn():
        a = rnd(N, C, H, W)
        w = rnd(1, C, 1, 1, scale=0.2)
        b = rnd(1)
        af = Feat.from_nchw(a)
        logits = torch.zeros(N, 1, H, W, device=dev)
        _lib.call("mtbc_head1x1_fwd", ops.ptr(af.t), N * H * W, af.Cp, C, ops.ptr(w), ops.ptr(b), ops.ptr(logits), None)
        ar = a.clone().requires_grad_(True); wr = w.clone().requires_grad_(True); br = b.clone().requires_grad_(True)
        ref = F.conv2d(ar, wr, br)
        report(name + " fwd", logits, ref.detach(), tol=1e-4)
        dl = rnd(N, 1, H, W)
        ref.backward(dl)
        base = rnd(N, C, H, W)
        dA = Feat.from_nchw(base)
        dw = torch.zeros(C, device=dev); db = torch.zeros(1, device=dev)
        _lib.call("mtbc_head1x1_bwd", ops.ptr(af.t), ops.ptr(dl), N * H * W, af.Cp, C, ops.ptr(w), ops.ptr(dA.t), 1,
                  ops.ptr(dw), ops.ptr(db), None)
        report(name + " dA(acc)", dA.to_nchw(), ar.grad + base, tol=1e-2)
        report(name + " dw", dw, wr.grad.flatten(), tol=1e-4)
        report(name + " db", db, br.grad, tol=1e-4)
        if af.Cp > C:
            report(name + " dA pad", dA.t[..., C:].float(), torch.zeros_like(dA.t[..., C:]).float(), tol=1e-6)

    run_case(fn, name)


def dshead_case(N, H, W, C, k):
    name = f"dshead N{N} {H}x{W} C{C} k{k}"

    def fn():
        a = rnd(N, C, H, W)
        wt = rnd(C, C, k, k, scale=0.1); bt = rnd(C, scale=0.1); w1 = rnd(1, C, 1, 1, scale=0.2); b1 = rnd(1)
        af = Feat.from_nchw(a)
        wc = torch.zeros(C, k * k, device=dev); bc = torch.zeros(1, device=dev)
        _lib.call("mtbc_dshead_compose", ops.ptr(wt), ops.ptr(bt), ops.ptr(w1), ops.ptr(b1), C, k, ops.ptr(wc), ops.ptr(bc), None)
        logits = torch.zeros(N, 1, H * k, W * k, device=dev)
        _lib.call("mtbc_dshead_fwd", ops.ptr(af.t), N, H, W, af.Cp, C, k, ops.ptr(wc), ops.ptr(bc), ops.ptr(logits), None)
        ps = [t.clone().requires_grad_(True) for t in (a, wt, bt, w1, b1)]
        ref = F.conv2d(F.conv_transpose2d(ps[0], ps[1], ps[2], stride=k), ps[3], ps[4])
        report(name + " fwd", logits, ref.detach(), tol=1e-4)
        dl = rnd(N, 1, H * k, W * k)
        ref.backward(dl)
        dA = Feat.empty(N, H, W, C)
        nparts = int(_lib.load().mtbc_dshead_bwd_parts(N, H, W))
        # partial rows are written whole by the kernel: poison them to prove nothing relies on a zeroed buffer
        dwc = torch.full((nparts + 1, C * k * k), float("nan"), device=dev); dbc = torch.full((nparts + 1,), float("nan"), device=dev)
        _lib.call("mtbc_dshead_bwd", ops.ptr(af.t), ops.ptr(dl), N, H, W, af.Cp, C, k, ops.ptr(wc), ops.ptr(dA.t), 0,
                  ops.ptr(dwc), ops.ptr(dbc), nparts, None)
        g = [torch.zeros_like(t) for t in (wt, bt, w1, b1)]
        _lib.call("mtbc_dshead_decompose", ops.ptr(dwc), ops.ptr(dbc), nparts, ops.ptr(wt), ops.ptr(bt), ops.ptr(w1), C, k,
                  ops.ptr(g[0]), ops.ptr(g[1]), ops.ptr(g[2]), ops.ptr(g[3]), None)
        report(name + " dA", dA.to_nchw(), ps[0].grad, tol=1e-2)
        for nm, mine, r in zip(("dwt", "dbt", "dw1", "db1"), g, ps[1:]):
            report(name + " " + nm, mine, r.grad, tol=2e-4)

    run_case(fn, name)


def gap_fc_case(N, H, W, Fd, Hd=256, K=3):
    name = f"gap_fc N{N} {H}x{W} F{Fd}"

    def fn():
        a = rnd(N, Fd, H, W)
        w1 = rnd(Hd, Fd, scale=0.05); b1 = rnd(Hd, scale=0.1); w2 = rnd(K, Hd, scale=0.1); b2 = rnd(K, scale=0.1)
        af = Feat.from_nchw(a)
        gap = torch.zeros(N, Fd, device=dev); hid = torch.zeros(N, Hd, device=dev); logits = torch.zeros(N, K, device=dev)
        _lib.call("mtbc_gap_fc_fwd", ops.ptr(af.t), N, H * W, af.Cp, Fd, ops.ptr(w1), ops.ptr(b1), Hd, ops.ptr(w2),
                  ops.ptr(b2), K, ops.ptr(gap), ops.ptr(hid), ops.ptr(logits), None)
        ps = [t.clone().requires_grad_(True) for t in (a, w1, b1, w2, b2)]
        ref = F.linear(F.relu(F.linear(ps[0].mean((2, 3)), ps[1], ps[2])), ps[3], ps[4])
        report(name + " fwd", logits, ref.detach(), tol=1e-4)
        dl = rnd(N, K)
        ref.backward(dl)
        dA = Feat.empty(N, H, W, Fd)
        g = [torch.zeros_like(t) for t in (w1, b1, w2, b2)]
        _lib.call("mtbc_gap_fc_bwd", ops.ptr(dl), N, H * W, af.Cp, Fd, ops.ptr(w1), Hd, ops.ptr(w2), K, ops.ptr(gap),
                  ops.ptr(hid), ops.ptr(dA.t), 0, ops.ptr(g[0]), ops.ptr(g[1]), ops.ptr(g[2]), ops.ptr(g[3]),
                  ops.ptr(torch.zeros(N, Fd, device=dev)), None)
        report(name + " dA", dA.to_nchw(), ps[0].grad, tol=1e-2)
        for nm, mine, r in zip(("dw1", "db1", "dw2", "db2"), g, ps[1:]):
            report(name + " " + nm, mine, r.grad, tol=2e-4)

    run_case(fn, name)


def flat_fc_case(N, H, W, C, Hd=256, K=3):
    name = f"flat_fc N{N} {H}x{W} C{C}"

    def fn():
        a = rnd(N, C, H, W)
        Fd = C * H * W
        w1 = rnd(Hd, Fd, scale=0.01); b1 = rnd(Hd, scale=0.1); w2 = rnd(K, Hd, scale=0.1); b2 = rnd(K, scale=0.1)
        af = Feat.from_nchw(a)
        hid = torch.zeros(N, Hd, device=dev); logits = torch.zeros(N, K, device=dev)
        _lib.call("mtbc_flat_fc_fwd", ops.ptr(af.t), N, H * W, af.Cp, C, ops.ptr(w1), ops.ptr(b1), Hd, ops.ptr(w2),
                  ops.ptr(b2), K, ops.ptr(hid), ops.ptr(logits), None)
        ps = [t.clone().requires_grad_(True) for t in (a, w1, b1, w2, b2)]
        ref = F.linear(F.relu(F.linear(ps[0].flatten(1), ps[1], ps[2])), ps[3], ps[4])
        report(name + " fwd", logits, ref.detach(), tol=1e-3)
        dl = rnd(N, K)
        ref.backward(dl)
        dA = Feat.empty(N, H, W, C)
        g = [torch.zeros_like(t) for t in (w1, b1, w2, b2)]
        scratch = torch.zeros(N, Hd, device=dev)
        _lib.call("mtbc_flat_fc_bwd", ops.ptr(af.t), ops.ptr(dl), N, H * W, af.Cp, C, ops.ptr(w1), Hd, ops.ptr(w2), K,
                  ops.ptr(hid), ops.ptr(dA.t), 0, ops.ptr(g[0]), ops.ptr(g[1]), ops.ptr(g[2]), ops.ptr(g[3]),
                  ops.ptr(scratch), None)
        report(name + " dA", dA.to_nchw(), ps[0].grad, tol=1e-2)
        for nm, mine, r in zip(("dw1", "db1", "dw2", "db2"), g, ps[1:]):
            report(name + " " + nm, mine, r.grad, tol=1e-3)

    run_case(fn, name)


def loss_cases():
    import ctypes as C
    name = "losses"

    def fn():
        N, H, W, K = 5, 24, 20, 3
        x = torch.randn(N, 1, H, W, device=dev) * 2
        t = (torch.rand(N, 1, H, W, device=dev) > 0.7).float()
        t[2] = 0
        xr = x.clone().requires_grad_(True)
        p = torch.sigmoid(xr)
        f = 1 - (2 * (t * p).sum((1, 2, 3)) + 1) / ((t * t).sum((1, 2, 3)) + (p * p).sum((1, 2, 3)) + 1)
        ref = f.mean()
        ref.backward()
        sums = torch.zeros(N, 3, device=dev); loss = torch.zeros(1, device=dev); d = torch.zeros_like(x)
        _lib.call("mtbc_dice_sums", ops.ptr(x), ops.ptr(t), N, H * W, ops.ptr(sums), None)
        _lib.call("mtbc_dice_finalize", ops.ptr(sums), N, ops.ptr(loss), None)
        gs = torch.full((1,), 0.35, device=dev)
        _lib.call("mtbc_dice_bwd", ops.ptr(x), ops.ptr(t), N, H * W, ops.ptr(sums), ops.ptr(gs), C.c_float(0.5), ops.ptr(d), None)
        report("dice fwd", loss, ref.detach().reshape(1), tol=1e-5)
        report("dice bwd", d, xr.grad * 0.35 * 0.5, tol=1e-4)
        # focal
        lg = torch.randn(7, K, device=dev)
        tg = F.one_hot(torch.arange(7, device=dev) % K, K).float()
        lr_ = lg.clone().requires_grad_(True)
        ce = F.cross_entropy(lr_, tg, reduction="none")
        fr = ((1 - torch.exp(-ce)) ** 2 * ce).mean()
        fr.backward()
        fl = torch.zeros(1, device=dev); dg = torch.zeros_like(lg)
        _lib.call("mtbc_focal_fwd", ops.ptr(lg), ops.ptr(tg), 7, K, C.c_float(1.0), C.c_float(2.0), ops.ptr(fl), None)
        _lib.call("mtbc_focal_bwd", ops.ptr(lg), ops.ptr(tg), 7, K, C.c_float(1.0), C.c_float(2.0), None, C.c_float(0.65),
                  ops.ptr(dg), None)
        report("focal fwd", fl, fr.detach().reshape(1), tol=1e-5)
        report("focal bwd", dg, lr_.grad * 0.65, tol=1e-4)
        zl = torch.zeros(2, 3, device=dev); zt = torch.tensor([[1., 0, 0], [0, 0, 1.]], device=dev)
        _lib.call("mtbc_focal_fwd", ops.ptr(zl), ops.ptr(zt), 2, 3, C.c_float(1.0), C.c_float(2.0), ops.ptr(fl), None)
        report("focal anchor (2/3)^2 ln3", fl, torch.tensor([0.488272], device=dev), tol=1e-5)
        # multitask mix
        dl_ = torch.tensor([0.5, 0.6, 0.7, 0.8], device=dev); out = torch.zeros(4, device=dev)
        _lib.call("mtbc_multitask_loss", ops.ptr(dl_), 4, 1, ops.ptr(fl), C.c_float(0.35), ops.ptr(out), None)
        seg = 0.5 + 0.6 / 2 + 0.7 / 3 + 0.8 / 4
        report("multitask mix", out[:3], torch.tensor([0.35 * seg + 0.65 * 0.488272, seg, 0.488272], device=dev), tol=1e-5)
        # adam vs torch
        prm = torch.randn(1000, device=dev); g = torch.randn(1000, device=dev) * 1e-3
        pt = prm.clone().requires_grad_(True); opt = torch.optim.Adam([pt], lr=1e-4, eps=1e-4)
        m = torch.zeros_like(prm); v = torch.zeros_like(prm); mine = prm.clone()
        for step in range(1, 4):
            pt.grad = g.clone() * step; opt.step()
            gg = g * step * 2.0
            _lib.call("mtbc_adam_step", ops.ptr(mine), ops.ptr(gg), ops.ptr(m), ops.ptr(v), 1000, C.c_float(1e-4),
                      C.c_float(0.9), C.c_float(0.999), C.c_float(1e-4), C.c_float(0.5), step, None)
        report("adam 3 steps (grad_scale 0.5)", mine - prm, pt.detach() - prm, tol=2e-3)

    run_case(fn, name)


GROUPS = ["fwd_identity", "fwd", "dgrad", "wgrad", "convT", "first", "norm", "heads", "loss"]


def run_group(group):
    lib = _lib.load()
    _lib.check(lib.mtbc_device_check(), "device_check")
    torch.manual_seed(0)
    if group == "fwd_identity":
        conv_fwd_case(2, 16, 16, [64], 64, bias=False, stats=False, identity=True)
        conv_fwd_case(2, 16, 16, [32], 32, bias=False, stats=False, identity=True)
    elif group == "fwd":
        fwd_cases()
    elif group == "dgrad":
        dgrad_cases()
    elif group == "wgrad":
        wgrad_cases()
    elif group == "convT":
        convT_cases()
    elif group == "first":
        first_conv_case(2, 32, 32, 24)
        first_conv_case(2, 32, 32, 32)
    elif group == "heads":
        head1x1_case(2, 32, 32, 24)
        head1x1_case(2, 16, 16, 16)
        dshead_case(2, 8, 8, 128, 8)
        dshead_case(2, 16, 16, 64, 4)
        dshead_case(3, 16, 16, 32, 2)
        dshead_case(2, 24, 40, 24, 2)
        dshead_case(3, 32, 32, 48, 4)
        gap_fc_case(3, 4, 4, 512)
        flat_fc_case(3, 16, 16, 256)
    elif group == "loss":
        loss_cases()
    elif group == "norm":
        norm_case(2, 32, 32, 24, True, True)
        norm_case(2, 16, 16, 96, True, False)
        norm_case(2, 16, 16, 320, False, True, slope=0.01)
        norm_case(3, 8, 8, 512, False, False, slope=0.01)
    bad = [r for r in RESULTS if not r[1]]
    print(f"==== group {group}: {len(RESULTS) - len(bad)}/{len(RESULTS)} OK", flush=True)
    return 1 if bad else 0


def main():
    import subprocess
    if len(sys.argv) > 2 and sys.argv[1] == "--group":
        return run_group(sys.argv[2])
    groups = sys.argv[1:] or GROUPS
    rc = 0
    for g in groups:
        print(f"\n######## group {g}", flush=True)
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--group", g], timeout=300)
            if p.returncode != 0:
                rc = 1
                print(f"group {g} exit code {p.returncode}", flush=True)
        except subprocess.TimeoutExpired:
            rc = 1
            print(f"group {g} TIMED OUT", flush=True)
    return rc


def fwd_cases():
    conv_fwd_case(2, 16, 16, [64], 64, bias=False, stats=False)
    conv_fwd_case(2, 16, 16, [32], 32, bias=False, stats=False)
    conv_fwd_case(2, 32, 32, [24, 24, 48], 24)
    conv_fwd_case(2, 32, 32, [48, 48, 48], 48)
    conv_fwd_case(2, 16, 16, [96, 96], 96)
    conv_fwd_case(3, 16, 16, [128], 512, stats=True)
    conv_fwd_case(2, 16, 16, [384, 384, 384], 512)
    conv_fwd_case(4, 8, 8, [320], 320, stats=False)
    conv_fwd_case(9, 4, 4, [64], 64, stats=False)
    conv_fwd_case(2, 64, 64, [24, 24, 24, 24, 48], 24)


def dgrad_cases():
    conv_dgrad_case(4, 64, 64, 24, 24, False, dyscale=1e-5)
    conv_dgrad_case(4, 64, 64, 24, 24, False, dyscale=1.0)
    conv_dgrad_case(4, 4, 4, 512, 512, False, dyscale=1e-3)
    conv_dgrad_case(2, 16, 16, 64, 64, False)
    conv_dgrad_case(2, 32, 32, 24, 48, True)
    conv_dgrad_case(2, 16, 16, 192, 96, False)
    conv_dgrad_multi_case(2, 32, 32, [24, 48], 24)
    conv_dgrad_multi_case(2, 64, 64, [24, 24, 24, 24, 48], 24)
    conv_dgrad_multi_case(2, 32, 32, [48, 48, 96], 48)
    conv_dgrad_multi_case(2, 16, 16, [96, 96, 192], 96)


def wgrad_cases():
    conv_wgrad_case(2, 16, 16, 64, 64)
    conv_wgrad_case(2, 16, 16, 32, 32)
    conv_wgrad_case(2, 32, 32, 24, 24)
    conv_wgrad_case(2, 32, 32, 48, 24)
    conv_wgrad_case(2, 16, 16, 192, 96)
    conv_wgrad_case(2, 16, 16, 128, 256)
    conv_wgrad_case(2, 16, 16, 384, 512)
    conv_wgrad_case(4, 8, 8, 320, 320)
    conv_wgrad_case(2, 64, 64, 24, 24, splits=7)


def convT_cases():
    convT_bwd_case(2, 16, 16, 48, 48, False)
    convT_bwd_case(2, 32, 16, 24, 48, True)
    convT_case(2, 16, 16, 64, 32)
    convT_case(2, 16, 16, 48, 48)
    convT_case(2, 8, 8, 384, 192)
    convT_case(4, 8, 8, 320, 320)


if __name__ == "__main__":
    sys.exit(main())
