"""Host-side builders that turn feature maps into libmtbc op handles / calls.

A feature map (`Feat`) is a bf16 NHWC torch tensor whose channel count is padded to a multiple of 32; PyTorch only owns
the memory.  The builders mirror the reference layers they replace:

* conv3x3 (+ folded skip concatenation)   MTnnUNet.py:12-16, MONAI Convolution (MTUNetPlusPlus.py:47-71,107-118)
* ConvTranspose2d k == stride             MTnnUNet.py:96-100, MONAI UpCat.upsample.deconv
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import OutSlice, ActView, ConvGemmDesc, ConvTBwdDesc, GemmSeg, WgradDesc, WgradMultiDesc, WgradTap


def pad32(c: int) -> int:
    return (c + 31) // 32 * 32


def _pitch_multiple() -> int:
    """Channel pitch of activation tensors: 8 (dense: a 24-channel tensor is stored with 24 channels, 48 bytes per
    pixel) or 32 (legacy padded layout, MTBC_PAD=32).  GEMM extents (K per source, N) are always padded to 32: TMA
    zero-fills the channels a dense tensor lacks and the epilogues only write the columns that exist."""
    import os
    return 32 if os.environ.get("MTBC_PAD", "8") == "32" else 8


def pitch_of(c: int) -> int:
    m = _pitch_multiple()
    return (c + m - 1) // m * m


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


class _TensorPtr(C.c_void_p):
    """c_void_p that keeps its tensor alive: bound launch closures hold raw device pointers, so the pointer object
    itself must own a reference or the caching allocator may hand the memory to someone else."""


def ptr(t: Optional[torch.Tensor]):
    if t is None:
        return None
    p = _TensorPtr(t.data_ptr())
    p._keep = t
    return p


@dataclass
class Feat:
    """NHWC activation: bf16 on the product path, fp32 in the TF32 parity mode.  Cp = channel pitch of the tensor
    (multiple of 8), Ck = its GEMM extent (multiple of 32)."""
    t: torch.Tensor  # [N, H, W, Cp] bf16 | fp32
    C: int           # true channels

    @property
    def fp32(self): return self.t.dtype == torch.float32

    @property
    def N(self): return self.t.shape[0]
    @property
    def H(self): return self.t.shape[1]
    @property
    def W(self): return self.t.shape[2]
    @property
    def Cp(self): return self.t.shape[3]
    @property
    def Ck(self): return pad32(self.t.shape[3])

    @staticmethod
    def empty(N, H, W, C, device="cuda", zero=True, dtype=torch.bfloat16) -> "Feat":
        f = torch.zeros if zero else torch.empty
        return Feat(f((N, H, W, pitch_of(C)), dtype=dtype, device=device), C)

    @staticmethod
    def from_nchw(x: torch.Tensor) -> "Feat":
        N, Cc, H, W = x.shape
        out = Feat.empty(N, H, W, Cc, device=x.device)
        out.t[..., :Cc] = x.permute(0, 2, 3, 1).to(torch.bfloat16)
        return out

    def to_nchw(self) -> torch.Tensor:
        return self.t[..., : self.C].permute(0, 3, 1, 2).float().contiguous()


def pair_view(f: Feat) -> Feat:
    """The same memory as (N, H, W/2, 2*Cp): one row of the view = two neighbouring pixels (dense tensors only: the
    channels that exist fill the pitch, so a pair is 2*C contiguous real channels)."""
    assert f.C == f.Cp and f.W % 2 == 0 and f.t.is_contiguous(), (f.C, f.Cp, f.W)
    return Feat(f.t.view(f.N, f.H, f.W // 2, 2 * f.Cp), 2 * f.C)


def _view(f: Feat) -> ActView:
    v = ActView()
    v.ptr = f.t.data_ptr()
    v.C, v.W, v.H, v.N = f.Cp, f.W, f.H, f.N
    v.sW, v.sH, v.sN = f.Cp, f.W * f.Cp, f.H * f.W * f.Cp
    return v


def _strided_view(f: Feat, k: int, i: int, j: int) -> ActView:
    """View of every k-th pixel of f starting at (i, j): the (i, j) sub-lattice of a k-times upsampled map."""
    v = ActView()
    v.ptr = f.t.data_ptr() + f.t.element_size() * ((i * f.W + j) * f.Cp)
    v.C, v.W, v.H, v.N = f.Cp, f.W // k, f.H // k, f.N
    v.sW, v.sH, v.sN = k * f.Cp, k * f.W * f.Cp, f.H * f.W * f.Cp
    return v


class Op:
    """Owning wrapper of an opaque mtbc_op handle."""

    def __init__(self, handle: C.c_void_p, keep: Sequence[object], kind: str):
        self.handle = handle
        self._keep = list(keep)  # tensors whose memory the encoded tensor maps point to
        self.kind = kind
        self.flops = _lib.load().mtbc_op_flops(handle)

    def launch(self, stream: Optional[int] = None):
        _lib.check(_lib.load().mtbc_op_launch(self.handle, C.c_void_p(stream if stream is not None else stream_ptr())),
                   self.kind)

    def __del__(self):
        try:
            if self.handle:
                _lib.load().mtbc_op_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def k_offsets(srcs: Sequence[Feat]):
    """Column offsets of each concat source inside a packed weight row, and the padded row length."""
    offs, o = [], 0
    for s in srcs:
        a = 64 if s.Ck % 64 == 0 else 32
        o = (o + a - 1) // a * a
        offs.append(o)
        o += s.Ck
    return offs, o


def _create_gemm(d: ConvGemmDesc, keep, kind) -> Op:
    h = C.c_void_p()
    _lib.check(_lib.load().mtbc_conv_gemm_create(C.byref(d), C.byref(h)), kind)
    return Op(h, keep, kind)


def _set_dtype(d: ConvGemmDesc, feats: Sequence[Feat], wpack: torch.Tensor, wpack_lo: Optional[torch.Tensor]):
    """dtype 0: bf16 / kind::f16; 1: fp32 storage / kind::tf32; 3: 3xTF32 (wpack_lo = TF32 remainder of the weights)."""
    fp32 = wpack.dtype == torch.float32
    assert all(f.fp32 == fp32 for f in feats), "activations and packed weights must share the precision mode"
    d.dtype = 0 if not fp32 else (3 if wpack_lo is not None else 1)
    d.wpack_lo = None if wpack_lo is None else wpack_lo.data_ptr()


def conv3x3_fwd_op(srcs: Sequence[Feat], wpack: torch.Tensor, out: Feat, bias: Optional[torch.Tensor] = None,
                   stat_sum: Optional[torch.Tensor] = None, stat_sq: Optional[torch.Tensor] = None,
                   accumulate: bool = False, ksz: int = 3, kind: str = "conv3x3_fwd",
                   wpack_lo: Optional[torch.Tensor] = None, bwd_fuse=None, stat_fold: int = 0) -> Op:
    """wpack: bf16 (or fp32 holding TF32 values) [ksz*ksz][out.Ck][Ktot] laid out by `k_offsets(srcs)`.
    stat_fold > 0: `srcs` / `out` are pixel-pair views (`pair_view`) and wpack the paired operand (JOB_PACK_CONV_PAIR);
    statistics / bias are indexed by column % stat_fold (mtbc_conv_gemm_desc.stat_fold).
    bwd_fuse = (y Feat, mean, rstd, gamma | None, beta | None, slope): fused InstanceNorm + LeakyReLU backward statistics
    (mtbc_conv_gemm_desc.bwd_y); stat_sum / stat_sq then receive s1 / s2.  Raises MtbcError when the shape is not served."""
    offs, ktot = k_offsets(srcs)
    assert wpack.shape == (ksz * ksz, out.Ck, ktot), (wpack.shape, (ksz * ksz, out.Ck, ktot))
    d = ConvGemmDesc()
    _set_dtype(d, [*srcs, out], wpack, wpack_lo)
    d.nviews = len(srcs)
    for i, s in enumerate(srcs):
        d.views[i] = _view(s)
    n = 0
    half = ksz // 2
    for r in range(ksz):
        for c in range(ksz):
            for i, s in enumerate(srcs):
                d.seg[n] = GemmSeg(i, r - half, c - half, offs[i], r * ksz + c)
                n += 1
    d.nseg = n
    d.wpack = wpack.data_ptr(); d.w_ntaps = ksz * ksz; d.w_ktot = ktot; d.ncols = out.Ck
    d.W, d.H, d.N = out.W, out.H, out.N
    d.epi_mode = 0; d.out = out.t.data_ptr(); d.out_C = out.Cp; d.up_k = 1; d.up_cp = out.Ck
    d.bias = None if bias is None else bias.data_ptr()
    d.stat_sum = None if stat_sum is None else stat_sum.data_ptr()
    d.stat_sq = None if stat_sq is None else stat_sq.data_ptr()
    d.stat_C = out.Cp
    d.stat_fold = int(stat_fold)
    d.accumulate = int(accumulate)
    d.nouts = 0
    keep = [*(s.t for s in srcs), wpack, wpack_lo, out.t, bias, stat_sum, stat_sq]
    if bwd_fuse is not None:
        yf, mean, rstd, gamma, beta, slope = bwd_fuse
        assert (yf.N, yf.H, yf.W, yf.Cp) == (out.N, out.H, out.W, out.Cp) and not yf.fp32
        assert mean.shape == (out.N, stat_fold or out.Cp) and rstd.shape == mean.shape
        assert stat_sum is not None and stat_sq is not None
        d.bwd_y = yf.t.data_ptr(); d.bwd_mean = mean.data_ptr(); d.bwd_rstd = rstd.data_ptr()
        d.bwd_gamma = None if gamma is None else gamma.data_ptr()
        d.bwd_beta = None if beta is None else beta.data_ptr()
        d.bwd_slope = float(slope)
        keep += [yf.t, mean, rstd, gamma, beta]
    return _create_gemm(d, keep, kind)


# stride-2 3x3 convolution (ResidualUNet.py:115-131): tap row r reads input row 2h + r - 1, i.e. the odd sub-lattice
# shifted by -1 (r = 0), the even one (r = 1) or the odd one (r = 2); likewise for columns.  Four strided views of the
# input and nine (view, shift) segments turn it into the generic implicit GEMM, zero padding included (TMA zero fill).
_S2 = {0: (1, -1), 1: (0, 0), 2: (1, 0)}


def conv3x3_s2_fwd_op(x: Feat, wpack: torch.Tensor, out: Feat, bias: Optional[torch.Tensor] = None,
                      stat_sum: Optional[torch.Tensor] = None, stat_sq: Optional[torch.Tensor] = None,
                      wpack_lo: Optional[torch.Tensor] = None) -> Op:
    """out (N, H/2, W/2) = conv3x3(x, stride 2, padding 1).  wpack: [9][out.Ck][x.Ck] as for the stride-1 conv."""
    assert x.H % 2 == 0 and x.W % 2 == 0 and (out.H, out.W) == (x.H // 2, x.W // 2)
    assert wpack.shape == (9, out.Ck, x.Ck), (wpack.shape, (9, out.Ck, x.Ck))
    d = ConvGemmDesc()
    _set_dtype(d, [x, out], wpack, wpack_lo)
    d.nviews = 4
    for i in range(2):
        for j in range(2):
            d.views[i * 2 + j] = _strided_view(x, 2, i, j)
    n = 0
    for r in range(3):
        for c in range(3):
            (i, dh), (j, dw) = _S2[r], _S2[c]
            d.seg[n] = GemmSeg(i * 2 + j, dh, dw, 0, r * 3 + c)
            n += 1
    d.nseg = n
    d.wpack = wpack.data_ptr(); d.w_ntaps = 9; d.w_ktot = x.Ck; d.ncols = out.Ck
    d.W, d.H, d.N = out.W, out.H, out.N
    d.epi_mode = 0; d.out = out.t.data_ptr(); d.out_C = out.Cp; d.up_k = 1; d.up_cp = out.Ck
    d.bias = None if bias is None else bias.data_ptr()
    d.stat_sum = None if stat_sum is None else stat_sum.data_ptr()
    d.stat_sq = None if stat_sq is None else stat_sq.data_ptr()
    d.stat_C = out.Cp
    d.accumulate = 0
    d.nouts = 0
    return _create_gemm(d, [x.t, wpack, wpack_lo, out.t, bias, stat_sum, stat_sq], "conv3x3_s2_fwd")


def conv3x3_s2_wgrad_op(x: Feat, dy: Feat, dw_acc: torch.Tensor, splits: int = 0) -> Op:
    """Weight gradient of the stride-2 conv: dw_acc fp32 [9][dy.Ck][x.Ck]; the reduction runs over dy's pixel grid."""
    assert dw_acc.dtype == torch.float32 and dw_acc.shape == (9, dy.Ck, x.Ck)
    d = WgradDesc()
    d.a_nviews = 4
    for i in range(2):
        for j in range(2):
            d.a_views[i * 2 + j] = _strided_view(x, 2, i, j)
    d.b_nviews = 1; d.b_views[0] = _view(dy)
    n = 0
    for r in range(3):
        for c in range(3):
            (i, dh), (j, dw) = _S2[r], _S2[c]
            d.taps[n] = WgradTap(i * 2 + j, dh, dw, 0)
            n += 1
    d.ntaps = n
    d.W, d.H, d.N = dy.W, dy.H, dy.N
    d.dw_acc = dw_acc.data_ptr(); d.n_rows = dw_acc.shape[1]; d.ld_k = dw_acc.shape[2]; d.k0 = 0; d.splits = splits
    assert x.fp32 == dy.fp32
    d.dtype = 1 if x.fp32 else 0
    h = C.c_void_p()
    _lib.check(_lib.load().mtbc_wgrad_create(C.byref(d), C.byref(h)), "conv3x3_s2_wgrad")
    return Op(h, [x.t, dy.t, dw_acc], "conv3x3_s2_wgrad")


def conv3x3_dgrad_op(dy: Feat, wd: torch.Tensor, dx: Feat, accumulate: bool, ksz: int = 3,
                     wd_lo: Optional[torch.Tensor] = None, bwd_fuse=None, s1: Optional[torch.Tensor] = None,
                     s2: Optional[torch.Tensor] = None, stat_fold: int = 0) -> Op:
    """wd: bf16 [ksz*ksz][dx.Ck][dy.Ck] (flipped taps, transposed channels) -> dx (+)= conv(dy, wd).
    bwd_fuse (see conv3x3_fwd_op): dx is the gradient of a = LeakyReLU(IN(y)); the launch stores gg instead and adds the
    two plane sums of the InstanceNorm backward to s1 / s2."""
    return conv3x3_fwd_op([dy], wd, dx, accumulate=accumulate, ksz=ksz, kind="conv3x3_dgrad", wpack_lo=wd_lo,
                          bwd_fuse=bwd_fuse, stat_sum=s1, stat_sq=s2, stat_fold=stat_fold)


def conv3x3_dgrad_multi_op(dy: Feat, wd_all: torch.Tensor, dxs: Sequence[Feat], accumulates: Sequence[bool]) -> Op:
    """Fused data gradient of a conv over a folded concat: ONE launch reads dy once and writes every source's gradient.
    wd_all: bf16 [9][sum(dx.Ck)][dy.Ck] (flipped taps, rows = the sources' padded channels back to back).
    Raises MtbcError when the shape is not served by the halo kernel (caller falls back to one launch per source)."""
    rows = sum(f.Ck for f in dxs)
    assert wd_all.shape == (9, rows, dy.Ck) and wd_all.is_contiguous(), (wd_all.shape, rows, dy.Ck)
    d = ConvGemmDesc()
    d.nviews = 1
    d.views[0] = _view(dy)
    n = 0
    for r in range(3):
        for c in range(3):
            d.seg[n] = GemmSeg(0, r - 1, c - 1, 0, r * 3 + c)
            n += 1
    d.nseg = n
    d.wpack = wd_all.data_ptr(); d.w_ntaps = 9; d.w_ktot = dy.Ck; d.ncols = rows
    d.W, d.H, d.N = dy.W, dy.H, dy.N
    d.epi_mode = 0; d.out = None; d.out_C = 0; d.up_k = 1; d.up_cp = rows
    d.bias = None; d.stat_sum = None; d.stat_sq = None; d.stat_C = 0; d.accumulate = 0
    d.nouts = len(dxs)
    col = 0
    for i, (f, acc) in enumerate(zip(dxs, accumulates)):
        assert (f.N, f.H, f.W) == (dy.N, dy.H, dy.W)
        d.outs[i] = OutSlice(f.t.data_ptr(), f.Cp, col, f.Ck, int(acc))   # (ptr, pitch, first column, GEMM columns)
        col += f.Ck
    return _create_gemm(d, [dy.t, wd_all, *(f.t for f in dxs)], "conv3x3_dgrad")


def convT_fwd_op(x: Feat, wf: torch.Tensor, out: Feat, k: int, bias: Optional[torch.Tensor],
                 wf_lo: Optional[torch.Tensor] = None) -> Op:
    """wf: bf16 (or fp32 / TF32 values) [1][k*k*out.Ck][x.Ck]; out spatial = k * x spatial."""
    assert wf.shape == (1, k * k * out.Ck, x.Ck)
    d = ConvGemmDesc()
    _set_dtype(d, [x, out], wf, wf_lo)
    d.nviews = 1; d.views[0] = _view(x)
    d.nseg = 1; d.seg[0] = GemmSeg(0, 0, 0, 0, 0)
    d.wpack = wf.data_ptr(); d.w_ntaps = 1; d.w_ktot = x.Ck; d.ncols = k * k * out.Ck
    d.W, d.H, d.N = x.W, x.H, x.N
    d.epi_mode = 1; d.out = out.t.data_ptr(); d.out_C = out.Cp; d.up_k = k; d.up_cp = out.Ck
    d.bias = None if bias is None else bias.data_ptr()
    d.stat_sum = None; d.stat_sq = None; d.stat_C = 0; d.accumulate = 0
    return _create_gemm(d, [x.t, wf, wf_lo, out.t, bias], "convT_fwd")


def convT_dgrad_op(dout: Feat, wd: torch.Tensor, dx: Feat, k: int, accumulate: bool,
                   wd_lo: Optional[torch.Tensor] = None) -> Op:
    """wd: bf16 [k*k][dx.Ck][dout.Ck]; dx (+)= sum_q dout_q * wd[q]."""
    assert wd.shape == (k * k, dx.Ck, dout.Ck)
    d = ConvGemmDesc()
    _set_dtype(d, [dout, dx], wd, wd_lo)
    d.nviews = k * k
    for q in range(k * k):
        d.views[q] = _strided_view(dout, k, q // k, q % k)
        d.seg[q] = GemmSeg(q, 0, 0, 0, q)
    d.nseg = k * k
    d.wpack = wd.data_ptr(); d.w_ntaps = k * k; d.w_ktot = dout.Ck; d.ncols = dx.Ck
    d.W, d.H, d.N = dx.W, dx.H, dx.N
    d.epi_mode = 0; d.out = dx.t.data_ptr(); d.out_C = dx.Cp; d.up_k = 1; d.up_cp = dx.Ck
    d.bias = None; d.stat_sum = None; d.stat_sq = None; d.stat_C = 0; d.accumulate = int(accumulate)
    return _create_gemm(d, [dout.t, wd, wd_lo, dx.t], "convT_dgrad")


def conv3x3_wgrad_op(x: Feat, dy: Feat, dw_acc: torch.Tensor, k0: int, ksz: int = 3, splits: int = 0) -> Op:
    """dw_acc: fp32 [ksz*ksz][dy.Ck][Ktot]; accumulates the slice of source x at column k0."""
    assert dw_acc.dtype == torch.float32 and dw_acc.shape[0] == ksz * ksz and dw_acc.shape[1] == dy.Ck
    d = WgradDesc()
    d.a_nviews = 1; d.a_views[0] = _view(x)
    d.b_nviews = 1; d.b_views[0] = _view(dy)
    half = ksz // 2
    n = 0
    for r in range(ksz):
        for c in range(ksz):
            d.taps[n] = WgradTap(0, r - half, c - half, 0)
            n += 1
    d.ntaps = n
    d.W, d.H, d.N = dy.W, dy.H, dy.N
    d.dw_acc = dw_acc.data_ptr(); d.n_rows = dw_acc.shape[1]; d.ld_k = dw_acc.shape[2]; d.k0 = k0; d.splits = splits
    assert x.fp32 == dy.fp32
    d.dtype = 1 if x.fp32 else 0     # fp32 views: exact fp32 reduction on the CUDA cores (parity modes)
    h = C.c_void_p()
    _lib.check(_lib.load().mtbc_wgrad_create(C.byref(d), C.byref(h)), "conv3x3_wgrad")
    return Op(h, [x.t, dy.t, dw_acc], "conv3x3_wgrad")


def conv3x3_wgrad_multi_op(xs: Sequence[Feat], dy: Feat, dw_acc: torch.Tensor, k0s: Sequence[int], splits: int = 0) -> Op:
    """Weight gradient of a 3x3 conv over a folded concat, all sources in one launch (dy is read once per pixel tile).
    dw_acc: fp32 [9][dy.Ck][Ktot]; source i accumulates at column k0s[i].  Raises MtbcError when the plane is not served
    by the halo kernels (caller falls back to conv3x3_wgrad_op per source)."""
    assert dw_acc.dtype == torch.float32 and dw_acc.shape[0] == 9 and dw_acc.shape[1] == dy.Ck and len(xs) <= 8
    d = WgradMultiDesc()
    d.nsrc = len(xs)
    for i, (x, k0) in enumerate(zip(xs, k0s)):
        d.x[i] = _view(x)
        d.k0[i] = k0
    d.dy = _view(dy)
    d.W, d.H, d.N = dy.W, dy.H, dy.N
    d.dw_acc = dw_acc.data_ptr(); d.n_rows = dw_acc.shape[1]; d.ld_k = dw_acc.shape[2]; d.splits = splits
    h = C.c_void_p()
    _lib.check(_lib.load().mtbc_wgrad_multi_create(C.byref(d), C.byref(h)), "conv3x3_wgrad")
    return Op(h, [*(x.t for x in xs), dy.t, dw_acc], "conv3x3_wgrad")


def convT_wgrad_op(x: Feat, dout: Feat, dw_acc: torch.Tensor, k: int, splits: int = 0) -> Op:
    """dw_acc: fp32 [k*k][dout.Ck][x.Ck]."""
    assert dw_acc.shape == (k * k, dout.Ck, x.Ck)
    d = WgradDesc()
    d.a_nviews = 1; d.a_views[0] = _view(x)
    d.b_nviews = k * k
    assert k * k <= 4, "tensor-core convT wgrad supports k = 2 (b_views[4])"
    for q in range(k * k):
        d.b_views[q] = _strided_view(dout, k, q // k, q % k)
        d.taps[q] = WgradTap(0, 0, 0, q)
    d.ntaps = k * k
    d.W, d.H, d.N = x.W, x.H, x.N
    d.dw_acc = dw_acc.data_ptr(); d.n_rows = dw_acc.shape[1]; d.ld_k = dw_acc.shape[2]; d.k0 = 0; d.splits = splits
    assert x.fp32 == dout.fp32
    d.dtype = 1 if x.fp32 else 0
    h = C.c_void_p()
    _lib.check(_lib.load().mtbc_wgrad_create(C.byref(d), C.byref(h)), "convT_wgrad")
    return Op(h, [x.t, dout.t, dw_acc], "convT_wgrad")


def convT_bwd_op(x: Feat, dout: Feat, wd: torch.Tensor, dw_acc: torch.Tensor, dbias: Optional[torch.Tensor], dx: Feat,
                 accumulate: bool) -> Op:
    """Fused backward of ConvTranspose2d(k = s = 2): dx (+)= data gradient, dw_acc fp32 [4][dout.Ck][x.Ck] += weight
    gradient, dbias fp32 [Cout] += bias gradient, one pass over dout (mtbc_convT_bwd_desc).  Raises MtbcError for shapes
    the kernel does not serve (the caller keeps convT_dgrad_op + convT_wgrad_op + mtbc_channel_sum)."""
    assert wd.shape == (4, dx.Ck, dout.Ck) and dw_acc.shape == (4, dout.Ck, x.Ck) and dw_acc.dtype == torch.float32
    assert (dx.N, dx.H, dx.W, dx.C) == (x.N, x.H, x.W, x.C) and (dout.H, dout.W) == (2 * x.H, 2 * x.W)
    if x.fp32 or dout.fp32:
        raise _lib.MtbcError("convT_bwd: bf16 only")
    d = ConvTBwdDesc()
    d.x = _view(x)
    for q in range(4):
        d.dy[q] = _strided_view(dout, 2, q // 2, q % 2)
    d.wd = wd.data_ptr(); d.wd_rows = wd.shape[1]; d.wd_ld = wd.shape[2]
    d.dw_acc = dw_acc.data_ptr(); d.n_rows = dw_acc.shape[1]; d.ld_k = dw_acc.shape[2]
    d.dbias = None if dbias is None else dbias.data_ptr()
    d.dx = dx.t.data_ptr(); d.dx_C = dx.Cp; d.accumulate = int(accumulate)
    d.Cout = dout.C
    h = C.c_void_p()
    _lib.check(_lib.load().mtbc_convT_bwd_create(C.byref(d), C.byref(h)), "convT_bwd")
    return Op(h, [x.t, dout.t, wd, dw_acc, dbias, dx.t], "convT_bwd")


# ----------------------------------------------------------------------------------------------------------------
# Weight packing helpers (fp32 parameters -> bf16 GEMM operands).  The packed buffers are allocated zeroed once; only
# true (unpadded) entries are ever rewritten, so pad rows/columns stay exactly zero.

def pack_conv_weight(w: torch.Tensor, srcs_C: Sequence[int], offs: Sequence[int], wf: torch.Tensor,
                     wds: Sequence[Optional[torch.Tensor]], stream: Optional[int] = None):
    Cout, Cin, ksz, _ = w.shape
    st = C.c_void_p(stream if stream is not None else stream_ptr())
    c0 = 0
    for cs, off, wd in zip(srcs_C, offs, wds):
        _lib.call("mtbc_pack_conv_weight", ptr(w), Cout, Cin, ksz, c0, cs, ptr(wf), wf.shape[1], wf.shape[2], off,
                  ptr(wd), 0 if wd is None else wd.shape[1], 0 if wd is None else wd.shape[2], st)
        c0 += cs
    assert c0 == Cin


def pack_convT_weight(w: torch.Tensor, cp: int, wf: torch.Tensor, wd: Optional[torch.Tensor],
                      stream: Optional[int] = None):
    Cin, Cout, k, _ = w.shape
    st = C.c_void_p(stream if stream is not None else stream_ptr())
    _lib.call("mtbc_pack_convT_weight", ptr(w), Cin, Cout, k, cp, ptr(wf), wf.shape[2], ptr(wd),
              0 if wd is None else wd.shape[1], 0 if wd is None else wd.shape[2], st)
