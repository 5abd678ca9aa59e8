"""Static execution plan for one (architecture, batch, H, W): every activation / gradient buffer is allocated once,
every tensor-core op has its TMA descriptors encoded once, and forward / backward are flat lists of C-ABI launches on
the current CUDA stream (capturable in a CUDA graph).  No autograd, no ATen kernels on the hot path.

Reverse-mode bookkeeping is explicit: each plan tensor owns an optional gradient buffer; the first backward
contribution stores, later ones accumulate inside the producing kernel's epilogue (dense nested skips of U-Net++:
x_0_0 feeds five consumers, MTUNetPlusPlus.py:101-118).  Shared modules applied twice (process_level_3,
MTUNetPlusPlus.py:128; upsample5, MTnnUNet.py:160,174) accumulate into one weight-gradient buffer.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib, ops
from .ops import Feat, pad32, pitch_of, ptr

EPS = 1e-5


PRECISIONS = ("bf16", "tf32", "tf32x3")

# Mode flags (mtbc_set_mode) of the plan being built: `_mk` binds them into every launch closure it creates.  Set by
# Plan.__init__, cleared by Plan.finalize -- a plan is built in one go (models._PlanModule._get_plan).
_build_mode = 0


class PTensor:
    """A bf16 NHWC activation of the plan plus its (lazily allocated) gradient."""

    def __init__(self, feat: Feat, name: str):
        self.feat = feat
        self.name = name
        self.g: Optional[Feat] = None
        self._g_init = False
        self.g_writes = 0       # launches that have written / accumulated into the gradient so far (backward emission)
        self.fuse_cand = None   # the single-source conv data gradient that wrote it first (see Plan._emit_dgrad)

    @property
    def g_init(self) -> bool:
        """Set during backward emission once some consumer has written the gradient."""
        return self._g_init

    @g_init.setter
    def g_init(self, v: bool):
        if v:
            self.g_writes += 1
        self._g_init = bool(v)

    def grad(self) -> Feat:
        if self.g is None:
            self.g = Feat.empty(self.feat.N, self.feat.H, self.feat.W, self.feat.C, device=self.feat.t.device,
                                dtype=self.feat.t.dtype)
        return self.g


class Arena:
    """fp32 scratch that must be zero at the start of a phase: carved from a few big chunks -> few memsets."""

    def __init__(self, device, chunk_floats=1 << 24):
        self.device, self.chunk_floats = device, chunk_floats
        self.chunks: List[torch.Tensor] = []
        self.used = 0

    def alloc(self, *shape) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= s
        n_al = (n + 63) // 64 * 64
        if not self.chunks or self.used + n_al > self.chunks[-1].numel():
            self.chunks.append(torch.zeros(max(self.chunk_floats, n_al), dtype=torch.float32, device=self.device))
            self.used = 0
        v = self.chunks[-1][self.used:self.used + n].view(*shape)
        self.used += n_al
        return v

    def zero_launches(self) -> List[Callable]:
        out = []
        for i, c in enumerate(self.chunks):
            nbytes = (self.used if i == len(self.chunks) - 1 else c.numel()) * 4
            if nbytes:
                out.append(_mk("mtbc_zero_bytes", ptr(c), nbytes))
        return out


def _mk(name: str, *args) -> Callable[[C.c_void_p], None]:
    """Bind a C-ABI call; the stream is supplied at launch time."""
    lib = _lib.load()
    fn = getattr(lib, name)
    mode = _build_mode

    if mode == 0:
        def launch(stream):
            rc = fn(*args, stream)
            if rc != 0:
                _lib.check(rc, name)
    else:
        # element-type generic entry points read the calling thread's mode (fp32 activations / deterministic
        # reductions); the product path (mode 0) pays nothing for it
        def launch(stream):
            lib.mtbc_set_mode(mode)
            try:
                rc = fn(*args, stream)
            finally:
                lib.mtbc_set_mode(0)
            if rc != 0:
                _lib.check(rc, name)
    launch.kind = name
    return launch


def _annot(l: Callable, desc: str, true_bytes: float) -> Callable:
    """Label a launch for the per-layer roofline table: `true_bytes` = ALGORITHMIC bytes (each tensor read once and
    written once in its stored dtype, true channel counts -- SURVEY 8d), not what the kernel happens to move."""
    l.desc = desc
    l.true_bytes = float(true_bytes)
    return l


class JobTable:
    """Parameter-side jobs (padded vector copies, weight packs, weight-gradient unpacks) run as ONE launch."""

    def __init__(self):
        self.jobs: List[_lib.ParamJob] = []
        self._keep: List[object] = []
        self.owner: List[Optional[str]] = []   # parameter name per job (gradient unpacks: decides the bucket)

    def add(self, kind: int, ints: Sequence[int], src: torch.Tensor, dst0: Optional[torch.Tensor],
            dst1: Optional[torch.Tensor] = None, owner: Optional[str] = None):
        self.owner.append(owner)
        j = _lib.ParamJob()
        j.kind = kind
        for k, v in enumerate(ints):
            j.i[k] = int(v)
        j.src = src.data_ptr()
        j.dst0 = None if dst0 is None else dst0.data_ptr()
        j.dst1 = None if dst1 is None else dst1.data_ptr()
        self.jobs.append(j)
        self._keep += [src, dst0, dst1]

    def launch(self, select: Optional[Sequence[int]] = None) -> List[Callable]:
        """One launch over all jobs, or over the jobs whose indices are in `select`."""
        jobs = self.jobs if select is None else [self.jobs[i] for i in select]
        if not jobs:
            return []
        arr = (_lib.ParamJob * len(jobs))(*jobs)
        h = C.c_void_p()
        _lib.check(_lib.load().mtbc_param_jobs_create(arr, len(jobs), C.byref(h)), "param_jobs")
        op = ops.Op(h, self._keep, "param_jobs")
        return [_mk_op(op, 0.0, f"{len(jobs)} parameter jobs")]


def _mk_op(op: ops.Op, true_flops: float = 0.0, desc: str = "", true_bytes: float = 0.0) -> Callable:
    """Bind a tensor-core op launch; `true_flops` = algorithmic 2*M*N*K with unpadded channels (roofline numerator)."""
    lib = _lib.load()
    cell = [op]

    def launch(stream):
        rc = lib.mtbc_op_launch(cell[0].handle, stream)
        if rc != 0:
            _lib.check(rc, op.kind)

    def swap(new_op):   # replace the op behind an already emitted launch (fused InstanceNorm backward statistics)
        cell[0] = new_op
        launch.op = new_op
    launch.swap = swap
    launch.kind = op.kind
    launch.op = op
    launch.true_flops = true_flops
    launch.true_bytes = float(true_bytes)
    launch.desc = desc
    return launch


class Plan:
    """Builder + executor.  Model files call the layer methods in forward order; `finalize()` emits the backward."""

    def __init__(self, B: int, H: int, W: int, device, params: Dict[str, torch.nn.Parameter], training: bool = True,
                 precision: str = "bf16", deterministic: bool = False):
        """precision: "bf16" = the product path (bf16 storage, tcgen05 kind::f16).  "tf32" / "tf32x3" = the parity modes
        north_star names ("1e-3 (TF32 mode)"): fp32 storage everywhere, tcgen05 kind::tf32 convolutions and data
        gradients through the generic implicit-GEMM kernel, single pass or with the 3xTF32 split (fp32-grade products),
        weight gradients as exact fp32 reductions on the CUDA cores.  Checker-grade paths: correct, not fast.
        deterministic: InstanceNorm statistics come from an order-independent reduction (mtbc_in_stats_det) instead of
        the conv epilogues' fp32 atomics, so the forward pass is bit-identical from run to run."""
        global _build_mode
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {PRECISIONS}")
        self.B, self.H, self.W, self.device = B, H, W, device
        self.params = params
        self.training = training
        self.precision, self.deterministic = precision, bool(deterministic)
        self.fp32 = precision != "bf16"
        self.act_dtype = torch.float32 if self.fp32 else torch.bfloat16
        self.mode_flags = (_lib.MODE_ACT_FP32 if self.fp32 else 0) | (_lib.MODE_DETERMINISTIC if deterministic else 0)
        _build_mode = self.mode_flags
        self.pack: List[Callable] = []
        self.pack_jobs = JobTable()     # vector copies + weight packs, one launch at the head of `pack`
        self.unpack_jobs = JobTable()   # weight-gradient unpacks, one launch at the tail of `bwd`
        self.fwd: List[Callable] = []
        self._bwd_blocks: List[List[Callable]] = []  # one block per forward op, reversed at finalize
        self.bwd: List[Callable] = []
        self.fwd_arena = Arena(device, 1 << 20)   # InstanceNorm statistics (zeroed every forward)
        self.bwd_arena = Arena(device, 1 << 24)   # s1/s2, weight-gradient accumulators (zeroed every backward)
        self._keep: List[object] = []
        self._scratch: Dict[Tuple, List[Feat]] = {}
        self._scratch_turn: Dict[Tuple, int] = {}
        self._slot_readers: Dict[Tuple, List[Callable]] = {}   # dy slot -> side-stream launches that read it
        self._cur_slot = None
        self._dy_slot_of: Dict[int, Tuple] = {}
        self.tc_flops_fwd = 0.0
        self.tc_flops_bwd = 0.0
        # outputs (fp32, plan-owned static buffers) and their incoming gradients
        self.outputs_cls: List[torch.Tensor] = []
        self.outputs_seg: List[torch.Tensor] = []
        self.g_cls: List[torch.Tensor] = []
        self.g_seg: List[Optional[torch.Tensor]] = []
        # flat parameter gradients (param-shaped views); registration order reversed ~ backward completion order
        names = list(params.keys())
        self.grad_range, total = flat_layout(params)
        self.grad_flat = torch.zeros(total, dtype=torch.float32, device=device)
        self.grad_view: Dict[str, torch.Tensor] = {
            n: self.grad_flat[a:b].view_as(params[n]) for n, (a, b) in self.grad_range.items()}
        self.has_grad: Dict[str, bool] = {n: False for n in names}
        # deferred conv weight-gradient unpacks: param name -> (pending applications, launches)
        self._wg_pending: Dict[str, int] = {}
        self._wg_acc: Dict[str, torch.Tensor] = {}
        self._late_unpacked: List[str] = []  # convT weights: unpacked by the batched job at the end of backward
        self._packed: Dict[str, dict] = {}
        self.finalized = False
        self._touched = set()
        self.tensors: Dict[str, PTensor] = {}  # name -> activation (debug / tests)
        self.debug: Dict[str, Feat] = {}
        self.param_done_at: Dict[str, int] = {}  # index into self.bwd after which the param's grad is final
        # F.dropout of the ResidualUNet sibling: masks (uint8, forward order), generator seed, test hook
        self.dropout_masks: List[torch.Tensor] = []
        self.dropout_seed = int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF
        self.dropout_external = False
        self._drop_ctr: Optional[torch.Tensor] = None

    # ------------------------------------------------------------------------------------------------ helpers
    def P(self, name: str) -> torch.nn.Parameter:
        return self.params[name]

    def new(self, N, H, W, Cc, name) -> PTensor:
        t = PTensor(Feat.empty(N, H, W, Cc, device=self.device, dtype=self.act_dtype), name)
        self.tensors[name] = t
        return t

    def scratch(self, N, H, W, Cc) -> Feat:
        """Gradient-of-conv-output buffer (dy) for one backward block.  Buffers of one shape rotate (two by default) so
        that a weight-gradient launch running on the side stream may still read the previous one while the next
        layer's InstanceNorm backward writes the other; `self._slot_readers[slot]` remembers the side launches that read
        a slot so the next writer can wait for exactly those."""
        nbuf = max(1, int(os.environ.get("MTBC_DY_BUFFERS", "2")))
        key = (N, H, W, pitch_of(Cc))
        if os.environ.get("MTBC_DEBUG_UNIQUE_SCRATCH"):
            key = key + (len(self._scratch),)
        ring = self._scratch.setdefault(key, [])
        turn = self._scratch_turn.get(key, 0)
        self._scratch_turn[key] = turn + 1
        if len(ring) < nbuf:
            ring.append(Feat.empty(N, H, W, Cc, device=self.device, dtype=self.act_dtype))
        idx = turn % len(ring) if len(ring) == nbuf else len(ring) - 1
        f = ring[idx]
        self._cur_slot = (key, idx)
        return Feat(f.t, Cc)

    def _padded_vec(self, name: Optional[str], Cp: int) -> Optional[torch.Tensor]:
        """fp32 [Cp] copy target for a per-channel parameter (bias / gamma / beta); refreshed in the pack phase."""
        if name is None:
            return None
        p = self.P(name)
        buf = torch.zeros(Cp, dtype=torch.float32, device=self.device)
        n = p.numel()
        self.pack_jobs.add(_lib.JOB_COPY_F32, [n], p, buf)
        self._keep.append(buf)
        return buf

    def _side(self, launch: Callable, dy: Optional[Feat] = None) -> Callable:
        """Mark a weight-gradient launch as runnable on the side stream (it only feeds the accumulators that the
        bucket unpack reads); if it reads a rotating dy buffer, register it as a reader of that slot."""
        launch.side = True
        if dy is not None:
            slot = self._dy_slot_of.get(id(dy.t))
            if slot is not None:
                self._slot_readers.setdefault(slot, []).append(launch)
        return launch

    def _side_fwd(self, launch: Callable) -> Callable:
        """Mask-head forward launches only feed the logits the objective reads after the whole forward pass: TrainStep
        can run them on the side stream, next to the class branch's convolutions (its objective launches wait for the
        side stream).  Measured, six alternating runs of 40 steps on one box: 11.82 / 11.80 / 11.81 ms with it,
        11.82 / 11.76 / 11.82 ms without -- no gain (the four head launches are 0.11 ms of HBM traffic that the
        class-branch convs do not hide), so it is off unless MTBC_SIDE_HEADS=1."""
        if os.environ.get("MTBC_SIDE_HEADS", "0") == "1":
            launch.side = True
        return launch

    def _mark_grad(self, *names):
        for n in names:
            if n is not None:
                self.has_grad[n] = True
                self._touched.add(n)

    # ------------------------------------------------------------------------------------------------ layers
    def input_conv_in_act(self, x_in: torch.Tensor, w: str, b: Optional[str], gamma: Optional[str],
                          beta: Optional[str], slope: float, pool: bool, name: str):
        """First layer: Conv2d(Cin<=4 -> C, 3x3) on the fp32 NCHW image + InstanceNorm + LeakyReLU (+pool)."""
        Wt = self.P(w)
        Cout, Cin = Wt.shape[0], Wt.shape[1]
        N, H, Wd = self.B, self.H, self.W
        y = self.new(N, H, Wd, Cout, name + ".y")
        ssum = self.fwd_arena.alloc(N, y.feat.Cp)
        ssq = self.fwd_arena.alloc(N, y.feat.Cp)
        bvec = None if b is None else self.P(b)
        xs = torch.zeros(N, Cin * 9, dtype=torch.float32, device=self.device)  # shifted plane sums (mean removal)
        self._keep.append(xs)
        npx = N * H * Wd
        det = self.deterministic
        self.fwd.append(_annot(_mk("mtbc_conv_first_fwd", ptr(x_in), N, Cin, H, Wd, ptr(Wt), ptr(bvec), Cout,
                                   ptr(y.feat.t), y.feat.Cp, None if det else ptr(ssum), None if det else ptr(ssq),
                                   ptr(xs)),
                               f"{name} first conv fwd {N}x{H}x{Wd} {Cin}->{Cout}", npx * (4 * Cin + 2 * Cout)))
        self.fwd[-1].true_flops = 2.0 * npx * Cout * Cin * 9
        if det:
            self.fwd.append(self._stats_launch(y, ssum, ssq, name))
        a, p, aux = self._norm_act(y, ssum, ssq, gamma, beta, slope, pool, name)

        def backward():
            blk: List[Callable] = []
            dy = self._norm_act_bwd(blk, y, a, p, aux, gamma, beta, slope)
            if dy is None:
                return blk
            l = _annot(_mk("mtbc_conv_first_wgrad", ptr(x_in), N, Cin, H, Wd, ptr(dy.t), dy.Cp, Cout,
                           ptr(self.grad_view[w])), f"{name} first conv wgrad {N}x{H}x{Wd} {Cin}x{Cout}",
                       npx * (4 * Cin + 2 * Cout))
            l.true_flops = 2.0 * npx * Cout * Cin * 9
            blk.append(self._side(l, dy))
            self._mark_grad(w, b)  # bias grad of a conv followed by InstanceNorm is identically zero
            return blk
        self._bwd_blocks.append(backward)
        return a, p

    def conv_in_act(self, srcs: Sequence[PTensor], w: str, b: Optional[str], gamma: Optional[str], beta: Optional[str],
                    slope: float, pool: bool, name: str):
        """Conv2d 3x3 over the (virtual) channel concatenation of `srcs` + InstanceNorm + LeakyReLU (+ 2x2 max-pool).
        MTnnUNet.py:19-39 / MONAI Convolution; the concat (MTUNetPlusPlus.py:107-118) is folded into the K loop."""
        Wt = self.P(w)
        Cout, Cin = Wt.shape[0], Wt.shape[1]
        feats = [s.feat for s in srcs]
        assert sum(f.C for f in feats) == Cin, (name, [f.C for f in feats], Cin)
        N, H, Wd = feats[0].N, feats[0].H, feats[0].W
        y = self.new(N, H, Wd, Cout, name + ".y")
        Cp = y.feat.Cp                      # channel pitch of y (statistics arrays follow it)
        pk = self._pack_conv(w, feats, y.feat.Ck)
        ssum = self.fwd_arena.alloc(N, Cp)
        ssq = self.fwd_arena.alloc(N, Cp)
        fused = H * Wd >= 128 and not self.deterministic
        # The conv bias (MONAI Convolution has bias=True, MTUNetPlusPlus.py:24) is NOT added: InstanceNorm follows
        # immediately and is invariant to a per-channel shift of its input, so `a` is unchanged in exact arithmetic, the
        # bias gradient is identically zero (see backward below), and y is stored closer to zero-mean (fewer bf16
        # mantissa bits spent on a DC level).  It also takes 16 adds per chunk out of the epilogue's critical path.
        bias = None
        op = None
        if self._pair_ok(feats, y.feat, fused):
            # pixel-pair view: the same convolution over (N, H, W/2, 2C) tensors -- one 96-byte TMA box row
            # per two pixels of a dense 24-channel tensor instead of two 48-byte ones, K = 48 = three full K steps
            try:
                op = ops.conv3x3_fwd_op([ops.pair_view(feats[0])], self._pair_pack(w, pk, feats[0], y.feat),
                                        ops.pair_view(y.feat), bias=bias, stat_sum=ssum, stat_sq=ssq, stat_fold=Cp)
            except _lib.MtbcError:
                op = None
        if op is None:
            op = ops.conv3x3_fwd_op(feats, pk["wf"], y.feat, bias=bias, stat_sum=ssum if fused else None,
                                    stat_sq=ssq if fused else None, wpack_lo=pk.get("wf_lo"))
        self.tc_flops_fwd += op.flops
        npx = N * H * Wd
        wbytes = 2.0 * 9 * Cin * Cout
        self.fwd.append(_mk_op(op, 2.0 * N * H * Wd * Cout * Cin * 9,
                               f"{name} fwd {N}x{H}x{Wd} {[f.C for f in feats]}->{Cout}",
                               2.0 * npx * (Cin + Cout) + wbytes))
        if not fused:
            self.fwd.append(self._stats_launch(y, ssum, ssq, name))
        a, p, aux = self._norm_act(y, ssum, ssq, gamma, beta, slope, pool, name)
        self._wg_pending[w] = self._wg_pending.get(w, 0) + 1

        def backward():
            blk: List[Callable] = []
            dy = self._norm_act_bwd(blk, y, a, p, aux, gamma, beta, slope)
            self._wg_pending[w] -= 1
            if dy is None:
                return blk
            acc = self._wg_accum(w, pk)
            o = None
            if H % 16 == 0 and Wd % 8 == 0 and not self.fp32 and not os.environ.get("MTBC_NO_FUSED_WGRAD"):
                try:   # one launch over every concat source: dy is read once per pixel tile
                    o = ops.conv3x3_wgrad_multi_op(feats, dy, acc, pk["offs"])
                except _lib.MtbcError:
                    o = None
            if o is not None:
                self.tc_flops_bwd += o.flops
                blk.append(self._side(_mk_op(o, 2.0 * N * H * Wd * Cout * Cin * 9,
                                             f"{name} wgrad {N}x{H}x{Wd} {[f.C for f in feats]}x{Cout} fused",
                                             2.0 * npx * (Cin + Cout) + 2 * wbytes), dy))
            else:
                for f, off in zip(feats, pk["offs"]):
                    o = ops.conv3x3_wgrad_op(f, dy, acc, off)
                    self.tc_flops_bwd += o.flops
                    blk.append(self._side(_mk_op(o, 2.0 * N * H * Wd * Cout * f.C * 9,
                                                 f"{name} wgrad {N}x{H}x{Wd} {f.C}x{Cout}",
                                                 2.0 * npx * (f.C + Cout) + 4.0 * 9 * f.C * Cout), dy))
            self._mark_grad(w, b)
            self._emit_dgrad(blk, srcs, dy, w, pk, name, Cout)
            return blk
        self._bwd_blocks.append(backward)
        return a, p

    def _stats_launch(self, y: PTensor, ssum, ssq, name: str) -> Callable:
        """Separate InstanceNorm statistics pass over a stored conv output (planes too small for the fused epilogue
        statistics, or deterministic mode)."""
        f = y.feat
        nbytes = 2.0 * f.N * f.H * f.W * f.C * (2 if self.fp32 else 1)
        if not self.deterministic:
            return _annot(_mk("mtbc_in_stats", ptr(f.t), f.N, f.H * f.W, f.Cp, ptr(ssum), ptr(ssq)),
                          f"{name} in_stats {f.N}x{f.H}x{f.W}x{f.C}", nbytes)
        need = int(_lib.load().mtbc_query_workspace_bytes(b"in_stats_det", f.N, f.H * f.W, f.Cp))
        if need <= 0:
            raise _lib.MtbcError(f"in_stats_det: no workspace size for N={f.N} HW={f.H * f.W} Cp={f.Cp}")
        ws = torch.zeros((need + 3) // 4, dtype=torch.int32, device=self.device)
        self._keep.append(ws)
        return _annot(_mk("mtbc_in_stats_det", ptr(f.t), f.N, f.H * f.W, f.Cp, ptr(ssum), ptr(ssq), ptr(ws)),
                      f"{name} in_stats_det {f.N}x{f.H}x{f.W}x{f.C}", nbytes)

    def _pack_conv(self, w: str, feats: Sequence[Feat], Cp: int) -> dict:
        """bf16 forward / data-gradient operands of a conv weight (shared modules are packed once per step)."""
        key = w
        if key in self._packed:
            pk = self._packed[key]
            assert [f.Ck for f in feats] == pk["src_cp"]
            return pk
        Wt = self.P(w)
        offs, ktot = ops.k_offsets(feats)
        # bf16 operands, or fp32 operands holding TF32 values (part 1 = rounded weight, part 2 = rounded remainder)
        wf = torch.zeros(9, Cp, ktot, dtype=torch.float32 if self.fp32 else torch.bfloat16, device=self.device)
        wf_lo = torch.zeros_like(wf) if self.precision == "tf32x3" else None
        st_c = [f.C for f in feats]
        c0 = 0
        for cs, off in zip(st_c, offs):
            for part, dst in ((1 if self.fp32 else 0, wf), (2, wf_lo)):
                if dst is not None:
                    self.pack_jobs.add(_lib.JOB_PACK_CONV, [Wt.shape[0], Wt.shape[1], 3, c0, cs, wf.shape[1],
                                                            wf.shape[2], off, 0, 0, part], Wt, dst, None)
            c0 += cs
        pk = {"wf": wf, "offs": offs, "ktot": ktot, "src_cp": [f.Ck for f in feats], "src_c": st_c, "Cp": Cp}
        if wf_lo is not None:
            pk["wf_lo"] = wf_lo
        self._packed[key] = pk
        return pk

    def _pair_ok(self, feats: Sequence[Feat], y: Feat, fused: bool) -> bool:
        """Single dense source and dense output whose pixel pairs fit one 128-byte row / 64 GEMM columns, planes the
        halo kernel tiles in the pair view (MTBC_PAIR=0 switches the view off)."""
        if os.environ.get("MTBC_PAIR", "1") == "0" or self.fp32 or not fused or len(feats) != 1:
            return False
        f = feats[0]
        # 24-channel tensors: a pair is 48 channels = three full K = 16 steps and 48 of 64 columns; with 32 channels the
        # pair operand would be half zeros at no saving in padding
        return (f.C == f.Cp and y.C == y.Cp and 32 < 2 * f.Cp <= 48 and 32 < 2 * y.Cp <= 48
                and y.H % 16 == 0 and y.W % 16 == 0)

    def _pair_pack(self, w: str, pk: dict, f: Feat, y: Feat) -> torch.Tensor:
        """Forward operand of the pixel-pair view: [9 = (dh, dq)][(op, co) -> 64][(par, ci) -> 64] (JOB_PACK_CONV_PAIR)."""
        if "wfp" not in pk:
            Wt = self.P(w)
            wfp = torch.zeros(9, ops.pad32(2 * y.Cp), ops.pad32(2 * f.Cp), dtype=torch.bfloat16, device=self.device)
            self.pack_jobs.add(_lib.JOB_PACK_CONV_PAIR, [Wt.shape[0], Wt.shape[1], 0, Wt.shape[1], wfp.shape[1],
                                                         wfp.shape[2], 0, 0, f.Cp, y.Cp, 0], Wt, wfp)
            pk["wfp"] = wfp
        return pk["wfp"]

    def _pair_pack_dgrad(self, cand: dict, g: Feat) -> torch.Tensor:
        """Data-gradient operand of the pixel-pair view for one concat source: [9][(op, ci) -> 64][(par, co) -> 64]."""
        pk, key = cand["pk"], ("wdp", cand["c0"])
        if key not in pk:
            Wt = self.P(cand["w"])
            dy = cand["dy"]
            wdp = torch.zeros(9, ops.pad32(2 * g.Cp), ops.pad32(2 * dy.Cp), dtype=torch.bfloat16, device=self.device)
            self.pack_jobs.add(_lib.JOB_PACK_CONV_PAIR, [Wt.shape[0], Wt.shape[1], cand["c0"], cand["cs"], wdp.shape[1],
                                                         wdp.shape[2], 0, 0, dy.Cp, g.Cp, 1], Wt, wdp)
            pk[key] = wdp
        return pk[key]

    def _dgrad_pack(self, w: str, pk: dict, fused: bool):
        """Data-gradient operand(s) of a conv weight, created when the backward is emitted: one tall pack
        [9][sum Cp_src][Cp_out] for the fused launch, or one [9][Cp_src][Cp_out] per source."""
        key = "wd_all" if fused else "wd"
        if key in pk:
            return pk[key]
        Wt = self.P(w)
        Cp = pk["Cp"]
        rows = sum(pk["src_cp"])
        if fused:
            wd_all = torch.zeros(9, rows, Cp, dtype=torch.bfloat16, device=self.device)
            c0, r0 = 0, 0
            for cs, cp in zip(pk["src_c"], pk["src_cp"]):
                # rows of this source start at r0 inside every tap plane: pass the offset pointer, plane stride = rows
                self.pack_jobs.add(_lib.JOB_PACK_CONV, [Wt.shape[0], Wt.shape[1], 3, c0, cs, 0, 0, 0, rows, Cp],
                                   Wt, None, wd_all[0, r0:])
                c0 += cs
                r0 += cp
            self._keep.append(wd_all)
            pk[key] = wd_all
        else:
            wds = []
            c0 = 0
            for cs, cp in zip(pk["src_c"], pk["src_cp"]):
                wd = torch.zeros(9, cp, Cp, dtype=torch.float32 if self.fp32 else torch.bfloat16, device=self.device)
                wd_lo = torch.zeros_like(wd) if self.precision == "tf32x3" else None
                for part, dst in ((1 if self.fp32 else 0, wd), (2, wd_lo)):
                    if dst is not None:
                        self.pack_jobs.add(_lib.JOB_PACK_CONV, [Wt.shape[0], Wt.shape[1], 3, c0, cs, 0, 0, 0, cp, Cp,
                                                                part], Wt, None, dst)
                c0 += cs
                wds.append((wd, wd_lo) if self.fp32 else wd)
            pk[key] = wds
        return pk[key]

    def _emit_dgrad(self, blk, srcs, dy: Feat, w: str, pk: dict, name: str, Cout: int):
        N, H, Wd = dy.N, dy.H, dy.W
        if len(srcs) > 1 and H % 16 == 0 and Wd % 8 == 0 and not self.fp32 and not os.environ.get("MTBC_NO_FUSED_DGRAD"):
            grads = [s.grad() for s in srcs]
            try:
                o = ops.conv3x3_dgrad_multi_op(dy, self._dgrad_pack(w, pk, True), grads, [s.g_init for s in srcs])
            except _lib.MtbcError:
                o = None  # weights of this width do not fit next to the halo ring: one launch per source below
            if o is not None:
                self.tc_flops_bwd += o.flops
                cin = sum(s.feat.C for s in srcs)
                blk.append(_mk_op(o, 2.0 * N * H * Wd * Cout * cin * 9,
                                  f"{name} dgrad {N}x{H}x{Wd} {[s.feat.C for s in srcs]}<-{Cout} fused",
                                  2.0 * N * H * Wd * (cin + Cout) + 2.0 * 9 * cin * Cout))
                for s in srcs:
                    s.g_init = True
                return
        c0 = 0
        for s, wd in zip(srcs, self._dgrad_pack(w, pk, False)):
            g = s.grad()
            wd, wd_lo = wd if isinstance(wd, tuple) else (wd, None)
            o = ops.conv3x3_dgrad_op(dy, wd, g, accumulate=s.g_init, wd_lo=wd_lo)
            self.tc_flops_bwd += o.flops
            l = _mk_op(o, 2.0 * N * H * Wd * Cout * s.feat.C * 9,
                       f"{name} dgrad {N}x{H}x{Wd} {s.feat.C}<-{Cout} acc={int(s.g_init)}",
                       2.0 * N * H * Wd * (s.feat.C + Cout) + 2.0 * 9 * s.feat.C * Cout)
            blk.append(l)
            if not s.g_init and not self.fp32 and not self.deterministic:
                # first writer of this gradient: if it stays the only one, the producer's InstanceNorm backward takes
                # its two plane sums from this launch's epilogue (_norm_act_bwd swaps the op)
                s.fuse_cand = {"launch": l, "dy": dy, "wd": wd, "w": w, "pk": pk, "c0": c0, "cs": s.feat.C}
            s.g_init = True
            c0 += s.feat.C

    def _wg_accum(self, w: str, pk: dict) -> torch.Tensor:
        if w not in self._wg_acc:
            Wt = self.P(w)
            acc = self.bwd_arena.alloc(*pk["wf"].shape)
            self._wg_acc[w] = acc
            c0 = 0
            for cs, off in zip(pk["src_c"], pk["offs"]):
                self.unpack_jobs.add(_lib.JOB_UNPACK_CONV, [acc.shape[1], acc.shape[2], off, Wt.shape[0], Wt.shape[1],
                                                            3, c0, cs, 0], acc, self.grad_view[w], owner=w)
                c0 += cs
        return self._wg_acc[w]

    def _norm_act(self, y: PTensor, ssum, ssq, gamma, beta, slope, pool, name):
        N, H, Wd, Cp, Cc = y.feat.N, y.feat.H, y.feat.W, y.feat.Cp, y.feat.C
        a = self.new(N, H, Wd, Cc, name)
        p = self.new(N, H // 2, Wd // 2, Cc, name + ".pool") if pool else None
        mean = torch.zeros(N, Cp, dtype=torch.float32, device=self.device)
        rstd = torch.zeros(N, Cp, dtype=torch.float32, device=self.device)
        gv = self._padded_vec(gamma, y.feat.Ck)
        bv = self._padded_vec(beta, y.feat.Ck)
        self.fwd.append(_annot(_mk("mtbc_in_apply", ptr(y.feat.t), N, H, Wd, Cp, ptr(ssum), ptr(ssq), ptr(gv), ptr(bv),
                                   Cc, C.c_float(EPS), C.c_float(slope), ptr(a.feat.t),
                                   None if p is None else ptr(p.feat.t), ptr(mean), ptr(rstd)),
                               f"{name} in_apply{'+pool' if pool else ''} {N}x{H}x{Wd}x{Cc}",
                               2.0 * N * H * Wd * Cc * (2.25 if pool else 2.0)))
        return a, p, (mean, rstd, gv, bv)

    def _norm_act_bwd(self, blk, y: PTensor, a: PTensor, p: Optional[PTensor], aux, gamma, beta, slope):
        """Emit pool / InstanceNorm / LeakyReLU backward; returns dy (gradient of the raw conv output) or None."""
        mean, rstd, gv, bv = aux
        N, H, Wd, Cp, Cc = y.feat.N, y.feat.H, y.feat.W, y.feat.Cp, y.feat.C
        if p is not None and p.g_init:
            g = a.grad()
            blk.append(_annot(_mk("mtbc_maxpool2_bwd", ptr(a.feat.t), ptr(p.g.t), N, H, Wd, Cp, ptr(g.t),
                                  int(a.g_init)), f"{a.name} maxpool2_bwd {N}x{H}x{Wd}x{Cc}",
                              2.0 * N * H * Wd * Cc * 2.25))
            a.g_init = True
        if not a.g_init:
            return None
        s1 = self.bwd_arena.alloc(N, Cp)
        s2 = self.bwd_arena.alloc(N, Cp)
        cnt = self.bwd_arena.alloc(N)   # zeroed with the arena; reinterpreted as int32 arrival counters
        dy = self.scratch(N, H, Wd, Cc)
        self.debug[a.name + ".dy"] = dy
        self.debug[a.name + ".aux"] = (mean, rstd, gv, bv, s1, s2)
        dg = self.grad_view[gamma] if gamma else None
        db = self.grad_view[beta] if beta else None
        fused = None
        cand = a.fuse_cand
        # (MTBC_FUSE_INBWD=0 keeps the two-pass backward everywhere; the kernel serves the channel pitches for which the
        #  trade pays: 24 channels @256^2, B = 32: reduction pass -40 us, data gradient +22 us -- DESIGN 9)
        if cand is not None and a.g_writes == 1 and os.environ.get("MTBC_FUSE_INBWD", "1") != "0":
            # `a` has ONE consumer, a 3x3 conv whose data gradient wrote a.g: that launch's epilogue already holds the
            # gradient in fp32, so it applies the LeakyReLU factor, stores gg and leaves sum(gg), sum(gg * xhat) in
            # s1 / s2 -- the reduction pass over (g, y) disappears (10 -> 6 B/elem + 2 B/elem read in the epilogue).
            # The same launch through the pixel-pair view (dy, gradient and y as (N, H, W/2, 48) tensors) is built and
            # tested but off: with the y tile in the epilogue this launch follows its bytes, not its box rows or MMAs --
            # 94 -> 98 us at 24 channels @256^2, B = 32 (gpurun_out/r04c, MTBC_PAIR_DGRAD=1 switches it on)
            if (os.environ.get("MTBC_PAIR_DGRAD", "0") == "1" and "w" in cand
                    and self._pair_ok([cand["dy"]], a.g, True) and y.feat.C == y.feat.Cp):
                try:
                    fused = ops.conv3x3_dgrad_op(ops.pair_view(cand["dy"]), self._pair_pack_dgrad(cand, a.g),
                                                 ops.pair_view(a.g), accumulate=False,
                                                 bwd_fuse=(ops.pair_view(y.feat), mean, rstd, gv, bv, slope),
                                                 s1=s1, s2=s2, stat_fold=Cp)
                except _lib.MtbcError:
                    fused = None
            try:
                if fused is None:
                    fused = ops.conv3x3_dgrad_op(cand["dy"], cand["wd"], a.g, accumulate=False,
                                                 bwd_fuse=(y.feat, mean, rstd, gv, bv, slope), s1=s1, s2=s2)
            except _lib.MtbcError:
                fused = None   # shape not served by the halo kernel's statistics epilogue: two-pass backward below
        if fused is not None:
            cand["launch"].swap(fused)
            cand["launch"].desc += " +in_bwd sums"
            l = _mk("mtbc_in_bwd_apply", ptr(a.g.t), ptr(y.feat.t), N, H * Wd, Cp, ptr(mean), ptr(rstd), ptr(gv), ptr(bv),
                    C.c_float(1.0), ptr(s1), ptr(s2), ptr(dy.t), ptr(dg), ptr(db), Cc)
            l.kind = "mtbc_in_bwd"   # same row of the per-kernel tables as the two-pass launches
            _annot(l, f"{a.name} in_bwd(apply) {N}x{H}x{Wd}x{Cc}", 6.0 * N * H * Wd * Cc)
        else:
            l = _mk("mtbc_in_bwd", ptr(a.g.t), ptr(y.feat.t), N, H * Wd, Cp, ptr(mean), ptr(rstd), ptr(gv), ptr(bv),
                    C.c_float(slope), ptr(s1), ptr(s2), ptr(dy.t), ptr(dg), ptr(db), Cc, ptr(cnt))
            # algorithmic: read g, read y, write dy once each (the two-pass kernel reads g and y twice: 10 B/elem moved)
            _annot(l, f"{a.name} in_bwd {N}x{H}x{Wd}x{Cc}", 6.0 * N * H * Wd * Cc)
        l.wait_side = self._slot_readers.pop(self._cur_slot, [])   # weight gradients still reading this dy buffer
        self._dy_slot_of[id(dy.t)] = self._cur_slot
        blk.append(l)
        self._mark_grad(gamma, beta)
        return dy

    def convT(self, x: PTensor, w: str, b: Optional[str], k: int, name: str) -> PTensor:
        """ConvTranspose2d(kernel = stride = k) as one GEMM with a pixel-shuffle epilogue (MTnnUNet.py:96-100)."""
        Wt = self.P(w)
        Cin, Cout = Wt.shape[0], Wt.shape[1]
        f = x.feat
        assert f.C == Cin
        out = self.new(f.N, f.H * k, f.W * k, Cout, name)
        cp = out.feat.Ck                    # GEMM columns per sub-pixel (the tensor itself may be denser: out.feat.Cp)
        wf = torch.zeros(1, k * k * cp, f.Ck, dtype=torch.float32 if self.fp32 else torch.bfloat16, device=self.device)
        wd = torch.zeros(k * k, f.Ck, cp, dtype=wf.dtype, device=self.device) if self.training else None
        x3 = self.precision == "tf32x3"
        wf_lo = torch.zeros_like(wf) if x3 else None
        wd_lo = torch.zeros_like(wd) if (x3 and wd is not None) else None
        for part, dst, dst1 in ((1 if self.fp32 else 0, wf, wd), (2, wf_lo, wd_lo)):
            if dst is not None:
                self.pack_jobs.add(_lib.JOB_PACK_CONVT, [Cin, Cout, k, cp, wf.shape[2], 0 if dst1 is None else wd.shape[1],
                                                         0 if dst1 is None else wd.shape[2], 0, 0, 0, part], Wt, dst, dst1)
        bias = self._padded_vec(b, cp)
        op = ops.convT_fwd_op(f, wf, out.feat, k, bias, wf_lo)
        self.tc_flops_fwd += op.flops
        t_flops = 2.0 * f.N * f.H * f.W * Cin * Cout * k * k
        t_bytes = 2.0 * f.N * f.H * f.W * (Cin + k * k * Cout) + 2.0 * Cin * Cout * k * k
        self.fwd.append(_mk_op(op, t_flops, f"{name} convT fwd {f.N}x{f.H}x{f.W} {Cin}->{Cout}", t_bytes))

        def backward():
            blk: List[Callable] = []
            if not out.g_init:
                return blk
            acc = self.bwd_arena.alloc(k * k, cp, f.Ck)
            if k == 2 and not self.fp32 and os.environ.get("MTBC_FUSE_CONVT_BWD", "1") != "0":
                # data, weight and bias gradient from ONE pass over the (N, 2H, 2W, Cout) gradient (convt_bwd.cu) instead
                # of three launches that each stream it from HBM; shapes the kernel does not serve keep the three below
                g = x.grad()
                try:
                    o = ops.convT_bwd_op(f, out.g, wd, acc, None if b is None else self.grad_view[b], g,
                                         accumulate=x.g_init)
                except _lib.MtbcError:
                    o = None
                if o is not None:
                    self.tc_flops_bwd += o.flops
                    blk.append(_mk_op(o, 2.0 * t_flops, f"{name} convT bwd {f.N}x{f.H}x{f.W} {Cin}<-{Cout} fused",
                                      2.0 * f.N * f.H * f.W * (2 * Cin + k * k * Cout) + 4.0 * Cin * Cout * k * k))
                    self.unpack_jobs.add(_lib.JOB_UNPACK_CONVT, [k * k * cp, f.Ck, Cin, Cout, k, 0], acc,
                                         self.grad_view[w], owner=w)
                    self._late_unpacked.append(w)
                    self._mark_grad(w, b)
                    x.g_init = True
                    return blk
            o = ops.convT_wgrad_op(f, out.g, acc, k)
            self.tc_flops_bwd += o.flops
            blk.append(self._side(_mk_op(o, t_flops, f"{name} convT wgrad {f.N}x{f.H}x{f.W} {Cin}->{Cout}", t_bytes)))
            self.unpack_jobs.add(_lib.JOB_UNPACK_CONVT, [k * k * cp, f.Ck, Cin, Cout, k, 0], acc, self.grad_view[w],
                                 owner=w)
            self._late_unpacked.append(w)
            if b is not None:
                l = _annot(_mk("mtbc_channel_sum", ptr(out.g.t), out.g.N * out.g.H * out.g.W, out.g.Cp, Cout,
                               ptr(self.grad_view[b]), 1), f"{name} convT bias grad",
                           2.0 * out.g.N * out.g.H * out.g.W * Cout)
                # like the weight gradients it only feeds the parameter-gradient buffer: side stream, next to the
                # weight gradient that reads the same tensor (MTBC_SIDE_BIAS=0 keeps it on the main stream)
                blk.append(self._side(l) if os.environ.get("MTBC_SIDE_BIAS", "1") != "0" else l)
            self._mark_grad(w, b)
            g = x.grad()
            o = ops.convT_dgrad_op(out.g, wd, g, k, accumulate=x.g_init, wd_lo=wd_lo)
            self.tc_flops_bwd += o.flops
            blk.append(_mk_op(o, t_flops, f"{name} convT dgrad {f.N}x{f.H}x{f.W} {Cin}<-{Cout}", t_bytes))
            x.g_init = True
            return blk
        self._bwd_blocks.append(backward)
        return out

    def upsample2(self, x: PTensor, name: str) -> PTensor:
        """nn.Upsample(scale_factor=2, mode='nearest') (Multi_BTS_UNet.py:100)."""
        f = x.feat
        out = self.new(f.N, f.H * 2, f.W * 2, f.C, name)
        self.fwd.append(_annot(_mk("mtbc_upsample2_fwd", ptr(f.t), f.N, f.H, f.W, f.Cp, ptr(out.feat.t)),
                               f"{name} upsample2 fwd", 2.0 * f.N * f.H * f.W * f.C * 5))

        def backward():
            if not out.g_init:
                return []
            g = x.grad()
            blk = [_annot(_mk("mtbc_upsample2_bwd", ptr(out.g.t), f.N, f.H, f.W, f.Cp, ptr(g.t), int(x.g_init)),
                          f"{name} upsample2 bwd", 2.0 * f.N * f.H * f.W * f.C * 5)]
            x.g_init = True
            return blk
        self._bwd_blocks.append(backward)
        return out

    def head1x1(self, a: PTensor, w: str, b: str, active: bool = True) -> Optional[torch.Tensor]:
        """Conv2d 1x1 C -> 1 mask head (MTnnUNet.py:6-9; MTUNetPlusPlus.py:73-76,120-123) -> fp32 logits (B,1,H,W)."""
        if not active:
            return None
        f = a.feat
        assert self.P(w).shape[0] == 1, "mask heads project to a single region"
        logits = torch.zeros(f.N, 1, f.H, f.W, dtype=torch.float32, device=self.device)
        dlog = torch.zeros_like(logits)
        npix = f.N * f.H * f.W
        self.fwd.append(self._side_fwd(_annot(_mk("mtbc_head1x1_fwd", ptr(f.t), npix, f.Cp, f.C, ptr(self.P(w)),
                                                  ptr(self.P(b)), ptr(logits)), f"{a.name} head1x1 fwd",
                                              npix * (2.0 * f.C + 4))))
        idx = len(self.outputs_seg)
        self.outputs_seg.append(logits)
        self.g_seg.append(dlog)

        def backward():
            if not self.seg_grad_active[idx]:
                return []
            g = a.grad()
            blk = [_annot(_mk("mtbc_head1x1_bwd", ptr(f.t), ptr(dlog), npix, f.Cp, f.C, ptr(self.P(w)), ptr(g.t),
                              int(a.g_init), ptr(self.grad_view[w]), ptr(self.grad_view[b])),
                          f"{a.name} head1x1 bwd", npix * (4.0 * f.C + 4))]
            a.g_init = True
            self._mark_grad(w, b)
            return blk
        self._bwd_blocks.append(backward)
        return logits

    def dshead(self, a: PTensor, wt: str, bt: str, w1: str, b1: str, k: int) -> torch.Tensor:
        """Deep-supervision head ConvTranspose2d(C,C,k,k) -> Conv2d 1x1 (C->1) (MTnnUNet.py:106-117), composed into a
        single C -> k*k projection so the k*k*C full-resolution intermediate never exists."""
        f = a.feat
        Cc, kk = f.C, k * k
        wc = torch.zeros(Cc, kk, dtype=torch.float32, device=self.device)
        bc = torch.zeros(1, dtype=torch.float32, device=self.device)
        logits = torch.zeros(f.N, 1, f.H * k, f.W * k, dtype=torch.float32, device=self.device)
        dlog = torch.zeros_like(logits)
        self.pack.append(_mk("mtbc_dshead_compose", ptr(self.P(wt)), ptr(self.P(bt)), ptr(self.P(w1)), ptr(self.P(b1)),
                             Cc, k, ptr(wc), ptr(bc)))
        self.fwd.append(self._side_fwd(_annot(_mk("mtbc_dshead_fwd", ptr(f.t), f.N, f.H, f.W, f.Cp, Cc, k, ptr(wc),
                                                  ptr(bc), ptr(logits)), f"{a.name} dshead k{k} fwd",
                                              f.N * f.H * f.W * (2.0 * Cc + 4 * kk))))
        idx = len(self.outputs_seg)
        self.outputs_seg.append(logits)
        self.g_seg.append(dlog)

        def backward():
            if not self.seg_grad_active[idx]:
                return []
            # per-block partial rows (+ one row of totals), written whole by the kernel: not in the zeroed arena
            nparts = int(_lib.load().mtbc_dshead_bwd_parts(f.N, f.H, f.W))
            dwc = torch.empty(nparts + 1, Cc * kk, dtype=torch.float32, device=self.device)
            dbc = torch.empty(nparts + 1, dtype=torch.float32, device=self.device)
            self._keep.extend((dwc, dbc))
            g = a.grad()
            blk = [_annot(_mk("mtbc_dshead_bwd", ptr(f.t), ptr(dlog), f.N, f.H, f.W, f.Cp, Cc, k, ptr(wc), ptr(g.t),
                              int(a.g_init), ptr(dwc), ptr(dbc), nparts), f"{a.name} dshead k{k} bwd",
                          f.N * f.H * f.W * (4.0 * Cc + 4 * kk)),
                   _mk("mtbc_dshead_decompose", ptr(dwc), ptr(dbc), nparts, ptr(self.P(wt)), ptr(self.P(bt)),
                       ptr(self.P(w1)), Cc, k, ptr(self.grad_view[wt]), ptr(self.grad_view[bt]), ptr(self.grad_view[w1]),
                       ptr(self.grad_view[b1]))]
            a.g_init = True
            self._mark_grad(wt, bt, w1, b1)
            return blk
        self._bwd_blocks.append(backward)
        return logits

    def gap_fc(self, a: PTensor, w1: str, b1: str, w2: str, b2: str) -> torch.Tensor:
        """AdaptiveAvgPool2d(1) -> Flatten -> Linear -> ReLU -> Linear (MTnnUNet.py:125-132)."""
        f = a.feat
        Fdim, Hd, K = self.P(w1).shape[1], self.P(w1).shape[0], self.P(w2).shape[0]
        assert Fdim == f.C
        gap = torch.zeros(f.N, Fdim, dtype=torch.float32, device=self.device)
        hid = torch.zeros(f.N, Hd, dtype=torch.float32, device=self.device)
        logits = torch.zeros(f.N, K, dtype=torch.float32, device=self.device)
        dlog = torch.zeros_like(logits)
        dgap = torch.zeros_like(gap)
        self.fwd.append(_annot(_mk("mtbc_gap_fc_fwd", ptr(f.t), f.N, f.H * f.W, f.Cp, Fdim, ptr(self.P(w1)),
                                   ptr(self.P(b1)), Hd, ptr(self.P(w2)), ptr(self.P(b2)), K, ptr(gap), ptr(hid),
                                   ptr(logits)), f"{a.name} gap_fc fwd", 2.0 * f.N * f.H * f.W * Fdim + 4.0 * Fdim * Hd))
        self.outputs_cls.append(logits)
        self.g_cls.append(dlog)

        def backward():
            g = a.grad()
            blk = [_annot(_mk("mtbc_gap_fc_bwd", ptr(dlog), f.N, f.H * f.W, f.Cp, Fdim, ptr(self.P(w1)), Hd,
                              ptr(self.P(w2)), K, ptr(gap), ptr(hid), ptr(g.t), int(a.g_init), ptr(self.grad_view[w1]),
                              ptr(self.grad_view[b1]), ptr(self.grad_view[w2]), ptr(self.grad_view[b2]), ptr(dgap)),
                          f"{a.name} gap_fc bwd", 2.0 * f.N * f.H * f.W * Fdim + 8.0 * Fdim * Hd)]
            a.g_init = True
            self._mark_grad(w1, b1, w2, b2)
            return blk
        self._bwd_blocks.append(backward)
        return logits

    def flat_fc(self, a: PTensor, w1: str, b1: str, w2: str, b2: str) -> torch.Tensor:
        """Flatten -> Linear(C*H*W, 256) -> ReLU -> Linear (Multi_BTS_UNet.py:107-115, BTS_UNET_classifier.py:89-95)."""
        f = a.feat
        Hd, K = self.P(w1).shape[0], self.P(w2).shape[0]
        assert self.P(w1).shape[1] == f.C * f.H * f.W, "the flatten head's Linear fixes the bottleneck extent"
        hid = torch.zeros(f.N, Hd, dtype=torch.float32, device=self.device)
        dh = torch.zeros(f.N, Hd, dtype=torch.float32, device=self.device)
        logits = torch.zeros(f.N, K, dtype=torch.float32, device=self.device)
        dlog = torch.zeros_like(logits)
        nfeat = f.C * f.H * f.W
        self.fwd.append(_annot(_mk("mtbc_flat_fc_fwd", ptr(f.t), f.N, f.H * f.W, f.Cp, f.C, ptr(self.P(w1)),
                                   ptr(self.P(b1)), Hd, ptr(self.P(w2)), ptr(self.P(b2)), K, ptr(hid), ptr(logits)),
                               f"{a.name} flat_fc fwd", 2.0 * f.N * nfeat + 4.0 * nfeat * Hd))
        self.outputs_cls.append(logits)
        self.g_cls.append(dlog)

        def backward():
            g = a.grad()
            blk = [_annot(_mk("mtbc_flat_fc_bwd", ptr(f.t), ptr(dlog), f.N, f.H * f.W, f.Cp, f.C, ptr(self.P(w1)), Hd,
                              ptr(self.P(w2)), K, ptr(hid), ptr(g.t), int(a.g_init), ptr(self.grad_view[w1]),
                              ptr(self.grad_view[b1]), ptr(self.grad_view[w2]), ptr(self.grad_view[b2]), ptr(dh)),
                          f"{a.name} flat_fc bwd", 4.0 * f.N * nfeat + 8.0 * nfeat * Hd)]
            a.g_init = True
            self._mark_grad(w1, b1, w2, b2)
            return blk
        self._bwd_blocks.append(backward)
        return logits

    def softmax_cls(self) -> torch.Tensor:
        """nn.Softmax(dim=1) on the class logits emitted last (nnUNet_classifier.py:165-166): the plan's class output
        becomes the probabilities, and their gradient is mapped back to the logits' before the head's backward."""
        logits, dlog = self.outputs_cls[-1], self.g_cls[-1]
        N, K = logits.shape
        probs = torch.zeros_like(logits)
        dprobs = torch.zeros_like(logits)
        self.fwd.append(_mk("mtbc_softmax_rows_fwd", ptr(logits), N, K, ptr(probs)))
        self.outputs_cls[-1] = probs
        self.g_cls[-1] = dprobs

        def backward():
            return [_mk("mtbc_softmax_rows_bwd", ptr(probs), ptr(dprobs), N, K, ptr(dlog))]
        self._bwd_blocks.append(backward)
        return probs


    # ------------------------------------------------------------------------------------------------ ResidualUNet
    # BatchNorm / residual / dropout sibling backbone (reference src/models/segmentation/ResidualUNet.py).  The tensor
    # kernels are the multi-task path's; BatchNorm2d is the InstanceNorm passes over batch-pooled sums (residual.cu).
    def conv_plain(self, srcs: Optional[Sequence[PTensor]], w: str, b: Optional[str], name: str, stride: int = 1,
                   stats: bool = False, first_input: Optional[torch.Tensor] = None):
        """Conv2d 3x3 (stride 1 or 2, padding 1, bias ADDED) whose output is used as is (a BatchNorm2d or a residual add
        follows, ResidualUNet.py:35-58,113-131).  Returns (y, stat_sum, stat_sq); the per-(sample, channel) sums are
        None unless `stats`."""
        Wt = self.P(w)
        Cout, Cin = Wt.shape[0], Wt.shape[1]
        if first_input is not None:
            N, H, Wd = self.B, self.H, self.W
            feats, x0 = None, None
        else:
            feats = [t.feat for t in srcs]
            assert sum(f.C for f in feats) == Cin and (stride == 1 or len(feats) == 1)
            x0 = srcs[0]
            N, H, Wd = feats[0].N, feats[0].H // stride, feats[0].W // stride
        y = self.new(N, H, Wd, Cout, name + ".y")
        Cp, npx = y.feat.Cp, N * H * Wd
        ssum = self.fwd_arena.alloc(N, Cp) if stats else None
        ssq = self.fwd_arena.alloc(N, Cp) if stats else None
        fused = stats and H * Wd >= 128 and not self.deterministic
        pk = None
        if first_input is not None:
            bvec = None if b is None else self.P(b)
            l = _annot(_mk("mtbc_conv_first_fwd", ptr(first_input), N, Cin, H, Wd, ptr(Wt), ptr(bvec), Cout,
                           ptr(y.feat.t), Cp, ptr(ssum) if fused else None, ptr(ssq) if fused else None, None),
                       f"{name} first conv fwd {N}x{H}x{Wd} {Cin}->{Cout}", npx * (4 * Cin + 2 * Cout))
            l.true_flops = 2.0 * npx * Cout * Cin * 9
            self.fwd.append(l)
        else:
            pk = self._pack_conv(w, feats, y.feat.Ck)
            bias = self._padded_vec(b, y.feat.Ck)
            kw = dict(bias=bias, stat_sum=ssum if fused else None, stat_sq=ssq if fused else None,
                      wpack_lo=pk.get("wf_lo"))
            op = (ops.conv3x3_fwd_op(feats, pk["wf"], y.feat, **kw) if stride == 1 else
                  ops.conv3x3_s2_fwd_op(feats[0], pk["wf"], y.feat, **kw))
            self.tc_flops_fwd += op.flops
            self.fwd.append(_mk_op(op, 2.0 * npx * Cout * Cin * 9,
                                   f"{name} fwd s{stride} {N}x{H}x{Wd} {[f.C for f in feats]}->{Cout}",
                                   2.0 * npx * (Cin * stride * stride + Cout) + 18.0 * Cin * Cout))
        if stats and not fused:
            self.fwd.append(self._stats_launch(y, ssum, ssq, name))

        def backward():
            blk: List[Callable] = []
            if not y.g_init:
                return blk
            dy = y.g
            if first_input is not None:
                blk.append(self._side(_annot(_mk("mtbc_conv_first_wgrad", ptr(first_input), N, Cin, H, Wd, ptr(dy.t),
                                                 dy.Cp, Cout, ptr(self.grad_view[w])),
                                             f"{name} first conv wgrad", npx * (4 * Cin + 2 * Cout))))
            elif stride == 2:
                acc = self._wg_accum(w, pk)
                o = ops.conv3x3_s2_wgrad_op(feats[0], dy, acc)
                self.tc_flops_bwd += o.flops
                blk.append(self._side(_mk_op(o, 2.0 * npx * Cout * Cin * 9, f"{name} wgrad s2 {Cin}x{Cout}")))
            else:
                acc = self._wg_accum(w, pk)
                for f, off in zip(feats, pk["offs"]):
                    o = ops.conv3x3_wgrad_op(f, dy, acc, off)
                    self.tc_flops_bwd += o.flops
                    blk.append(self._side(_mk_op(o, 2.0 * npx * Cout * f.C * 9, f"{name} wgrad {f.C}x{Cout}")))
            if b is not None:
                blk.append(_annot(_mk("mtbc_channel_sum", ptr(dy.t), npx, dy.Cp, Cout, ptr(self.grad_view[b]), 1),
                                  f"{name} bias grad", 2.0 * npx * Cout))
            self._mark_grad(w, b)
            if first_input is not None:
                return blk                      # the image needs no gradient
            if stride == 2:
                up = self._keep_feat(N, 2 * H, 2 * Wd, Cout)
                blk.append(_annot(_mk("mtbc_zero_stuff2", ptr(dy.t), N, H, Wd, dy.Cp, ptr(up.t)),
                                  f"{name} zero-stuff dy", 2.0 * npx * Cout * 5))
                self._emit_dgrad(blk, [x0], up, w, pk, name, Cout)
            else:
                self._emit_dgrad(blk, srcs, dy, w, pk, name, Cout)
            return blk
        self._bwd_blocks.append(backward)
        return y, ssum, ssq

    def _keep_feat(self, N, H, W, Cc) -> Feat:
        f = Feat.empty(N, H, W, Cc, device=self.device, dtype=self.act_dtype)
        self._keep.append(f.t)
        return f

    def bn_act(self, x: PTensor, bn: torch.nn.Module, prefix: str, slope: float, p_drop: float, name: str,
               stats=None, training: bool = True) -> PTensor:
        """nn.BatchNorm2d (eps 1e-5, momentum 0.1, affine) -> leaky_relu(slope; slope 1 = no activation) ->
        F.dropout(p) on tensor x (ResidualUNet.py:58-61,137-145).  `stats` = (sum, sum of squares) per (sample, channel)
        when the producing conv delivered them, else they are reduced here."""
        f = x.feat
        N, H, Wd, Cp, Cc = f.N, f.H, f.W, f.Cp, f.C
        gamma, beta = prefix + ".weight", prefix + ".bias"
        if stats is None or stats[0] is None:
            ssum, ssq = self.fwd_arena.alloc(N, Cp), self.fwd_arena.alloc(N, Cp)
            self.fwd.append(self._stats_launch(x, ssum, ssq, name))
        else:
            ssum, ssq = stats
        self.fwd.append(_mk("mtbc_bn_pool_fwd", ptr(ssum), ptr(ssq), N, Cp, Cc, H * Wd, int(training),
                            C.c_float(float(bn.momentum if bn.momentum is not None else 0.1)), ptr(bn.running_mean),
                            ptr(bn.running_var), ptr(bn.num_batches_tracked)))
        a = self.new(N, H, Wd, Cc, name)
        mean = torch.zeros(N, Cp, dtype=torch.float32, device=self.device)
        rstd = torch.zeros(N, Cp, dtype=torch.float32, device=self.device)
        gv, bv = self._padded_vec(gamma, f.Ck), self._padded_vec(beta, f.Ck)
        nbytes = 2.0 * N * H * Wd * Cc * (2 if self.fp32 else 1)
        self.fwd.append(_annot(_mk("mtbc_in_apply", ptr(f.t), N, H, Wd, Cp, ptr(ssum), ptr(ssq), ptr(gv), ptr(bv), Cc,
                                   C.c_float(float(bn.eps)), C.c_float(slope), ptr(a.feat.t), None, ptr(mean),
                                   ptr(rstd)), f"{name} bn_apply {N}x{H}x{Wd}x{Cc}", 2 * nbytes))
        out, mask = a, None
        nel = N * H * Wd * Cp
        if p_drop > 0.0:
            out = self.new(N, H, Wd, Cc, name + ".drop")
            mask = torch.zeros(nel, dtype=torch.uint8, device=self.device)
            self.dropout_masks.append(mask)
            self.fwd.append(_annot(_mk("mtbc_dropout_fwd", ptr(a.feat.t), ptr(out.feat.t), ptr(mask), nel,
                                       C.c_float(p_drop), C.c_uint64(self.dropout_seed), ptr(self._drop_counter()),
                                       len(self.dropout_masks), int(self.dropout_external)),
                                   f"{name} dropout", 2 * nbytes))

        def backward():
            blk: List[Callable] = []
            if not out.g_init:
                return blk
            if mask is not None:
                g = a.grad()
                blk.append(_mk("mtbc_dropout_bwd", ptr(out.g.t), ptr(mask), ptr(g.t), nel, C.c_float(p_drop), 0))
                a.g_init = True
            s1, s2 = self.bwd_arena.alloc(N, Cp), self.bwd_arena.alloc(N, Cp)
            dx = self.scratch(N, H, Wd, Cc)
            blk.append(_mk("mtbc_in_bwd_reduce", ptr(a.g.t), ptr(f.t), N, H * Wd, Cp, ptr(mean), ptr(rstd), ptr(gv),
                           ptr(bv), C.c_float(slope), ptr(s1), ptr(s2)))
            blk.append(_mk("mtbc_bn_pool_bwd", ptr(s1), ptr(s2), N, Cp, Cc, int(training), ptr(self.grad_view[gamma]),
                           ptr(self.grad_view[beta])))
            l = _annot(_mk("mtbc_in_bwd_apply", ptr(a.g.t), ptr(f.t), N, H * Wd, Cp, ptr(mean), ptr(rstd), ptr(gv),
                           ptr(bv), C.c_float(slope), ptr(s1), ptr(s2), ptr(dx.t), None, None, Cc),
                       f"{name} bn_bwd", 3 * nbytes)
            l.wait_side = self._slot_readers.pop(self._cur_slot, [])
            blk.append(l)
            g = x.grad()
            blk.append(_mk("mtbc_accumulate", ptr(dx.t), ptr(g.t), nel, int(x.g_init)))
            x.g_init = True
            self._mark_grad(gamma, beta)
            return blk
        self._bwd_blocks.append(backward)
        return out

    def _drop_counter(self) -> torch.Tensor:
        """Device-side draw counter of this plan's dropout masks: advanced once at the head of every forward (also
        under CUDA-graph replay), so every step draws fresh masks."""
        if self._drop_ctr is None:
            self._drop_ctr = torch.zeros(1, dtype=torch.int32, device=self.device)
            self.fwd.insert(0, _mk("mtbc_increment_i32", ptr(self._drop_ctr)))
        return self._drop_ctr

    def add(self, a: PTensor, b: PTensor, name: str) -> PTensor:
        """`path + residual` (ResidualUNet.py:69,155)."""
        fa, fb = a.feat, b.feat
        assert fa.t.shape == fb.t.shape
        out = self.new(fa.N, fa.H, fa.W, fa.C, name)
        nel = fa.t.numel()
        self.fwd.append(_annot(_mk("mtbc_add", ptr(fa.t), ptr(fb.t), ptr(out.feat.t), nel), f"{name} add",
                               3.0 * nel * fa.t.element_size()))

        def backward():
            blk: List[Callable] = []
            if not out.g_init:
                return blk
            for t in (a, b):
                g = t.grad()
                blk.append(_mk("mtbc_accumulate", ptr(out.g.t), ptr(g.t), nel, int(t.g_init)))
                t.g_init = True
            return blk
        self._bwd_blocks.append(backward)
        return out

    # ------------------------------------------------------------------------------------------------ finalize / run
    def finalize(self, seg_grad_active: Optional[Sequence[bool]] = None):
        """Emit the backward launch list (reverse order of the forward ops)."""
        assert not self.finalized
        self.seg_grad_active = list(seg_grad_active) if seg_grad_active is not None else [True] * len(self.outputs_seg)
        if self.training:
            body: List[Callable] = []
            n_pre = 1 + len(self.bwd_arena.chunks) + 1  # zero launches precede the body (upper bound fixed below)
            for mk_block in reversed(self._bwd_blocks):
                self._touched = set()
                body.extend(mk_block())
                for n in self._touched:  # the last block touching a parameter finalises its gradient
                    self.param_done_at[n] = len(body)
            body = self._bucketize(body)
            self.bwd = [_mk("mtbc_zero_bytes", ptr(self.grad_flat), self.grad_flat.numel() * 4)]
            self.bwd += self.bwd_arena.zero_launches()
            n_pre = len(self.bwd)
            self.bwd += body
            self.param_done_at = {n: i + n_pre for n, i in self.param_done_at.items()}
        pack_jobs = self.pack_jobs.launch()
        if pack_jobs and os.environ.get("MTBC_SIDE_PACK", "1") != "0":
            # The parameter pack (one launch, latency bound: ~0.1 ms for 15 M parameters; bf16 conv operands and the padded
            # gamma / beta / bias vectors) is not needed by the first layer's conv, which reads the fp32 image and the
            # fp32 master weights on the CUDA cores.  TrainStep forks the pack onto the side stream; the first forward
            # launch that may read packed data (anything after the first-layer conv) waits for it.
            free = ("mtbc_zero_bytes", "mtbc_conv_first_fwd", "mtbc_increment_i32")
            waiter = next((l for l in self.fwd if l.kind not in free), None)
            if waiter is not None and getattr(waiter, "wait_side", None) is None and self.fwd and \
                    any(l.kind == "mtbc_conv_first_fwd" for l in self.fwd):
                for l in pack_jobs:
                    l.side = True
                waiter.wait_side = list(pack_jobs)
        self.pack = pack_jobs + self.pack
        self.fwd = self.fwd_arena.zero_launches() + self.fwd
        self._bwd_blocks = []
        self.finalized = True
        global _build_mode
        _build_mode = 0

    def _bucketize(self, body: List[Callable]) -> List[Callable]:
        """Split the flat gradient buffer into `n_buckets` contiguous ranges of about equal size and make each range
        final as early as the backward order allows: the weight-gradient unpack jobs of a bucket's parameters run as
        one launch right after the last backward block that touches any of them, followed by a `bucket_ready` marker
        (a no-op launch that TrainStep turns into the fork point of that bucket's gradient all-reduce).  The flat layout
        is the reverse registration order, i.e. about the order in which backward finishes parameters, so bucket 0
        (heads, last decoder node) is ready after a fraction of the backward pass."""
        nb = max(1, int(os.environ.get("MTBC_DP_BUCKETS", "4")))
        names = [n for n in self.grad_range if self.has_grad.get(n)]
        self.buckets, bucket_idx, ready = plan_buckets({n: self.grad_range[n] for n in self.grad_range},
                                                       {n: self.param_done_at.get(n, len(body)) for n in names},
                                                       self.grad_flat.numel(), nb)

        def bucket_of(n):
            return bucket_idx[n]
        jobs_of = [[i for i, o in enumerate(self.unpack_jobs.owner) if bucket_of(o) == k] for k in range(nb)]
        out: List[Callable] = []
        emitted = [False] * nb
        done_after: Dict[int, int] = {}

        def flush(i):
            for k in range(nb):
                if not emitted[k] and ready[k] <= i:
                    for ul in self.unpack_jobs.launch(jobs_of[k]):
                        ul.wait_side = "all"       # the accumulators it reads are written by side-stream launches
                        out.append(ul)
                    mk = _mk_marker(k)
                    mk.wait_side = "all"           # (a bucket without unpack jobs may still hold first-layer grads)
                    out.append(mk)
                    done_after[k] = len(out)
                    emitted[k] = True
        flush(0)
        for i, l in enumerate(body):
            out.append(l)
            flush(i + 1)
        self.param_done_at = {n: done_after[bucket_of(n)] for n in names}
        return out

    def run_pack(self, stream=None):
        st = C.c_void_p(stream if stream is not None else ops.stream_ptr())
        for l in self.pack:
            l(st)

    def run_forward(self, stream=None):
        st = C.c_void_p(stream if stream is not None else ops.stream_ptr())
        for l in self.fwd:
            l(st)

    def run_backward(self, stream=None):
        st = C.c_void_p(stream if stream is not None else ops.stream_ptr())
        for l in self.bwd:
            l(st)

    def launch_counts(self) -> Dict[str, int]:
        return {"pack": len(self.pack), "fwd": len(self.fwd), "bwd": len(self.bwd)}


def plan_buckets(ranges: Dict[str, Tuple[int, int]], done_at: Dict[str, int], total: int, nb: int):
    """Pure bucket arithmetic (unit-tested on the CPU): split [0, total) into nb contiguous ranges of about equal size;
    a parameter belongs to the bucket that contains its first element; a bucket is ready once every one of its
    parameters that receives a gradient is done (done_at[name] = position in the backward launch list), and never
    before the bucket in front of it (all-reduces are issued in bucket order on every rank).
    Returns ([(lo, hi)], {name: bucket}, [ready position])."""
    cuts = [total * (k + 1) // nb for k in range(nb)]
    lo = [0] + cuts[:-1]
    idx = {}
    for n, (a, _) in ranges.items():
        k = 0
        while k < nb - 1 and a >= cuts[k]:
            k += 1
        idx[n] = k
    ready = [0] * nb
    for n, d in done_at.items():
        ready[idx[n]] = max(ready[idx[n]], d)
    for k in range(1, nb):
        ready[k] = max(ready[k], ready[k - 1])
    return [(lo[k], cuts[k]) for k in range(nb)], idx, ready


def _mk_marker(k: int) -> Callable:
    """No-op entry of a launch list: 'bucket k of the flat gradient buffer is final from here on'."""
    def launch(stream):
        return None
    launch.kind = "bucket_ready"
    launch.bucket = k
    return launch


def flat_layout(params) -> Tuple[Dict[str, Tuple[int, int]], int]:
    """Offsets of every parameter inside the flat fp32 gradient (and, for TrainStep, parameter / Adam-state) buffers:
    reversed registration order (~ the order in which backward finishes them), each slot padded to 64 floats."""
    ranges: Dict[str, Tuple[int, int]] = {}
    off = 0
    for n in reversed(list(params.keys())):
        k = params[n].numel()
        ranges[n] = (off, off + k)
        off += (k + 63) // 64 * 64
    return ranges, off


def _mk_copy_f32(dst: torch.Tensor, src: torch.Tensor, n: int) -> Callable:
    """Refresh a zero-padded fp32 copy of a small parameter vector (device-to-device, graph-capturable)."""
    l = _mk("mtbc_copy_f32", ptr(dst), ptr(src), n)
    l._keep = (dst, src)
    return l
