"""Sync-free training step for the multi-task models (the loop body of src/training_multitask.py:79-103).

    zero_grad -> forward -> Dice(+deep supervision weights) + focal -> alpha mix -> backward -> Adam(lr, eps=1e-4)

Everything between the host->device copy of the batch and the optimizer update is a flat list of C-ABI launches on one
stream (plan.py), captured once into a CUDA graph and replayed.  The reference's per-step host syncs (`torch.isnan` in a
Python `if`, `.item()` on the loss: criterions.py:72, training_multitask.py:99) become a device-side NaN flag and a
4-float device buffer that the caller reads when it wants to.

Data parallel (new functionality, the reference is single device): one process per GPU; the flat fp32 gradient buffer
is split into a few contiguous buckets (plan.Plan._bucketize), each bucket's NCCL all-reduce (sum) is forked onto a
communication stream at the point of the backward pass where the bucket becomes final, so it overlaps the remaining
data- / weight-gradient kernels, and the optimizer joins all of them; the 1/world scaling is folded into the Adam
kernel.  The step is captured as one CUDA graph per segment between bucket boundaries (forward + loss + the first part
of backward, ..., the optimizer); the NCCL calls are issued eagerly between the replays, on the communication stream
(capturing the collectives themselves inside the graph hung with the NCCL 2.28.9 / torch 2.11 of this image).  No
operation couples samples
(InstanceNorm is per sample, Dice is per sample then mean, focal is per sample then mean), so R ranks x B/R samples
equals 1 rank x B samples up to fp32 summation order.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib
from .ops import ptr, stream_ptr
from .plan import Plan, _mk, flat_layout


def bind_to_gpu_numa(device_index: int) -> Optional[List[int]]:
    """Pin the calling process to the host cores NVML reports as local to GPU `device_index` (call before allocating
    pinned batches: first touch then places them on the GPU's NUMA node).  One process per GPU, all started by torchrun
    on whatever cores the launcher ran on, otherwise stage every rank's batch through the same socket.  Returns the
    core list, or None when NVML or the affinity call is unavailable (nothing is changed then)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cores = [64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1]
        cores = [c for c in cores if c in os.sched_getaffinity(0)] or cores
        if cores:
            os.sched_setaffinity(0, cores)
        return cores
    except Exception:  # noqa: BLE001 - best effort: an unbound process is still correct
        return None


class TrainStep:
    def __init__(self, model: torch.nn.Module, batch_shape: Sequence[int], lr: float = 1e-4, betas=(0.9, 0.999),
                 eps: float = 1e-4, alpha: float = 0.35, inversely_weighted: bool = True, focal_alpha: float = 1.0,
                 focal_gamma: float = 2.0, process_group=None, use_graph: bool = True, device=None,
                 refine: bool = False, refine_flags=(True, True, 0), normal_id: int = 2,
                 share_state_with: Optional["TrainStep"] = None, external_cotangent: bool = False):
        """`model` is one of models.{MTUNetPlusPlus, MTnnUNet, Multi_BTS_UNet} already on its CUDA device.
        Hyper-parameters keep the meaning of src/config.yaml (optimizer.lr, training.alpha, loss.inversely_weighted).

        `share_state_with`: a TrainStep of the SAME model planned for another batch shape (the loader's ragged last
        batch, BUSI_dataloader.py:146-148: no drop_last).  The new step gets its own plan / activations / graph but
        updates the very same flat parameter buffer, Adam moments, step counter and learning-rate scalar."""
        self.model = model
        self.device = device or next(model.parameters()).device
        if self.device.type != "cuda":
            raise _lib.MtbcError("TrainStep needs a CUDA (sm_100a) device; there is no CPU fallback")
        self.B, self.Cin, self.H, self.W = (int(v) for v in batch_shape)
        self.alpha, self.inv_w = float(alpha), bool(inversely_weighted)
        self.focal_alpha, self.focal_gamma = float(focal_alpha), float(focal_gamma)
        self.betas, self.eps = (float(betas[0]), float(betas[1])), float(eps)
        self.pg = process_group
        self.world = 1
        if process_group is not None:
            import torch.distributed as dist
            self.world = dist.get_world_size(process_group)
        self.use_graph = use_graph
        # refine=True: the prediction-refining module (utils/models.py:316-332,366-386) runs inside the step (and its
        # CUDA graph) on the full-decoder logits: `refined_mask` (uint8), `refined_class`, `pixel_count` (int32).
        # Any eager launch between two graph replays exposes ~0.35 ms of graph start-up per step (tools/diag_e2e.py),
        # which is why it lives in the graph and not in the caller's loop.
        self.refine, self.refine_flags, self.normal_id = bool(refine), tuple(refine_flags), int(normal_id)
        # external_cotangent=True (gradient-wiring / data-parallel checks): the fused objective is left out and the
        # caller fills plan.g_cls / plan.g_seg itself, i.e. the step back-propagates sum(out * g) for fixed g
        self.external_cotangent = bool(external_cotangent)
        self._shared = share_state_with
        if share_state_with is not None and share_state_with.model is not model:
            raise ValueError("share_state_with: the other TrainStep drives a different model")
        with torch.cuda.device(self.device):
            if share_state_with is None:
                self._flatten_params()
            else:
                self.flat_p, self.param_ranges = share_state_with.flat_p, share_state_with.param_ranges
            x = torch.zeros(self.B, self.Cin, self.H, self.W, dtype=torch.float32, device=self.device)
            self.plan: Plan = model._get_plan(x, True)
            self.x = self.plan.x_in
            self.mask = torch.zeros(self.B, 1, self.H, self.W, dtype=torch.float32, device=self.device)
            K = self.plan.outputs_cls[0].shape[1]
            self.onehot = torch.zeros(self.B, K, dtype=torch.float32, device=self.device)
            self.K = K
            nh = len(self.plan.outputs_seg)
            self.nheads = nh
            self.dice_sums = torch.zeros(nh, self.B, 3, dtype=torch.float32, device=self.device)
            self.dice_loss = torch.zeros(nh, dtype=torch.float32, device=self.device)  # [0] = full decoder head
            self.focal_loss = torch.zeros(1, dtype=torch.float32, device=self.device)
            self.loss_out = torch.zeros(4, dtype=torch.float32, device=self.device)  # total, seg, cls, nan flag
            self.lr = float(lr)   # host copy (exact double) of the device scalar the captured Adam launch reads
            if share_state_with is None:
                self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=self.device)
                self.step_dev = torch.zeros(1, dtype=torch.int32, device=self.device)
                self.exp_avg = torch.zeros_like(self.flat_p)
                self.exp_avg_sq = torch.zeros_like(self.flat_p)
            else:
                o = share_state_with
                self.lr, self.lr_dev, self.step_dev = o.lr, o.lr_dev, o.step_dev
                self.exp_avg, self.exp_avg_sq = o.exp_avg, o.exp_avg_sq
            if self.refine:
                self.refined_mask = torch.zeros(self.B, 1, self.H, self.W, dtype=torch.uint8, device=self.device)
                self.refined_class = torch.zeros(self.B, dtype=torch.int32, device=self.device)
                self.pixel_count = torch.zeros(self.B, dtype=torch.int32, device=self.device)
            import os as _os
            self.overlap = self.world > 1 and _os.environ.get("MTBC_DP_OVERLAP", "1") != "0"
            # Per-bucket optimizer (MTBC_DP_BUCKET_ADAM=1, off by default): bucket k's Adam runs on the communication
            # stream right behind its all-reduce, i.e. while bucket k+1 is still being reduced and the backward pass is
            # still running; only the last bucket's update is left exposed instead of one Adam launch over all
            # parameters that waits for every all-reduce.  Safe because a bucket is only marked ready once every kernel
            # that reads its fp32 master parameters in the backward pass (heads, FC layers) has run; the convolutions
            # read the bf16 operands packed at the head of the step.  MEASURED, same 8 x B200 box, back to back
            # (profiles/r02g_bench_n8*.json): 11.96 ms/step with it, 11.87 ms without -- the four small Adam launches
            # compete with the backward kernels for HBM and sit between the all-reduces on the communication stream,
            # which costs more than the ~0.05 ms of exposed tail it removes.  Kept as a switch, not as the default.
            self.bucket_adam = self.overlap and _os.environ.get("MTBC_DP_BUCKET_ADAM", "0") == "1"
            self._build_launches()
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self._heads: Dict[Optional[int], torch.cuda.CUDAGraph] = {}   # first graph of a step, per staging slot (None: no copy)
        # read-back of the loss vector (losses_to_host): stream, two snapshots, completion events; with a prefetched host
        # batch the snapshot is the last node of the step's head graph (MTBC_SNAP_IN_GRAPH=0: an eager copy after it)
        self._read_stream: Optional[torch.cuda.Stream] = None
        self._loss_snap: List[torch.Tensor] = []
        self._read_done: List[Optional[torch.cuda.Event]] = [None, None]
        self._read_turn = 0
        self._snap_slot: Optional[int] = None
        import os as _os2
        self._snap_in_graph = _os2.environ.get("MTBC_SNAP_IN_GRAPH", "1") != "0"
        self.comm_stream: Optional[torch.cuda.Stream] = None
        self.side_stream: Optional[torch.cuda.Stream] = None
        self.side_enabled = _os2.environ.get("MTBC_SIDE_WGRAD", "1") != "0"
        self._copy_stream: Optional[torch.cuda.Stream] = None
        self._pending = None
        self._load_turn = 0
        self.seg_graphs = None   # data parallel: [(CUDAGraph, bucket index or None)] + opt_graph
        self.opt_graph = None
        self.steps_done = 0
        self._pinned: Dict[str, torch.Tensor] = {}

    # ------------------------------------------------------------------------------------------------ set-up
    def _flatten_params(self):
        """Re-home every parameter as a view of one flat fp32 buffer laid out like the plan's gradient buffer, so Adam
        and the gradient all-reduce are single launches.  Values, shapes, dtypes and state_dict keys are unchanged."""
        params = dict(self.model.named_parameters())
        ranges, total = flat_layout(params)
        self.flat_p = torch.zeros(total, dtype=torch.float32, device=self.device)
        for n, p in params.items():
            a, b = ranges[n]
            v = self.flat_p[a:b].view_as(p)
            v.copy_(p.data)
            p.data = v
        self.param_ranges = ranges

    def _build_launches(self):
        plan = self.plan
        nh, B, HW = self.nheads, self.B, self.H * self.W
        L: List = []
        L += plan.pack
        L += plan.fwd
        if self.external_cotangent:
            L += plan.bwd
            self.launches_fb = L
            self._build_opt_launches()
            return
        # ---- fused objective (criterions.py:52-76 + training_multitask.py:98)
        L.append(_mk("mtbc_zero_bytes", ptr(self.dice_sums), self.dice_sums.numel() * 4))
        L[-1].wait_side = "all"     # the mask heads' forward launches run on the side stream (plan._side_fwd)
        for i, logits in enumerate(plan.outputs_seg):
            j = nh - 1 - i  # reversed list: the last head (full decoder) gets weight 1
            L.append(_mk("mtbc_dice_sums", ptr(logits), ptr(self.mask), B, HW, ptr(self.dice_sums[i])))
            L.append(_mk("mtbc_dice_finalize", ptr(self.dice_sums[i]), B, ptr(self.dice_loss[j:j + 1])))
        L.append(_mk("mtbc_focal_fwd", ptr(plan.outputs_cls[0]), ptr(self.onehot), B, self.K,
                     C.c_float(self.focal_alpha), C.c_float(self.focal_gamma), ptr(self.focal_loss)))
        L.append(_mk("mtbc_multitask_loss", ptr(self.dice_loss), nh, int(self.inv_w), ptr(self.focal_loss),
                     C.c_float(self.alpha), ptr(self.loss_out)))
        if self.refine:
            sbc, cbs, thr = self.refine_flags
            L.append(_mk("mtbc_refine_predictions", ptr(plan.outputs_seg[-1]), ptr(plan.outputs_cls[0]), B, HW, self.K,
                         self.normal_id, int(bool(sbc)), int(bool(cbs)), int(thr), ptr(self.refined_mask),
                         ptr(self.refined_class), ptr(self.pixel_count)))
        # ---- d(total)/d(logits)
        for i, logits in enumerate(plan.outputs_seg):
            j = nh - 1 - i
            wgt = self.alpha * (1.0 / (j + 1) if self.inv_w else 1.0)
            L.append(_mk("mtbc_dice_bwd", ptr(logits), ptr(self.mask), B, HW, ptr(self.dice_sums[i]), None,
                         C.c_float(wgt), ptr(plan.g_seg[i])))
        L.append(_mk("mtbc_focal_bwd", ptr(plan.outputs_cls[0]), ptr(self.onehot), B, self.K,
                     C.c_float(self.focal_alpha), C.c_float(self.focal_gamma), None, C.c_float(1.0 - self.alpha),
                     ptr(plan.g_cls[0])))
        L += plan.bwd
        self.launches_fb = L
        self._build_opt_launches()

    def _adam_launch(self, lo: int, hi: int):
        plan = self.plan
        return _mk("mtbc_adam_step_dev", ptr(self.flat_p[lo:hi]), ptr(plan.grad_flat[lo:hi]), ptr(self.exp_avg[lo:hi]),
                   ptr(self.exp_avg_sq[lo:hi]), hi - lo, ptr(self.lr_dev), C.c_float(self.betas[0]),
                   C.c_float(self.betas[1]), C.c_float(self.eps), C.c_float(1.0 / self.world), ptr(self.step_dev))

    def _build_opt_launches(self):
        # ---- optimizer (after the all-reduce when data parallel)
        inc = _mk("mtbc_increment_i32", ptr(self.step_dev))
        if self.bucket_adam:
            # the step counter moves at the head of the step; each bucket's update follows its own all-reduce
            self.launches_fb = [inc] + self.launches_fb
            self.launches_opt = []
            self.launches_opt_bucket = {k: [self._adam_launch(lo, hi)] for k, (lo, hi) in enumerate(self.plan.buckets)
                                        if hi > lo}
            self._bucket_opt_graphs = {}
            return
        self.launches_opt = [inc, self._adam_launch(0, self.flat_p.numel())]

    # ------------------------------------------------------------------------------------------------ running
    @property
    def n_launches(self) -> int:
        """Kernel / memset launches of one step (our own kernels only; the NCCL all-reduces are not counted)."""
        extra = sum(len(v) for v in getattr(self, "launches_opt_bucket", {}).values())
        return sum(1 for l in self.launches_fb + self.launches_opt if l.kind != "bucket_ready") + extra

    def _run_list(self, launches):
        """Launch a list on the current stream.  Launches marked `side` (weight gradients: they only feed accumulators
        read by the bucket unpack) are forked onto a second stream so they overlap the HBM-bound InstanceNorm backward
        and the data gradients of the following layers; `wait_side` on a launch lists the side launches it must not
        overtake (a rotating dy buffer about to be overwritten) or "all".  Same code eagerly and under graph capture."""
        main = torch.cuda.current_stream(self.device)
        st = C.c_void_p(main.cuda_stream)
        if not self.side_enabled:
            for l in launches:
                l(st)
            return
        if self.side_stream is None:
            self.side_stream = torch.cuda.Stream(device=self.device)
        side = self.side_stream
        sst = C.c_void_p(side.cuda_stream)
        used = False
        # Events recorded by an EARLIER call are already covered by that call's final join (and, under the data-parallel
        # per-segment graph capture, belong to another capture: waiting on them is cudaErrorInvalidValue), so a
        # dependency is only waited for when its side launch was issued by this very call.
        self._run_gen = getattr(self, "_run_gen", 0) + 1
        gen = self._run_gen
        for l in launches:
            if getattr(l, "side", False):
                side.wait_stream(main)
                l(sst)
                ev = torch.cuda.Event()
                ev.record(side)
                l.done_event = ev
                l.done_gen = gen
                used = True
                continue
            ws = getattr(l, "wait_side", None)
            if ws == "all":
                if used:
                    main.wait_stream(side)
            elif ws:
                for dep in ws:
                    ev = getattr(dep, "done_event", None)
                    if ev is not None and getattr(dep, "done_gen", -1) == gen:
                        main.wait_event(ev)
            l(st)
        if used:
            main.wait_stream(side)

    def _segments(self):
        """launches_fb split at the bucket markers: [(launches, bucket index or None)]."""
        segs, cur = [], []
        for l in self.launches_fb:
            if l.kind == "bucket_ready":
                segs.append((cur, l.bucket))
                cur = []
            else:
                cur.append(l)
        segs.append((cur, None))
        return segs

    def _fork_allreduce(self, k):
        """All-reduce of gradient bucket k on the communication stream, ordered after everything enqueued so far."""
        import torch.distributed as dist
        lo, hi = self.plan.buckets[k]
        if hi <= lo:
            return
        main = torch.cuda.current_stream(self.device)
        if self.comm_stream is None:
            self.comm_stream = torch.cuda.Stream(device=self.device)
        self.comm_stream.wait_stream(main)
        with torch.cuda.stream(self.comm_stream):
            dist.all_reduce(self.plan.grad_flat[lo:hi], op=dist.ReduceOp.SUM, group=self.pg)
            if self.bucket_adam and k in self.launches_opt_bucket:
                g = self._bucket_opt_graphs.get(k)
                if g is not None:
                    g.replay()
                else:
                    st = C.c_void_p(self.comm_stream.cuda_stream)
                    for l in self.launches_opt_bucket[k]:
                        l(st)

    def _run_step_dp(self, head=None):
        """forward + loss + backward with each gradient bucket's all-reduce forked onto the communication stream as soon
        as the bucket is final, then the optimizer after all of them.  `head` replaces the first segment's graph (the
        variant that starts with the staging-slot copies)."""
        main = torch.cuda.current_stream(self.device)
        if self.seg_graphs is not None:
            for i, (g, k) in enumerate(self.seg_graphs):
                g = head if (i == 0 and head is not None) else g
                if g is not None:
                    g.replay()
                if k is not None:
                    self._fork_allreduce(k)
            main.wait_stream(self.comm_stream)
            if self.opt_graph is not None:
                self.opt_graph.replay()
            return
        for launches, k in self._segments():
            self._run_list(launches)
            if k is not None:
                self._fork_allreduce(k)
        if self.comm_stream is not None:
            main.wait_stream(self.comm_stream)
        self._run_list(self.launches_opt)

    def _capture(self):
        # warm up on a side stream, then capture forward+loss+backward (+ optimizer when single GPU)
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        saved = (self.flat_p.clone(), self.exp_avg.clone(), self.exp_avg_sq.clone(), self.step_dev.clone())
        with torch.cuda.stream(s):
            for _ in range(2):
                if self.overlap:
                    self._run_step_dp()
                else:
                    self._run_list(self.launches_fb)
                    if self.world == 1:
                        self._run_list(self.launches_opt)
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        self.flat_p.copy_(saved[0]); self.exp_avg.copy_(saved[1]); self.exp_avg_sq.copy_(saved[2])
        self.step_dev.copy_(saved[3])
        if self.overlap:
            # one graph per segment between bucket boundaries + one for the optimizer; NCCL stays outside the graphs
            graphs = []
            for launches, k in self._segments():
                g = None   # buckets that become final at the same point leave an empty segment: nothing to replay
                if launches:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, capture_error_mode="thread_local"):
                        self._run_list(launches)
                graphs.append((g, k))
            if self.bucket_adam:
                for k, ls in self.launches_opt_bucket.items():
                    bg = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(bg, capture_error_mode="thread_local"):
                        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
                        for l in ls:
                            l(st)
                    self._bucket_opt_graphs[k] = bg
            else:
                og = torch.cuda.CUDAGraph()
                with torch.cuda.graph(og, capture_error_mode="thread_local"):
                    self._run_list(self.launches_opt)
                self.opt_graph = og
            self.seg_graphs = graphs
            self.graph = graphs[0][0]
            self._heads[None] = self.graph
            return
        self.graph = self._capture_head(None)
        self._heads[None] = self.graph

    def _head_launches(self):
        if self.overlap:
            return self._segments()[0][0]
        return self.launches_fb + (self.launches_opt if self.world == 1 else [])

    def _capture_head(self, slot):
        """The first graph of a step (the whole step when single GPU).  slot = 0 / 1: the graph starts with the three
        device copies staging slot -> static inputs, so a prefetched host batch costs no eager launch between replays."""
        g = torch.cuda.CUDAGraph()
        kw = dict(capture_error_mode="thread_local") if self.overlap else {}
        snap = slot is not None and self._snap_in_graph
        if snap:
            self._ensure_readback()
        with torch.cuda.graph(g, **kw):
            if slot is not None:
                sx, sm, so = self._stage[slot]
                self.x.copy_(sx, non_blocking=True)
                self.mask.copy_(sm, non_blocking=True)
                self.onehot.copy_(so, non_blocking=True)
            self._run_list(self._head_launches())
            if snap:
                # the loss vector of this step, parked for `losses_to_host` (the objective is part of every head graph):
                # as an eager 16-byte copy between two replays it exposed the next graph's start-up
                self._loss_snap[slot].copy_(self.loss_out, non_blocking=True)
        return g

    def load_batch(self, x: torch.Tensor, mask: torch.Tensor, onehot: torch.Tensor):
        """Hand the next batch to the step.  Device tensors are copied into the static input buffers on the current
        stream.  Host (pinned) tensors are prefetched: the host->device copy runs on a copy stream into one of two
        staging slots, so the transfer of batch i+1 overlaps the kernels of step i; `step()` waits for the slot and
        moves it into the static buffers with three small device copies."""
        if x.is_cuda:
            self.x.copy_(x, non_blocking=True)
            self.mask.copy_(mask, non_blocking=True)
            self.onehot.copy_(onehot, non_blocking=True)
            self._pending = None
            return
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._stage = [(torch.empty_like(self.x), torch.empty_like(self.mask), torch.empty_like(self.onehot))
                           for _ in range(2)]
            self._stage_ready = [torch.cuda.Event(), torch.cuda.Event()]
            self._stage_free = [None, None]
        slot = self._load_turn & 1
        self._load_turn += 1
        cs = self._copy_stream
        if self._stage_free[slot] is not None:
            cs.wait_event(self._stage_free[slot])      # the step that consumed this slot has copied it out
        with torch.cuda.stream(cs):
            sx, sm, so = self._stage[slot]
            sx.copy_(x, non_blocking=True)
            sm.copy_(mask, non_blocking=True)
            so.copy_(onehot, non_blocking=True)
            self._stage_ready[slot].record(cs)
        self._pending = slot

    def _consume_pending(self, in_graph: bool = False):
        """Make the compute stream wait for the prefetched batch; copy it into the static inputs here (eager paths) or
        leave that to the slot's head graph (returns the slot; the caller records `_stage_free` after the replay)."""
        if self._pending is None:
            return None
        slot, self._pending = self._pending, None
        main = torch.cuda.current_stream(self.device)
        main.wait_event(self._stage_ready[slot])
        if in_graph:
            return slot
        sx, sm, so = self._stage[slot]
        self.x.copy_(sx, non_blocking=True)
        self.mask.copy_(sm, non_blocking=True)
        self.onehot.copy_(so, non_blocking=True)
        self._mark_stage_free(slot)
        return None

    def _mark_stage_free(self, slot):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._stage_free[slot] = ev

    def step(self):
        """One optimisation step on the currently loaded batch.  Asynchronous; losses stay on the device."""
        with torch.cuda.device(self.device):
            if self.use_graph:
                if self.graph is None:
                    self._consume_pending()      # first step: copy eagerly so the warm-up runs on the real batch
                    self._capture()
                slot = self._consume_pending(in_graph=True)
                head = self._heads.get(slot)
                if head is None:
                    head = self._heads[slot] = self._capture_head(slot)
                self._snap_slot = slot if (slot is not None and self._snap_in_graph) else None
                if self._snap_slot is not None and self._read_done[slot] is not None:
                    # the host read of two steps ago must have left the snapshot this replay overwrites
                    torch.cuda.current_stream(self.device).wait_event(self._read_done[slot])
                if self.overlap:
                    self._run_step_dp(head)
                else:
                    head.replay()
                    if self.world > 1:
                        self._allreduce_then_opt()
                if slot is not None:
                    self._mark_stage_free(slot)
                self.steps_done += 1
                return
            self._consume_pending()
            self._snap_slot = None
            if self.overlap:
                self._run_step_dp()
            else:
                self._run_list(self.launches_fb)
                if self.world == 1:
                    self._run_list(self.launches_opt)
                else:
                    self._allreduce_then_opt()
        self.steps_done += 1

    def _allreduce_then_opt(self):
        """MTBC_DP_OVERLAP=0: one all-reduce over the whole flat gradient buffer after the backward pass."""
        import torch.distributed as dist
        dist.all_reduce(self.plan.grad_flat, op=dist.ReduceOp.SUM, group=self.pg)
        self._run_list(self.launches_opt)

    def set_lr(self, lr: float):
        self.lr = float(lr)
        self.lr_dev.fill_(float(lr))
        if self._shared is not None:
            self._shared.lr = self.lr

    def losses(self) -> torch.Tensor:
        """Device tensor [total, seg, cls, nan_flag] of the last step (reading it synchronises)."""
        return self.loss_out

    def losses_to_host(self, out: torch.Tensor):
        """Asynchronous device->host read of [total, seg, cls, nan_flag] of the step just enqueued into the pinned
        tensor `out` (valid after `torch.cuda.synchronize()` or once the returned event has completed).  The copy runs
        on a read-back stream from a snapshot taken on the compute stream, so the next step's launches never queue
        behind a copy engine that is busy with the next batch's host->device transfer."""
        main = torch.cuda.current_stream(self.device)
        self._ensure_readback()
        if self._snap_slot is not None:
            slot = self._snap_slot      # the step's graph has already parked the losses in this slot's snapshot
        else:
            slot = self._read_turn & 1
            self._read_turn += 1
            if self._read_done[slot] is not None:
                main.wait_event(self._read_done[slot])       # the read of two steps ago has left this snapshot
            self._loss_snap[slot].copy_(self.loss_out, non_blocking=True)
        ready = torch.cuda.Event()
        ready.record(main)
        rs = self._read_stream
        rs.wait_event(ready)
        with torch.cuda.stream(rs):
            out.copy_(self._loss_snap[slot], non_blocking=True)
            done = torch.cuda.Event()
            done.record(rs)
        self._read_done[slot] = done
        return done

    def _ensure_readback(self):
        if self._read_stream is None:
            self._read_stream = torch.cuda.Stream(device=self.device)
            self._loss_snap = [torch.zeros_like(self.loss_out) for _ in range(2)]

    def forward_backward_only(self):
        """Forward + loss + backward without the optimizer (parity tests)."""
        with torch.cuda.device(self.device):
            self._consume_pending()
            self._run_list(self.launches_fb)

    def deliver_grads(self):
        self.model._deliver_grads(self.plan)
