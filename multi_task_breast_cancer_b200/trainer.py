"""Epoch-level drivers on either side of the hot path.

    init_optimizer / init_lr_scheduler   src/utils/experiment_init.py:175-196, 266-283
    FlatAdam                             torch.optim.Adam facade over TrainStep's flat fp32 buffers: `param_groups[0]['lr']`
                                         is what torch's own ReduceLROnPlateau / CosineAnnealingLR drive, and
                                         `state_dict()` / `load_state_dict()` speak torch.optim.Adam's format, i.e. the
                                         'optimizer_state_dict' of the reference checkpoints
    EpochRunner.train_one_epoch          src/training_multitask.py:74-116 without its 4 + 2B host syncs per step
    EpochRunner.validate_one_epoch       src/training_multitask.py:119-159
    save_checkpoint / load_pretrained_model   src/training_multitask.py:240-249, src/utils/models.py:19-36
    inference_multitask                  src/utils/models.py:273-397 batched: ONE forward per batch under no_grad instead
                                         of two per image, refinement + per-image confusion counts on the device

The per-step bookkeeping the reference does with `.item()` (loss sums, per-batch hard Dice, predicted / true class
lists) is two small launches (`mtbc_hard_dice_counts`, `mtbc_metrics_accumulate`) into device accumulators that are read
ONCE per epoch.  CUDA only: everything here drives the C ABI of libmtbc.so; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import logging
import math
import os
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch

from . import _lib
from .ops import ptr, stream_ptr

METRIC_KEYS = ["Haussdorf distance", "DICE", "Sensitivity", "Specificity", "Accuracy", "Jaccard index", "Precision"]


# ======================================================================================================================
# pure host helpers (no device needed)
# ======================================================================================================================
def classification_scores(confusion: Sequence[Sequence[int]]) -> Tuple[float, float]:
    """(accuracy, weighted F1) from a confusion matrix indexed [ground truth][prediction]: what
    sklearn.metrics.accuracy_score and f1_score(labels=[0, 1, 2], average='weighted') return for the label lists the
    reference collects (training_multitask.py:112-113); a class without predictions and truths scores F1 = 0."""
    K = len(confusion)
    total = sum(sum(int(v) for v in row) for row in confusion)
    if total == 0:
        return float("nan"), 0.0
    acc = sum(int(confusion[k][k]) for k in range(K)) / total
    f1w = 0.0
    for k in range(K):
        tp = int(confusion[k][k])
        support = sum(int(v) for v in confusion[k])
        pred = sum(int(confusion[g][k]) for g in range(K))
        f1 = 2.0 * tp / (support + pred) if (support + pred) > 0 else 0.0
        f1w += support * f1
    return acc, f1w / total


def hausdorff_from_row_distances(d2_seg_to_gt: int, d2_gt_to_seg: int, n_seg: int, n_gt: int) -> float:
    """haussdorf_distance (src/utils/metrics.py:236-252) from the integers of `mtbc_row_hausdorff`: the reference hands
    the (H, W) boolean images to scipy's directed_hausdorff, so rows are the points and distances are sqrt(Hamming);
    NaN when exactly one mask is empty, 0 when both are (the reference's `hd = 0` is then recomputed as 0 by its
    else branch)."""
    if (n_seg == 0) != (n_gt == 0):
        return float("nan")
    return math.sqrt(float(max(int(d2_seg_to_gt), int(d2_gt_to_seg))))


def segmentation_metrics_from_counts(tp: float, fp: float, fn: float, tn: float,
                                     hausdorff: float = float("nan")) -> Dict[str, float]:
    """calculate_metrics (src/utils/metrics.py:26-76) from the four cardinalities; the Hausdorff distance is passed in
    (hausdorff_from_row_distances) or NaN."""
    tp, fp, fn, tn = float(tp), float(fp), float(fn), float(tn)
    gt_empty, seg_empty = (tp + fn) == 0, (tp + fp) == 0
    nan = float("nan")
    return {
        "Haussdorf distance": hausdorff,
        "DICE": (1.0 if seg_empty else 0.0) if gt_empty else 2 * tp / (2 * tp + fp + fn),
        "Sensitivity": nan if tp == 0 else tp / (tp + fn),
        "Specificity": tn / (tn + fp) if (tn + fp) > 0 else nan,
        "Accuracy": (tp + tn) / (tp + tn + fp + fn),
        "Jaccard index": (1.0 if seg_empty else 0.0) if gt_empty else tp / (tp + fp + fn),
        "Precision": nan if tp == 0 else tp / (tp + fp),
    }


def adam_state_dict_from_flat(names: Sequence[str], ranges: Dict[str, Tuple[int, int]], shapes: Dict[str, torch.Size],
                              exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, step: int, group: Dict) -> Dict:
    """torch.optim.Adam.state_dict() layout ({'state': {i: {'step', 'exp_avg', 'exp_avg_sq'}}, 'param_groups': [...]})
    from the flat moment buffers; parameter i is the i-th entry of model.parameters().  Before the first step the
    state is empty, as in torch."""
    state = {}
    if step > 0:
        for i, n in enumerate(names):
            a, b = ranges[n]
            state[i] = {"step": torch.tensor(float(step)),
                        "exp_avg": exp_avg[a:b].detach().clone().view(shapes[n]),
                        "exp_avg_sq": exp_avg_sq[a:b].detach().clone().view(shapes[n])}
    g = dict(group)
    g["params"] = list(range(len(names)))
    return {"state": state, "param_groups": [g]}


def adam_state_dict_to_flat(sd: Dict, names: Sequence[str], ranges: Dict[str, Tuple[int, int]],
                            exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor) -> Tuple[int, Dict]:
    """Inverse of adam_state_dict_from_flat: fills the flat moment buffers in place and returns (step, param group).
    Accepts the reference's checkpoints (torch 2.0.1 Adam: 'step' is a 0-dim tensor)."""
    groups = sd["param_groups"]
    if len(groups) != 1 or len(groups[0]["params"]) != len(names):
        raise ValueError("optimizer state_dict does not describe one parameter group over this model's parameters")
    exp_avg.zero_()
    exp_avg_sq.zero_()
    steps = set()
    for i, n in enumerate(names):
        st = sd["state"].get(i, sd["state"].get(str(i)))
        if st is None:
            continue
        a, b = ranges[n]
        exp_avg[a:b].copy_(st["exp_avg"].reshape(-1))
        exp_avg_sq[a:b].copy_(st["exp_avg_sq"].reshape(-1))
        steps.add(int(float(st["step"])))
    if len(steps) > 1:
        raise ValueError(f"parameters were stepped a different number of times: {sorted(steps)}")
    group = {k: v for k, v in groups[0].items() if k != "params"}
    return (steps.pop() if steps else 0), group


# ======================================================================================================================
# optimizer + LR schedulers (f1)
# ======================================================================================================================
class FlatAdam(torch.optim.Optimizer):
    """torch.optim.Adam-shaped handle on a TrainStep (whose CUDA graph contains the fused Adam kernel).

    `step()` is the reference's `optimizer.step()`: with the fused path the update already happened inside
    `TrainStep.step()`, so it only keeps torch's scheduler bookkeeping consistent.  The learning rate lives in
    `param_groups[0]['lr']`; TrainStep pushes it to the device scalar the captured Adam launch reads whenever it
    changed, so `torch.optim.lr_scheduler.*` work unmodified."""

    def __init__(self, train_step):
        self.ts = train_step
        params = [p for p in train_step.model.parameters()]
        # the group keys of the installed torch's Adam (they differ between torch versions), with this step's values
        defaults = dict(torch.optim.Adam([torch.nn.Parameter(torch.zeros(1))], lr=float(train_step.lr),
                                         betas=tuple(train_step.betas), eps=train_step.eps).defaults)
        super().__init__(params, defaults)
        self._names = list(dict(train_step.model.named_parameters()).keys())
        self._shapes = {n: p.shape for n, p in train_step.model.named_parameters()}
        self._pushed_lr = self.param_groups[0]["lr"]

    def sync_lr(self):
        lr = float(self.param_groups[0]["lr"])
        if lr != self._pushed_lr:
            self.ts.set_lr(lr)
            self._pushed_lr = lr

    def step(self, closure=None):
        self.sync_lr()
        return None

    def zero_grad(self, set_to_none: bool = True):
        return None  # the plan's backward stores (first contribution) or accumulates in-kernel: nothing to clear

    def state_dict(self):
        ts = self.ts
        group = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        return adam_state_dict_from_flat(self._names, ts.param_ranges, self._shapes, ts.exp_avg.cpu(),
                                         ts.exp_avg_sq.cpu(), int(ts.step_dev.item()), group)

    def load_state_dict(self, sd):
        ts = self.ts
        ea, es = torch.zeros_like(ts.exp_avg, device="cpu"), torch.zeros_like(ts.exp_avg_sq, device="cpu")
        step, group = adam_state_dict_to_flat(sd, self._names, ts.param_ranges, ea, es)
        if tuple(group.get("betas", ts.betas)) != tuple(ts.betas) or float(group.get("eps", ts.eps)) != ts.eps:
            raise ValueError("betas / eps of the checkpoint differ from the ones captured in the training step")
        ts.exp_avg.copy_(ea)
        ts.exp_avg_sq.copy_(es)
        ts.step_dev.fill_(step)
        self.param_groups[0]["lr"] = float(group.get("lr", self.param_groups[0]["lr"]))
        self.sync_lr()


def init_optimizer(model: torch.nn.Module, optimizer: str, learning_rate: float = 0.001):
    """src/utils/experiment_init.py:175-196, for the module API (gradients land in param.grad, any torch optimizer
    works).  The fused path is `TrainStep(...)` + `FlatAdam(train_step)` and implements the 'Adam' branch."""
    if optimizer == "Adam":
        return torch.optim.Adam(model.parameters(), lr=learning_rate, eps=1e-4)
    if optimizer == "SGD":
        return torch.optim.SGD(model.parameters(), lr=learning_rate, momentum=0.9, nesterov=True)
    if optimizer == "AdamW":
        return torch.optim.AdamW(model.parameters(), lr=learning_rate)
    logging.info(f"The optimizer '{optimizer}' is not recognized. SGD will be used instead.")
    return torch.optim.SGD(model.parameters(), lr=0.001, momentum=0.9, nesterov=True)


def init_lr_scheduler(optimizer, scheduler: str = "cosine", t_max: int = 20, factor: float = 0.5, min_lr: float = 1e-6,
                      patience: int = 20):
    """src/utils/experiment_init.py:266-283 (`verbose=True` no longer exists in torch >= 2.7 and only printed)."""
    if scheduler == "plateau":
        return torch.optim.lr_scheduler.ReduceLROnPlateau(optimizer, mode="min", factor=factor, patience=patience,
                                                          min_lr=min_lr)
    if scheduler == "cosine":
        return torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, T_max=t_max, eta_min=min_lr)
    raise SystemExit("Select a scheduler allowed: ['plateau', 'cosine']")


# ======================================================================================================================
# epoch drivers (f1: training, f2: validation / test-time inference)
# ======================================================================================================================
def _batch_of(data, num_classes: int):
    """Reference loader protocol (dict with 'image', 'mask', 'label': BUSI_dataset.py:97-111) or a plain tuple
    (image, mask, one-hot); integer labels become float one-hot as at training_multitask.py:83-84."""
    if isinstance(data, dict):
        img, mask, label = data["image"], data["mask"], data["label"]
    else:
        img, mask, label = data[0], data[1], data[2]
    if label.dim() < 2 or label.shape[-1] != num_classes or label.dtype not in (torch.float32, torch.float64):
        label = torch.nn.functional.one_hot(label.flatten().to(torch.int64), num_classes=num_classes)
    return img.float(), mask.float(), label.to(torch.float32)


class _Accumulators:
    """Device-side epoch accumulators: acc[6] (double), confusion[K*K] and the tp/fp/fn scratch (int64)."""

    def __init__(self, K: int, device):
        self.K = K
        self.acc = torch.zeros(6, dtype=torch.float64, device=device)
        self.confusion = torch.zeros(K * K, dtype=torch.int64, device=device)
        self.counts = torch.zeros(3, dtype=torch.int64, device=device)

    def reset(self):
        self.acc.zero_(); self.confusion.zero_(); self.counts.zero_()

    def add_step(self, loss4: torch.Tensor, mask_logits: torch.Tensor, mask: torch.Tensor, class_logits: torch.Tensor,
                 onehot: torch.Tensor):
        st = C.c_void_p(stream_ptr())
        _lib.call("mtbc_hard_dice_counts", ptr(mask_logits), ptr(mask), mask_logits.numel(), ptr(self.counts), st)
        _lib.call("mtbc_metrics_accumulate", ptr(loss4), ptr(self.counts), ptr(class_logits), ptr(onehot),
                  class_logits.shape[0], self.K, ptr(self.acc), ptr(self.confusion), st)

    def read(self):
        """The one host sync of the epoch."""
        acc = self.acc.cpu().tolist()
        conf = self.confusion.cpu().view(self.K, self.K).tolist()
        return acc, conf


class EpochRunner:
    """train_one_epoch / validate_one_epoch of src/training_multitask.py on top of a TrainStep.

    The TrainStep is a static CUDA graph for one batch shape.  The reference's loaders have no `drop_last`
    (BUSI_dataloader.py:146-148) and `train_one_epoch` trains on the ragged last batch (training_multitask.py:79), so a
    batch with FEWER samples than planned gets a second plan + graph keyed on its size (`TrainStep(share_state_with=)`:
    same parameters, Adam moments and step counter) instead of an error; any other shape mismatch is rejected.  Data
    parallel runs must hand every rank full batches (data.shard_indices pads the tail by wrapping around)."""

    def __init__(self, train_step, num_classes: int = 3):
        self.ts = train_step
        self.K = int(num_classes)
        if self.K != train_step.K:
            raise ValueError(f"the model emits {train_step.K} class logits, not {num_classes}")
        self.dev = train_step.device
        self.optimizer = FlatAdam(train_step)
        self._acc = _Accumulators(self.K, self.dev)
        self._eval = {}     # batch size -> EvalStep
        self._tail = {}     # batch size -> TrainStep sharing the optimizer state (ragged last batch)

    # -------------------------------------------------------------------------------------------------- training
    def train_one_epoch(self, loader: Iterable) -> Tuple[float, float, float, float]:
        """-> (avg_training_loss, avg_training_dice, training_acc, training_f1), training_multitask.py:74-116."""
        ts = self.ts
        self.ts.model.train(True)
        self.optimizer.sync_lr()
        self._acc.reset()
        n = 0
        with torch.cuda.device(self.dev):
            for data in loader:
                img, mask, onehot = _batch_of(data, self.K)
                ts = self._step_for(img)
                ts.load_batch(img if img.is_cuda or img.is_pinned() else img.pin_memory(),
                              mask if mask.is_cuda or mask.is_pinned() else mask.pin_memory(),
                              onehot if onehot.is_cuda or onehot.is_pinned() else onehot.pin_memory())
                ts.step()
                self.optimizer.step()
                self._acc.add_step(ts.loss_out, ts.plan.outputs_seg[-1], ts.mask, ts.plan.outputs_cls[0], ts.onehot)
                n += 1
        acc, conf = self._acc.read()
        if acc[3] != 0.0:
            logging.info("NaN in model loss!!")           # criterions.py:72-76, checked once per epoch instead of per step
            raise SystemExit(1)
        if n == 0:
            raise ValueError("empty loader")
        accuracy, f1w = classification_scores(conf)
        return acc[0] / n, acc[4] / n, accuracy, f1w

    # -------------------------------------------------------------------------------------------------- validation
    def _eval_step(self, B: Optional[int] = None):
        B = self.ts.B if B is None else int(B)
        if B not in self._eval:
            self._eval[B] = EvalStep(self.ts.model, (B, self.ts.Cin, self.ts.H, self.ts.W), alpha=self.ts.alpha,
                                     inversely_weighted=self.ts.inv_w, focal_alpha=self.ts.focal_alpha,
                                     focal_gamma=self.ts.focal_gamma)
        return self._eval[B]

    def _step_for(self, img):
        """The TrainStep for this batch: the planned one, or a sibling for a ragged (smaller) last batch."""
        self._check_shape(img)
        B = int(img.shape[0])
        if B == self.ts.B:
            return self.ts
        if self.ts.world > 1:
            raise ValueError("data-parallel steps need full batches on every rank (equal step counts and unbiased "
                             "1/world averaging): shard the epoch with data.shard_indices")
        if B not in self._tail:
            t = self.ts
            self._tail[B] = type(t)(t.model, (B, t.Cin, t.H, t.W), lr=t.lr, betas=t.betas, eps=t.eps, alpha=t.alpha,
                                    inversely_weighted=t.inv_w, focal_alpha=t.focal_alpha, focal_gamma=t.focal_gamma,
                                    use_graph=t.use_graph, device=t.device, share_state_with=t)
        return self._tail[B]

    def validate_one_epoch(self, loader: Iterable) -> Tuple[float, float, float, float, float, float]:
        """-> (avg_val_loss, avg_val_dice, val_acc, val_f1, avg_seg_val_loss, avg_cls_val_loss),
        training_multitask.py:119-159 (forward + objective only, no gradients)."""
        self.ts.model.train(False)
        self._acc.reset()
        n = 0
        with torch.cuda.device(self.dev):
            for data in loader:
                img, mask, onehot = _batch_of(data, self.K)
                self._check_shape(img)
                ev = self._eval_step(img.shape[0])
                ev.run(img, mask, onehot)
                self._acc.add_step(ev.loss_out, ev.plan.outputs_seg[-1], ev.mask, ev.plan.outputs_cls[0], ev.onehot)
                n += 1
        acc, conf = self._acc.read()
        if n == 0:
            raise ValueError("empty loader")
        accuracy, f1w = classification_scores(conf)
        return acc[0] / n, acc[4] / n, accuracy, f1w, acc[1] / n, acc[2] / n

    def _check_shape(self, img):
        want = (self.ts.B, self.ts.Cin, self.ts.H, self.ts.W)
        if img.dim() != 4 or tuple(img.shape[1:]) != want[1:] or not 0 < img.shape[0] <= want[0]:
            raise ValueError(f"batch of shape {tuple(img.shape)} handed to a step planned for {want} (only the batch "
                             "dimension may be smaller: the loader's last batch)")


class EvalStep:
    """Forward + fused multi-task objective on a static forward-only plan, captured in a CUDA graph.  Shares the
    parameters with the training plan (weights are re-packed to bf16 at the head of every run)."""

    def __init__(self, model, batch_shape: Sequence[int], alpha: float = 0.35, inversely_weighted: bool = True,
                 focal_alpha: float = 1.0, focal_gamma: float = 2.0, use_graph: bool = True):
        from .plan import _mk
        self.model = model
        self.device = next(model.parameters()).device
        if self.device.type != "cuda":
            raise _lib.MtbcError("EvalStep needs a CUDA (sm_100a) device; there is no CPU fallback")
        B, Cin, H, W = (int(v) for v in batch_shape)
        self.B = B
        with torch.cuda.device(self.device):
            x = torch.zeros(B, Cin, H, W, dtype=torch.float32, device=self.device)
            self.plan = model._get_plan(x, False)
            plan = self.plan
            self.x = plan.x_in
            self.mask = torch.zeros(B, 1, H, W, dtype=torch.float32, device=self.device)
            K = plan.outputs_cls[0].shape[1]
            self.K = K
            self.onehot = torch.zeros(B, K, dtype=torch.float32, device=self.device)
            nh = len(plan.outputs_seg)
            self.dice_sums = torch.zeros(nh, B, 3, dtype=torch.float32, device=self.device)
            self.dice_loss = torch.zeros(nh, dtype=torch.float32, device=self.device)
            self.focal_loss = torch.zeros(1, dtype=torch.float32, device=self.device)
            self.loss_out = torch.zeros(4, dtype=torch.float32, device=self.device)
            L = list(plan.pack) + list(plan.fwd)
            L.append(_mk("mtbc_zero_bytes", ptr(self.dice_sums), self.dice_sums.numel() * 4))
            for i, logits in enumerate(plan.outputs_seg):
                j = nh - 1 - i
                L.append(_mk("mtbc_dice_sums", ptr(logits), ptr(self.mask), B, H * W, ptr(self.dice_sums[i])))
                L.append(_mk("mtbc_dice_finalize", ptr(self.dice_sums[i]), B, ptr(self.dice_loss[j:j + 1])))
            L.append(_mk("mtbc_focal_fwd", ptr(plan.outputs_cls[0]), ptr(self.onehot), B, K, C.c_float(focal_alpha),
                         C.c_float(focal_gamma), ptr(self.focal_loss)))
            L.append(_mk("mtbc_multitask_loss", ptr(self.dice_loss), nh, int(bool(inversely_weighted)),
                         ptr(self.focal_loss), C.c_float(alpha), ptr(self.loss_out)))
            self.launches = L
        self.use_graph = use_graph
        self.graph: Optional[torch.cuda.CUDAGraph] = None

    def _run_list(self):
        st = C.c_void_p(stream_ptr())
        for l in self.launches:
            l(st)

    def run(self, img: torch.Tensor, mask: Optional[torch.Tensor] = None, onehot: Optional[torch.Tensor] = None):
        """Copies the batch into the static inputs (host or device tensors) and runs forward (+ objective).  Outputs:
        `plan.outputs_cls`, `plan.outputs_seg` (fp32 logits), `loss_out` = [total, seg, cls, nan flag]."""
        with torch.cuda.device(self.device):
            self.x.copy_(img, non_blocking=True)
            if mask is not None:
                self.mask.copy_(mask, non_blocking=True)
            if onehot is not None:
                self.onehot.copy_(onehot, non_blocking=True)
            if not self.use_graph:
                self._run_list()
                return
            if self.graph is None:
                s = torch.cuda.Stream(device=self.device)
                s.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(s):
                    self._run_list()
                torch.cuda.current_stream(self.device).wait_stream(s)
                torch.cuda.synchronize(self.device)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._run_list()
                self.graph = g
            self.graph.replay()


# ======================================================================================================================
# checkpoints (f2)
# ======================================================================================================================
def save_checkpoint(path: str, epoch: int, model: torch.nn.Module, optimizer, val_loss: float):
    """The dictionary the reference writes at src/training_multitask.py:240-246 (same keys, the literal 'scheduler'
    string included), tensors on the host so that the file loads anywhere."""
    torch.save({
        "epoch": epoch,
        "model_state_dict": {k: v.detach().cpu().clone() for k, v in model.state_dict().items()},
        "optimizer_state_dict": optimizer.state_dict(),
        "scheduler": "scheduler",
        "val_loss": val_loss,
    }, path)


def load_pretrained_model(model: torch.nn.Module, ckpt_path: str):
    """src/utils/models.py:19-36.  Works on a model whose parameters a TrainStep has re-homed into its flat buffer
    (load_state_dict copies in place, so the step sees the restored weights)."""
    if os.path.isfile(ckpt_path):
        checkpoint = torch.load(ckpt_path, map_location="cpu", weights_only=False)
        model.load_state_dict(checkpoint["model_state_dict"])
        logging.info(f"Loaded checkpoint '{ckpt_path}'. Last epoch: {checkpoint['epoch']}")
    else:
        raise ValueError(f"\n\t-> No checkpoint found at '{ckpt_path}'")
    return model


# ======================================================================================================================
# test-time inference with prediction refinement (f2)
# ======================================================================================================================
@torch.no_grad()
def inference_multitask(model: torch.nn.Module, test_loader: Iterable, device=None, threshold: int = 0,
                        overlap_seg_based_on_class: bool = False, overlap_class_based_on_seg: bool = False,
                        num_classes: int = 3, normal_id: int = 2, keep: Optional[Dict[str, List]] = None,
                        hausdorff: bool = True):
    """Batched inference_multitask_multiclass_classification_segmentation (src/utils/models.py:273-397): one forward per
    batch, refinement (`mtbc_refine_predictions`), per-image tp/fp/fn/tn (`mtbc_confusion_counts`) and the reference's
    row-wise Hausdorff distance (`mtbc_row_hausdorff`; `hausdorff=False` reports NaN instead) on the device, one
    device->host read at the end.  Returns (segmentation_rows, classification_rows): lists of dicts with the
    columns of results_segmentation.csv and results_classification.csv.  File output (PNG
    masks, CSVs) stays with the caller: pass a dict as `keep` to receive, per batch, the refined uint8 masks
    (`keep['masks']`), the full-decoder mask logits and the averaged class logits (device tensors)."""
    from .criterions import refine_predictions
    dev = torch.device(device) if device is not None else next(model.parameters()).device
    if dev.type != "cuda":
        raise _lib.MtbcError("inference runs on CUDA sm_100a only (no CPU fallback)")
    model.train(False)
    per_batch = []
    for data in test_loader:
        if isinstance(data, dict):
            img, mask, label = data["image"], data["mask"], data["label"]
            ids = data.get("patient_id")
        else:
            img, mask, label = data[0], data[1], data[2]
            ids = data[3] if len(data) > 3 else None
        img, mask = img.to(dev, torch.float32), mask.to(dev, torch.float32).contiguous()
        B = img.shape[0]
        with torch.cuda.device(dev):
            logits, outs = model(img)
            seg_logits = outs[-1] if isinstance(outs, list) else outs
            cls_logits = torch.mean(torch.stack(logits, dim=0), dim=0) if isinstance(logits, list) else logits
            # models.py:316-332: the segmentation table applies only the class->mask overlap; :366-386: the
            # classification table applies only the mask->class overlap; both read the initial predictions
            rmask, _, _ = refine_predictions(seg_logits, cls_logits, normal_id, overlap_seg_based_on_class, False, threshold)
            _, rcls, cnt = refine_predictions(seg_logits, cls_logits, normal_id, False, overlap_class_based_on_seg, threshold)
            counts = torch.zeros(B, 4, dtype=torch.int64, device=dev)
            _lib.call("mtbc_confusion_counts", ptr(rmask), ptr(mask), B, mask[0].numel(), ptr(counts),
                      C.c_void_p(stream_ptr()))
            hd = None
            if hausdorff:
                hd = torch.zeros(B, 4, dtype=torch.int32, device=dev)
                _lib.call("mtbc_row_hausdorff", ptr(rmask), ptr(mask), B, mask.shape[-2], mask.shape[-1], ptr(hd),
                          C.c_void_p(stream_ptr()))
        if label.dim() >= 2 and label.shape[-1] == num_classes:
            gt = label.argmax(-1).flatten()
        else:
            gt = label.flatten().to(torch.int64)
        per_batch.append((ids, counts, rcls, cls_logits.float(), gt, hd))
        if keep is not None:
            keep.setdefault("masks", []).append(rmask)
            keep.setdefault("mask_logits", []).append(seg_logits)
            keep.setdefault("class_logits", []).append(cls_logits)
    seg_rows, cls_rows = [], []
    k = 0
    for ids, counts, rcls, cl, gt, hd in per_batch:
        counts, rcls, cl, gt = counts.cpu().tolist(), rcls.cpu().tolist(), cl.cpu().tolist(), gt.cpu().tolist()
        hd = hd.cpu().tolist() if hd is not None else None
        for b in range(len(rcls)):
            pid = (ids[b].item() if torch.is_tensor(ids) else ids[b]) if ids is not None else k
            row = {"patient_id": pid}
            row.update(segmentation_metrics_from_counts(
                *counts[b], hausdorff=hausdorff_from_row_distances(*hd[b]) if hd is not None else math.nan))
            row["class"] = int(gt[b])
            seg_rows.append(row)
            cls_rows.append({"patient_id": pid, "ground_truth": int(gt[b]), "predicted_label": int(rcls[b]),
                             "prob_benign": cl[b][0], "prob_malignant": cl[b][1],
                             "prob_normal": cl[b][2] if len(cl[b]) > 2 else math.nan})
            k += 1
    return seg_rows, cls_rows
