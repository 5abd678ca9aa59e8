#!/usr/bin/env python
"""Benchmark of the hot path: one training step (forward + Dice/focal loss + backward + Adam) of the multi-task
U-Net++ on synthetic 1x256x256 batches (BASELINE.json configs[1]), in images/s.

    python bench.py --gpus N --steps K --warmup W                 our arm (one process per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K --warmup W  reference arm: the reference's CPU training loop

Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for the definition of every key.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train_images_per_s"
UNIT = "img/s"
MODEL_NAMES = {"unetpp": "MTUNetPlusPlus(deep_supervision=True)", "nnunet": "MTnnUNet", "bts": "Multi_BTS_UNet(width=32, deep_supervision=True)"}
# GFLOP per image of one training step (true channels, fwd + dgrad + wgrad; SURVEY section 8d / BASELINE.md section 3)
TRAIN_GFLOP = {("unetpp", 256): 128.29, ("unetpp", 512): 513.16, ("nnunet", 256): 68.90, ("nnunet", 128): 17.23, ("bts", 128): 15.29}


def config_of(arch: str, batch: int, size: int, world: int) -> dict:
    """`config` of the JSON line -- the SAME dictionary for both arms (the reference arm times the reference's CPU
    loop 'on our arm's config'; what it actually sampled is in its cpu_baseline.sample)."""
    return {"workload": workload(arch, batch, size), "arch": arch, "batch_per_gpu": batch, "global_batch": world * batch,
            "size": size, "parallelism": f"dp{world}", "l2": "per-step working set (GBs of activations) >> 126 MB L2",
            "cuda_graph": True}


def workload(arch: str, batch: int, size: int) -> str:
    """The default (unetpp, 32, 256) is BASELINE.json configs[1]; --arch nnunet is configs[2], --size 512 configs[4]."""
    return (f"{MODEL_NAMES[arch]} training step, batch {batch}/GPU, 1x{size}x{size}, 3 classes, Dice+focal, "
            f"Adam(eps=1e-4), + prediction refinement")




def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="per-GPU batch")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--arch", default="unetpp", choices=["unetpp", "nnunet", "bts"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true")
    ap.add_argument("--cpu-batch", type=int, default=0, help="batch of the CPU arm (0: the workload's own batch at "
                    "<= 256x256, 8 at 512x512 where one CPU step of 32 images takes ~25 s)")
    ap.add_argument("--cpu-steps", type=int, default=2)
    a = ap.parse_args()
    if a.cpu_batch <= 0:
        a.cpu_batch = a.batch if a.size <= 256 else min(a.batch, 8)
    return a


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_models(arch):
    from oracle import torch_oracle as O  # used ONLY for synthetic data + the CPU baseline / reference arm
    return O


def cpu_reference_loop(arch, B, size, steps, warmup):
    """The reference's own training loop body (training_multitask.py:87-103) on the host cores, fp32, all threads.
    /root/reference cannot travel to the GPU box, so this is the oracle restatement ("port"), which is pinned
    bit-exactly against the reference modules by tests/golden/make_golden.py."""
    import torch
    from oracle import torch_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(1993)
    model = {"unetpp": lambda: O.MTUNetPlusPlus(in_channels=1, out_channels=1, n_classes=3, deep_supervision=True),
             "nnunet": lambda: O.MTnnUNet(1, 1, 3), "bts": lambda: O.Multi_BTS_UNet(1, 1, 3, 32, True)}[arch]()
    opt = O.make_optimizer(model, 1e-4)
    img, mask, onehot, _ = O.synthetic_batch(B, size, size)
    for _ in range(warmup):
        O.train_step(model, opt, img, mask, onehot)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        O.train_step(model, opt, img, mask, onehot)
        times.append(time.perf_counter() - t0)
    return {"value": B / statistics.median(times), "best": B / min(times), "ms_per_step": 1e3 * statistics.median(times),
            "cores": cores, "batch": B, "steps": steps}


def library_baseline(arch, B, size, steps, warmup, dev):
    """'The Blackwell kernel to beat' (SURVEY 2.2 / 8d): the reference's modules (oracle restatement, pinned bit-exactly
    against /root/reference) under plain torch eager on the SAME GPU, same batch, same step (forward + Dice/focal +
    backward + Adam(eps 1e-4)), i.e. cuDNN / cuBLAS / ATen kernels -- once in fp32 with TF32 tensor cores allowed, once
    under bf16 autocast with channels_last.  cudnn.benchmark is on (the library gets to pick its fastest algorithm).
    Checker leg: runs after, and outside, every timed region of our arm; nothing of it is on the product path."""
    import torch
    from oracle import torch_oracle as O
    out = {"what": "reference modules under torch eager on this GPU (cuDNN/cuBLAS/ATen), same step and batch",
           "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(), "batch": B, "steps": steps}
    mk = {"unetpp": lambda: O.MTUNetPlusPlus(in_channels=1, out_channels=1, n_classes=3, deep_supervision=True),
          "nnunet": lambda: O.MTnnUNet(1, 1, 3), "bts": lambda: O.Multi_BTS_UNet(1, 1, 3, 32, True)}[arch]
    img, mask, onehot, _ = O.synthetic_batch(B, size, size, device=dev)
    dice, focal = O.DiceLoss(), O.FocalLoss(alpha=1, gamma=2)
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    try:
        for key, autocast in (("tf32", False), ("bf16_autocast_channels_last", True)):
            torch.manual_seed(1993)
            torch.backends.cudnn.allow_tf32 = True
            torch.backends.cuda.matmul.allow_tf32 = True
            torch.backends.cudnn.benchmark = True
            model = mk().to(dev)
            x = img
            if autocast:
                model = model.to(memory_format=torch.channels_last)
                x = img.contiguous(memory_format=torch.channels_last)
            opt = O.make_optimizer(model, 1e-4)

            def step():
                opt.zero_grad(set_to_none=True)
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                    logits, outs = model(x)
                    seg, cls = O.multitask_criterion(dice, mask, outs, focal, onehot, logits, True)
                    total = 0.35 * seg + 0.65 * cls
                total.backward()
                opt.step()
                return total
            for _ in range(max(3, warmup)):
                step()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / steps
            out[key + "_img_s"] = B / (ms * 1e-3)
            out[key + "_ms_per_step"] = ms
            del model, opt
            torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001 - a baseline that cannot run is reported, it never fails our arm
        out["error"] = f"{type(e).__name__}: {e}"[:300]
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = saved
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    B = args.cpu_batch
    r = cpu_reference_loop(args.arch, B, args.size, max(1, min(args.steps, 3)), max(1, min(args.warmup, 1)))
    sample = f"{r['steps']} timed steps of batch {B} (the workload's batch is {args.batch}/GPU) after 1 warm-up, median"
    out = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
           "steps": r["steps"], "warmup": 1, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": config_of(args.arch, args.batch, args.size, args.gpus),
           "note": f"reference CPU training loop (oracle port of the reference modules, fp32, all host threads), batch {B} per step",
           "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": sample},
           "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(out)
    return 0


def per_kernel_times(ts, reps=3):
    """CUDA-event time of every launch of one step (eager, on the launching stream), aggregated by kernel kind."""
    import ctypes as C
    import torch
    from multi_task_breast_cancer_b200.ops import stream_ptr
    launches = [l for l in ts.launches_fb + (ts.launches_opt if ts.world == 1 else []) if l.kind != "bucket_ready"]
    st = C.c_void_p(stream_ptr())
    agg = {}
    for rep in range(reps):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(launches) + 1)]
        evs[0].record()
        for i, l in enumerate(launches):
            l(st)
            evs[i + 1].record()
        torch.cuda.synchronize()
        if rep == 0:
            continue  # warm
        for i, l in enumerate(launches):
            k = agg.setdefault(l.kind, {"ms": 0.0, "n": 0, "flops": 0.0})
            k["ms"] += evs[i].elapsed_time(evs[i + 1]) / (reps - 1)
            if rep == 1:
                k["n"] += 1
                k["flops"] += getattr(l, "true_flops", 0.0)
    return agg


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = None
    if world > 1 and os.environ.get("MTBC_NUMA_BIND", "1") != "0":
        # every rank stages 16.8 MB of pinned host memory per step: keep it on its own GPU's socket
        from multi_task_breast_cancer_b200.train import bind_to_gpu_numa
        try:   # NVML numbers the physical devices; CUDA_VISIBLE_DEVICES may remap the local rank
            gpu = int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local])
        except (KeyError, ValueError, IndexError):
            gpu = local
        numa = bind_to_gpu_numa(gpu)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    from oracle import torch_oracle as O  # synthetic data generator + CPU baseline only (never on the GPU path)
    from multi_task_breast_cancer_b200 import models as M
    from multi_task_breast_cancer_b200.train import TrainStep

    B, S = args.batch, args.size
    torch.manual_seed(1993)
    model = {"unetpp": lambda: M.MTUNetPlusPlus(in_channels=1, out_channels=1, n_classes=3, deep_supervision=True),
             "nnunet": lambda: M.MTnnUNet(1, 1, 3), "bts": lambda: M.Multi_BTS_UNet(1, 1, 3, 32, True)}[args.arch]().to(dev)
    # refine=True: the prediction-refining module runs inside the step's CUDA graph (2 kernels + a memset)
    ts = TrainStep(model, (B, 1, S, S), lr=1e-4, eps=1e-4, alpha=0.35, inversely_weighted=True, process_group=pg,
                   refine=True)
    img, mask, onehot, _ = O.synthetic_batch(B, S, S, seed=1993 + rank)
    h_img, h_mask, h_onehot = img.pin_memory(), mask.pin_memory(), onehot.pin_memory()
    h_loss = torch.zeros(4).pin_memory()
    ts.load_batch(h_img, h_mask, h_onehot)
    torch.cuda.synchronize()

    def one_step():
        ts.step()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        one_step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- device-resident timing (value)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        one_step()
    e1.record()
    barrier()
    ms_dev = e0.elapsed_time(e1) / args.steps
    # ---- end to end through the public API: pinned host batch -> H2D -> step -> D2H of the loss, every step
    for _ in range(2):
        ts.load_batch(h_img, h_mask, h_onehot); one_step(); ts.losses_to_host(h_loss)
    barrier()
    e0.record()
    for _ in range(args.steps):
        ts.load_batch(h_img, h_mask, h_onehot)
        one_step()
        done = ts.losses_to_host(h_loss)   # D2H of [total, seg, cls, nan] every step, on the read-back stream
    torch.cuda.current_stream().wait_event(done)   # the last read is inside the timed region too
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_dev, ms_e2e], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = t.tolist()

    out = None
    if rank == 0:
        peaks = measured_peaks()
        agg = per_kernel_times(ts)
        tc_kinds = [k for k in agg if agg[k]["flops"] > 0]
        top = max(tc_kinds, key=lambda k: agg[k]["ms"])
        tot_ms = sum(v["ms"] for v in agg.values())
        tc_ms = sum(agg[k]["ms"] for k in tc_kinds)
        tc_fl = sum(agg[k]["flops"] for k in tc_kinds)
        achieved = agg[top]["flops"] / (agg[top]["ms"] * 1e-3) / 1e12
        peak = peaks["bf16_tflops_sustained"]
        traffic, traffic_note = None, "profiles/traffic.json absent"
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            from multi_task_breast_cancer_b200 import build as _build
            tj = json.load(open(tp))
            t = tj.get("by_kind", {}).get(top)
            # the ncu pass is a separate run: only quote it when it profiled THIS build of the kernels
            if tj.get("build_digest") != _build.lib_digest():
                traffic_note = f"profiles/traffic.json ({tj.get('tag')}) was captured on another build: not quoted"
            elif t and agg[top]["n"]:
                traffic = t["dram_bytes_per_step"] / agg[top]["n"]  # mean DRAM bytes per launch (ncu, profiles/)
                traffic_note = f"ncu launch list {tj.get('tag')} of this build (digest {str(tj.get('build_digest'))[:12]})"
        roofline = {"bound": "tensor", "kernel": top, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_note,
                    "peak_source": peaks["source"] + " sustained",
                    "launches_per_step": agg[top]["n"], "kernel_ms_per_step": agg[top]["ms"],
                    "share_of_step": agg[top]["ms"] / tot_ms,
                    "all_tensor_kernels": {"achieved": tc_fl / (tc_ms * 1e-3) / 1e12, "frac": tc_fl / (tc_ms * 1e-3) / 1e12 / peak,
                                           "ms_per_step": tc_ms, "share_of_step": tc_ms / tot_ms},
                    "by_kernel_ms": {k: round(v["ms"], 4) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])[:12]}}
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            r = cpu_reference_loop(args.arch, args.cpu_batch, S, args.cpu_steps, 1)
            cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                   "sample": f"{r['steps']} timed steps of batch {r['batch']} of the same workload on the host CPU after 1 warm-up "
                             f"(median {r['ms_per_step']:.0f} ms/step)"}
        lib = None
        if not args.no_library_baseline and world == 1:
            lib = library_baseline(args.arch, B, S, 10, 3, dev)
            if lib.get("bf16_autocast_channels_last_img_s"):
                lib["ours_over_bf16"] = (B / (ms_dev * 1e-3)) / lib["bf16_autocast_channels_last_img_s"]
            if lib.get("tf32_img_s"):
                lib["ours_over_tf32"] = (B / (ms_dev * 1e-3)) / lib["tf32_img_s"]
        h2d = (h_img.numel() + h_mask.numel() + h_onehot.numel()) * 4
        n_launch = (ts.n_launches + 2) * args.steps   # + the two extra device launches behind the refinement entry point
        out = {"metric": METRIC, "value": world * B / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": max(args.warmup, 3), "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
               "config": config_of(args.arch, B, S, world),
               "tflops_true": world * B * TRAIN_GFLOP[(args.arch, S)] / ms_dev if (args.arch, S) in TRAIN_GFLOP else None,
               "e2e": {"value": world * B / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 16,
                       "ms_per_step": ms_e2e},
               "host_cores_bound": None if numa is None else len(numa),
               "gpu_launches": n_launch, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
               "library_baseline": lib}
        _emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


_RESULT = sys.stdout


def _emit(out: dict):
    """The ONE JSON line of the contract, on the process's original stdout."""
    _RESULT.write(json.dumps(out) + "\n")
    _RESULT.flush()


def main():
    global _RESULT
    args = parse()
    # Native libraries write to file descriptor 1 (NCCL prints its version banner there at communicator creation):
    # keep the original stdout for the result line only and send everything else to stderr.
    sys.stdout.flush()
    _RESULT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
