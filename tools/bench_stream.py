"""GPU tool: achieved HBM GB/s of the streaming (InstanceNorm / channel-sum) kernels at the U-Net++ level shapes, with the
bulk-copy pipelined kernels (stream_pipe.cu) and with the register-file kernels (MTBC_NO_PIPE=1), next to a plain
device copy of the same bytes.

    python tools/bench_stream.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from multi_task_breast_cancer_b200 import _lib, ops
from multi_task_breast_cancer_b200.ops import Feat, ptr

dev = "cuda"
SHAPES = [(32, 256, 256, 24), (32, 128, 128, 48), (32, 64, 64, 96), (32, 32, 32, 192), (32, 16, 16, 384)]
if len(sys.argv) > 1:
    SHAPES = SHAPES[:int(sys.argv[1])]


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    # rotate over enough distinct buffers? the tensors here are >= L2 at level 0/1; deeper levels are L2 resident
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


def cases_for(y, g, a, N, H, W, C):
    dy = Feat.empty(N, H, W, C)
    Cp = y.Cp
    ssum = torch.zeros(N, Cp, device=dev); ssq = torch.ones(N, Cp, device=dev) * H * W
    mean = torch.zeros(N, Cp, device=dev); rstd = torch.ones(N, Cp, device=dev)
    s1 = torch.zeros(N, Cp, device=dev); s2 = torch.zeros(N, Cp, device=dev)
    gam = torch.ones(Cp, device=dev); bet = torch.zeros(Cp, device=dev)
    dg = torch.zeros(C, device=dev); db = torch.zeros(C, device=dev)
    out = torch.zeros(C, device=dev)
    cnt = torch.zeros(N, dtype=torch.int32, device=dev)
    return {
        "in_apply (4 B/el)": (4, lambda: _lib.call("mtbc_in_apply", ptr(y.t), N, H, W, Cp, ptr(ssum), ptr(ssq), ptr(gam), ptr(bet), C,
                                                  1e-5, 0.1, ptr(a.t), None, ptr(mean), ptr(rstd), None)),
        "in_bwd_reduce (4 B/el)": (4, lambda: _lib.call("mtbc_in_bwd_reduce", ptr(g.t), ptr(y.t), N, H * W, Cp, ptr(mean), ptr(rstd),
                                                       ptr(gam), ptr(bet), 0.1, ptr(s1), ptr(s2), None)),
        "in_bwd_apply (6 B/el)": (6, lambda: _lib.call("mtbc_in_bwd_apply", ptr(g.t), ptr(y.t), N, H * W, Cp, ptr(mean), ptr(rstd),
                                                      ptr(gam), ptr(bet), 0.1, ptr(s1), ptr(s2), ptr(dy.t), ptr(dg), ptr(db), C, None)),
        "in_bwd (reduce+apply, 10 B/el unfused)": (10, lambda: (cnt.zero_(), _lib.call("mtbc_in_bwd", ptr(g.t), ptr(y.t), N, H * W, Cp, ptr(mean), ptr(rstd),
                                                      ptr(gam), ptr(bet), 0.1, ptr(s1), ptr(s2), ptr(dy.t), ptr(dg), ptr(db), C, ptr(cnt), None))),
        "channel_sum (2 B/el)": (2, lambda: _lib.call("mtbc_channel_sum", ptr(g.t), N * H * W, Cp, C, ptr(out), 1, None)),
    }


for (N, H, W, C) in ([] if os.environ.get("SWEEP") else SHAPES):
    y = Feat.empty(N, H, W, C); y.t.normal_()
    g = Feat.empty(N, H, W, C); g.t.normal_()
    a = Feat.empty(N, H, W, C)
    Cp = y.Cp
    elems = N * H * W * Cp
    cases = cases_for(y, g, a, N, H, W, C)
    t_copy = timeit(lambda: a.t.copy_(y.t))
    print(f"--- {N}x{H}x{W}x{C} (Cp {Cp}, {elems * 2 / 1e6:.0f} MB per tensor); torch copy: {t_copy:.1f} us = {elems * 4 / t_copy / 1e3:.0f} GB/s")
    for name, (bpe, fn) in cases.items():
        res = []
        for mode in ("0", "1"):
            os.environ["MTBC_NO_PIPE"] = mode
            t = timeit(fn)
            res.append((t, elems * bpe / t / 1e3))
        os.environ["MTBC_NO_PIPE"] = "0"
        print(f"  {name:26s} pipe {res[0][0]:8.1f} us {res[0][1]:6.0f} GB/s | regs {res[1][0]:8.1f} us {res[1][1]:6.0f} GB/s")

if os.environ.get("SWEEP"):
    N, H, W, C = 32, 256, 256, 24
    y = Feat.empty(N, H, W, C); y.t.normal_()
    g = Feat.empty(N, H, W, C); g.t.normal_()
    a = Feat.empty(N, H, W, C)
    Cp = y.Cp; elems = N * H * W * Cp
    t = timeit(lambda: y.t.sum())
    print(f"torch bf16 sum (read only): {t:.1f} us = {elems * 2 / t / 1e3:.0f} GB/s")
    t = timeit(lambda: torch.add(y.t, g.t, out=a.t))
    print(f"torch add (2 reads 1 write): {t:.1f} us = {elems * 6 / t / 1e3:.0f} GB/s")
    for stages in (2, 3, 4, 6):
        for vpt in (4,):
            for ctas in (1, 2, 3, 4):
                os.environ.update(MTBC_PIPE_STAGES=str(stages), MTBC_PIPE_VPT=str(vpt), MTBC_PIPE_CTAS=str(ctas), MTBC_NO_PIPE="0")
                line = f"stages {stages} vpt {vpt} ctas {ctas}:"
                for name, (bpe, fn) in cases_for(y, g, a, N, H, W, C).items():
                    try:
                        t = timeit(fn, 10)
                        line += f"  {name.split()[0]} {elems * bpe / t / 1e3:5.0f}"
                    except Exception as e:
                        line += f"  {name.split()[0]} ERR"
                print(line, flush=True)
