"""GPU tool: the single-source 3x3 data gradient with and without the fused InstanceNorm-backward statistics epilogue,
next to the forward conv of the same shape (same statistics epilogue without the backward extras), under the debug
switches of conv_halo.cu (read at op creation): time per launch and the MMA lane's cycle breakdown of CTA 0."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from multi_task_breast_cancer_b200 import _lib, ops
from multi_task_breast_cancer_b200.ops import Feat
lib = _lib.load()
lib.mtbc_debug_halo_times.argtypes = [C.c_void_p, C.c_int]


def timed(op, reps=10):
    for _ in range(3):
        op.launch()
    torch.cuda.synchronize()
    buf = (C.c_longlong * 8)()
    lib.mtbc_debug_halo_times(buf, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        op.launch()
    e1.record()
    torch.cuda.synchronize()
    lib.mtbc_debug_halo_times(buf, 1)
    acc, data, issue, commit, tiles, chunks, total = list(buf)[:7]
    t = max(tiles, 1)
    return e0.elapsed_time(e1) * 1e3 / reps, f"cyc/tile {total / t:.0f}: wait-acc {acc / t:.0f} wait-data {data / t:.0f} issue {issue / t:.0f} commit {commit / t:.0f}"


def case(N, H, W, Cc):
    y = Feat.empty(N, H, W, Cc); y.t.normal_()
    dy = Feat.empty(N, H, W, Cc); dy.t.normal_()
    ga = Feat.empty(N, H, W, Cc)
    wd = (torch.randn(9, ga.Ck, dy.Ck, device="cuda") * 0.1).to(torch.bfloat16)
    mean = torch.zeros(N, y.Cp, device="cuda"); rstd = torch.ones(N, y.Cp, device="cuda")
    s1 = torch.zeros(N, y.Cp, device="cuda"); s2 = torch.zeros(N, y.Cp, device="cuda")
    envs = [("plain", {}, "plain"), ("plain 1cta", {"MTBC_HALO_CTAS": "1"}, "plain"),
            ("fwd-stats", {}, "fwd"), ("fwd-stats epi2", {"MTBC_HALO_EPI": "2"}, "fwd"),
            ("fused", {}, "fused"), ("fused epi2", {"MTBC_HALO_EPI": "2"}, "fused")]
    for name, env, kind in envs:
        env = dict(env)
        os.environ["MTBC_HALO_DBG"] = str(1 | int(env.pop("D", 0)))
        for k, v in env.items():
            os.environ[k] = v
        if kind == "plain":
            op = ops.conv3x3_dgrad_op(dy, wd, ga, False)
        elif kind == "fwd":
            op = ops.conv3x3_fwd_op([dy], wd, ga, stat_sum=s1, stat_sq=s2)
        else:
            op = ops.conv3x3_dgrad_op(dy, wd, ga, False, bwd_fuse=(y, mean, rstd, None, None, 0.1), s1=s1, s2=s2)
        for k in env:
            os.environ.pop(k, None)
        us, brk = timed(op)
        print(f"{N}x{H}x{W}x{Cc} {name:22s} {us:7.1f} us   {brk}", flush=True)


case(32, 256, 256, 24)
case(32, 128, 128, 48)
