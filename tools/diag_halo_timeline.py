"""GPU tool: per-tile timeline of CTA 0 of the halo conv kernel (producer / MMA lanes / first and last epilogue warp),
tiles 24..39 of its range, in SM cycles.  Needs the HALO_TL instrumentation compiled into conv_halo.cu (a temporary
patch: `g_halo_tl`, `mtbc_debug_halo_timeline`, MTBC_HALO_DBG=128 -- see DESIGN 9 "per-tile timeline"; the shipped
kernel does not carry it).  Output of the round-2 run: profiles/r02z_halo_timeline_cta0.txt."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from multi_task_breast_cancer_b200 import _lib, ops
from multi_task_breast_cancer_b200.ops import Feat
lib = _lib.load()
lib.mtbc_debug_halo_timeline.argtypes = [C.c_void_p, C.c_int]
N, H, W, Cc = 32, 256, 256, 24
dy = Feat.empty(N, H, W, Cc); dy.t.normal_(); ga = Feat.empty(N, H, W, Cc)
wd = (torch.randn(9, ga.Ck, dy.Ck, device="cuda") * 0.1).to(torch.bfloat16)
s1 = torch.zeros(N, ga.Cp, device="cuda"); s2 = torch.zeros(N, ga.Cp, device="cuda")
names = ["prod.wait", "prod.tma", "mma.top", "mma.acc", "mma.data", "mma.issued", "mma.commit", "epi.top", "epi.full", "epi.rel", "epi.done", "epiL.full", "epiL.done"]
for label, env, kind in [("plain 1cta (P=2, lanes 2)", {"MTBC_HALO_CTAS": "1"}, "plain"), ("fwd-stats (P=4, lanes 2)", {}, "fwd")]:
    os.environ["MTBC_HALO_DBG"] = "128"
    for k, v in env.items(): os.environ[k] = v
    op = ops.conv3x3_dgrad_op(dy, wd, ga, False) if kind == "plain" else ops.conv3x3_fwd_op([dy], wd, ga, stat_sum=s1, stat_sq=s2)
    for k in env: os.environ.pop(k, None)
    for _ in range(3): op.launch()
    torch.cuda.synchronize()
    buf = (C.c_longlong * 512)()
    lib.mtbc_debug_halo_timeline(buf, 512)
    tl = [[buf[i * 32 + e] for e in range(13)] for i in range(16)]
    base = min(v for row in tl for v in row if v > 0)
    print("==", label)
    print("tile " + " ".join(f"{n:>10s}" for n in names))
    for i, row in enumerate(tl):
        print(f"{24 + i:4d} " + " ".join(f"{(v - base) if v > 0 else -1:10d}" for v in row))
