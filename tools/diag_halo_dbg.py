"""GPU tool: cycle breakdown of the MMA-issuing lane of CTA 0 for a few level-0 forward convs (MTBC_HALO_DBG=1)."""
import ctypes as C, os, sys
os.environ.setdefault("MTBC_HALO_DBG", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from multi_task_breast_cancer_b200 import _lib, ops
from multi_task_breast_cancer_b200.ops import Feat
lib = _lib.load()
lib.mtbc_debug_halo_times.argtypes = [C.c_void_p, C.c_int]
N, H, W = 32, 256, 256
for src_C, Cout in [([24], 24), ([24, 48], 24), ([24, 24, 24, 24, 48], 24), ([48], 48)]:
    h, w = (H, W) if Cout == 24 else (128, 128)
    srcs = [Feat.empty(N, h, w, c) for c in src_C]
    for f in srcs: f.t.normal_()
    out = Feat.empty(N, h, w, Cout)
    offs, ktot = ops.k_offsets(srcs)
    wf = torch.randn(9, out.Ck, ktot, device="cuda").to(torch.bfloat16)
    ssum = torch.zeros(N, out.Cp, device="cuda"); ssq = torch.zeros(N, out.Cp, device="cuda")
    op = ops.conv3x3_fwd_op(srcs, wf, out, stat_sum=ssum, stat_sq=ssq)
    for _ in range(2): op.launch()
    torch.cuda.synchronize()
    buf = (C.c_longlong * 8)()
    lib.mtbc_debug_halo_times(buf, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); op.launch(); e1.record(); torch.cuda.synchronize()
    lib.mtbc_debug_halo_times(buf, 1)
    acc, data, issue, commit, tiles, chunks, total = list(buf)[:7]
    print(f"{src_C}->{Cout} {h}x{w}: {e0.elapsed_time(e1)*1e3:.1f} us; CTA0 MMA lane: tiles {tiles} chunks {chunks} total {total} cyc "
          f"({total/max(tiles,1):.0f}/tile): wait-acc {acc/max(tiles,1):.0f}/tile, wait-data {data/max(chunks,1):.0f}/chunk, "
          f"issue {issue/max(chunks,1):.0f}/chunk, commit {commit/max(chunks,1):.0f}/chunk")
