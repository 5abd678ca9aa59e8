"""GPU tool (for ncu): one fused InstanceNorm backward at the level-0 shape.  python tools/ncu_inbwd.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from multi_task_breast_cancer_b200 import _lib
from multi_task_breast_cancer_b200.ops import Feat, ptr
N, H, W, C = 32, 256, 256, 24
y = Feat.empty(N, H, W, C); y.t.normal_()
g = Feat.empty(N, H, W, C); g.t.normal_()
dy = Feat.empty(N, H, W, C)
Cp = y.Cp
mean = torch.zeros(N, Cp, device="cuda"); rstd = torch.ones(N, Cp, device="cuda")
gam = torch.ones(Cp, device="cuda"); bet = torch.zeros(Cp, device="cuda")
dg = torch.zeros(C, device="cuda"); db = torch.zeros(C, device="cuda")
for _ in range(3):
    s1 = torch.zeros(N, Cp, device="cuda"); s2 = torch.zeros(N, Cp, device="cuda")
    cnt = torch.zeros(N, dtype=torch.int32, device="cuda")
    _lib.call("mtbc_in_bwd", ptr(g.t), ptr(y.t), N, H * W, Cp, ptr(mean), ptr(rstd), ptr(gam), ptr(bet), 0.1, ptr(s1), ptr(s2),
              ptr(dy.t), ptr(dg), ptr(db), C, ptr(cnt), None)
    torch.cuda.synchronize()
print("ok")
