"""GPU diagnostic: run the U-Net++ backward launch list one launch at a time and watch a debug buffer."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MTBC_DEBUG_UNIQUE_SCRATCH"] = "1"
import ctypes as C
import torch
from oracle import torch_oracle as O
from multi_task_breast_cancer_b200 import models as M, criterions as Cr, ops

torch.manual_seed(1993)
ref = O.MTUNetPlusPlus(deep_supervision=True).cuda()
new = M.MTUNetPlusPlus(deep_supervision=True).cuda()
new.load_state_dict(ref.state_dict())
img, mask, onehot, label = O.synthetic_batch(4, 64, 64, device="cuda")
if os.environ.get("RUN_ORACLE", "1") == "1":
    rl, ro = ref(img)
    seg_r, cls_r = O.multitask_criterion(O.DiceLoss(), mask, ro, O.FocalLoss(), onehot, rl, True)
    (0.35 * seg_r + 0.65 * cls_r).backward()
    del rl, ro, seg_r, cls_r
nl, no = new(img)
seg_n, cls_n = Cr.apply_criterion_multitask_segmentation_classification(
    Cr.DiceLoss(sigmoid=True, squared_pred=True, smooth_nr=1, smooth_dr=1), mask, no, Cr.FocalLoss(), onehot, nl, True)
(0.35 * seg_n + 0.65 * cls_n).backward()
torch.cuda.synchronize()
plan = next(iter(new._plans.values()))
watch = plan.debug[sys.argv[1] if len(sys.argv) > 1 else "upcat_0_4.convs.conv_1.dy"]
x04g = plan.tensors["upcat_0_4.convs.conv_1"].g
st = C.c_void_p(ops.stream_ptr())
print("after autograd backward: dy absmax/sample", [watch.t[n].float().abs().max().item() for n in range(4)])
for rep in range(3):
    plan.run_backward()
    torch.cuda.synchronize()
    print("after plain re-run", rep, ": dy absmax/sample", [watch.t[n].float().abs().max().item() for n in range(4)])
prev = None
for i, l in enumerate(plan.bwd):
    l(st)
    torch.cuda.synchronize()
    v = [watch.t[n].float().abs().max().item() for n in range(4)]
    g = [x04g.t[n].float().abs().max().item() for n in range(4)]
    if v != prev:
        print(i, l.kind, "dy absmax/sample", ["%.3g" % a for a in v], " x04.g absmax/sample", ["%.3g" % a for a in g])
        prev = v
