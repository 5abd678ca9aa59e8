"""GPU tool: BASELINE.json configs[3] -- inference-only forward + prediction refinement (utils/models.py:316-332,366-386),
batch 256 at 1x256x256: throughput and latency through the drop-in module API under torch.inference_mode().

    python tools/bench_infer.py [arch] [batch] [size]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import torch_oracle as O   # synthetic batch generator only
from multi_task_breast_cancer_b200 import models as M
from multi_task_breast_cancer_b200.criterions import refine_predictions

arch = sys.argv[1] if len(sys.argv) > 1 else "unetpp"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
S = int(sys.argv[3]) if len(sys.argv) > 3 else 256
torch.manual_seed(1993)
model = {"unetpp": lambda: M.MTUNetPlusPlus(deep_supervision=True), "nnunet": lambda: M.MTnnUNet(1, 1, 3),
         "bts": lambda: M.Multi_BTS_UNet(1, 1, 3, 32, True)}[arch]().cuda().eval()
res = {}
for b in (B, 1):
    img, _, _, _ = O.synthetic_batch(b, S, S, device="cuda")
    h_img = img.cpu().pin_memory()
    d_img = torch.empty_like(img)
    with torch.inference_mode():
        def run():
            d_img.copy_(h_img, non_blocking=True)                 # H2D of the batch
            logits, outs = model(d_img)
            mask, cls, cnt = refine_predictions(outs[-1], logits[0])
            return mask, cls
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10 if b > 1 else 50
        e0.record()
        for _ in range(reps):
            mask, cls = run()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    res[f"batch_{b}"] = {"ms_per_batch": round(ms, 3), "img_per_s": round(b / ms * 1e3, 1)}
print(json.dumps({"workload": f"{arch} inference forward + prediction refinement, 1x{S}x{S}, H2D of the batch included",
                  **res}))
