"""Pin the oracle's scalar segmentation metrics and whole-batch hard Dice against the UNMODIFIED reference functions
(src/utils/metrics.py: calculate_metrics, dice_score_from_tensor) on seeded random masks, edge cases included (empty
ground truth, empty prediction, both empty, full overlap).  Run in the build container only:

    python tests/golden/make_metrics_golden.py      -> tests/golden/metrics.json

The fixture holds the inputs as packed bits and the reference's outputs (the Hausdorff distance under "hausdorff":
the reference's haussdorf_distance through scipy, None for NaN); tests/test_oracle_golden.py re-checks oracle and host
helpers against it anywhere."""
import json
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from src.utils.metrics import calculate_metrics, dice_score_from_tensor  # noqa: E402  (the reference itself)
from oracle import torch_oracle as O  # noqa: E402

rng = np.random.default_rng(1993)
H = W = 24
cases = []
for i in range(24):
    gt = (rng.random((1, 1, H, W)) < rng.choice([0.0, 0.05, 0.3, 0.7])).astype(np.float32)
    seg = (rng.random((1, 1, H, W)) < rng.choice([0.0, 0.05, 0.3, 0.7])).astype(np.float32)
    if i == 0:
        gt[:] = 0; seg[:] = 0
    if i == 1:
        seg = gt.copy(); gt[0, 0, 0, 0] = 1; seg[0, 0, 0, 0] = 1
    if gt.sum() + (1 - seg).sum() == 0 or (1 - gt).sum() == 0:
        continue  # specificity 0/0 raises a numpy warning in the reference; not a case the loop meets
    ref = calculate_metrics(gt, seg, patient=i)
    hdist = ref["Haussdorf distance"]
    hdist = None if math.isnan(hdist) else float(hdist)
    o = O.hausdorff_rows(gt, seg)
    assert (hdist is None and math.isnan(o)) or (hdist is not None and o == hdist), (i, hdist, o)   # bit-exact
    ref = {k: (None if isinstance(v, float) and math.isnan(v) else float(v)) for k, v in ref.items()
           if k not in ("patient_id", "Haussdorf distance")}
    ora = O.segmentation_metrics(gt, seg)
    for k, v in ref.items():
        o = ora[k]
        assert (v is None and math.isnan(o)) or (v is not None and abs(float(o) - v) < 1e-12), (i, k, v, o)
    hd = float(dice_score_from_tensor(torch.from_numpy(gt), torch.from_numpy(seg).bool()))
    logits = torch.from_numpy(seg) * 2 - 1
    assert abs(O.hard_dice(torch.from_numpy(gt), logits) - hd) < 1e-12
    cases.append({"gt": np.packbits(gt.astype(np.uint8)).tolist(), "seg": np.packbits(seg.astype(np.uint8)).tolist(),
                  "metrics": ref, "hard_dice": hd, "hausdorff": hdist})
json.dump({"H": H, "W": W, "cases": cases}, open(os.path.join(HERE, "metrics.json"), "w"))
print(f"wrote {len(cases)} cases")
