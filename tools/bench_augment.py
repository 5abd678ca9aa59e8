"""Achieved HBM bandwidth of the input-pipeline launch (mtbc_augment_batch): B=32 (and 256) x 256x256, all samples
flipped + rotated.  Algorithmic bytes per pixel: 2 B gathered (uint8 image + mask) + 8 B written (fp32 image + mask)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_task_breast_cancer_b200.data import DeviceBUSI, draw_transform_params  # noqa: E402

n, H, W = 780, 256, 256
g = torch.Generator().manual_seed(0)
ds = DeviceBUSI(torch.randint(0, 256, (n, H, W), generator=g, dtype=torch.uint8),
                (torch.rand(n, H, W, generator=g) < 0.2).to(torch.uint8), torch.randint(0, 3, (n,), generator=g))
for B in (32, 256):
    ids = torch.randint(0, n, (B,), generator=g).tolist()
    hf, vf, ang = draw_transform_params(B, generator=g)
    out = ds.batch(ids, hf, vf, ang)
    torch.cuda.synchronize()
    # time only the launch: parameters staged once
    import ctypes as C
    from multi_task_breast_cancer_b200 import _lib
    from multi_task_breast_cancer_b200.ops import ptr, stream_ptr
    from multi_task_breast_cancer_b200.data import rotation_theta
    idx_d = torch.tensor(ids, dtype=torch.int32, device="cuda")
    fl_d = torch.tensor([(1 if h else 0) | (2 if v else 0) | 4 for h, v in zip(hf, vf)], dtype=torch.uint8, device="cuda")
    th_d = torch.tensor([rotation_theta(a, H, W) for a in ang], dtype=torch.float32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for it in range(12):
        flush.zero_()                                   # > 126 MB L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.call("mtbc_augment_batch", ptr(ds.images), ptr(ds.masks), ptr(ds.labels), ptr(idx_d), ptr(fl_d), ptr(th_d),
                  B, H, W, 3, ptr(out[0]), ptr(out[1]), ptr(out[2]), C.c_void_p(stream_ptr()))
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    by = B * H * W * 10
    print(f"augment_batch B={B}: {ms * 1e3:.1f} us, {by / ms / 1e6:.0f} GB/s algorithmic ({by / 1e6:.1f} MB), "
          f"{B / ms * 1e3:.0f} img/s")
