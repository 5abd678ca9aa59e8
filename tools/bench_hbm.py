"""GPU tool: what do plain PyTorch kernels reach on this B200 for read-only, write-only and mixed streams (the practical
roofs beside MEASURED_PEAKS.json's copy figure)?   python tools/bench_hbm.py"""
import torch

def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

for mb in (268, 1072):
    n = mb * 1000 * 1000 // 2
    a = torch.randn(n, device="cuda", dtype=torch.bfloat16)
    b = torch.empty_like(a)
    c = torch.empty_like(a)
    af = a.view(torch.float32)
    print(f"--- {mb} MB tensors")
    for name, nbytes, fn in [
        ("write only  zero_()", 2 * n, lambda: b.zero_()),
        ("write only  fill_(1)", 2 * n, lambda: b.fill_(1.0)),
        ("read only   sum bf16", 2 * n, lambda: a.sum()),
        ("read only   sum fp32 view", 2 * n, lambda: af.sum()),
        ("read only   max fp32 view", 2 * n, lambda: af.max()),
        ("1r + 1w     copy_", 4 * n, lambda: b.copy_(a)),
        ("2r + 1w     add", 6 * n, lambda: torch.add(a, b, out=c)),
        ("1r + 1w     relu", 4 * n, lambda: torch.relu(a, out=b) if False else torch.clamp_min(a, 0, out=b)),
    ]:
        t = timeit(fn)
        print(f"  {name:28s} {t:8.1f} us  {nbytes / t / 1e3:7.0f} GB/s")
