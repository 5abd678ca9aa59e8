mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "fused_norm_backward" > gpurun_out/r02x_pytest_k.log 2>&1; echo "pytest kernels exit $?"; tail -3 gpurun_out/r02x_pytest_k.log; grep "BAD\|EXC" gpurun_out/r02x_pytest_k.log | head -20
timeout 300 python tools/diag_fused_dgrad.py > gpurun_out/r02x_diag.txt 2>&1; cat gpurun_out/r02x_diag.txt
