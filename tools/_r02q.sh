mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "composed or head1x1" > gpurun_out/r02q_pytest_k.log 2>&1; echo "pytest kernels exit $?"; tail -5 gpurun_out/r02q_pytest_k.log
timeout 900 python -m pytest tests/test_models_gpu.py tests/test_grad_wiring_gpu.py -m gpu -q -x -k "nnunet or bts or BTS" > gpurun_out/r02q_pytest_m.log 2>&1; echo "pytest models exit $?"; tail -5 gpurun_out/r02q_pytest_m.log
timeout 300 python tools/profile_plan.py nnunet 32 256 400 > gpurun_out/r02q_nnunet.txt 2>&1; head -12 gpurun_out/r02q_nnunet.txt
timeout 300 python tools/profile_plan.py unetpp 32 256 400 > gpurun_out/r02q_unetpp.txt 2>&1; head -6 gpurun_out/r02q_unetpp.txt
