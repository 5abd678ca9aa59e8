mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "fused_norm_backward or data_gradient" > gpurun_out/r02s_pytest_k.log 2>&1; echo "pytest kernels exit $?"; tail -3 gpurun_out/r02s_pytest_k.log; grep "BAD\|EXC" gpurun_out/r02s_pytest_k.log | head -20
for v in 0 1 0 1; do
  MTBC_FUSE_INBWD=$v timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-library-baseline 2>gpurun_out/r02s_bench_f$v.err | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('fuse=$v ms %.4f e2e %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step']), {k:v for k,v in d['roofline']['by_kernel_ms'].items() if 'in_bwd' in k or 'dgrad' in k})"
done
timeout 900 python -m pytest tests/test_models_gpu.py tests/test_grad_wiring_gpu.py tests/test_trainer_gpu.py -m gpu -q -x > gpurun_out/r02s_pytest_m.log 2>&1; echo "pytest models exit $?"; tail -3 gpurun_out/r02s_pytest_m.log
