mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x > gpurun_out/r02r_pytest_k.log 2>&1; echo "pytest kernels exit $?"; tail -3 gpurun_out/r02r_pytest_k.log
for v in 0 1 0 1; do
  MTBC_PDL=$v timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-library-baseline 2>gpurun_out/r02r_bench_pdl$v.err | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('pdl=$v ms %.4f e2e %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step']))"
done
for v in 0 1; do
  MTBC_PDL=$v timeout 300 python bench.py --arch nnunet --steps 40 --warmup 5 --no-cpu-baseline --no-library-baseline 2>gpurun_out/r02r_bench_nn_pdl$v.err | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('nnunet pdl=$v ms %.4f e2e %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step']))"
done
timeout 900 python -m pytest tests/test_models_gpu.py tests/test_grad_wiring_gpu.py tests/test_trainer_gpu.py -m gpu -q -x > gpurun_out/r02r_pytest_m.log 2>&1; echo "pytest models exit $?"; tail -3 gpurun_out/r02r_pytest_m.log
