mkdir -p gpurun_out
timeout 300 python tools/diag_fused_dgrad.py > gpurun_out/r02y_diag.txt 2>&1; cat gpurun_out/r02y_diag.txt
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02y_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r02y_pytest.log
MTBC_FUSE_INBWD=1 timeout 900 python -m pytest tests/test_models_gpu.py tests/test_grad_wiring_gpu.py -m gpu -q -x > gpurun_out/r02y_pytest_fused.log 2>&1; echo "pytest fused exit $?"; tail -3 gpurun_out/r02y_pytest_fused.log
for v in 0 1; do
  MTBC_FUSE_INBWD=$v timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-library-baseline 2>gpurun_out/r02y_bench_f$v.err | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('fuse=$v ms %.4f e2e %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step']))"
done
