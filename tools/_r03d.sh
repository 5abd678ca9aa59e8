mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "first_layer or composed" 2>&1 | tail -3
timeout 300 python tools/profile_plan.py unetpp 32 256 400 > gpurun_out/r03d_unetpp.txt 2>&1; grep -E "conv_first|gap_fc|adam|param_jobs|dice|focal|head1x1" gpurun_out/r03d_unetpp.txt | head -30
for v in 1 0 1 0; do
  if [ $v = 0 ]; then export MTBC_FIRST_WGRAD_PIXELS=1; else unset MTBC_FIRST_WGRAD_PIXELS; fi
  timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-library-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('rows=$v ms %.4f e2e %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step']))"
done
