mkdir -p gpurun_out
for v in 1 0 1 0 1 0; do
  MTBC_SIDE_HEADS=$v timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-library-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('side_heads=$v ms %.4f e2e %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step']))"
done
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02n_pytest.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/r02n_pytest.log
