# loss snapshot inside the head graph (MTBC_SNAP_IN_GRAPH): prefetch test + e2e A/B
mkdir -p gpurun_out
O=gpurun_out
timeout 120 python -m pytest tests/test_models_gpu.py -m gpu -q -x -k "prefetches_host_batches or train_step_matches" > $O/r04g_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $O/r04g_pytest.log
for v in 0 1 0 1; do
  MTBC_SNAP_IN_GRAPH=$v timeout 100 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-library-baseline 2>$O/r04g_bench_$v.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('snap_in_graph=$v ms %.4f e2e %.4f gap %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['ms_per_step'] - d['ms_per_step']))"
done
