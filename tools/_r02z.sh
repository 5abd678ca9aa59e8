mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "transposed_conv" > gpurun_out/r02z_pytest_k.log 2>&1; echo "pytest kernels exit $?"; tail -3 gpurun_out/r02z_pytest_k.log; grep "BAD\|EXC" gpurun_out/r02z_pytest_k.log | head -20
timeout 300 python tools/profile_plan.py unetpp 32 256 400 > gpurun_out/r02z_unetpp.txt 2>&1; head -3 gpurun_out/r02z_unetpp.txt; grep -E "convT bwd" gpurun_out/r02z_unetpp.txt | head -12
for v in 0 1 0 1; do
  MTBC_FUSE_CONVT_BWD=$v timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-library-baseline 2>gpurun_out/r02z_bench_f$v.err | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('convT fuse=$v ms %.4f e2e %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step']))"
done
