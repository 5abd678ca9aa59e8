mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "fused_norm_backward" > gpurun_out/r02u_pytest_k.log 2>&1; echo "pytest kernels exit $?"; tail -3 gpurun_out/r02u_pytest_k.log; grep "BAD\|EXC" gpurun_out/r02u_pytest_k.log | head -20
run() { tag=$1; shift
  env "$@" timeout 300 python tools/profile_plan.py unetpp 32 256 400 > gpurun_out/r02u_$tag.txt 2>&1
  grep -E "^  (conv3x3_dgrad|mtbc_in_bwd)" gpurun_out/r02u_$tag.txt | head -4
  grep -E "24<-24 acc=0|48<-48 acc=0" gpurun_out/r02u_$tag.txt | head -3
}
run f1 MTBC_FUSE_INBWD=1
run f1_epi2 MTBC_FUSE_INBWD=1 MTBC_HALO_EPI=2
for v in 0 1 0 1; do
  MTBC_FUSE_INBWD=$v timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-library-baseline 2>gpurun_out/r02u_bench_f$v.err | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('fuse=$v ms %.4f e2e %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step']))"
done
