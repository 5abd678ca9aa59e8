# pair view with / without the zero blocks of the pair operand (MTBC_PAIR_SKIP).  gpurun --timeout 600 -- 'bash tools/_r04b.sh'
mkdir -p gpurun_out
O=gpurun_out
timeout 200 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "pixel_pair" > $O/r04b_pytest.log 2>&1; echo "pytest exit $?"; tail -5 $O/r04b_pytest.log
pp() { # name, env...
  n=$1; shift
  env "$@" timeout 120 python tools/profile_plan.py unetpp 32 256 400 > $O/r04b_pp_$n.txt 2>&1
  echo "== $n: $(head -1 $O/r04b_pp_$n.txt)"
  grep -E "conv3x3_fwd +[0-9]" $O/r04b_pp_$n.txt
  grep -E "fwd 32x256x256 \[24\]->24" $O/r04b_pp_$n.txt | head -12
}
pp skip MTBC_PAIR=1
pp noskip MTBC_PAIR=1 MTBC_PAIR_SKIP=0
pp skip_epi2 MTBC_PAIR=1 MTBC_HALO_EPI=2
for v in 0 1 0 1; do
  MTBC_PAIR=$v timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-library-baseline 2>$O/r04b_bench_$v.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pair=$v ms %.4f e2e %.4f fwd %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['by_kernel_ms']['conv3x3_fwd']))"
done
