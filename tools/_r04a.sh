# Round-2 late experiment: pixel-pair view of the level-0 24 -> 24 forward convs (MTBC_PAIR) and 16 epilogue warps on
# G = 1 statistics layers of 64 columns (MTBC_HALO_EPI4_G1).  gpurun --timeout 600 -- 'bash tools/_r04a.sh'
mkdir -p gpurun_out
O=gpurun_out
timeout 200 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "pixel_pair or conv3x3_forward" > $O/r04a_pytest.log 2>&1; echo "pytest exit $?"; tail -5 $O/r04a_pytest.log
pp() { # name, env...
  n=$1; shift
  env "$@" timeout 120 python tools/profile_plan.py unetpp 32 256 400 > $O/r04a_pp_$n.txt 2>&1
  echo "== $n: $(head -1 $O/r04a_pp_$n.txt)"
  grep -E "conv3x3_fwd +[0-9]" $O/r04a_pp_$n.txt
  grep -E "fwd 32x256x256 \[24\]->24|fwd 32x128x128 \[24\]->48|fwd 32x128x128 \[48\]->48" $O/r04a_pp_$n.txt | head -12
}
pp base MTBC_PAIR=0
pp pair MTBC_PAIR=1
pp pair_epi2 MTBC_PAIR=1 MTBC_HALO_EPI=2
pp epi4g1 MTBC_PAIR=0 MTBC_HALO_EPI4_G1=1
for v in 0 1 0 1; do
  MTBC_PAIR=$v timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-library-baseline 2>$O/r04a_bench_$v.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pair=$v ms %.4f e2e %.4f fwd %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['by_kernel_ms']['conv3x3_fwd']))"
done
