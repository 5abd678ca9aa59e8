# pair view forward + fused data gradient: kernel tests, whole GPU suite with the view on, per-launch times, A/B.
mkdir -p gpurun_out
O=gpurun_out
timeout 200 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "pixel_pair" > $O/r04c_pytest_k.log 2>&1; echo "kernel pytest exit $?"; tail -4 $O/r04c_pytest_k.log
MTBC_PAIR=1 timeout 600 python -m pytest tests -m gpu -q > $O/r04c_pytest_pair.log 2>&1; echo "suite (MTBC_PAIR=1) exit $?"; tail -6 $O/r04c_pytest_pair.log
pp() { # name, env...
  n=$1; shift
  env "$@" timeout 120 python tools/profile_plan.py unetpp 32 256 400 > $O/r04c_pp_$n.txt 2>&1
  echo "== $n: $(head -1 $O/r04c_pp_$n.txt)"
  grep -E "(conv3x3_fwd|conv3x3_dgrad|mtbc_in_bwd) +[0-9]" $O/r04c_pp_$n.txt
  grep -E "dgrad 32x256x256 24<-24" $O/r04c_pp_$n.txt | head -6
}
pp base MTBC_PAIR=0
pp pair MTBC_PAIR=1
for v in 0 1 0 1 0 1; do
  MTBC_PAIR=$v timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-library-baseline 2>$O/r04c_bench_$v.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['roofline']['by_kernel_ms']; print('pair=$v ms %.4f e2e %.4f fwd %.4f dgrad %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step'], k['conv3x3_fwd'], k['conv3x3_dgrad']))"
done
