# trajectory spread (nnU-Net) + two micro-optimisations that were measured and NOT kept (32-bit index math in three
# streaming kernels, cooperative InstanceNorm backward for small planes behind MTBC_FUSED_INBWD_SMALL: DESIGN 9); the
# profile / bench legs below therefore only reproduce on the tree of that experiment
mkdir -p gpurun_out
O=gpurun_out
timeout 200 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "instance_norm or head1x1 or pixel_pair" > $O/r04d_pytest_k.log 2>&1; echo "kernel pytest exit $?"; tail -3 $O/r04d_pytest_k.log
MTBC_FUSED_INBWD_SMALL=1 timeout 200 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "instance_norm" > $O/r04d_pytest_k2.log 2>&1; echo "kernel pytest (fused small) exit $?"; tail -3 $O/r04d_pytest_k2.log
( timeout 100 python tools/diag_traj.py nnunet 4 64 200 4
  timeout 100 python tools/diag_traj.py nnunet 4 64 200 3 det
  timeout 100 python tools/diag_traj.py nnunet 8 64 200 3
  timeout 100 python tools/diag_traj.py nnunet 4 128 200 3 ) > $O/r04d_traj.txt 2>&1
cat $O/r04d_traj.txt | grep -v Warning
pp() { # name, env...
  n=$1; shift
  env "$@" timeout 120 python tools/profile_plan.py unetpp 32 256 400 > $O/r04d_pp_$n.txt 2>&1
  echo "== $n: $(head -1 $O/r04d_pp_$n.txt)"
  grep -E "(mtbc_in_bwd|mtbc_head1x1_bwd|mtbc_maxpool2_bwd|mtbc_in_apply) +[0-9]" $O/r04d_pp_$n.txt
}
pp base
pp fusedsmall MTBC_FUSED_INBWD_SMALL=1
for v in 0 1 0 1; do
  MTBC_FUSED_INBWD_SMALL=$v timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-library-baseline 2>$O/r04d_bench_$v.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['roofline']['by_kernel_ms']; print('fusedsmall=$v ms %.4f e2e %.4f in_bwd %.4f head_bwd %.4f pool_bwd %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step'], k['mtbc_in_bwd'], k['mtbc_head1x1_bwd'], k['mtbc_maxpool2_bwd']))"
done
