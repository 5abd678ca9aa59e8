#!/bin/bash
# One gpurun call: GPU tests, bench (both arms), per-launch timing, ncu launch list of one step (+ optional full capture).
#   gpurun --timeout 1800 -- 'bash tools/gpu_round.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
  tail -5 $OUT/${TAG}_pytest_gpu.log
  timeout 300 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/${TAG}_smoke.log
fi
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
cat $OUT/${TAG}_bench.json
if [ -z "$SKIP_REF" ]; then
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "ref exit $?"
  cat $OUT/${TAG}_bench_ref.json
fi
timeout 300 python tools/profile_plan.py unetpp 32 256 400 > $OUT/${TAG}_profile_plan.txt 2>&1; echo "profile exit $?"
head -30 $OUT/${TAG}_profile_plan.txt
if [ -z "$SKIP_NCU" ]; then
  # launch list of ONE eager step, kernels renamed by the NVTX label of their plan launch (tools/ncu_step.py)
  timeout 300 python tools/ncu_step.py $OUT/${TAG}_step_launches.json > $OUT/${TAG}_plain.log 2>&1 &&
  timeout 900 ncu --profile-from-start off --nvtx --print-nvtx-rename kernel --print-units base \
      --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum \
      --clock-control none --csv --log-file $OUT/${TAG}_launches.csv python tools/ncu_step.py /dev/null > $OUT/${TAG}_ncu1.log 2>&1
  echo "ncu launches exit $?"
  if [ -n "$NCU_FULL" ]; then
    timeout 300 python tools/ncu_step.py /dev/null > $OUT/${TAG}_plain2.log 2>&1 &&
    timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
        -k regex:${NCU_KERNEL:-conv_halo} -s ${NCU_SKIP:-0} -c ${NCU_COUNT:-3} \
        -f -o $OUT/${TAG}_top python tools/ncu_step.py /dev/null > $OUT/${TAG}_ncu2.log 2>&1
    echo "ncu full exit $?"
  fi
fi
ls -la $OUT | tail -20
