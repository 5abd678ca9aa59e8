#!/bin/bash
# One gpurun call: GPU tests, bench (both arms), per-launch timing, ncu launch list + one --set full capture.
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
timeout 300 python tools/profile_plan.py unetpp 32 256 70 > $OUT/${TAG}_profile_plan.txt 2>&1; echo "profile exit $?"
head -30 $OUT/${TAG}_profile_plan.txt
if [ -z "$SKIP_NCU" ]; then
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
  timeout 300 $CMD > $OUT/${TAG}_plain.log 2>&1 &&
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 420 --csv \
      --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu1.log 2>&1
  echo "ncu launches exit $?"
  timeout 300 $CMD > $OUT/${TAG}_plain2.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-conv_halo} -s ${NCU_SKIP:-60} -c ${NCU_COUNT:-4} \
      -f -o $OUT/${TAG}_top $CMD > $OUT/${TAG}_ncu2.log 2>&1
  echo "ncu full exit $?"
fi
ls -la $OUT
