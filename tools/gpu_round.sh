#!/bin/bash
# One gpurun call: GPU tests, smoke, bench (both arms), per-layer dual-roof table, ncu launch list of one step
# (+ optional --set full capture of selected kernels).
#   gpurun --timeout 1800 -- 'bash tools/gpu_round.sh [tag]'
# env: SKIP_TESTS, SKIP_REF, SKIP_NCU, NCU_FULL="regex1 regex2 ..." (one .ncu-rep per regex, NCU_COUNT launches each)
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
python -c "from multi_task_breast_cancer_b200 import build; print('lib digest', build.lib_digest()); print('src digest', build._digest())" > $OUT/${TAG}_digest.txt 2>&1
cat $OUT/${TAG}_digest.txt
if [ -z "$SKIP_TESTS" ]; then
  timeout 1200 python -m pytest tests -m gpu -q -s --durations=8 ${PYTEST_ARGS} > $OUT/${TAG}_pytest_gpu.log 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
  grep -E "passed|failed|FAILED|ERROR|pytest exit" $OUT/${TAG}_pytest_gpu.log | tail -30
  timeout 300 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/${TAG}_smoke.log
fi
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
cat $OUT/${TAG}_bench.json
if [ -z "$SKIP_REF" ]; then
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "ref exit $?"
  cat $OUT/${TAG}_bench_ref.json
fi
timeout 300 python tools/per_layer_table.py unetpp 32 256 > $OUT/${TAG}_per_layer.md 2> $OUT/${TAG}_per_layer.err; echo "per-layer exit $?"
head -40 $OUT/${TAG}_per_layer.md
if [ -z "$SKIP_NCU" ]; then
  # launch list of ONE eager step, kernels renamed by the NVTX label of their plan launch (tools/ncu_step.py)
  timeout 300 python tools/ncu_step.py $OUT/${TAG}_step_launches.json > $OUT/${TAG}_plain.log 2>&1 &&
  timeout 900 ncu --profile-from-start off --nvtx --print-nvtx-rename kernel --print-units base \
      --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum \
      --clock-control none --csv --log-file $OUT/${TAG}_launches.csv python tools/ncu_step.py /dev/null > $OUT/${TAG}_ncu1.log 2>&1
  echo "ncu launches exit $?"
  i=0
  for RX in $NCU_FULL; do
    i=$((i+1))
    timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
        -k regex:${RX} -s ${NCU_SKIP:-0} -c ${NCU_COUNT:-2} \
        -f -o $OUT/${TAG}_full_${i}_${RX//[^a-zA-Z0-9_]/} python tools/ncu_step.py /dev/null > $OUT/${TAG}_ncu_full_${i}.log 2>&1
    echo "ncu full ${RX} exit $?"
  done
fi
ls -la $OUT | tail -30
