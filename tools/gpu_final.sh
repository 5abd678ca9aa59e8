#!/bin/bash
# Final evidence run of a round on one B200:  gpurun --timeout 2400 -- 'bash tools/gpu_final.sh r02z'
# GPU tests + smoke, bench (both arms), other configs, per-layer table, per-launch times, ncu launch list of one step,
# ncu --set full of selected plan launches (ONLY=<label substring> of tools/ncu_step.py).
TAG=${1:-r02z}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
python -c "from multi_task_breast_cancer_b200 import build; print('lib digest', build.lib_digest()); print('src digest', build._digest())" > $OUT/${TAG}_digest.txt 2>&1
cat $OUT/${TAG}_digest.txt
timeout 1200 python -m pytest tests -m gpu -q --durations=8 > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
grep -E "passed|failed|FAILED|ERROR|pytest exit" $OUT/${TAG}_pytest_gpu.log | tail -12
timeout 300 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/${TAG}_smoke.log
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
cat $OUT/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "ref exit $?"
cat $OUT/${TAG}_bench_ref.json
timeout 300 python bench.py --arch nnunet --no-cpu-baseline --no-library-baseline > $OUT/${TAG}_bench_nnunet.json 2> $OUT/${TAG}_bench_nnunet.err; echo "nnunet exit $?"
timeout 300 python bench.py --arch bts --size 128 --no-cpu-baseline --no-library-baseline > $OUT/${TAG}_bench_bts.json 2> $OUT/${TAG}_bench_bts.err; echo "bts exit $?"
timeout 300 python bench.py --size 512 --no-cpu-baseline --no-library-baseline > $OUT/${TAG}_bench_512.json 2> $OUT/${TAG}_bench_512.err; echo "512 exit $?"
for a in nnunet bts 512; do python -c "
import json,sys; d=json.loads(open('$OUT/${TAG}_bench_$a.json').read().strip().splitlines()[-1]); print('$a', round(d['value'],1), d['unit'], round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value'],1))"; done
timeout 300 python tools/per_layer_table.py unetpp 32 256 > $OUT/${TAG}_per_layer.md 2> $OUT/${TAG}_per_layer.err; echo "per-layer exit $?"
head -36 $OUT/${TAG}_per_layer.md
timeout 300 python tools/profile_plan.py unetpp 32 256 400 > $OUT/${TAG}_profile_plan.txt 2>&1
timeout 300 python tools/profile_plan.py nnunet 32 256 400 > $OUT/${TAG}_profile_plan_nnunet.txt 2>&1; head -12 $OUT/${TAG}_profile_plan_nnunet.txt
timeout 300 python tools/diag_fused_dgrad.py > $OUT/${TAG}_diag_fused_dgrad.txt 2>&1
timeout 300 python tools/ncu_step.py $OUT/${TAG}_step_launches.json > $OUT/${TAG}_plain.log 2>&1 &&
timeout 900 ncu --profile-from-start off --nvtx --print-nvtx-rename kernel --print-units base \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum \
    --clock-control none --csv --log-file $OUT/${TAG}_launches.csv python tools/ncu_step.py /dev/null > $OUT/${TAG}_ncu1.log 2>&1
echo "ncu launches exit $?"
i=0
while IFS= read -r SEL; do
  [ -z "$SEL" ] && continue
  [ -n "$SKIP_FULL" ] && continue
  i=$((i+1))
  ONLY="$SEL" timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
      -f -o $OUT/${TAG}_full_${i} python tools/ncu_step.py /dev/null > $OUT/${TAG}_ncu_full_${i}.log 2>&1
  echo "ncu full [$SEL] exit $?"
  echo "$i $SEL" >> $OUT/${TAG}_full_index.txt
done <<'SEL_EOF'
conv_0_0.conv_1 fwd
upcat_0_4.convs.conv_0 fwd
upcat_0_4.up convT bwd
upcat_0_4.convs.conv_1 dgrad
upcat_0_4.convs.conv_0 wgrad
upcat_0_4.convs.conv_0 dgrad
SEL_EOF
ls -la $OUT | grep ${TAG} | tail -40
