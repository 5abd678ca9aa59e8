mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_residual_unet.py -m gpu -q -s > gpurun_out/r02i_pytest_residual.log 2>&1; echo "residual pytest exit $?"
grep -E "passed|failed|FAILED|ERROR|Error|rel L2|running stat|worst|Dice over|assert" gpurun_out/r02i_pytest_residual.log | tail -30
# ncu evidence on this build: launch list + --set full for the five tensor kernels the verdict names
timeout 300 python tools/ncu_step.py gpurun_out/r02i_step_launches.json > gpurun_out/r02i_plain.log 2>&1 &&
timeout 900 ncu --profile-from-start off --nvtx --print-nvtx-rename kernel --print-units base \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum \
  --clock-control none --csv --log-file gpurun_out/r02i_launches.csv python tools/ncu_step.py /dev/null > gpurun_out/r02i_ncu1.log 2>&1
echo "ncu launches exit $?"
for RX in "upcat_0_4.convs.conv_0 fwd" "upcat_0_4.convs.conv_0 dgrad" "upcat_0_4.convs.conv_0 wgrad" "upcat_1_3.convs.conv_0 fwd" "upcat_0_4.up convT wgrad" "upcat_0_4.up convT fwd"; do
  tag=$(echo "$RX" | tr ' .' '__')
  timeout 600 ncu --profile-from-start off --nvtx --nvtx-include "regex:.*${RX// /.}.*" --set full --clock-control none --import-source on -c 1 \
     -f -o gpurun_out/r02i_full_${tag} python tools/ncu_step.py /dev/null > gpurun_out/r02i_ncu_full_${tag}.log 2>&1
  echo "ncu full [$RX] exit $?"
done
ls -la gpurun_out | grep r02i
