mkdir -p gpurun_out
i=0
for RX in "conv3x3_fwd|upcat_0_4.convs.conv_0 fwd" "conv3x3_dgrad|upcat_0_4.convs.conv_0 dgrad" "conv3x3_wgrad|upcat_0_4.convs.conv_0 wgrad" "conv3x3_fwd|upcat_1_3.convs.conv_0 fwd" "convT_wgrad|upcat_0_4.up convT wgrad" "convT_fwd|upcat_0_4.up convT fwd" "conv3x3_wgrad|upcat_1_3.convs.conv_0 wgrad"; do
  i=$((i+1))
  tag=$(echo "$RX" | sed 's/.*|//' | tr ' .' '__')
  ONLY="$RX" timeout 600 ncu --profile-from-start off --nvtx --set full --clock-control none --import-source on \
     -f -o gpurun_out/r02j_full_${tag} python tools/ncu_step.py /dev/null > gpurun_out/r02j_ncu_full_${tag}.log 2>&1
  echo "ncu full [$RX] exit $?"; tail -2 gpurun_out/r02j_ncu_full_${tag}.log
done
ls -la gpurun_out | grep r02j
