mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 900 $TR tools/dp_check.py unetpp 32 256 > gpurun_out/r02f_dp_check_unetpp.txt 2>&1; echo "dp_check unetpp exit $?"; grep -E "==|ok\]|ok tol|FAIL|DP CHECK|Error|error" gpurun_out/r02f_dp_check_unetpp.txt | tail -40
timeout 600 $TR tools/dp_check.py nnunet 16 128 > gpurun_out/r02f_dp_check_nnunet.txt 2>&1; echo "dp_check nnunet exit $?"; grep -E "tol|DP CHECK|FAIL" gpurun_out/r02f_dp_check_nnunet.txt | tail -12
