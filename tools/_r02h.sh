mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_residual_unet.py -m gpu -q -s -x > gpurun_out/r02h_pytest_residual.log 2>&1; echo "residual pytest exit $?"
grep -E "passed|failed|FAILED|ERROR|Error|rel L2|running stat|worst|Dice over" gpurun_out/r02h_pytest_residual.log | tail -30
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_residual_unet.py > gpurun_out/r02h_pytest_rest.log 2>&1; echo "rest pytest exit $?"
tail -4 gpurun_out/r02h_pytest_rest.log
