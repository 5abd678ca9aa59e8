# three full passes of the GPU suite on the final build: anything that fails in one of them is a flaky margin
mkdir -p gpurun_out
for i in 1 2 3; do
  timeout 300 python -m pytest tests -m gpu -q -s > gpurun_out/r04e_pytest_$i.log 2>&1; echo "pass $i exit $?"
  grep -E "passed|failed" gpurun_out/r04e_pytest_$i.log | tail -2
  grep -E "^FAILED" gpurun_out/r04e_pytest_$i.log
  grep -E "worst relative deviation|masks identical with" gpurun_out/r04e_pytest_$i.log
done
