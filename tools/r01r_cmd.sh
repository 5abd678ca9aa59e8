mkdir -p gpurun_out
timeout 200 python tools/diag_classifiers.py > gpurun_out/r01r_classifiers.txt 2>&1; echo "diag exit $?"; tail -15 gpurun_out/r01r_classifiers.txt
timeout 200 python -m pytest tests/test_widen_classifiers.py -m gpu -q > gpurun_out/r01r_pytest_classifiers.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/r01r_pytest_classifiers.log
timeout 100 python __graft_entry__.py --smoke > gpurun_out/r01r_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r01r_smoke.log
timeout 150 python bench.py > gpurun_out/r01r_bench.json 2> gpurun_out/r01r_bench.err; echo "bench exit $?"; cat gpurun_out/r01r_bench.json
