python -m pytest tests -m gpu -q > gpurun_out/r2_full_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_full_tests.log
tail -8 gpurun_out/r2_full_tests.log
