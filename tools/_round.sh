mkdir -p gpurun_out/r1t
timeout 100 python -m pytest tests -m gpu -x -q > gpurun_out/r1t/gpu_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r1t/gpu_tests.log; tail -6 gpurun_out/r1t/gpu_tests.log
