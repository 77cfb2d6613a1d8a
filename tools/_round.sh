mkdir -p gpurun_out/r1l
timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/r1l/gpu_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r1l/gpu_tests.log; tail -25 gpurun_out/r1l/gpu_tests.log
