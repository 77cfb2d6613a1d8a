mkdir -p gpurun_out/r1h
timeout 240 python -m pytest tests -m gpu -x -q -rs > gpurun_out/r1h/gpu_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r1h/gpu_tests.log; tail -25 gpurun_out/r1h/gpu_tests.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1h/smoke.log 2>&1; tail -2 gpurun_out/r1h/smoke.log
