# round-1g GPU call: smoke fix + new API tests
mkdir -p gpurun_out/r1g
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1g/smoke.log 2>&1; tail -3 gpurun_out/r1g/smoke.log
timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/r1g/gpu_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r1g/gpu_tests.log; tail -12 gpurun_out/r1g/gpu_tests.log
