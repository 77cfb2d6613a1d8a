mkdir -p gpurun_out/r1s
timeout 120 python -m pytest tests -m gpu -x -q > gpurun_out/r1s/gpu_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r1s/gpu_tests.log; tail -8 gpurun_out/r1s/gpu_tests.log
timeout 40 python bench.py --workload mnist --batch 64 --res 28 --no-rooflines --no-cpu-baseline --steps 50 --warmup 5 2>/dev/null | cut -c1-130 | tee gpurun_out/r1s/bench_mnist.txt
timeout 40 python bench.py --workload cifar --batch 128 --res 32 --no-rooflines --no-cpu-baseline --steps 50 --warmup 5 2>/dev/null | cut -c1-130 | tee gpurun_out/r1s/bench_cifar.txt
