mkdir -p gpurun_out/r1p
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r1p/gpu_tests.log 2>&1; tail -3 gpurun_out/r1p/gpu_tests.log
for i in 1 2; do
timeout 300 python bench.py --no-cpu-baseline --no-rooflines > gpurun_out/r1p/bench_n1.json 2> gpurun_out/r1p/bench_n1.err; head -c 250 gpurun_out/r1p/bench_n1.json; echo
BCNN_B200_NO_L2_HINTS=1 timeout 300 python bench.py --no-cpu-baseline --no-rooflines > gpurun_out/r1p/bench_n1_nohint.json 2> gpurun_out/r1p/bench_n1.err; head -c 250 gpurun_out/r1p/bench_n1_nohint.json; echo
done
