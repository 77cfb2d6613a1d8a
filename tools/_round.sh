# round-1f GPU call (short budget): parity suite first, then smoke, optimizer roofline, adam ncu, short bench
mkdir -p gpurun_out/r1f
timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/r1f/gpu_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r1f/gpu_tests.log; tail -8 gpurun_out/r1f/gpu_tests.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1f/smoke.log 2>&1; tail -2 gpurun_out/r1f/smoke.log
timeout 60 python tools/optim_bench.py > gpurun_out/r1f/optim_bench.txt 2>&1; cat gpurun_out/r1f/optim_bench.txt
timeout 90 ncu --set full --clock-control none --import-source on -k regex:adam_kernel -s 2 -c 1 -f -o gpurun_out/r1f/full_adam_kernel python tools/optim_bench.py 25557032 > gpurun_out/r1f/ncu_adam.log 2>&1; echo "ncu rc=$?"
timeout 200 python bench.py --steps 5 --warmup 3 > gpurun_out/r1f/bench_n1.json 2> gpurun_out/r1f/bench_n1.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r1f/bench_n1.json
