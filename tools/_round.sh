O=gpurun_out/r2zf; mkdir -p $O
timeout 900 python -m pytest tests/test_nhwc_bf16_gpu.py tests/test_kernels_gpu.py -x -q > $O/t_kernels.log 2>&1; tail -3 $O/t_kernels.log
timeout 300 python tools/resident_sweep.py 256 5 "" sdw > $O/sweep_default.txt 2>&1; tail -4 $O/sweep_default.txt; head -5 $O/sweep_default.txt
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; head -c 400 $O/bench_n1.json; echo
timeout 300 python tools/node_profile.py resnet50 256 resident > $O/nodes_resident_b256.txt 2>&1; tail -8 $O/nodes_resident_b256.txt
