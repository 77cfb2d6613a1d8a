O=gpurun_out/r2zo; mkdir -p $O
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; head -c 250 $O/bench_n1.json; echo; tail -2 $O/bench_n1.err
timeout 900 python -m pytest tests/test_nets_gpu.py tests/test_kernels_gpu.py -x -q > $O/t.log 2>&1; tail -3 $O/t.log
BCNN_B200_DW_PLANE=0 timeout 200 python tools/dw_sweep.py 64 > $O/dw_rows.txt 2>&1
timeout 200 python tools/dw_sweep.py 64 > $O/dw_default.txt 2>&1
BCNN_B200_DW_PLANE=12544 timeout 200 python tools/dw_sweep.py 64 > $O/dw_all.txt 2>&1
paste -d'|' <(cut -c1-58 $O/dw_rows.txt) <(cut -c20-58 $O/dw_default.txt) <(cut -c20-58 $O/dw_all.txt)
