O=gpurun_out/r2zg; mkdir -p $O
timeout 900 python -m pytest tests/test_nhwc_bf16_gpu.py -x -q > $O/t_kernels.log 2>&1; tail -3 $O/t_kernels.log
BCNN_B200_HALO=1 timeout 900 python -m pytest tests/test_nhwc_bf16_gpu.py -x -q > $O/t_halo1.log 2>&1; tail -2 $O/t_halo1.log
timeout 300 python tools/resident_sweep.py 256 5 "" sd > $O/sweep_default.txt 2>&1; grep "3x3\|7x7" $O/sweep_default.txt; tail -3 $O/sweep_default.txt
BCNN_B200_HALO=0 timeout 300 python tools/resident_sweep.py 256 5 "64,56,64,3,1,1" sd > $O/sweep_nohalo.txt 2>&1; grep "3x3" $O/sweep_nohalo.txt
BCNN_B200_HALO=1 timeout 300 python tools/resident_sweep.py 256 5 "" sd > $O/sweep_halo1.txt 2>&1; grep "3x3" $O/sweep_halo1.txt
