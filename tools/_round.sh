O=gpurun_out/r2zm; mkdir -p $O
timeout 900 python -m pytest tests/test_nhwc_bf16_gpu.py -x -q > $O/t.log 2>&1; tail -2 $O/t.log
timeout 600 python -m pytest tests/test_baseline_parity_gpu.py -x -q -k "layer_shapes" > $O/t2.log 2>&1; tail -2 $O/t2.log
timeout 300 python tools/resident_sweep.py 256 5 "" w > $O/sweep_w.txt 2>&1; tail -2 $O/sweep_w.txt
BCNN_B200_WG_NO_PAIR=1 timeout 300 python tools/resident_sweep.py 256 5 "" w > $O/sweep_w_nopair.txt 2>&1; tail -2 $O/sweep_w_nopair.txt
paste <(awk '{print $1,$2,$3,$4,$5,$6}' $O/sweep_w.txt) <(awk '{print $6}' $O/sweep_w_nopair.txt) | grep wgrad
