O=gpurun_out/r2zx; mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_nets_gpu.py -x -q -k "sgd_update_multi or pack_table" > $O/t.log 2>&1; tail -3 $O/t.log
bash tools/sanitize.sh $O 420
