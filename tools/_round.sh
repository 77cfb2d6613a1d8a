mkdir -p gpurun_out/r1k /tmp/ncu_r1k
timeout 900 python tools/ncu_rooflines.py capture /tmp/ncu_r1k 256 > /dev/null
python tools/ncu_rooflines.py parse /tmp/ncu_r1k gpurun_out/r1k/ncu_traffic.json > gpurun_out/r1k/ncu_full_summary.md 2> gpurun_out/r1k/parse.err
cat gpurun_out/r1k/ncu_full_summary.md; tail -3 gpurun_out/r1k/parse.err
cp /tmp/ncu_r1k/bn_apply.ncu-rep gpurun_out/r1k/
