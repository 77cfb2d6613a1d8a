mkdir -p gpurun_out/r1i
timeout 200 python tools/infer_bench.py > gpurun_out/r1i/infer.jsonl 2> gpurun_out/r1i/infer.err; echo "infer rc=$?"; cut -c1-420 gpurun_out/r1i/infer.jsonl; tail -5 gpurun_out/r1i/infer.err
