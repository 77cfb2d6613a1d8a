mkdir -p gpurun_out/r1k
timeout 200 python tools/infer_bench.py --cpu > gpurun_out/r1k/infer.jsonl 2> gpurun_out/r1k/infer.err; echo "infer rc=$?"; cut -c1-200 gpurun_out/r1k/infer.jsonl; grep -o '"cpu_baseline.*' gpurun_out/r1k/infer.jsonl; tail -5 gpurun_out/r1k/infer.err
