#!/bin/bash
# One gpurun call: GPU parity suite, bench line, per-layer conv sweep, ncu launch list and
# full captures of the dominant kernels.  Usage (on the GPU box): bash tools/gpu_round.sh TAG [parts]
# parts: any of t(ests) b(ench) s(weep) l(aunch list) n(cu full)   default: tbsln
TAG=${1:-rX}
PARTS=${2:-tbsln}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
if [[ $PARTS == *t* ]]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/gpu_tests.log 2>&1
  echo "tests rc=$?" >> $OUT/gpu_tests.log
  tail -3 $OUT/gpu_tests.log
fi
if [[ $PARTS == *b* ]]; then
  timeout 900 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err
  echo "bench rc=$?"; head -c 600 $OUT/bench_n1.json; echo
fi
if [[ $PARTS == *s* ]]; then
  timeout 600 python tools/conv_sweep.py 256 5 > $OUT/conv_sweep_b256.txt 2>&1
  tail -4 $OUT/conv_sweep_b256.txt
fi
if [[ $PARTS == *l* ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
     --log-file $OUT/launches.csv python tools/profile_step.py resnet50 64 2 > $OUT/launches.log 2>&1
  python tools/summarize_launches.py $OUT/launches.csv > $OUT/launches_resnet50_b64.md 2>&1
  head -20 $OUT/launches_resnet50_b64.md
fi
if [[ $PARTS == *n* ]]; then
  for K in conv_tma_fwd_kernel conv_tma_wgrad_kernel bn_bwd_apply_kernel bn_apply_kernel; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 20 -c 2 \
       -f -o $OUT/full_$K python tools/profile_step.py resnet50 64 1 > $OUT/full_$K.log 2>&1
    echo "ncu $K rc=$?"
  done
fi
ls -la $OUT
