#!/bin/bash
# One gpurun call that produces the round's evidence: GPU parity suite, smoke, bench line, per-layer
# convolution sweep, per-node profile, ncu launch list of one eager step and `ncu --set full` captures
# of the dominant kernels.   Usage (on the GPU box): bash tools/gpu_round.sh TAG [parts]
# parts: any of t(ests) b(ench) s(weep) p(er-node) l(aunch list) n(cu full)   default: tbspln
TAG=${1:-rX}
PARTS=${2:-tbspln}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
if [[ $PARTS == *t* ]]; then
  timeout 1500 python -m pytest tests -m gpu -q > $OUT/gpu_tests.log 2>&1
  echo "tests rc=$?" >> $OUT/gpu_tests.log
  tail -3 $OUT/gpu_tests.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1
  tail -1 $OUT/smoke.log
fi
if [[ $PARTS == *b* ]]; then
  timeout 900 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err
  echo "bench rc=$?"; head -c 300 $OUT/bench_n1.json; echo
fi
if [[ $PARTS == *s* ]]; then
  timeout 600 python tools/resident_sweep.py 256 5 "" sdw > $OUT/resident_sweep_b256.txt 2>&1
  tail -5 $OUT/resident_sweep_b256.txt
fi
if [[ $PARTS == *p* ]]; then
  timeout 300 python tools/node_profile.py resnet50 256 resident > $OUT/nodes_resident_b256.txt 2>&1
  tail -1 $OUT/nodes_resident_b256.txt
fi
if [[ $PARTS == *l* ]]; then
  BCNN_B200_GRAPHS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
     --log-file $OUT/launches.csv python tools/profile_step.py resnet50 256 2 resident > $OUT/launches.log 2>&1
  # the summary is the second step alone: it starts at the last pack_jobs_kernel launch (the first
  # kernel of a TRAIN forward once the pack table exists), else behind the first step's launch count
  N=$(python -c "
import csv
rows=[r for r in csv.DictReader(l for l in open('$OUT/launches.csv') if l.startswith('\"')) if r.get('Metric Name')=='gpu__time_duration.sum']
idx=[i for i,r in enumerate(rows) if 'pack_jobs_kernel' in r['Kernel Name']]
print(idx[-1] if idx else len(rows)//2)")
  python tools/summarize_launches.py $OUT/launches.csv $N > $OUT/launches_resident_b256.md 2>&1
  head -16 $OUT/launches_resident_b256.md
fi
if [[ $PARTS == *n* ]]; then
  # one launch of the main kernel of a roofline entry per capture: "entry name|shape|pass|kernel regex"
  while IFS='|' read -r NAME SHAPE PASS KERN; do
    STEM=$(echo "$NAME" | tr ' /<>@,+-' '_________' | tr -s '_')
    # the .ncu-rep files (~35 MB each) stay on the box: gpurun brings back at most 64 MiB, so only
    # the raw-page metrics of each capture travel
    mkdir -p /tmp/ncu_$TAG
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:$KERN -s 3 -c 1 -f \
       -o /tmp/ncu_$TAG/full_$STEM python tools/resident_sweep.py 256 2 "$SHAPE" $PASS > $OUT/full_$STEM.log 2>&1
    echo "ncu $NAME rc=$?"
    echo "## $NAME" >> $OUT/ncu_full_metrics.txt
    python tools/ncu_metrics.py /tmp/ncu_$TAG/full_$STEM.ncu-rep >> $OUT/ncu_full_metrics.txt 2>&1
    # the traffic entry bench.py reads (merged into profiles/ncu_traffic.json back home)
    python tools/ncu_traffic.py $OUT/ncu_traffic.json "$NAME=/tmp/ncu_$TAG/full_$STEM.ncu-rep" > /dev/null 2>&1
  done <<'EOF'
conv_fprop 7x7/2 3->64 @224|3,224,64,7,2,3|s|conv_tma_fwd
conv_fprop 1x1/1 64->256 @56|64,56,256,1,1,0|s|conv_tma_fwd
conv_fprop 3x3/1 64->64 @56|64,56,64,3,1,1|s|conv_tma_fwd
conv_fprop 3x3/1 256->256 @14|256,14,256,3,1,1|s|conv_tma_fwd
conv_dgrad 1x1/1 256->64 @56|256,56,64,1,1,0|d|conv_tma_fwd
conv_wgrad 3x3/1 256->256 @14|256,14,256,3,1,1|w|conv_tma_wgrad
conv_wgrad 1x1/1 64->256 @56|64,56,256,1,1,0|w|conv_tma_wgrad
conv_wgrad 3x3/1 64->64 @56|64,56,64,3,1,1|w|conv_tma_wgrad
EOF
fi
find $OUT -size +8M -delete
ls -la $OUT | head -40
