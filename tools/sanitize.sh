#!/bin/bash
# compute-sanitizer over the kernel-level GPU tests: memcheck (out-of-bounds /
# misaligned accesses, leaks) and racecheck (shared-memory hazards) on the tcgen05 / TMA convolution
# kernels, the NHWC batch-norm / pooling kernels and the FP32 reductions.
#   bash tools/sanitize.sh OUTDIR [seconds per tool] ["test files"]
# Writes OUTDIR/memcheck.log, OUTDIR/racecheck.log and OUTDIR/summary.txt (copy into profiles/).
OUT=${1:-gpurun_out/sanitize}
LIMIT=${2:-420}
mkdir -p $OUT
FILES=${3:-tests/test_nhwc_bf16_gpu.py tests/test_kernels_gpu.py}
export BCNN_B200_CONV_MATH=fp32
for TOOL in memcheck racecheck; do
  timeout $LIMIT compute-sanitizer --tool $TOOL --error-exitcode 1 \
      python -m pytest $FILES -q -m gpu -x > $OUT/$TOOL.log 2>&1
  echo "$TOOL rc=$?" >> $OUT/$TOOL.log
done
{
  for TOOL in memcheck racecheck; do
    echo "== $TOOL"
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=" $OUT/$TOOL.log | tail -5
    grep -c "=========.*(Invalid|Race|hazard)" $OUT/$TOOL.log | sed 's/^/error lines: /'
  done
} > $OUT/summary.txt
cat $OUT/summary.txt
