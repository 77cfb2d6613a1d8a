bash tools/gpu_round.sh r2final5 tb
