bash tools/gpu_round.sh r2final3 tb
