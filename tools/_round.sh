bash tools/gpu_round.sh r2zb pl
