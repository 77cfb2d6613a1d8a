bash tools/gpu_round.sh r2final4 pl
