bash tools/gpu_round.sh r2final2 tbpl
