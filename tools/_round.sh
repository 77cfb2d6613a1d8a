bash tools/gpu_round.sh r2zk tb
