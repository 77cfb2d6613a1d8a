# scratch entry point of `gpurun -- 'bash tools/_round.sh'`: the round's evidence in one call
bash tools/gpu_round.sh rX tbspln
