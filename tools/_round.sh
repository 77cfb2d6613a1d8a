O=gpurun_out/r2ze; mkdir -p $O
timeout 300 python tools/epi_decomp.py > $O/epi_decomp.txt 2>&1; cat $O/epi_decomp.txt
BCNN_B200_FWD_STAGES=2 timeout 300 python tools/epi_decomp.py > $O/epi_decomp_s2.txt 2>&1; cat $O/epi_decomp_s2.txt
