set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:score_div -c 1 -f -o gpurun_out/r1j_scorediv13_full python profiles/run_kernels.py 13 37888 1 > gpurun_out/r1j_ncu_sd13.log 2>&1
ncu -i gpurun_out/r1j_scorediv13_full.ncu-rep --page raw --csv > gpurun_out/r1j_scorediv13_raw.csv 2>/dev/null
tail -3 gpurun_out/r1j_ncu_sd13.log
ls -la gpurun_out
