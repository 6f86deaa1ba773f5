#!/bin/bash
# round 2, call p: ncu --set full of the two LJ-55 kernels (why do three different mappings stop at 0.5 of the FP32 peak?);
# compute-sanitizer memcheck + racecheck over every kernel on small batches
mkdir -p gpurun_out
for cfg in 0 7; do
  PITA_LJ_CFG=$cfg timeout 600 ncu --set full --clock-control none --import-source on -k regex:lj_ -s 2 -c 1 -o gpurun_out/r2p_lj55_cfg$cfg -f python profiles/run_lj.py 262144 > gpurun_out/r2p_ncu_lj_cfg$cfg.log 2>&1
  tail -2 gpurun_out/r2p_ncu_lj_cfg$cfg.log
done
for part in lj egnn13 egnn55 ad2 loop; do
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python profiles/run_small.py $part > gpurun_out/r2p_memcheck_$part.txt 2>&1; echo "memcheck $part rc=$?"; tail -3 gpurun_out/r2p_memcheck_$part.txt
done
for part in lj egnn13 ad2 loop; do
  timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python profiles/run_small.py $part > gpurun_out/r2p_racecheck_$part.txt 2>&1; echo "racecheck $part rc=$?"; tail -3 gpurun_out/r2p_racecheck_$part.txt
done
