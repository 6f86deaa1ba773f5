#!/bin/bash
# round 2, call r: ncu --set full of the blocked LJ-55 kernel after the staging fix
mkdir -p gpurun_out
PITA_LJ_CFG=7 timeout 600 ncu --set full --clock-control none --import-source on -k regex:lj_ -s 2 -c 1 -o gpurun_out/r2r_lj55_blocked -f python profiles/run_lj.py 1048576 > gpurun_out/r2r_ncu.log 2>&1
tail -2 gpurun_out/r2r_ncu.log
