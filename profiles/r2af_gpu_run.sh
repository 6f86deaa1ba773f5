#!/bin/bash
# round 2, call af: phase-B experiment — how much of the item period is the finish stage (upper bound of build / finish warp specialisation)
mkdir -p gpurun_out
echo "baseline" | tee gpurun_out/r2af_trib_finish_experiment.txt
timeout 60 python profiles/time_scorediv.py 2>&1 | tail -1 | tee -a gpurun_out/r2af_trib_finish_experiment.txt
cp pita_b200/libpita_b200.so /tmp/lib_orig.so; cp gpurun_exp1.so pita_b200/libpita_b200.so
echo "generic finish skipped (results wrong; timing only)" | tee -a gpurun_out/r2af_trib_finish_experiment.txt
timeout 60 python profiles/time_scorediv.py 2>&1 | tail -1 | tee -a gpurun_out/r2af_trib_finish_experiment.txt
cp /tmp/lib_orig.so pita_b200/libpita_b200.so
