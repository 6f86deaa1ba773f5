#!/bin/bash
# round 2, call a: first hardware run of the bilinear score/divergence engine (table-by-table check against the fp64 oracle)
cd "$GRAFT_REPO_ROOT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 300 python tests/tri_debug.py 13 20 > gpurun_out/r2a_tri13.txt 2>&1; echo "rc13=$?" >> gpurun_out/r2a_tri13.txt
timeout 300 python tests/tri_debug.py 55 5 > gpurun_out/r2a_tri55.txt 2>&1; echo "rc55=$?" >> gpurun_out/r2a_tri55.txt
tail -30 gpurun_out/r2a_tri13.txt; tail -30 gpurun_out/r2a_tri55.txt
