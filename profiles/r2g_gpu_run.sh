#!/bin/bash
# round 2, call g: ncu --set full of the current phase B (after packed fp32 / schedulable loads / 3-stage ring)
cd "$GRAFT_REPO_ROOT"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tri_phase_b -s 1 -c 1 -o gpurun_out/r2g_trib55 -f python profiles/run_kernels.py 55 592 2 > gpurun_out/r2g_ncu.log 2>&1; tail -2 gpurun_out/r2g_ncu.log
