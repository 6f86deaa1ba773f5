#!/bin/bash
# round 2, call h: schedulable constant loads in the row-engine helpers: full GPU suite + per-kernel times
cd "$GRAFT_REPO_ROOT"
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.txt 2>&1; tail -5 gpurun_out/r2h_pytest.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2h_launches_lj55.csv python profiles/run_kernels.py 55 2368 2 > gpurun_out/r2h_ncu.log 2>&1
grep -E "tri_phase|energy_rows" gpurun_out/r2h_launches_lj55.csv | awk -F'","' '{print substr($5,1,40), $NF}' | head -8
