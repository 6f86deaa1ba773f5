#!/bin/bash
# round 2, call f: timing experiments on phase B (what bounds the item period?): skip waits / D read-back / operand store / MMAs
cd "$GRAFT_REPO_ROOT"
for d in 0 8 16 24 31; do
PITA_TRI_DEBUG=$d timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:tri_phase_b -c 3 --csv --log-file gpurun_out/r2f_dbg$d.csv python profiles/run_kernels.py 55 1184 1 > /dev/null 2>&1
echo "dbg=$d $(grep tri_phase_b gpurun_out/r2f_dbg$d.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')"
done
