#!/bin/bash
# round 2, call ai: racecheck of the Laplacian kernel on one LJ-13 particle (the two-size run of call ah did not finish in 20 min)
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python profiles/run_small.py lap13 > gpurun_out/r2ai_racecheck_lap13.txt 2>&1; echo "racecheck lap13 rc=$?"; tail -2 gpurun_out/r2ai_racecheck_lap13.txt
