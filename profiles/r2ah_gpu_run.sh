#!/bin/bash
# round 2, call ah: compute-sanitizer over the kernels that changed after call p (blocked LJ, alanine-dipeptide, Laplacian) and
# the racecheck re-run of the fp32 engine after the __syncwarp fix
mkdir -p gpurun_out
for part in lj ad2 lap; do
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python profiles/run_small.py $part > gpurun_out/r2ah_memcheck_$part.txt 2>&1; echo "memcheck $part rc=$?"; tail -2 gpurun_out/r2ah_memcheck_$part.txt
done
for part in lj ad2 lap egnn13; do
  timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python profiles/run_small.py $part > gpurun_out/r2ah_racecheck_$part.txt 2>&1; echo "racecheck $part rc=$?"; tail -2 gpurun_out/r2ah_racecheck_$part.txt
done
