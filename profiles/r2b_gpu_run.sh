#!/bin/bash
# round 2, call b: bilinear engine — parity tests, first timing (bench default LJ-55 262144 and LJ-13 2^20), launch list
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_egnn.py -x -q -k "bilinear" > gpurun_out/r2b_pytest_bilinear.txt 2>&1; tail -5 gpurun_out/r2b_pytest_bilinear.txt
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2b_bench_lj55.json 2> gpurun_out/r2b_bench_lj55.err; tail -c 1500 gpurun_out/r2b_bench_lj55.json; tail -3 gpurun_out/r2b_bench_lj55.err
timeout 600 python bench.py --workload lj13 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2b_bench_lj13.json 2> gpurun_out/r2b_bench_lj13.err; tail -c 1500 gpurun_out/r2b_bench_lj13.json; tail -3 gpurun_out/r2b_bench_lj13.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2b_launches_lj55.csv python bench.py --particles 16384 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2b_ncu_bench.log 2>&1
tail -3 gpurun_out/r2b_ncu_bench.log
