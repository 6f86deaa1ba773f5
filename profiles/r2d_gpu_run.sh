#!/bin/bash
# round 2, call d: packed-fp32 phase B + prefetching phase A: parity (egnn + post tests) and timing
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_egnn.py tests/test_gpu_post.py -x -q -k "bilinear or post or mala or descent" > gpurun_out/r2d_pytest.txt 2>&1; tail -4 gpurun_out/r2d_pytest.txt
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2d_bench_lj55.json 2> gpurun_out/r2d_bench_lj55.err; python -c "
import json; d=json.loads(open('gpurun_out/r2d_bench_lj55.json').read().strip().splitlines()[-1]); print('value', d['value'], 'ms/step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'])"; tail -3 gpurun_out/r2d_bench_lj55.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2d_launches_lj55.csv python bench.py --particles 16384 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2d_ncu_bench.log 2>&1
tail -2 gpurun_out/r2d_ncu_bench.log
