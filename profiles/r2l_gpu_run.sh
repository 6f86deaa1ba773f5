#!/bin/bash
# round 2, call l: alanine-dipeptide path: parity tests, bench line (N=1), kernel launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ad2.py -q -m gpu 2>&1 | tail -30 | tee gpurun_out/r2l_pytest_ad2.txt
timeout 900 python bench.py --workload aldp22 --steps 2 --warmup 3 > gpurun_out/r2l_bench_aldp22.json 2> gpurun_out/r2l_bench_aldp22.err
tail -c 3000 gpurun_out/r2l_bench_aldp22.json; tail -5 gpurun_out/r2l_bench_aldp22.err
