#!/bin/bash
# round 2, call ad: final state — full GPU suite, smoke(), the three bench workloads
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/r2ad_pytest_gpu.txt 2>&1; tail -4 gpurun_out/r2ad_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2ad_smoke.txt
timeout 900 python bench.py > gpurun_out/r2ad_bench_lj55.json 2> gpurun_out/r2ad_bench_lj55.err; tail -c 400 gpurun_out/r2ad_bench_lj55.json; tail -3 gpurun_out/r2ad_bench_lj55.err
timeout 600 python bench.py --workload lj13 > gpurun_out/r2ad_bench_lj13.json 2> gpurun_out/r2ad_bench_lj13.err; tail -3 gpurun_out/r2ad_bench_lj13.err
timeout 600 python bench.py --workload aldp22 > gpurun_out/r2ad_bench_aldp22.json 2> gpurun_out/r2ad_bench_aldp22.err; tail -3 gpurun_out/r2ad_bench_aldp22.err
