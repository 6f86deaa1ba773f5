#!/bin/bash
# round 2, call aj: alanine-dipeptide bench line with the reference's inference chunk (2048)
mkdir -p gpurun_out
timeout 600 python bench.py --workload aldp22 > gpurun_out/r2aj_bench_aldp22.json 2> gpurun_out/r2aj_bench_aldp22.err; tail -c 500 gpurun_out/r2aj_bench_aldp22.json; tail -3 gpurun_out/r2aj_bench_aldp22.err
