#!/bin/bash
# round 2, call ak (2 GPUs): final-state regression of the N > 1 path: NCCL test, LJ-55 and ALDP-22 bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2ak_pytest_multi.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29731 bench.py --gpus 2 --steps 2 --warmup 3 --particles 65536 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2ak_bench_lj55_n2.json 2> gpurun_out/r2ak_bench_lj55_n2.err
tail -c 300 gpurun_out/r2ak_bench_lj55_n2.json; tail -2 gpurun_out/r2ak_bench_lj55_n2.err
timeout 600 $TR --master-port 29732 bench.py --gpus 2 --steps 2 --warmup 3 --workload aldp22 --particles 8192 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2ak_bench_aldp22_n2.json 2> gpurun_out/r2ak_bench_aldp22_n2.err
tail -c 300 gpurun_out/r2ak_bench_aldp22_n2.json; tail -2 gpurun_out/r2ak_bench_aldp22_n2.err
