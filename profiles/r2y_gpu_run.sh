#!/bin/bash
# round 2, call y (8 GPUs): the driver's own multi-GPU invocation of bench.py (default workload, weak scaling) + reference arm
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/r2y_gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29731 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2y_bench_lj55_n8.json 2> gpurun_out/r2y_bench_lj55_n8.err
tail -c 900 gpurun_out/r2y_bench_lj55_n8.json; tail -3 gpurun_out/r2y_bench_lj55_n8.err
timeout 600 $TR --master-port 29732 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/r2y_bench_reference_n8.json 2> gpurun_out/r2y_bench_reference_n8.err
tail -c 300 gpurun_out/r2y_bench_reference_n8.json; tail -2 gpurun_out/r2y_bench_reference_n8.err
