#!/bin/bash
# round 2, call x (4 GPUs): the driver's own multi-GPU invocation of bench.py (default workload, weak scaling) + reference arm
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/r2x_gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29721 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/r2x_bench_lj55_n4.json 2> gpurun_out/r2x_bench_lj55_n4.err
tail -c 900 gpurun_out/r2x_bench_lj55_n4.json; tail -3 gpurun_out/r2x_bench_lj55_n4.err
timeout 600 $TR --master-port 29722 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 > gpurun_out/r2x_bench_reference_n4.json 2> gpurun_out/r2x_bench_reference_n4.err
tail -c 300 gpurun_out/r2x_bench_reference_n4.json; tail -2 gpurun_out/r2x_bench_reference_n4.err
