#!/bin/bash
# round 2, call m (2 GPUs): world-size-2 NCCL test of the sharded resampler + loop, bench at N=2 (weak, strong), aldp22 at N=2
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/r2m_gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r2m_pytest_multi.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29711 bench.py --gpus 2 --steps 2 --warmup 3 --particles 65536 --no-cpu-baseline --no-gpu-baseline \
   > gpurun_out/r2m_bench_lj55_n2_weak64k.json 2> gpurun_out/r2m_bench_lj55_n2_weak64k.err
tail -c 1500 gpurun_out/r2m_bench_lj55_n2_weak64k.json; tail -3 gpurun_out/r2m_bench_lj55_n2_weak64k.err
timeout 900 $TR --master-port 29712 bench.py --gpus 2 --steps 2 --warmup 3 --scaling strong --total-particles 131072 --no-cpu-baseline --no-gpu-baseline \
   > gpurun_out/r2m_bench_lj55_n2_strong128k.json 2> gpurun_out/r2m_bench_lj55_n2_strong128k.err
tail -c 600 gpurun_out/r2m_bench_lj55_n2_strong128k.json; tail -3 gpurun_out/r2m_bench_lj55_n2_strong128k.err
timeout 900 $TR --master-port 29713 bench.py --gpus 2 --steps 2 --warmup 3 --workload aldp22 --particles 8192 --no-cpu-baseline --no-gpu-baseline \
   > gpurun_out/r2m_bench_aldp22_n2.json 2> gpurun_out/r2m_bench_aldp22_n2.err
tail -c 600 gpurun_out/r2m_bench_aldp22_n2.json; tail -3 gpurun_out/r2m_bench_aldp22_n2.err
timeout 900 $TR --master-port 29714 bench.py --gpus 2 --steps 3 --warmup 3 --workload lj13 --particles 262144 --no-cpu-baseline --no-gpu-baseline \
   > gpurun_out/r2m_bench_lj13_n2.json 2> gpurun_out/r2m_bench_lj13_n2.err
tail -c 600 gpurun_out/r2m_bench_lj13_n2.json; tail -3 gpurun_out/r2m_bench_lj13_n2.err
