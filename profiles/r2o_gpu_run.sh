#!/bin/bash
# round 2, call o: LJ pair-chain microbenchmark (FMA-pipe occupancy vs warps/SM and chains in flight)
mkdir -p gpurun_out
timeout 300 ./profiles/ubench/lj_chain.bin 2>&1 | tee gpurun_out/r2o_ubench_lj_chain.txt
