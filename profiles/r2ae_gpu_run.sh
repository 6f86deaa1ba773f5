#!/bin/bash
# round 2, call ae: the bench-size property test of the default engine
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_egnn.py -q -m gpu -k "bench_size" 2>&1 | tail -15 | tee gpurun_out/r2ae_pytest_bench_size.txt
