#!/bin/bash
# round 2, call ag: the new denoiser tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_egnn.py -q -m gpu -k "denoisers" 2>&1 | tail -15 | tee gpurun_out/r2ag_pytest_denoisers.txt
