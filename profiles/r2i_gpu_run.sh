#!/bin/bash
# round 2, call i: remaining GPU tests after the pin fix; LJ kernel after the streamed-pair prefetch (bench_lj sweep)
cd "$GRAFT_REPO_ROOT"
timeout 1800 python -m pytest tests/test_gpu_sde.py tests/test_gpu_lj.py tests/test_gpu_resample.py tests/test_gpu_umma.py tests/test_gpu_multi.py -x -q > gpurun_out/r2i_pytest.txt 2>&1; tail -4 gpurun_out/r2i_pytest.txt
timeout 600 python bench_lj.py --n 55 --batches 16384,262144,1048576,4194304 > gpurun_out/r2i_bench_lj55.jsonl 2>&1; cat gpurun_out/r2i_bench_lj55.jsonl | cut -c1-330
timeout 600 python bench_lj.py --n 13 --batches 1048576,4194304 > gpurun_out/r2i_bench_lj13.jsonl 2>&1; cat gpurun_out/r2i_bench_lj13.jsonl | cut -c1-330
