#!/bin/bash
# round 2, call e: helper warps at low warp ids + suspend hints; timing of the two phases (launch list) and the bench
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_gpu_egnn.py -x -q -k "bilinear" > gpurun_out/r2e_pytest.txt 2>&1; tail -3 gpurun_out/r2e_pytest.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2e_launches_lj55.csv python profiles/run_kernels.py 55 2368 2 > gpurun_out/r2e_ncu.log 2>&1
grep -E "tri_phase|energy_rows" gpurun_out/r2e_launches_lj55.csv | awk -F'","' '{print $5, $(NF-1), $NF}' | head -20
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2e_bench_lj55.json 2> gpurun_out/r2e_bench_lj55.err; tail -c 2500 gpurun_out/r2e_bench_lj55.json; tail -5 gpurun_out/r2e_bench_lj55.err
