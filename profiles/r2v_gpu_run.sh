#!/bin/bash
# round 2, call v: full GPU suite, smoke(), the three bench workloads + reference arm, launch list, ncu --set full of the engine
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/r2v_pytest_gpu.txt 2>&1; tail -6 gpurun_out/r2v_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2v_smoke.txt
timeout 900 python bench.py > gpurun_out/r2v_bench_lj55.json 2> gpurun_out/r2v_bench_lj55.err; tail -c 2500 gpurun_out/r2v_bench_lj55.json; tail -3 gpurun_out/r2v_bench_lj55.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2v_bench_reference_lj55.json 2> gpurun_out/r2v_bench_reference_lj55.err; tail -c 600 gpurun_out/r2v_bench_reference_lj55.json
timeout 600 python bench.py --workload lj13 > gpurun_out/r2v_bench_lj13.json 2> gpurun_out/r2v_bench_lj13.err; tail -c 1200 gpurun_out/r2v_bench_lj13.json; tail -3 gpurun_out/r2v_bench_lj13.err
timeout 600 python bench.py --workload aldp22 > gpurun_out/r2v_bench_aldp22.json 2> gpurun_out/r2v_bench_aldp22.err; tail -c 1200 gpurun_out/r2v_bench_aldp22.json; tail -3 gpurun_out/r2v_bench_aldp22.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2v_launches_lj55.csv python bench.py --particles 16384 --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2v_ncu_bench.log 2>&1
tail -2 gpurun_out/r2v_ncu_bench.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tri_phase|egnn_energy_rows" -s 3 -c 3 -o gpurun_out/r2v_engine55 -f python profiles/run_kernels.py 55 592 2 > gpurun_out/r2v_ncu55.log 2>&1; tail -1 gpurun_out/r2v_ncu55.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tri_phase|egnn_energy_rows" -s 3 -c 3 -o gpurun_out/r2v_engine13 -f python profiles/run_kernels.py 13 2664 2 > gpurun_out/r2v_ncu13.log 2>&1; tail -1 gpurun_out/r2v_ncu13.log
