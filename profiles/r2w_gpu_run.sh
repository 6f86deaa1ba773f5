#!/bin/bash
# round 2, call w: the three bench workloads + reference arm, launch list, ncu --set full of the engine kernels summarised ON the box
# (the reports themselves exceed the 64 MiB return limit and are deleted after summarising)
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2w_bench_lj55.json 2> gpurun_out/r2w_bench_lj55.err; tail -c 600 gpurun_out/r2w_bench_lj55.json; tail -3 gpurun_out/r2w_bench_lj55.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2w_bench_reference_lj55.json 2> gpurun_out/r2w_bench_reference_lj55.err
timeout 600 python bench.py --workload lj13 > gpurun_out/r2w_bench_lj13.json 2> gpurun_out/r2w_bench_lj13.err; tail -3 gpurun_out/r2w_bench_lj13.err
timeout 600 python bench.py --workload aldp22 > gpurun_out/r2w_bench_aldp22.json 2> gpurun_out/r2w_bench_aldp22.err; tail -3 gpurun_out/r2w_bench_aldp22.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2w_launches_lj55.csv python bench.py --particles 16384 --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2w_ncu_bench.log 2>&1
tail -2 gpurun_out/r2w_ncu_bench.log
timeout 900 ncu --set full --clock-control none -k regex:"tri_phase|egnn_energy_rows" -s 3 -c 3 -o /tmp/r2w_engine55 -f python profiles/run_kernels.py 55 592 2 > gpurun_out/r2w_ncu55.log 2>&1; tail -1 gpurun_out/r2w_ncu55.log
timeout 900 ncu --set full --clock-control none -k regex:"tri_phase|egnn_energy_rows" -s 3 -c 3 -o /tmp/r2w_engine13 -f python profiles/run_kernels.py 13 2664 2 > gpurun_out/r2w_ncu13.log 2>&1; tail -1 gpurun_out/r2w_ncu13.log
python profiles/ncu_summary.py /tmp/r2w_engine55.ncu-rep > gpurun_out/r2w_ncu_engine55.txt
python profiles/ncu_summary.py /tmp/r2w_engine13.ncu-rep > gpurun_out/r2w_ncu_engine13.txt
cp profiles/r2_ncu_summary.json /tmp/old_summary.json
python profiles/ncu_to_json.py n55:592:/tmp/r2w_engine55.ncu-rep n13:2664:/tmp/r2w_engine13.ncu-rep > /dev/null && cp profiles/r2_ncu_summary.json gpurun_out/r2_ncu_summary.json
ls -la gpurun_out | head -30; du -sh gpurun_out
