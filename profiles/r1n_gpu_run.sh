set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r1n_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r1n_pytest.log
ncu --set full --clock-control none --import-source on -k regex:score_div -c 1 -f -o gpurun_out/r1n_scorediv13_full python profiles/run_kernels.py 13 37888 1 > gpurun_out/r1n_ncu_sd13.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:score_div -c 1 -f -o gpurun_out/r1n_scorediv55_full python profiles/run_kernels.py 55 4736 1 > gpurun_out/r1n_ncu_sd55.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1n_launches_lj55.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --particles 8192 > gpurun_out/r1n_ncu_bench.log 2>&1
timeout 300 python bench.py --workload lj13 --steps 5 --warmup 3 > gpurun_out/r1n_bench_lj13.json 2> gpurun_out/r1n_bench_lj13.err; echo "bench13 rc=$?"
cat gpurun_out/r1n_bench_lj13.json
ls -la gpurun_out | tail -12
