set -x
nvidia-smi -L
python -m pytest tests -m gpu -x -q > gpurun_out/r1b_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r1b_pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r1b_bench_lj13.json 2> gpurun_out/r1b_bench_lj13.err; echo "bench rc=$?"
cat gpurun_out/r1b_bench_lj13.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1b_bench_ref_lj13.json 2>&1
cat gpurun_out/r1b_bench_ref_lj13.json
python bench.py --workload lj55 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1b_bench_lj55.json 2> gpurun_out/r1b_bench_lj55.err; echo "bench55 rc=$?"
cat gpurun_out/r1b_bench_lj55.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1b_launches_lj13.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --particles 131072 > gpurun_out/r1b_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -f -o gpurun_out/r1b_full_n13 python profiles/run_kernels.py 13 37888 1 > gpurun_out/r1b_full_n13.log 2>&1
ncu --set full --clock-control none --import-source on -f -o gpurun_out/r1b_full_n55 python profiles/run_kernels.py 55 4736 1 > gpurun_out/r1b_full_n55.log 2>&1
ls -la gpurun_out
