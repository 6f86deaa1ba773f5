set -x
mkdir -p gpurun_out
./profiles/ubench/fp32_pipes.bin > gpurun_out/r1c_ubench_fp32.txt 2>&1
python -m pytest tests/test_gpu_lj.py -q --tb=short > gpurun_out/r1c_pytest_lj.log 2>&1; echo "pytest lj rc=$?"
tail -15 gpurun_out/r1c_pytest_lj.log
for k in paired ordered; do
  python bench_lj.py --n 55 --kernel $k --batches 1024,16384,262144,1048576 >> gpurun_out/r1c_bench_lj.jsonl 2>&1
  python bench_lj.py --n 13 --kernel $k --batches 16384,1048576,4194304 >> gpurun_out/r1c_bench_lj.jsonl 2>&1
done
cat gpurun_out/r1c_bench_lj.jsonl
python -m pytest tests/test_gpu_egnn.py -q --tb=short -k "oracle_random or full_size" > gpurun_out/r1c_pytest_egnn.log 2>&1; echo "pytest egnn rc=$?"
grep -E "AssertionError|passed|failed|FAILED" gpurun_out/r1c_pytest_egnn.log | head -40
ncu --set full --clock-control none --import-source on -k regex:lj_pairs -c 1 -f -o gpurun_out/r1c_lj55_full python bench_lj.py --n 55 --batches 262144 --reps 1 > gpurun_out/r1c_ncu_lj.log 2>&1
ncu -i gpurun_out/r1c_lj55_full.ncu-rep --page raw --csv > gpurun_out/r1c_lj55_full_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:score_div -c 1 -f -o gpurun_out/r1c_scorediv13_full python profiles/run_kernels.py 13 37888 1 > gpurun_out/r1c_ncu_sd13.log 2>&1
ncu -i gpurun_out/r1c_scorediv13_full.ncu-rep --page raw --csv > gpurun_out/r1c_scorediv13_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:score_div -c 1 -f -o gpurun_out/r1c_scorediv55_full python profiles/run_kernels.py 55 4736 1 > gpurun_out/r1c_ncu_sd55.log 2>&1
ncu -i gpurun_out/r1c_scorediv55_full.ncu-rep --page raw --csv > gpurun_out/r1c_scorediv55_raw.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1c_launches_lj13.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --particles 131072 > gpurun_out/r1c_ncu_bench.log 2>&1
ls -la gpurun_out; du -sh gpurun_out
