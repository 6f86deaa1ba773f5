set -x
mkdir -p gpurun_out
nvidia-smi -L
./profiles/ubench/fp32_pipes.bin > gpurun_out/r1d_ubench_fp32.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r1d_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r1d_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r1d_bench_lj13.json 2> gpurun_out/r1d_bench_lj13.err; echo "bench rc=$?"
cat gpurun_out/r1d_bench_lj13.json
for k in paired ordered; do
  python bench_lj.py --n 55 --kernel $k --batches 1024,16384,262144,1048576,4194304 >> gpurun_out/r1d_bench_lj.jsonl 2>&1
  python bench_lj.py --n 13 --kernel $k --batches 16384,1048576,4194304,16777216 >> gpurun_out/r1d_bench_lj.jsonl 2>&1
done
python bench_lj.py --n 55 --batches 1024 --reps 2 --cpu >> gpurun_out/r1d_bench_lj.jsonl 2>&1
python bench_lj.py --n 13 --batches 1024 --reps 2 --cpu >> gpurun_out/r1d_bench_lj.jsonl 2>&1
cat gpurun_out/r1d_bench_lj.jsonl
ncu --set full --clock-control none --import-source on -k regex:lj_ -c 1 -f -o gpurun_out/r1d_lj55_full python bench_lj.py --n 55 --batches 262144 --reps 1 > gpurun_out/r1d_ncu_lj.log 2>&1
ncu -i gpurun_out/r1d_lj55_full.ncu-rep --page raw --csv > gpurun_out/r1d_lj55_full_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:score_div -c 1 -f -o gpurun_out/r1d_scorediv13_full python profiles/run_kernels.py 13 37888 1 > gpurun_out/r1d_ncu_sd13.log 2>&1
ncu -i gpurun_out/r1d_scorediv13_full.ncu-rep --page raw --csv > gpurun_out/r1d_scorediv13_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:energy -c 1 -f -o gpurun_out/r1d_energy13_full python profiles/run_kernels.py 13 37888 1 > gpurun_out/r1d_ncu_en13.log 2>&1
ncu -i gpurun_out/r1d_energy13_full.ncu-rep --page raw --csv > gpurun_out/r1d_energy13_raw.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1d_launches_lj13.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --particles 131072 > gpurun_out/r1d_ncu_bench.log 2>&1
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1d_bench_ref_lj13.json 2>&1
cat gpurun_out/r1d_bench_ref_lj13.json
ls -la gpurun_out; du -sh gpurun_out
