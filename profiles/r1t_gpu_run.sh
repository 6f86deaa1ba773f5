set -x
mkdir -p gpurun_out
timeout 100 python profiles/err_by_mode.py > gpurun_out/r1t_err_by_mode.jsonl 2>&1; grep 3xtf32 gpurun_out/r1t_err_by_mode.jsonl
timeout 100 python bench.py --workload lj55 --particles 65536 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1t_bench_lj55_64k.json 2>&1
cut -c1-400 gpurun_out/r1t_bench_lj55_64k.json
