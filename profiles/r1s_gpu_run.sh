set -x
mkdir -p gpurun_out
PITA_DIV_MODE=tf32 timeout 120 python bench.py --workload lj55 --particles 65536 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1s_bench_lj55_64k_tf32mode.json 2>&1
cat gpurun_out/r1s_bench_lj55_64k_tf32mode.json | cut -c1-600
PITA_DIV_MODE=tf32 timeout 100 python bench.py --workload lj13 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1s_bench_lj13_tf32mode.json 2>&1
cat gpurun_out/r1s_bench_lj13_tf32mode.json | cut -c1-600
