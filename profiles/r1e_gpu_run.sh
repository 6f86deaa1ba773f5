set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r1e_pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r1e_pytest.log
timeout 300 python profiles/err_by_mode.py > gpurun_out/r1e_err_by_mode.jsonl 2>&1; cat gpurun_out/r1e_err_by_mode.jsonl
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r1e_bench_lj13.json 2> gpurun_out/r1e_bench_lj13.err; echo "bench rc=$?"
cat gpurun_out/r1e_bench_lj13.json; tail -3 gpurun_out/r1e_bench_lj13.err
PITA_DIV_MODE=tf32 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1e_bench_lj13_tf32.json 2>&1
cat gpurun_out/r1e_bench_lj13_tf32.json
timeout 400 python bench.py --workload lj55 --particles 65536 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1e_bench_lj55_64k.json 2>&1
cat gpurun_out/r1e_bench_lj55_64k.json
ls -la gpurun_out | head -50
