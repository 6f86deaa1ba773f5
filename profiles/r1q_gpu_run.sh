set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_egnn.py tests/test_gpu_sde.py -m gpu -q --tb=short > gpurun_out/r1q_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "AssertionError|passed|failed|Error" gpurun_out/r1q_pytest.log | cut -c1-400 | head -20
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1q_bench_lj13.json 2> gpurun_out/r1q_bench_lj13.err; echo "bench rc=$?"
cat gpurun_out/r1q_bench_lj13.json; tail -3 gpurun_out/r1q_bench_lj13.err
timeout 400 python bench.py --workload lj55 --particles 65536 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1q_bench_lj55_64k.json 2>&1
cat gpurun_out/r1q_bench_lj55_64k.json
