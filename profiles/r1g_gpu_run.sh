set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_egnn.py tests/test_gpu_lj.py -m gpu -q --tb=short > gpurun_out/r1g_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "MULTI_GPU_RESULT|AssertionError|passed|failed" gpurun_out/r1g_pytest.log | cut -c1-700
python bench_lj.py --n 55 --kernel paired --batches 262144,1048576,4194304 > gpurun_out/r1g_bench_lj55_minb2.jsonl 2>&1
PITA_LJ_MINB=3 python bench_lj.py --n 55 --kernel paired --batches 262144,1048576,4194304 > gpurun_out/r1g_bench_lj55_minb3.jsonl 2>&1
cat gpurun_out/r1g_bench_lj55_minb2.jsonl gpurun_out/r1g_bench_lj55_minb3.jsonl
python bench_lj.py --n 13 --kernel paired --batches 4194304,16777216 > gpurun_out/r1g_bench_lj13.jsonl 2>&1; cat gpurun_out/r1g_bench_lj13.jsonl
