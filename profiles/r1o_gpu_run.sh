set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sde.py -m gpu -q --tb=short > gpurun_out/r1o_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r1o_pytest.log | cut -c1-300
date
timeout 900 python bench.py > gpurun_out/r1o_bench_default_lj55_256k.json 2> gpurun_out/r1o_bench_default.err; echo "bench rc=$?"
date
cat gpurun_out/r1o_bench_default_lj55_256k.json; tail -3 gpurun_out/r1o_bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1o_bench_reference_lj55.json 2>&1
date
cat gpurun_out/r1o_bench_reference_lj55.json
