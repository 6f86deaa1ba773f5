set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --tb=short > gpurun_out/r1p_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "MULTI_GPU_RESULT|passed|failed" gpurun_out/r1p_pytest.log | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1p_bench_lj55_2gpu.json 2> gpurun_out/r1p_bench_lj55_2gpu.err; echo "bench2 rc=$?"
cat gpurun_out/r1p_bench_lj55_2gpu.json; tail -3 gpurun_out/r1p_bench_lj55_2gpu.err
