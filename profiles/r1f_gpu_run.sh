set -x
mkdir -p gpurun_out
nvidia-smi -L
./profiles/ubench/roundtrip.bin > gpurun_out/r1f_ubench_roundtrip.txt 2>&1; cat gpurun_out/r1f_ubench_roundtrip.txt
timeout 300 python profiles/err_by_mode.py > gpurun_out/r1f_err_by_mode.jsonl 2>&1; cat gpurun_out/r1f_err_by_mode.jsonl
timeout 600 python -m pytest tests/test_gpu_egnn.py tests/test_gpu_multi.py tests/test_gpu_sde.py -m gpu -q --tb=short > gpurun_out/r1f_pytest.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/r1f_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1f_bench_lj13_2gpu.json 2> gpurun_out/r1f_bench_lj13_2gpu.err; echo "bench2 rc=$?"
cat gpurun_out/r1f_bench_lj13_2gpu.json; tail -5 gpurun_out/r1f_bench_lj13_2gpu.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1f_bench_lj13_1gpu.json 2>&1; cat gpurun_out/r1f_bench_lj13_1gpu.json
