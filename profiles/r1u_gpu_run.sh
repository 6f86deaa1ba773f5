mkdir -p gpurun_out
timeout 95 python -m pytest tests/test_gpu_egnn.py tests/test_gpu_sde.py -x -q -m gpu > gpurun_out/r1u_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r1u_pytest.log | cut -c1-300
