set -x
mkdir -p gpurun_out
python -m pytest tests/ -x -q -m gpu > gpurun_out/r1r_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r1r_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1r_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r1r_smoke.log
