#!/bin/bash
# round 2, call c: full parity suite on the new default engine + ncu --set full of the two phases of the bilinear engine and of the energy kernel
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.txt 2>&1; tail -8 gpurun_out/r2c_pytest.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tri_phase -s 2 -c 2 -o gpurun_out/r2c_tri55 -f python profiles/run_kernels.py 55 592 2 > gpurun_out/r2c_ncu_tri.log 2>&1; tail -2 gpurun_out/r2c_ncu_tri.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:egnn_energy_rows -s 1 -c 1 -o gpurun_out/r2c_energy55 -f python profiles/run_kernels.py 55 592 2 > gpurun_out/r2c_ncu_en.log 2>&1; tail -2 gpurun_out/r2c_ncu_en.log
