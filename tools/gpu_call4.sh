#!/bin/bash
# round 2, GPU call 4: Jacobi code-layout experiment (A: diag+exact, B: rowmajor+exact, C: diag, D: rowmajor; no exact)
mkdir -p gpurun_out
timeout 600 python tools/tune2.py > gpurun_out/c4_tune2.txt 2>&1
FLUID_B200_SO=$PWD/webgl_fluid_simulation_b200/libfluid_b200_A.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "jacobi or 4096 or subnormal" > gpurun_out/c4_pytest_A.log 2>&1
FLUID_B200_SO=$PWD/webgl_fluid_simulation_b200/libfluid_b200_A.so timeout 300 ncu --set full --clock-control none -k regex:jacobi_tb -s 6 -c 1 -o gpurun_out/c4_tb_A python tools/prof_jacobi.py > gpurun_out/c4_ncu.log 2>&1
FLUID_B200_SO=$PWD/webgl_fluid_simulation_b200/libfluid_b200_C.so timeout 300 ncu --set full --clock-control none -k regex:jacobi_tb -s 6 -c 1 -o gpurun_out/c4_tb_C python tools/prof_jacobi.py >> gpurun_out/c4_ncu.log 2>&1
cat gpurun_out/c4_tune2.txt; tail -3 gpurun_out/c4_pytest_A.log
