#!/bin/bash
# round 2, GPU call 2: generation-7 Jacobi — parity (both stagings), tuning sweep, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c2_pytest.log
FLUID_TB_STAGE=ldgsts timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "jacobi or 4096 or subnormal" > gpurun_out/c2_pytest_ldgsts.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c2_pytest_ldgsts.log
timeout 900 python tools/tune2.py > gpurun_out/c2_tune2.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:jacobi_tb -s 6 -c 1 -o gpurun_out/c2_tb7_tma python tools/prof_jacobi.py > gpurun_out/c2_ncu.log 2>&1
FLUID_TB_STAGE=ldgsts timeout 300 ncu --set full --clock-control none -k regex:jacobi_tb -s 6 -c 1 -o gpurun_out/c2_tb7_ldgsts python tools/prof_jacobi.py >> gpurun_out/c2_ncu.log 2>&1
tail -4 gpurun_out/c2_pytest.log; tail -3 gpurun_out/c2_pytest_ldgsts.log; cat gpurun_out/c2_tune2.txt
