#!/bin/bash
# round 2, GPU call 9: TRANSPARENT display, fused halo kernel (single-GPU regression), per-pass timing + ncu of CVD / advection
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c9_pytest.log
timeout 600 python bench.py --no-cpu > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'cvd_stream|advect_velocity4|advect_dye4' -s 3 -c 3 -o gpurun_out/c9_step python tools/prof_jacobi.py step > gpurun_out/c9_ncu.log 2>&1
tail -12 gpurun_out/c9_pytest.log; tail -c 300 gpurun_out/c9_bench.err
