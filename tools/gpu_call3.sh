#!/bin/bash
# round 2, GPU call 3: streaming CVD / gradient, leaner advect / splat — parity, per-pass timing, ncu; Jacobi code-size experiment
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c3_pytest.log
timeout 300 python tools/tune2.py > gpurun_out/c3_tune2.txt 2>&1
timeout 120 tools/ubench/fp32_lat > gpurun_out/c3_fp32_lat.txt 2>&1
timeout 600 python bench.py --no-cpu > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'cvd_stream|gradient_stream|advect_|splat_' -c 12 -o gpurun_out/c3_step python tools/prof_jacobi.py step > gpurun_out/c3_ncu.log 2>&1
tail -4 gpurun_out/c3_pytest.log; cat gpurun_out/c3_tune2.txt; tail -12 gpurun_out/c3_fp32_lat.txt
