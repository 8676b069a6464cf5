#!/bin/bash
# round 2, GPU call 1: tests, fp32 pipe micro-benchmark, ncu baseline of the non-Jacobi passes, bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
timeout 120 tools/ubench/fp32_pipe > gpurun_out/c1_fp32_pipe.txt 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'curl_vorticity|gradient_subtract|advect_|splat_' -c 12 -o gpurun_out/c1_step_base python tools/prof_jacobi.py step > gpurun_out/c1_ncu.log 2>&1
timeout 600 python bench.py > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
tail -3 gpurun_out/c1_pytest.log; tail -5 gpurun_out/c1_fp32_pipe.txt; tail -c 600 gpurun_out/c1_bench.json
