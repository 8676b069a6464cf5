#!/bin/bash
# round 2, GPU call 6: scaled branch-free CVD, 4-cell advection, post-FX — parity + per-pass timing; operand-bandwidth micro-benchmark
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c6_pytest.log
timeout 120 tools/ubench/fp32_pipe > gpurun_out/c6_fp32_pipe.txt 2>&1
timeout 600 python bench.py --no-cpu > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'cvd_stream|advect_' -c 6 -o gpurun_out/c6_step python tools/prof_jacobi.py step > gpurun_out/c6_ncu.log 2>&1
tail -15 gpurun_out/c6_pytest.log; grep -E "distinct|FADD2 \(|FADD r,r,r" gpurun_out/c6_fp32_pipe.txt
