#!/bin/bash
# round 2, GPU call 5: branch-free CVD / prefetching gradient parity + timing; Jacobi occupancy variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c5_pytest.log
timeout 600 python tools/tune2.py > gpurun_out/c5_tune2.txt 2>&1
timeout 600 python bench.py --no-cpu > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'cvd_stream|gradient_stream' -c 4 -o gpurun_out/c5_step python tools/prof_jacobi.py step > gpurun_out/c5_ncu.log 2>&1
tail -4 gpurun_out/c5_pytest.log; cat gpurun_out/c5_tune2.txt
