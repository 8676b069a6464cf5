#!/bin/bash
# round 2, GPU call 7: half-float storage mode, strided 4-cell advection, CVD occupancy — parity + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c7_pytest.log
timeout 600 python bench.py --no-cpu > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err
tail -15 gpurun_out/c7_pytest.log; tail -c 400 gpurun_out/c7_bench.err
