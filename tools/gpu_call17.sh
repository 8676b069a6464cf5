#!/bin/bash
# round 2, final single-GPU call: full GPU test suite, the driver's bench command, its ncu launch list, one
# ncu --set full capture of the Jacobi kernel and of the step kernels, memcheck of a small step
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c17_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c17_pytest.log
timeout 900 python bench.py > gpurun_out/c17_bench.json 2> gpurun_out/c17_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c17_bench_reference.json 2> gpurun_out/c17_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c17_launches_bench.csv python bench.py --steps 20 --warmup 3 --quick > gpurun_out/c17_ncu_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/c17_launches_step.csv python tools/prof_jacobi.py step > gpurun_out/c17_ncu_step.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'jacobi_tb' -s 6 -c 2 -o gpurun_out/c17_jacobi python tools/prof_jacobi.py > gpurun_out/c17_ncu_jacobi.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'cvd_stream|gradient_stream|advect_velocity4|advect_dye4' -s 4 -c 4 -o gpurun_out/c17_step python tools/prof_jacobi.py step > gpurun_out/c17_ncu_stepk.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize.py > gpurun_out/c17_memcheck.txt 2>&1
tail -5 gpurun_out/c17_pytest.log; tail -c 400 gpurun_out/c17_bench.err; tail -3 gpurun_out/c17_memcheck.txt
python tools/show_step.py gpurun_out/c17_bench.json
