#!/bin/bash
# round 2, GPU call 11 (1 GPU): regression after the render-band / PDL / fused-halo changes; PDL on vs off
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c11_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c11_pytest.log
for i in 1 2; do
python bench.py --steps 200 --warmup 10 --quick > gpurun_out/c11_bench_pdl_$i.json 2>/dev/null
FLUID_PDL=0 python bench.py --steps 200 --warmup 10 --quick > gpurun_out/c11_bench_nopdl_$i.json 2>/dev/null
done
FLUID_TB_STAGE=tma python bench.py --steps 200 --warmup 10 --quick > gpurun_out/c11_bench_tma.json 2>/dev/null
tail -6 gpurun_out/c11_pytest.log
python - <<'P'
import json
for n in ("pdl_1", "nopdl_1", "pdl_2", "nopdl_2", "tma"):
    try:
        d = json.load(open(f"gpurun_out/c11_bench_{n}.json"))
        print(n, "ms/step", round(d["ms_per_step"], 4), "T updates/s", round(d["value"] / 1e12, 3))
    except Exception as e:
        print(n, "failed", e)
P
