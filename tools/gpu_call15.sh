#!/bin/bash
# round 2, GPU call 15 (1 GPU): cvd_stream_kernel occupancy / ring variants (full-step per-kernel times)
mkdir -p gpurun_out
for v in base cvd_mb4 cvd_mb6 cvd_ring8 cvd_mb6r4; do
  so=""; [ $v != base ] && so=/root/repo/tools/ab/libfluid_b200_$v.so
  FLUID_B200_SO=$so python bench.py --steps 100 --warmup 10 --no-cpu > gpurun_out/c15_$v.json 2> gpurun_out/c15_$v.err
done
for i in 1 2; do
python bench.py --steps 200 --warmup 10 --quick > gpurun_out/c15_plain_$i.json 2>/dev/null
FLUID_TB_FORCE_SYNC=1 python bench.py --steps 200 --warmup 10 --quick > gpurun_out/c15_forcesync_$i.json 2>/dev/null
done
timeout 600 python -m pytest tests -m gpu -x -q -k "host_pressure or update_loop or config1" > gpurun_out/c15_pytest.log 2>&1; tail -3 gpurun_out/c15_pytest.log
FLUID_E2E_BANDS=1 python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/c15_e2e_onepiece.json 2>/dev/null
for nb in 4 16; do FLUID_E2E_BANDS=$nb python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/c15_e2e_bands$nb.json 2>/dev/null; done
python - <<'P'
import json
for n in ("plain_1", "forcesync_1", "plain_2", "forcesync_2"):
    d = json.load(open(f"gpurun_out/c15_{n}.json")); print(n, round(d["ms_per_step"], 4))
for n in ("base", "e2e_onepiece", "e2e_bands4", "e2e_bands16"):
    d = json.load(open(f"gpurun_out/c15_{n}.json")); print(n, "e2e ms/step", round(d["e2e"]["ms_per_step"], 3), "G updates/s", round(d["e2e"]["value"] / 1e9, 1))
P
python tools/show_step.py gpurun_out/c15_base.json gpurun_out/c15_cvd_mb4.json gpurun_out/c15_cvd_mb6.json gpurun_out/c15_cvd_ring8.json gpurun_out/c15_cvd_mb6r4.json
