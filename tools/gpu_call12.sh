#!/bin/bash
# round 2, GPU call 12 (1 GPU): does the slab hand-off plumbing (TbSync) cost the single-GPU kernel anything?
# A = library built from the commit before it, B = current.  Then: PDL edges inside the captured step graph.
mkdir -p gpurun_out
for i in 1 2; do
for v in A B; do
  so=""; [ $v = A ] && so=/root/repo/tools/ab/libfluid_b200_A.so
  FLUID_B200_SO=$so python bench.py --steps 200 --warmup 10 --quick > gpurun_out/c12_${v}_pdl_$i.json 2>/dev/null
  FLUID_B200_SO=$so FLUID_PDL=0 python bench.py --steps 200 --warmup 10 --quick > gpurun_out/c12_${v}_nopdl_$i.json 2>/dev/null
done
done
python bench.py --steps 200 --warmup 10 --no-cpu > gpurun_out/c12_full.json 2> gpurun_out/c12_full.err
FLUID_PDL_GRAPH=1 python bench.py --steps 200 --warmup 10 --no-cpu > gpurun_out/c12_full_pdlgraph.json 2> gpurun_out/c12_full_pdlgraph.err
python - <<'P'
import json
for n in ("A_pdl_1", "B_pdl_1", "A_nopdl_1", "B_nopdl_1", "A_pdl_2", "B_pdl_2", "A_nopdl_2", "B_nopdl_2"):
    try:
        d = json.load(open(f"gpurun_out/c12_{n}.json"))
        print(n, "ms/step", round(d["ms_per_step"], 4), "T updates/s", round(d["value"] / 1e12, 3))
    except Exception as e:
        print(n, "failed", e)
for n in ("full", "full_pdlgraph"):
    try:
        d = json.load(open(f"gpurun_out/c12_{n}.json"))
        fs = d.get("full_step", {})
        print(n, round(d["ms_per_step"], 4), {k: (v.get("ms_per_step") if isinstance(v, dict) else v) for k, v in fs.items()}, d.get("default_config_step"))
    except Exception as e:
        print(n, "failed", e)
P
tail -3 gpurun_out/c12_full_pdlgraph.err
