#!/bin/bash
# round 2, GPU call 13 (2 GPUs): mirrored pressure ghost rows with acq_rel fences (the fence.sc version cost 35 us/solve)
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
{
echo "mirror on"; FLUID_HALO_MIRROR=1 SLAB_W=4096 SLAB_H=1024 SLAB_WD=4096 SLAB_HD=1024 SLAB_ITERS=50 timeout 180 $R --nproc-per-node=2 --master-port 29756 tools/slab_check.py 2>&1 | grep -E "SLAB_CHECK|rror" | head -3
echo "mirror on, 20 iterations (2 launches)"; FLUID_HALO_MIRROR=1 SLAB_W=1024 SLAB_H=1024 SLAB_WD=1024 SLAB_HD=1024 SLAB_ITERS=20 timeout 180 $R --nproc-per-node=2 --master-port 29757 tools/slab_check.py 2>&1 | grep -E "SLAB_CHECK|rror" | head -3
} > gpurun_out/c13_slab_check.log 2>&1
for i in 1 2; do
timeout 300 $R --nproc-per-node=2 --master-port 2976$i bench.py --gpus 2 --steps 200 --warmup 10 --quick 2>&1 | grep -E "^\{" > gpurun_out/c13_bench2_base_$i.json
FLUID_HALO_MIRROR=1 timeout 300 $R --nproc-per-node=2 --master-port 2977$i bench.py --gpus 2 --steps 200 --warmup 10 --quick 2>&1 | grep -E "^\{" > gpurun_out/c13_bench2_mirror_$i.json
done
FLUID_HALO_MIRROR=1 timeout 300 $R --nproc-per-node=2 --master-port 29781 bench.py --gpus 2 --steps 200 --warmup 10 --no-cpu > gpurun_out/c13_bench2_mirror_full.log 2>&1
grep -E "^\{" gpurun_out/c13_bench2_mirror_full.log > gpurun_out/c13_bench2_mirror_full.json
python bench.py --steps 200 --warmup 10 --quick > gpurun_out/c13_bench1.json 2>/dev/null
cat gpurun_out/c13_slab_check.log
python - <<'P'
import json
for n in ("1", "2_base_1", "2_mirror_1", "2_base_2", "2_mirror_2", "2_mirror_full"):
    try:
        d = json.load(open(f"gpurun_out/c13_bench{n}.json"))
        print(n, "ms/step", round(d["ms_per_step"], 4), "T updates/s", round(d["value"] / 1e12, 3), d.get("parity"), d.get("strong", {}).get("ms_per_step"))
    except Exception as e:
        print(n, "failed", e)
P
