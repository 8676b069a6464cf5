#!/bin/bash
# round 2, GPU call 10 (2 GPUs): fused halo kernel + PDL + 64-row ghost zone: slab parity, N=1/2 bench, halo timing
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
{
for m in p2p nccl; do
  echo "halo=$m"; FLUID_HALO=$m SLAB_W=4096 SLAB_H=1024 SLAB_WD=4096 SLAB_HD=1024 SLAB_ITERS=50 timeout 180 $R --nproc-per-node=2 --master-port 29751 tools/slab_check.py 2>&1 | grep -E "SLAB_CHECK|rror" | head -3
done
echo "fast splat at the slab edge (2000 texel/s)"; SLAB_FAST=2000 SLAB_W=1024 SLAB_H=1024 SLAB_WD=1024 SLAB_HD=1024 SLAB_ITERS=20 timeout 180 $R --nproc-per-node=2 --master-port 29752 tools/slab_check.py 2>&1 | grep -E "SLAB_CHECK|rror" | head -3
echo "resize on slabs (512x768/1024x1536 -> 640x960/1280x1920)"; SLAB_RESIZE=640,960,1280,1920 timeout 180 $R --nproc-per-node=2 --master-port 29754 tools/slab_check.py 2>&1 | grep -E "SLAB_CHECK|rror" | head -3
echo "resize on slabs, NCCL"; FLUID_HALO=nccl SLAB_RESIZE=384,576,512,768 timeout 180 $R --nproc-per-node=2 --master-port 29755 tools/slab_check.py 2>&1 | grep -E "SLAB_CHECK|rror" | head -3
echo "mirror on (last launch stores the neighbours' ghost rows)"; FLUID_HALO_MIRROR=1 SLAB_W=4096 SLAB_H=1024 SLAB_WD=4096 SLAB_HD=1024 SLAB_ITERS=50 timeout 180 $R --nproc-per-node=2 --master-port 29756 tools/slab_check.py 2>&1 | grep -E "SLAB_CHECK|rror" | head -3
echo "mirror on, 20 iterations (2 launches)"; FLUID_HALO_MIRROR=1 SLAB_W=1024 SLAB_H=1024 SLAB_WD=1024 SLAB_HD=1024 SLAB_ITERS=20 timeout 180 $R --nproc-per-node=2 --master-port 29757 tools/slab_check.py 2>&1 | grep -E "SLAB_CHECK|rror" | head -3
echo "overlap on"; FLUID_HALO_OVERLAP=1 SLAB_W=4096 SLAB_H=1024 SLAB_WD=4096 SLAB_HD=1024 SLAB_ITERS=50 timeout 180 $R --nproc-per-node=2 --master-port 29753 tools/slab_check.py 2>&1 | grep -E "SLAB_CHECK|rror" | head -3
} > gpurun_out/c10_slab_check.log 2>&1
python bench.py --steps 200 --warmup 10 --quick > gpurun_out/c10_bench1.json 2>/dev/null
FLUID_PDL=0 python bench.py --steps 200 --warmup 10 --quick > gpurun_out/c10_bench1_nopdl.json 2>/dev/null
FLUID_DEBUG_HALO_TIMING=1 timeout 300 $R --nproc-per-node=2 --master-port 29761 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/c10_bench2.log 2>&1
grep -E "^\{" gpurun_out/c10_bench2.log > gpurun_out/c10_bench2.json
FLUID_HALO_OVERLAP=1 timeout 300 $R --nproc-per-node=2 --master-port 29762 bench.py --gpus 2 --steps 200 --warmup 10 --quick 2>&1 | grep -E "^\{" > gpurun_out/c10_bench2_overlap.json
FLUID_HALO_MIRROR=1 FLUID_DEBUG_HALO_TIMING=1 timeout 300 $R --nproc-per-node=2 --master-port 29764 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/c10_bench2_mirror.log 2>&1
grep -E "^\{" gpurun_out/c10_bench2_mirror.log > gpurun_out/c10_bench2_mirror.json
FLUID_HALO=nccl timeout 300 $R --nproc-per-node=2 --master-port 29763 bench.py --gpus 2 --steps 200 --warmup 10 --quick 2>&1 | grep -E "^\{" > gpurun_out/c10_bench2_nccl.json
cat gpurun_out/c10_slab_check.log
python - <<'P'
import json
for n in ("1", "1_nopdl", "2", "2_mirror", "2_overlap", "2_nccl"):
    try:
        d = json.load(open(f"gpurun_out/c10_bench{n}.json"))
        print(n, "ms/step", round(d["ms_per_step"], 4), "T updates/s", round(d["value"] / 1e12, 3), d.get("parity"), d.get("strong", {}).get("ms_per_step"))
    except Exception as e:
        print(n, "failed", e)
P
grep -h "halo rank" gpurun_out/c10_bench2.log gpurun_out/c10_bench2_mirror.log | head -8
