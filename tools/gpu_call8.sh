#!/bin/bash
# round 2, GPU call 8 (2 GPUs): slab parity (p2p + nccl, overlap on/off, a width that is not a multiple of 4), bench at N=2
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/c8_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c8_pytest_multi.log
{
for ov in 1 0; do for m in p2p nccl; do
  echo "overlap=$ov halo=$m"; FLUID_HALO_OVERLAP=$ov FLUID_HALO=$m SLAB_W=4096 SLAB_H=1024 SLAB_WD=4096 SLAB_HD=1024 SLAB_ITERS=50 timeout 180 $R --nproc-per-node=2 --master-port 29751 tools/slab_check.py 2>&1 | grep -E "SLAB_CHECK|rror" | head -3
done; done
echo "width 510 (not a multiple of 4: peer-memory refused, NCCL fallback)"; SLAB_W=510 SLAB_H=512 SLAB_WD=1020 SLAB_HD=1024 SLAB_ITERS=20 timeout 180 $R --nproc-per-node=2 --master-port 29752 tools/slab_check.py 2>&1 | grep -E "SLAB_CHECK|rror" | head -3
} > gpurun_out/c8_slab_check.log 2>&1
timeout 300 $R --nproc-per-node=2 --master-port 29761 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/c8_bench2.log 2>&1
grep -E "^\{" gpurun_out/c8_bench2.log > gpurun_out/c8_bench2.json
FLUID_HALO_OVERLAP=0 timeout 300 $R --nproc-per-node=2 --master-port 29762 bench.py --gpus 2 --steps 200 --warmup 10 --quick 2>&1 | grep -E "^\{" > gpurun_out/c8_bench2_nooverlap.json
python bench.py --steps 200 --warmup 10 --quick > gpurun_out/c8_bench1.json 2>/dev/null
tail -3 gpurun_out/c8_pytest_multi.log; cat gpurun_out/c8_slab_check.log
python - <<'P'
import json
for n in ("1", "2", "2_nooverlap"):
    try:
        d = json.load(open(f"gpurun_out/c8_bench{n}.json"))
        print(n, "ms/step", round(d["ms_per_step"], 4), "T updates/s", round(d["value"] / 1e12, 3), d.get("parity"), d.get("strong", {}).get("ms_per_step"))
    except Exception as e:
        print(n, "failed", e)
P
tail -5 gpurun_out/c8_bench2.log | cut -c1-300
