"""Round-2 tuning sweep of the generation-7 blocked Jacobi kernel: staging (TMA boxes / LDGSTS ring) x
chunking (resident-stream cap, explicit rows per stream) at 4096^2 x 50; ms per solve.
usage (under gpurun): python tools/tune2.py > gpurun_out/tune2.txt
FLUID_TB_STAGE / FLUID_JACOBI_WARPS are read once per process, so every point runs in a child."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys
import numpy as np
sys.path.insert(0, %r)
import webgl_fluid_simulation_b200 as pkg
W = 4096; ITERS = 50
rng = np.random.default_rng(0)
p = rng.standard_normal((W, W)).astype(np.float32); d = rng.uniform(-1, 1, (W, W)).astype(np.float32)
for kb in [int(x) for x in os.environ.get("TUNE_KB", "10").split(",")]:
    s = pkg.FluidSimulation({"SIM_RESOLUTION": W, "DYE_RESOLUTION": 64, "PRESSURE_ITERATIONS": ITERS}, 1024, 1024, jacobi_block=kb)
    s.writeField("pressure", p); s.writeField("divergence", d)
    for _ in range(3): s.pass_("pressure_solve")
    s.sync(); n = 30; s.mark(0)
    for _ in range(n): s.pass_("pressure_solve")
    s.mark(1); ms = s.elapsed_ms() / n
    print(f"kb={kb:2d} {ms:8.4f} ms {W*W*ITERS/ms/1e6:9.1f} G/s", flush=True)
    s.close()
''' % ROOT

def run(tag, env):
    e = dict(os.environ); e.update(env)
    out = subprocess.run([sys.executable, "-c", CHILD], env=e, capture_output=True, text=True, timeout=300)
    for l in out.stdout.splitlines():
        print(f"{tag:48s} {l}", flush=True)
    if out.returncode:
        print(f"{tag}: FAILED {out.stderr[-400:]}", flush=True)

libs = {"default": None}
import glob
for pth in sorted(glob.glob(os.path.join(ROOT, "webgl_fluid_simulation_b200", "libfluid_b200_*.so"))):
    libs[os.path.basename(pth)[len("libfluid_b200_"):-3]] = pth
for lib, pth in libs.items():
    for stage in ("tma", "ldgsts"):
        for warps in ((0,) if "192" not in lib else (0, 8, 9)):
            env = {"FLUID_TB_STAGE": stage, "TUNE_KB": "10", "FLUID_JACOBI_WARPS": str(warps)}
            if pth: env["FLUID_B200_SO"] = pth
            run(f"lib={lib} stage={stage} warps={warps}", env)
