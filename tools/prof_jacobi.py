"""Minimal driver for ncu: a few 4096^2 pressure solves (7 blocked launches each) and one full step.
usage: ncu ... python tools/prof_jacobi.py [naive|step]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import webgl_fluid_simulation_b200 as pkg  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "tb"
W = 4096
rng = np.random.default_rng(0)
if mode == "step":
    s = pkg.FluidSimulation({"SIM_RESOLUTION": W, "DYE_RESOLUTION": W, "PRESSURE_ITERATIONS": 50}, 1024, 1024,
                            random=np.random.RandomState(1234).random_sample)
    s.multipleSplats(16)                       # the same field bench.py's full_step leg times
    for _ in range(3):
        s.step(0.016666)
    s.sync()
else:
    flags = pkg.FLAG_NAIVE_JACOBI if mode == "naive" else 0
    s = pkg.FluidSimulation({"SIM_RESOLUTION": W, "DYE_RESOLUTION": 64, "PRESSURE_ITERATIONS": 50}, 1024, 1024,
                            flags=flags, jacobi_block=1 if mode == "naive" else 0)
    s.writeField("pressure", rng.standard_normal((W, W)).astype(np.float32))
    s.writeField("divergence", rng.uniform(-1, 1, (W, W)).astype(np.float32))
    for _ in range(3):
        s.pass_("pressure_solve")
    s.sync()
s.close()
