"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool memcheck python tools/sanitize.py
Exercises every kernel on grids that hit the wall / mirror / ragged-tile paths."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import webgl_fluid_simulation_b200 as pkg  # noqa: E402

for (W, H, Wd, Hd, flags, jb) in [(132, 45, 200, 70, 0, 5), (64, 64, 128, 128, pkg.FLAG_NO_GRAPH, 10),
                                  (36, 20, 54, 30, pkg.FLAG_UNFUSED, 3), (256, 130, 256, 130, 0, 12)]:
    rs = np.random.RandomState(3)
    s = pkg.FluidSimulation({"PRESSURE_ITERATIONS": 23}, W, H, flags=flags, jacobi_block=jb,
                            sizes=(W, H, Wd, Hd), random=rs.random_sample)
    s.multipleSplats(3)
    for _ in range(3):
        s.step(0.016666)
    s.config["SIM_RESOLUTION"] = 48
    for n in ("velocity", "dye", "pressure", "divergence", "curl"):
        assert np.isfinite(s.readField(n)).all(), n
    s.close()
for stage in ("tma",):
    os.environ["FLUID_TB_STAGE"] = stage
print("sanitize run complete")
