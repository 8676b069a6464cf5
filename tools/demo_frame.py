"""Visual parity artefact: the same seeded scenario (reference defaults: 128^2 sim, 1024^2 dye, 20
iterations, CURL 30; 10 random splats) run through the B200 library and through the CPU oracle,
rendered with render() (SHADING, black background) and written as PNGs.
    python tools/demo_frame.py [steps] [outdir]
Writes <outdir>/demo_gpu_<steps>.png, demo_oracle_<steps>.png and prints the 8-bit image difference."""
import os
import sys

import numpy as np
from PIL import Image

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import webgl_fluid_simulation_b200 as pkg  # noqa: E402
from oracle import oracle as O  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out"
os.makedirs(out, exist_ok=True)
SIM, DYE, VIEW = 128, 1024, 512

rs = np.random.RandomState(2024)
rnd = lambda: float(rs.random_sample())
sim = pkg.FluidSimulation({"SIM_RESOLUTION": SIM, "DYE_RESOLUTION": DYE}, 1024, 1024, random=rnd)
ref = O.OracleSim(SIM, SIM, DYE, DYE)
rs2 = np.random.RandomState(2024)
rnd2 = lambda: float(rs2.random_sample())
for _ in range(10):                                    # multipleSplats(10), S:1427-1439, same random stream
    c = pkg.HSVtoRGB(rnd2(), 1.0, 1.0)
    col = [c[k] * 0.15 * 10.0 for k in "rgb"]
    x, y = rnd2(), rnd2()
    dx, dy = 1000 * (rnd2() - 0.5), 1000 * (rnd2() - 0.5)
    ref.splat(x, y, dx, dy, *col)
sim.multipleSplats(10)
for _ in range(steps):
    sim.step(0.016666)
    ref.step(0.016666)
g = sim.textureToCanvas(sim.render(VIEW, VIEW))
o = sim.textureToCanvas(O.display(ref.dye, VIEW, VIEW, True, (0.0, 0.0, 0.0)))
Image.fromarray(g[..., :3]).save(os.path.join(out, f"demo_gpu_{steps}.png"))
Image.fromarray(o[..., :3]).save(os.path.join(out, f"demo_oracle_{steps}.png"))
d = np.abs(g[..., :3].astype(int) - o[..., :3].astype(int))
rel = np.abs(sim.readField("dye") - ref.dye).max() / np.abs(ref.dye).max()
print(f"steps={steps} view={VIEW}x{VIEW}: max 8-bit difference {d.max()}, pixels differing {int((d > 0).any(-1).sum())} of {VIEW * VIEW}, "
      f"dye max-rel {rel:.2e}")
sim.close()
