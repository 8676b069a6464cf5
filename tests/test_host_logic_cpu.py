"""Host-mirror logic that needs no device: pointer bookkeeping, aspect corrections, colour cycling and
the dt clamp, against values worked out by hand from script.js (S:1188-1229, S:1527-1597)."""
import numpy as np

import webgl_fluid_simulation_b200 as pkg
from webgl_fluid_simulation_b200.sim import FluidSimulation, Pointer


def bare(width, height, **cfg):
    """A FluidSimulation without a device handle: only the pure host logic is exercised."""
    s = object.__new__(FluidSimulation)
    s.config = pkg.default_config(); s.config.update(cfg)
    s.canvas = {"width": width, "height": height}
    seq = iter(np.linspace(0.05, 0.95, 64))
    s.random = lambda: float(next(seq))
    s.pointers = [Pointer()]; s.splatStack = []
    s.lastUpdateTime = 1000.0; s.colorUpdateTimer = 0.0
    s.calls = []
    s.splat = lambda x, y, dx, dy, color: s.calls.append((x, y, dx, dy, dict(color) if isinstance(color, dict) else color))
    s.step = lambda dt: s.calls.append(("step", dt))
    return s


def test_pointer_down_and_move_wide_canvas():
    s = bare(1600, 800)                                   # aspect 2 > 1: deltaY is divided (S:1559-1563)
    p = s.pointers[0]
    s.updatePointerDownData(p, 7, 400, 200)
    assert (p.id, p.down, p.moved) == (7, True, False)
    assert p.texcoordX == 0.25 and p.texcoordY == 0.75    # y flipped, S:1531
    assert set(p.color) == {"r", "g", "b"}
    s.updatePointerMoveData(p, 480, 120)
    assert abs(p.deltaX - 0.05) < 1e-12                   # aspect >= 1: x untouched (S:1553-1557)
    assert abs(p.deltaY - 0.1 / 2) < 1e-12 and p.moved
    s.updatePointerMoveData(p, 480, 120)
    assert p.deltaX == 0 and p.deltaY == 0 and not p.moved


def test_pointer_move_tall_canvas_and_splat_pointer():
    s = bare(500, 1000, SPLAT_FORCE=6000)                 # aspect .5 < 1: deltaX is multiplied
    p = s.pointers[0]
    s.updatePointerDownData(p, 0, 100, 100)
    s.updatePointerMoveData(p, 150, 100)
    assert abs(p.deltaX - 0.1 * 0.5) < 1e-12 and p.deltaY == 0
    s.applyInputs()                                       # S:1219-1229 -> splatPointer, S:1421-1425
    x, y, dx, dy, _ = s.calls[-1]
    assert (x, y) == (p.texcoordX, p.texcoordY) and abs(dx - 0.05 * 6000) < 1e-9 and dy == 0 and not p.moved


def test_multiple_splats_draw_order_and_scaling():
    s = bare(1024, 1024)
    r = list(np.linspace(0.05, 0.95, 64))
    s.multipleSplats(2)                                   # per splat: colour, x, y, dx, dy (S:1429-1437)
    (x0, y0, dx0, dy0, c0), (x1, *_rest) = s.calls
    assert (x0, y0) == (r[1], r[2]) and abs(dx0 - 1000 * (r[3] - 0.5)) < 1e-9 and abs(dy0 - 1000 * (r[4] - 0.5)) < 1e-9
    base = pkg.HSVtoRGB(r[0], 1.0, 1.0)
    assert all(abs(c0[k] - base[k] * 0.15 * 10.0) < 1e-12 for k in "rgb")      # generateColor x10
    assert x1 == r[6]


def test_dt_clamp_update_colors_and_pause():
    s = bare(1024, 1024)
    assert s.calcDeltaTime(now_ms=1005.0) == 0.005
    assert s.calcDeltaTime(now_ms=2005.0) == 0.016666     # S:1191
    s.colorUpdateTimer = 0.95
    before = dict(s.pointers[0].color)
    s.updateColors(0.01)                                  # 0.95 + 0.01*10 >= 1 -> wrap + new colours (S:1207-1217)
    assert abs(s.colorUpdateTimer - 0.05) < 1e-12 and s.pointers[0].color != before
    s.splatStack.append(3)
    s.config["PAUSED"] = True
    s.update(now_ms=2010.0)                               # applyInputs still runs, step() does not (S:1182-1184)
    assert sum(1 for c in s.calls if c[0] == "step") == 0 and len(s.calls) == 3 and not s.splatStack
    s.config["PAUSED"] = False
    s.update(now_ms=2020.0)
    assert s.calls[-1] == ("step", 0.01)
