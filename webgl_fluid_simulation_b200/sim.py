"""Host-side mirror of the reference's simulation surface (script.js), over the C ABI.

The reference's interface is a set of globals in one classic script: `config` (S:59-85),
`step(dt)` (S:1231-1294), `splat(x,y,dx,dy,color)` (S:1441-1455), `multipleSplats(n)` (S:1427-1439),
`splatPointer(p)` (S:1421-1425), `initFramebuffers()` (S:982-1010), `generateColor()` (S:1565-1571),
`calcDeltaTime()` (S:1188-1194), `update()` (S:1176-1186) and the field objects
`velocity/dye/pressure/divergence/curl` (S:950-954).  `FluidSimulation` keeps those names,
argument meanings and the live-`config` behaviour (keys are read on every call), and forwards the
work to libfluid_b200 (hand-written sm_100a kernels).  js/fluid-sim.js is the same mirror in the
reference's own language for when Node is available; this Python twin is what the tests drive.

There is no CPU or PyTorch fallback: constructing a simulation without the built library or
without a B200 raises.

The host-side helpers (Pointer, HSVtoRGB, wrap, generateColor, multipleSplats, splatPointer,
calcDeltaTime, updateColors, applyInputs, updatePointer*Data, correctDelta*) follow script.js line
for line because mirroring that interface is the point; those parts are Copyright (c) 2017 Pavel
Dobryakov, MIT License (see LICENSE).
"""
from __future__ import annotations

import ctypes as C
import math
import random as _random
import time

import numpy as np

from . import _lib
from ._lib import FIELD, PARAM, STAT, Config, FluidError, PostFX, Timing


def default_config() -> dict:
    """The simulation keys of the reference `config` literal, S:59-69 + S:73 (render keys are out of
    scope and kept only so a reference config object can be passed through unchanged)."""
    return {
        "SIM_RESOLUTION": 128,
        "DYE_RESOLUTION": 1024,
        "DENSITY_DISSIPATION": 1,
        "VELOCITY_DISSIPATION": 0.2,
        "PRESSURE": 0.8,
        "PRESSURE_ITERATIONS": 20,
        "CURL": 30,
        "SPLAT_RADIUS": 0.25,
        "SPLAT_FORCE": 6000,
        "SHADING": True,
        "COLORFUL": True,
        "COLOR_UPDATE_SPEED": 10,
        "PAUSED": False,
        "BACK_COLOR": {"r": 0, "g": 0, "b": 0},
        "TRANSPARENT": False,
        "BLOOM": False,          # reference default is true; needs `sim.dithering` (the 64x64 RGB texture), see render()
        "BLOOM_ITERATIONS": 8,
        "BLOOM_RESOLUTION": 256,
        "BLOOM_INTENSITY": 0.8,
        "BLOOM_THRESHOLD": 0.6,
        "BLOOM_SOFT_KNEE": 0.7,
        "SUNRAYS": False,        # reference default is true
        "SUNRAYS_RESOLUTION": 196,
        "SUNRAYS_WEIGHT": 1.0,
    }


class Pointer:
    """pointerPrototype, S:87-98."""

    def __init__(self):
        self.id = -1
        self.texcoordX = 0.0
        self.texcoordY = 0.0
        self.prevTexcoordX = 0.0
        self.prevTexcoordY = 0.0
        self.deltaX = 0.0
        self.deltaY = 0.0
        self.down = False
        self.moved = False
        self.color = {"r": 30, "g": 0, "b": 300}


def HSVtoRGB(h, s, v):
    """S:1573-1597."""
    i = math.floor(h * 6)
    f = h * 6 - i
    p = v * (1 - s)
    q = v * (1 - f * s)
    t = v * (1 - (1 - f) * s)
    r, g, b = [(v, t, p), (q, v, p), (p, v, t), (p, q, v), (t, p, v), (v, p, q)][i % 6]
    return {"r": r, "g": g, "b": b}


def getResolution(resolution, canvas_w, canvas_h):
    """S:1612-1624 (via the C ABI so both hosts share one implementation)."""
    w, h = C.c_int(), C.c_int()
    _lib.lib().fluid_get_resolution(int(resolution), int(canvas_w), int(canvas_h), C.byref(w), C.byref(h))
    return {"width": w.value, "height": h.value}


class _Field:
    """What the reference exposes as `velocity.read`, `dye.read`, ...: size + texel size + data."""

    def __init__(self, sim, name):
        self._sim, self._name = sim, name

    @property
    def width(self):
        return self._sim._dims(self._name)[0]

    @property
    def height(self):
        return self._sim._dims(self._name)[1]

    @property
    def texelSizeX(self):
        return 1.0 / self.width

    @property
    def texelSizeY(self):
        return 1.0 / self.height

    def read(self) -> np.ndarray:
        return self._sim.readField(self._name)


class FluidSimulation:
    def __init__(self, config: dict | None = None, canvas_width: int = 1024, canvas_height: int = 1024,
                 device: int = -1, flags: int = 0, jacobi_block: int = 0, random=None,
                 rank: int = 0, world: int = 1, nccl_uid: bytes | None = None,
                 sizes: tuple | None = None):
        self.config = default_config()
        if config:
            self.config.update(config)
        self.canvas = {"width": int(canvas_width), "height": int(canvas_height)}
        self.random = random or _random.random          # Math.random stand-in (injectable, seedable)
        self.pointers = [Pointer()]                     # S:100-102
        self.splatStack = []
        self.lastUpdateTime = time.time() * 1000.0      # S:1172
        self.colorUpdateTimer = 0.0
        self._L = _lib.lib()
        self._h = C.c_void_p()
        self._device, self._flags, self._jb = device, flags, jacobi_block
        self._rank, self._world, self._uid = rank, world, nccl_uid
        # test hook: explicit (sim_w, sim_h, dye_w, dye_h) instead of getResolution(config.*)
        self._sizes = sizes
        self._pushed = {}
        self.velocity = _Field(self, "velocity")
        self.dye = _Field(self, "dye")
        self.pressure = _Field(self, "pressure")
        self.divergence = _Field(self, "divergence")
        self.curl = _Field(self, "curl")
        self.initFramebuffers()

    # ---- plumbing ----------------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            msg = self._L.fluid_last_error(self._h if self._h else None)
            raise FluidError(rc, msg.decode() if msg else "")

    def _aspect(self):
        return self.canvas["width"] / self.canvas["height"]

    def _push_config(self):
        """The reference reads `config` live inside step()/splat(); mirror that by pushing every
        changed scalar before the call."""
        vals = {
            "DENSITY_DISSIPATION": self.config["DENSITY_DISSIPATION"],
            "VELOCITY_DISSIPATION": self.config["VELOCITY_DISSIPATION"],
            "PRESSURE": self.config["PRESSURE"],
            "PRESSURE_ITERATIONS": self.config["PRESSURE_ITERATIONS"],
            "CURL": self.config["CURL"],
            "SPLAT_RADIUS": self.config["SPLAT_RADIUS"],
            "ASPECT": self._aspect(),
        }
        for k, v in vals.items():
            if self._pushed.get(k) != v:
                # JS numbers are doubles: SPLAT_RADIUS and the aspect enter double arithmetic in
                # correctRadius (S:1457-1462) before the single fp32 narrowing of gl.uniform1f
                self._check(self._L.fluid_set_param_f64(self._h, PARAM[k], float(v)))
                self._pushed[k] = v

    def _dims(self, name):
        w, h, c, r0 = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._check(self._L.fluid_field_dims(self._h, FIELD[name], C.byref(w), C.byref(h), C.byref(c), C.byref(r0)))
        return w.value, h.value, c.value, r0.value

    # ---- initFramebuffers, S:982-1010 --------------------------------------------------------------
    def initFramebuffers(self):
        simRes = getResolution(self.config["SIM_RESOLUTION"], self.canvas["width"], self.canvas["height"])
        dyeRes = getResolution(self.config["DYE_RESOLUTION"], self.canvas["width"], self.canvas["height"])
        if self._sizes is not None:
            simRes = {"width": self._sizes[0], "height": self._sizes[1]}
            dyeRes = {"width": self._sizes[2], "height": self._sizes[3]}
        if not self._h:                                  # `dye == null` branch: create
            cfg = Config()
            self._L.fluid_config_default(C.byref(cfg))
            cfg.sim_w, cfg.sim_h = simRes["width"], simRes["height"]
            cfg.dye_w, cfg.dye_h = dyeRes["width"], dyeRes["height"]
            cfg.aspect = self._aspect()
            cfg.device, cfg.flags, cfg.jacobi_block = self._device, self._flags, self._jb
            # the simulation keys as they stand now (a slab handle sizes its ghost zone from
            # PRESSURE_ITERATIONS); later changes are pushed live by _push_config()
            cfg.density_dissipation = float(self.config["DENSITY_DISSIPATION"])
            cfg.velocity_dissipation = float(self.config["VELOCITY_DISSIPATION"])
            cfg.pressure = float(self.config["PRESSURE"])
            cfg.pressure_iterations = int(self.config["PRESSURE_ITERATIONS"])
            cfg.curl = float(self.config["CURL"])
            cfg.splat_radius = float(self.config["SPLAT_RADIUS"])
            if self._world > 1:
                uid = C.create_string_buffer(self._uid, len(self._uid))
                self._check(self._L.fluid_create_slab(C.byref(cfg), self._rank, self._world, uid,
                                                      len(self._uid), C.byref(self._h)))
            else:
                self._check(self._L.fluid_create(C.byref(cfg), C.byref(self._h)))
        else:                                            # resizeDoubleFBO branch, S:1116-1126
            self._check(self._L.fluid_resize(self._h, simRes["width"], simRes["height"],
                                             dyeRes["width"], dyeRes["height"]))
            if self._world > 1 and getattr(self, "halo_transport", "nccl") == "p2p":
                # the peer mappings died with the old arena: export / connect again (collective)
                from .distributed import connect_peers
                self.halo_transport = "p2p" if connect_peers(self) else "nccl"
        self._pushed = {}
        self._push_config()

    # ---- the hot path ------------------------------------------------------------------------------
    def step(self, dt):
        """step(dt), S:1231-1294."""
        self._push_config()
        self._check(self._L.fluid_step(self._h, float(dt)))

    def splat(self, x, y, dx, dy, color):
        """splat(x, y, dx, dy, color), S:1441-1455; color is {r,g,b} like the reference's."""
        self._push_config()
        if isinstance(color, dict):
            r, g, b = color["r"], color["g"], color["b"]
        else:
            r, g, b = color
        self._check(self._L.fluid_splat(self._h, float(x), float(y), float(dx), float(dy),
                                        float(r), float(g), float(b)))

    def splatPointer(self, pointer):
        """S:1421-1425."""
        dx = pointer.deltaX * self.config["SPLAT_FORCE"]
        dy = pointer.deltaY * self.config["SPLAT_FORCE"]
        self.splat(pointer.texcoordX, pointer.texcoordY, dx, dy, pointer.color)

    def multipleSplats(self, amount):
        """S:1427-1439 (same draw order from the random stream as the reference)."""
        for _ in range(int(amount)):
            color = self.generateColor()
            color["r"] *= 10.0
            color["g"] *= 10.0
            color["b"] *= 10.0
            x = self.random()
            y = self.random()
            dx = 1000 * (self.random() - 0.5)
            dy = 1000 * (self.random() - 0.5)
            self.splat(x, y, dx, dy, color)

    def generateColor(self):
        """S:1565-1571."""
        c = HSVtoRGB(self.random(), 1.0, 1.0)
        c["r"] *= 0.15
        c["g"] *= 0.15
        c["b"] *= 0.15
        return c

    # ---- frame driver (the caller of the hot path), S:1176-1229 --------------------------------------
    def calcDeltaTime(self, now_ms=None):
        """S:1188-1194: dt = min(wall delta, 0.016666)."""
        now = time.time() * 1000.0 if now_ms is None else now_ms
        dt = (now - self.lastUpdateTime) / 1000.0
        dt = min(dt, 0.016666)
        self.lastUpdateTime = now
        return dt

    def updateColors(self, dt):
        """S:1207-1217."""
        if not self.config["COLORFUL"]:
            return
        self.colorUpdateTimer += dt * self.config["COLOR_UPDATE_SPEED"]
        if self.colorUpdateTimer >= 1:
            self.colorUpdateTimer = wrap(self.colorUpdateTimer, 0, 1)
            for p in self.pointers:
                p.color = self.generateColor()

    def applyInputs(self):
        """S:1219-1229."""
        if self.splatStack:
            self.multipleSplats(self.splatStack.pop())
        for p in self.pointers:
            if p.moved:
                p.moved = False
                self.splatPointer(p)

    def update(self, now_ms=None):
        """One iteration of update(), S:1176-1186, minus render() (out of scope) and rAF."""
        dt = self.calcDeltaTime(now_ms)
        self.updateColors(dt)
        self.applyInputs()
        if not self.config["PAUSED"]:
            self.step(dt)
        return dt

    # pointer helpers, S:1527-1563
    def updatePointerDownData(self, pointer, id_, posX, posY):
        pointer.id = id_
        pointer.down = True
        pointer.moved = False
        pointer.texcoordX = posX / self.canvas["width"]
        pointer.texcoordY = 1.0 - posY / self.canvas["height"]
        pointer.prevTexcoordX = pointer.texcoordX
        pointer.prevTexcoordY = pointer.texcoordY
        pointer.deltaX = 0
        pointer.deltaY = 0
        pointer.color = self.generateColor()

    def updatePointerMoveData(self, pointer, posX, posY):
        pointer.prevTexcoordX = pointer.texcoordX
        pointer.prevTexcoordY = pointer.texcoordY
        pointer.texcoordX = posX / self.canvas["width"]
        pointer.texcoordY = 1.0 - posY / self.canvas["height"]
        pointer.deltaX = self.correctDeltaX(pointer.texcoordX - pointer.prevTexcoordX)
        pointer.deltaY = self.correctDeltaY(pointer.texcoordY - pointer.prevTexcoordY)
        pointer.moved = abs(pointer.deltaX) > 0 or abs(pointer.deltaY) > 0

    def correctDeltaX(self, delta):
        a = self._aspect()
        return delta * a if a < 1 else delta

    def correctDeltaY(self, delta):
        a = self._aspect()
        return delta / a if a > 1 else delta

    # ---- data access (the reference hands textures to GL; a host mirror hands arrays) ----------------
    def readField(self, name, out: np.ndarray | None = None) -> np.ndarray:
        """framebufferToTexture (S:301-307) generalised: rows x width [x channels] float32, row 0 =
        bottom.  On a slab rank: the rows this rank owns.  `out`: a caller-owned float32 buffer of
        the right size (e.g. pinned host memory) to read into instead of allocating."""
        w, h, c, _ = self._dims(name)
        shape = (h, w, c) if c > 1 else (h, w)
        if out is None:
            out = np.empty(shape, np.float32)
        else:
            assert out.dtype == np.float32 and out.flags.c_contiguous and out.size == h * w * c
            out = out.reshape(shape)
        self._check(self._L.fluid_read(self._h, FIELD[name], out.ctypes.data_as(C.c_void_p), out.size))
        return out

    def writeField(self, name, array):
        a = np.ascontiguousarray(array, np.float32)
        self._check(self._L.fluid_write(self._h, FIELD[name], a.ctypes.data_as(C.c_void_p), a.size))

    def render(self, width=None, height=None, target=False) -> np.ndarray:
        """render(target), S:1296-1317: (height, width, 4) float32 RGBA, row 0 = bottom.  Defaults
        to the canvas size (target == null branch, S:1332-1333).  config.TRANSPARENT: the display is
        drawn over the checkerboard on the screen (target=False, S:1311-1312) and bare, un-blended,
        into a capture target (target=True: what captureScreenshot renders into, S:287-299)."""
        w = int(width or self.canvas["width"]); h = int(height or self.canvas["height"])
        bc = self.config["BACK_COLOR"]                       # normalizeColor, S:1599-1606
        out = np.empty((h, w, 4), np.float32)
        cfg = self.config
        bg = 0 if not cfg.get("TRANSPARENT") else (2 if target else 1)
        if self._pushed.get("BACKGROUND") != bg:
            self._check(self._L.fluid_set_param(self._h, PARAM["BACKGROUND"], float(bg)))
            self._pushed["BACKGROUND"] = bg
        self._push_config()                                  # the checkerboard uses canvas.width / canvas.height (S:1327)
        if cfg.get("BLOOM") or cfg.get("SUNRAYS"):
            # built for the reference's default keyword set only: SHADING + BLOOM + SUNRAYS
            if not (cfg.get("BLOOM") and cfg.get("SUNRAYS") and cfg.get("SHADING")):
                raise NotImplementedError("post-FX are built for SHADING + BLOOM + SUNRAYS together (the reference defaults)")
            dith = getattr(self, "dithering", None)
            if dith is None:
                raise ValueError("set sim.dithering to the (64, 64, 3) float RGB dithering texture (LDR_LLL1_0.png / 255)")
            dith = np.ascontiguousarray(dith, np.float32)
            fx = PostFX(int(cfg["BLOOM_ITERATIONS"]), int(cfg["BLOOM_RESOLUTION"]), float(cfg["BLOOM_INTENSITY"]),
                        float(cfg["BLOOM_THRESHOLD"]), float(cfg["BLOOM_SOFT_KNEE"]), int(cfg["SUNRAYS_RESOLUTION"]),
                        float(cfg["SUNRAYS_WEIGHT"]))
            b = getResolution(cfg["BLOOM_RESOLUTION"], w, h); sr = getResolution(cfg["SUNRAYS_RESOLUTION"], w, h)
            self.last_bloom = np.empty((b["height"], b["width"], 4), np.float32)
            self.last_sunrays = np.empty((sr["height"], sr["width"]), np.float32)
            self._check(self._L.fluid_render_postfx(
                self._h, w, h, C.byref(fx), dith.ctypes.data_as(C.c_void_p), dith.shape[1], dith.shape[0],
                bc["r"] / 255, bc["g"] / 255, bc["b"] / 255, out.ctypes.data_as(C.c_void_p), out.size,
                self.last_bloom.ctypes.data_as(C.c_void_p), self.last_sunrays.ctypes.data_as(C.c_void_p)))
            return out
        y0, y1 = C.c_int(), C.c_int()                        # a slab rank draws its band of the target only
        self._check(self._L.fluid_render_band(self._h, h, C.byref(y0), C.byref(y1)))
        out = out[: y1.value - y0.value]
        self.render_band = (y0.value, y1.value)
        self._check(self._L.fluid_render(self._h, w, h, 1 if self.config["SHADING"] else 0,
                                         bc["r"] / 255, bc["g"] / 255, bc["b"] / 255,
                                         out.ctypes.data_as(C.c_void_p), out.size))
        return out

    @staticmethod
    def textureToCanvas(rgba: np.ndarray) -> np.ndarray:
        """normalizeTexture + textureToCanvas (S:309-349): clamp01 * 255 truncated into a Uint8
        image, flipped so that row 0 is the TOP (canvas order)."""
        img = (np.clip(rgba, 0.0, 1.0) * 255.0).astype(np.uint8)
        return img[::-1].copy()

    def sync(self):
        self._check(self._L.fluid_sync(self._h))

    # ---- per-pass entry points (test surface; one reference blit each) -------------------------------
    def pass_(self, name, *args):
        self._push_config()
        fn = getattr(self._L, "fluid_pass_" + name)
        conv = [C.c_int(int(a)) if name == "jacobi" else C.c_float(float(a)) for a in args]
        self._check(fn(self._h, *conv))

    def pressure_solve_host(self, div_host: np.ndarray, p_host: np.ndarray, iters: int):
        assert div_host.dtype == np.float32 and p_host.dtype == np.float32
        self._push_config()
        self._check(self._L.fluid_pressure_solve_host(self._h, div_host.ctypes.data_as(C.c_void_p),
                                                      p_host.ctypes.data_as(C.c_void_p), int(iters)))

    def mark(self, slot):
        self._check(self._L.fluid_mark(self._h, slot))

    def elapsed_ms(self) -> float:
        ms = C.c_float()
        self._check(self._L.fluid_elapsed_ms(self._h, C.byref(ms)))
        return ms.value

    def timing(self) -> dict:
        t = Timing()
        self._check(self._L.fluid_timing_last(self._h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in Timing._fields_}

    def launch_count(self) -> int:
        return int(self._L.fluid_launch_count(self._h))

    def stat(self, name) -> int:
        """Library counters (fluid_stat): launches, jacobi_launches, halo_launches, halo_exchanges,
        graph_captures, graph_launches, halo_transport_p2p."""
        return int(self._L.fluid_stat(self._h, STAT[name]))

    def device_ptr(self, name) -> int:
        return int(self._L.fluid_device_ptr(self._h, FIELD[name]) or 0)

    def set_param(self, key, value):
        self._check(self._L.fluid_set_param(self._h, PARAM[key], float(value)))

    def close(self):
        if self._h:
            self._L.fluid_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def wrap(value, mn, mx):
    """S:1599-1603."""
    rng = mx - mn
    if rng == 0:
        return mn
    return (value - mn) % rng + mn
