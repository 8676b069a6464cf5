"""B200-native stable-fluids hot path behind the step()/splat()/config surface of
PavelDoGreat/WebGL-Fluid-Simulation.  See DESIGN.md; C ABI in include/fluid.h.

(The directory is `webgl_fluid_simulation_b200`: a hyphenated name cannot be imported in Python.)
"""
from ._lib import FIELD, PARAM, FLAG_HALF_STORAGE, FLAG_NAIVE_JACOBI, FLAG_NO_GRAPH, FLAG_TILED_PASSES, FLAG_UNFUSED, FluidError, build, lib  # noqa: F401
from .sim import FluidSimulation, Pointer, default_config, getResolution, HSVtoRGB, wrap  # noqa: F401
