"""Executes the reference's GLSL ES 1.00 shader SOURCE TEXT on the CPU (numpy) to cut golden vectors.

TEST INFRASTRUCTURE ONLY (generation time, in the build container only — /root/reference does not
exist on the GPU box; the vectors it produces are committed under tests/golden/).

The reference cannot run here (needs a browser DOM + WebGL; no node / GL in the image), so its
own arithmetic is recovered the only way available: the shader strings are read out of
/root/reference/script.js AT RUN TIME (nothing is copied into this repo), transpiled statement by
statement to numpy expressions, and drawn "full-screen" by evaluating the vertex shader at every
fragment centre and the fragment shader on the resulting varyings.  Textures are emulated per the
GL ES sampler rules the reference sets up (createFBO S:1045-1077): CLAMP_TO_EDGE, NEAREST or
LINEAR, optional half-float storage (S:138-147).  The JS orchestration (step S:1231-1294, splat
S:1441-1462) is restated below call for call (program, uniforms, target, swap).

What this pins: the formulas, operation order, neighbour geometry, wall rules and pass order of the
C oracle against the shader text itself.  What it cannot pin: a real driver's exp()/mix()/LINEAR
weight precision — those are implementation-defined in WebGL (see DESIGN.md "Parity").
"""
from __future__ import annotations

import re

import numpy as np

F = np.float32
REFERENCE_JS = "/root/reference/script.js"

# ------------------------------------------------------------------------------------------------
# vector runtime


class V:
    """GLSL vecN over a fragment grid: arr has shape (..., N) float32."""

    __slots__ = ("arr",)
    __array_ufunc__ = None      # numpy scalars (uniforms) defer to V.__r*__ instead of iterating it
    _IDX = {c: i for s in ("xyzw", "rgba", "stpq") for i, c in enumerate(s)}

    def __init__(self, arr):
        object.__setattr__(self, "arr", np.asarray(arr, dtype=F))

    def _co(self, o):
        if isinstance(o, V):
            return o.arr
        if isinstance(o, np.ndarray) and o.ndim >= 1:
            return o[..., None]
        return o

    def __getattr__(self, name):
        if not all(c in V._IDX for c in name):
            raise AttributeError(name)
        idx = [V._IDX[c] for c in name]
        if len(idx) == 1:
            return self.arr[..., idx[0]].copy()
        return V(self.arr[..., idx])

    def __setattr__(self, name, value):
        idx = [V._IDX[c] for c in name]
        if len(idx) == 1:
            self.arr[..., idx[0]] = value
        else:
            self.arr[..., idx] = value.arr if isinstance(value, V) else value

    def __add__(self, o): return V(self.arr + self._co(o))
    def __radd__(self, o): return V(self._co(o) + self.arr)
    def __sub__(self, o): return V(self.arr - self._co(o))
    def __rsub__(self, o): return V(self._co(o) - self.arr)
    def __mul__(self, o): return V(self.arr * self._co(o))
    def __rmul__(self, o): return V(self._co(o) * self.arr)
    def __truediv__(self, o): return V(self.arr / self._co(o))
    def __rtruediv__(self, o): return V(self._co(o) / self.arr)
    def __neg__(self): return V(-self.arr)
    # augmented assignment rebinds (GLSL value semantics), never aliases
    __iadd__ = __add__
    __isub__ = __sub__
    __imul__ = __mul__
    __itruediv__ = __truediv__


def _bc(*xs):
    """Broadcast a mix of python floats / (H,W) arrays / V to a common fragment shape."""
    shp = ()
    for x in xs:
        a = x.arr[..., 0] if isinstance(x, V) else np.asarray(x)
        shp = np.broadcast_shapes(shp, a.shape)
    return shp


def _vecn(n):
    def ctor(*args):
        comps = []
        for a in args:
            if isinstance(a, V):
                comps += [a.arr[..., k] for k in range(a.arr.shape[-1])]
            else:
                comps.append(np.asarray(a, dtype=F))
        if len(comps) == 1:
            comps = comps * n
        assert len(comps) == n, (n, len(comps))
        shp = np.broadcast_shapes(*[c.shape for c in comps])
        return V(np.stack([np.broadcast_to(c, shp) for c in comps], axis=-1))
    return ctor


def _map(fn):
    def g(x):
        return V(fn(x.arr)) if isinstance(x, V) else fn(np.asarray(x, dtype=F))
    return g


def _floor(x): return np.floor(x)
def _fract(x): return x - np.floor(x)                      # GLSL: x - floor(x)


def g_mix(x, y, a):                                         # GLSL: x*(1-a) + y*a
    if isinstance(x, V):
        aa = x._co(a)
        return V(x.arr * (F(1.0) - aa) + y.arr * aa)
    return x * (F(1.0) - a) + y * a


def g_dot(a, b):
    p = a.arr * b.arr
    s = p[..., 0]
    for k in range(1, p.shape[-1]):
        s = s + p[..., k]
    return s


def g_length(a): return np.sqrt(g_dot(a, a))


def g_min(a, b):
    if isinstance(a, V): return V(np.minimum(a.arr, a._co(b)))
    return np.minimum(a, b)


def g_max(a, b):
    if isinstance(a, V): return V(np.maximum(a.arr, a._co(b)))
    return np.maximum(a, b)


def g_normalize(a): return V(a.arr / g_length(a)[..., None])        # v / length(v)


def g_clamp(x, lo, hi): return g_min(g_max(x, lo), hi)              # min(max(x, lo), hi)


def g_pow(x, y):
    if isinstance(x, V): return V(np.power(x.arr, x._co(y)).astype(F))
    return np.power(x, y).astype(F)


def g_mod(x, y):                                                    # GLSL: x - y * floor(x / y)
    if isinstance(x, V): return V(x.arr - x._co(y) * np.floor(x.arr / x._co(y)))
    x = np.asarray(x, F); return (x - y * np.floor(x / y)).astype(F)


def g_where(c, a, b):
    if isinstance(a, V) or isinstance(b, V):
        aa = a.arr if isinstance(a, V) else a
        bb = b.arr if isinstance(b, V) else b
        return V(np.where(np.asarray(c)[..., None], aa, bb))
    return np.where(c, a, b).astype(F)


def g_copy(x):
    if isinstance(x, V): return V(x.arr.copy())
    return np.array(x, dtype=F)


# ------------------------------------------------------------------------------------------------
# textures


class Texture:
    """createFBO (S:1045-1077): w x h texels, `ch` stored channels (R / RG / RGBA), filter NEAREST
    or LINEAR, CLAMP_TO_EDGE, cleared to clearColor (0,0,0,1) (S:136, S:1059)."""

    def __init__(self, w, h, ch, linear, half=False, repeat=False):
        self.w, self.h, self.ch, self.linear, self.half = w, h, ch, linear, half
        self.repeat = repeat                # TEXTURE_WRAP REPEAT (the dithering texture, S:1133-1134)
        self.data = np.zeros((h, w, 4), F)
        self.data[..., 3] = 1.0
        self.texelSizeX = 1.0 / w          # JS doubles (S:1061-1062); narrowed by gl.uniform2f
        self.texelSizeY = 1.0 / h

    def store(self, rgba):
        d = np.zeros((self.h, self.w, 4), F)
        d[..., 3] = 1.0                     # channels a format lacks read back as (0,0,1)
        src = rgba.arr[..., : self.ch]
        if self.half:
            src = src.astype(np.float16).astype(F)
        d[..., : self.ch] = src
        self.data = d

    def _fetch(self, ix, iy):
        if self.repeat:
            ix = np.mod(ix, self.w).astype(np.int64); iy = np.mod(iy, self.h).astype(np.int64)
        else:
            ix = np.clip(ix, 0, self.w - 1).astype(np.int64)
            iy = np.clip(iy, 0, self.h - 1).astype(np.int64)
        return self.data[iy, ix]

    def sample(self, uv):
        u = uv.arr[..., 0] * F(self.w)
        v = uv.arr[..., 1] * F(self.h)
        if not self.linear:
            return V(self._fetch(np.floor(u), np.floor(v)))
        # GL ES 2.0 §3.7.7 LINEAR: i0 = floor(u - .5), alpha = frac(u - .5), exact fp32 weights
        u = u - F(0.5); v = v - F(0.5)
        i0 = np.floor(u); j0 = np.floor(v)
        a = (u - i0)[..., None]; b = (v - j0)[..., None]
        t00 = self._fetch(i0, j0); t10 = self._fetch(i0 + 1, j0)
        t01 = self._fetch(i0, j0 + 1); t11 = self._fetch(i0 + 1, j0 + 1)
        one = F(1.0)
        return V((one - a) * (one - b) * t00 + a * (one - b) * t10 + (one - a) * b * t01 + a * b * t11)


class DoubleTexture:
    """createDoubleFBO (S:1079-1106)."""

    def __init__(self, *a, **k):
        self.read = Texture(*a, **k)
        self.write = Texture(*a, **k)
        self.texelSizeX, self.texelSizeY = self.read.texelSizeX, self.read.texelSizeY

    def swap(self):
        self.read, self.write = self.write, self.read


def g_texture2D(sam, uv):
    return sam.sample(uv)


# ------------------------------------------------------------------------------------------------
# GLSL -> python


def extract_shader(js: str, name: str):
    """Source of `const <name> = compileShader(gl.X_SHADER, `...`[, keywords]);` or of a plain
    template-string constant (`const displayShaderSource = `...`;`, compiled later by Material)."""
    m = re.search(r"const\s+" + re.escape(name) + r"\s*=\s*compileShader\(\s*gl\.\w+,\s*`(.*?)`", js, re.S)
    if not m:
        m = re.search(r"const\s+" + re.escape(name) + r"\s*=\s*`(.*?)`", js, re.S)
    if not m:
        raise KeyError(name)
    return m.group(1)


def _preprocess(src: str, defines):
    out, stack, macros = [], [True], {}
    for line in src.splitlines():
        s = line.strip()
        if s.startswith("#define"):
            parts = s.split(None, 2)
            if len(parts) == 3 and all(stack):
                macros[parts[1]] = parts[2]
            continue
        if s.startswith("#ifdef"):
            stack.append(s.split()[1] in defines)
        elif s.startswith("#else"):
            stack[-1] = not stack[-1]
        elif s.startswith("#endif"):
            stack.pop()
        elif all(stack):
            line = re.sub(r"//.*", "", line)
            for k, v in macros.items():
                line = re.sub(r"\b" + re.escape(k) + r"\b", v, line)
            out.append(line)
    return "\n".join(out)


_TYPES = r"(?:float|vec2|vec3|vec4|int|bool)"
_RENAME = {"texture2D": "g_texture2D", "mix": "g_mix", "dot": "g_dot", "length": "g_length",
           "min": "g_min", "max": "g_max", "floor": "g_floor", "fract": "g_fract", "exp": "g_exp",
           "abs": "g_abs", "normalize": "g_normalize", "clamp": "g_clamp", "pow": "g_pow", "mod": "g_mod"}


def _expr(e: str) -> str:
    e = e.strip()
    for k, v in _RENAME.items():
        e = re.sub(r"\b" + k + r"\s*\(", v + "(", e)
    return e


def _match(s, i, open_c, close_c):
    d = 0
    for k in range(i, len(s)):
        if s[k] == open_c: d += 1
        elif s[k] == close_c:
            d -= 1
            if d == 0: return k
    raise ValueError("unbalanced")


def _stmts(body: str, ind: str):
    py, i, n = [], 0, len(body)
    while i < n:
        while i < n and body[i].isspace(): i += 1
        if i >= n: break
        if body.startswith("if", i) and re.match(r"if\s*\(", body[i:]):
            p0 = body.index("(", i); p1 = _match(body, p0, "(", ")")
            cond = _expr(body[p0 + 1:p1])
            b0 = body.index("{", p1); b1 = _match(body, b0, "{", "}")
            for st in body[b0 + 1:b1].split(";"):
                st = st.strip()
                if not st: continue
                m = re.match(r"([\w.]+)\s*=\s*(.*)$", st, re.S)
                assert m, "only plain assignments are supported under if: " + st
                py.append(f"{ind}{m.group(1)} = g_where({cond}, {_expr(m.group(2))}, {m.group(1)})")
            i = b1 + 1
            continue
        mf = re.match(r"for\s*\(\s*int\s+(\w+)\s*=\s*(\w+)\s*;\s*\1\s*<\s*(\w+)\s*;\s*\1\+\+\s*\)", body[i:])
        if mf:                                    # for (int i = A; i < B; i++) { ... }
            b0 = body.index("{", i + mf.end() - 1); b1 = _match(body, b0, "{", "}")
            py.append(f"{ind}for {mf.group(1)} in range({mf.group(2)}, {mf.group(3)}):")
            py += _stmts(body[b0 + 1:b1], ind + "    ") or [ind + "    pass"]
            i = b1 + 1
            continue
        j = body.index(";", i)
        st = body[i:j].strip(); i = j + 1
        if not st: continue
        m = re.match(r"return\s+(.*)$", st, re.S)
        if m:
            py.append(f"{ind}return {_expr(m.group(1))}"); continue
        m = re.match(_TYPES + r"\s+(\w+)\s*=\s*(.*)$", st, re.S)
        if m:
            py.append(f"{ind}{m.group(1)} = g_copy({_expr(m.group(2))})"); continue
        if re.match(_TYPES + r"\s+\w+$", st):
            continue
        m = re.match(r"([\w.]+)\s*(=|\+=|-=|\*=|/=)\s*(.*)$", st, re.S)
        assert m, "unsupported statement: " + st
        py.append(f"{ind}{m.group(1)} {m.group(2)} {_expr(m.group(3))}")
    return py


def transpile(src: str, defines=()):
    """Returns (python_source, uniforms, varyings, attributes).  main() becomes main(E) where E is
    a namespace object holding uniforms / varyings / gl_FragColor."""
    src = _preprocess(src, set(defines))
    src = re.sub(r"precision\s+\w+\s+\w+\s*;", "", src)
    decl = {"uniform": [], "varying": [], "attribute": []}
    for kind in decl:
        for m in re.finditer(kind + r"\s+(?:highp\s+|mediump\s+|lowp\s+)?(\w+)\s+(\w+)\s*;", src):
            decl[kind].append((m.group(1), m.group(2)))
        src = re.sub(kind + r"\s+(?:highp\s+|mediump\s+|lowp\s+)?\w+\s+\w+\s*;", "", src)
    glob = [n for _, n in decl["uniform"] + decl["varying"] + decl["attribute"]]
    glob += ["gl_FragColor", "gl_Position"]
    py, i = [], 0
    for m in re.finditer(r"(\w+)\s+(\w+)\s*\(([^)]*)\)\s*\{", src):
        if m.start() < i: continue
        b1 = _match(src, m.end() - 1, "{", "}")
        name = m.group(2)
        params = [p.split()[-1] for p in m.group(3).split(",") if p.strip()]
        body = _stmts(src[m.end():b1], "    ")
        if name == "main":
            py.append("def main(E):")
            py += [f"    {g} = getattr(E, '{g}', None)" for g in glob]
            py += body
            py += [f"    E.{g} = {g}" for g in [n for _, n in decl["varying"]] + ["gl_FragColor", "gl_Position"]]
        else:
            py.append(f"def {name}({', '.join(params)}):")
            # helper functions see uniforms through the module-level namespace set before each draw
            py += body
        i = b1
    return "\n".join(py), decl["uniform"], decl["varying"], decl["attribute"]


class _NS:
    pass


class Program:
    """`new Program(vertexShader, fragmentShader)` (S:376-394) + blit (S:915-942)."""

    def __init__(self, js, vs_name, fs_name, defines=()):
        self.name = fs_name
        self.vs_src, _, self.varyings, _ = transpile(extract_shader(js, vs_name))
        self.fs_src, self.uniforms, _, _ = transpile(extract_shader(js, fs_name), defines)
        base = {"vec2": _vecn(2), "vec3": _vecn(3), "vec4": _vecn(4), "g_texture2D": g_texture2D,
                "g_mix": g_mix, "g_dot": g_dot, "g_length": g_length, "g_min": g_min,
                "g_max": g_max, "g_floor": _map(_floor), "g_fract": _map(_fract),
                "g_exp": _map(np.exp), "g_abs": _map(np.abs), "g_where": g_where, "g_copy": g_copy,
                "g_normalize": g_normalize, "g_clamp": g_clamp, "g_pow": g_pow, "g_mod": g_mod}
        self.vs_env = dict(base); exec(self.vs_src, self.vs_env)
        self.fs_env = dict(base); exec(self.fs_src, self.fs_env)
        self.u = {}

    def set(self, **uniforms):
        for k, val in uniforms.items():
            if isinstance(val, (tuple, list)):
                val = V(np.array(val, dtype=F))      # gl.uniform2f / uniform3f narrow to fp32
            elif isinstance(val, (float, int, np.floating)):
                val = F(val)                         # gl.uniform1f
            self.u[k] = val

    def blit(self, target: Texture, blend=None):
        """blend: None (gl.disable(BLEND)), "add" (blendFunc(ONE, ONE)) or "premult"
        (blendFunc(ONE, ONE_MINUS_SRC_ALPHA))."""
        w, h = target.w, target.h
        # fragment centres in NDC: the quad spans [-1,1]^2 over the w x h viewport (S:917, S:931)
        xs = (np.arange(w, dtype=F) + F(0.5)) / F(w) * F(2.0) - F(1.0)
        ys = (np.arange(h, dtype=F) + F(0.5)) / F(h) * F(2.0) - F(1.0)
        pos = V(np.stack(np.broadcast_arrays(xs[None, :], ys[:, None]), axis=-1))
        E = _NS()
        E.aPosition = pos
        for k, val in self.u.items():
            setattr(E, k, val)
        self.vs_env["main"](E)
        for k, val in self.u.items():               # helper functions (bilerp) read uniforms as globals
            self.fs_env[k] = val
        self.fs_env["main"](E)
        src = E.gl_FragColor
        if blend is not None:
            dst = target.data
            full = np.broadcast_to(src.arr, dst.shape).astype(F)
            if blend == "add":
                src = V(full + dst)
            elif blend == "premult":
                src = V(full + dst * (F(1.0) - full[..., 3:4]))
        target.store(src)


# ------------------------------------------------------------------------------------------------
# the reference's simulation, orchestrated exactly as script.js does


class GLSLSim:
    """initFramebuffers (S:982-1010) + step (S:1231-1294) + splat (S:1441-1462) over the executed
    shaders.  linear_filtering=False is the reference's `!ext.supportLinearFiltering` path
    (MANUAL_FILTERING keyword S:783, NEAREST samplers S:988): its arithmetic is fully defined by
    the shader text.  half=True stores every field as fp16 like the real textures (S:986)."""

    def __init__(self, sim_w, sim_h, dye_w, dye_h, linear_filtering=False, half=False, config=None,
                 js_path=REFERENCE_JS, canvas_aspect=None):
        js = open(js_path).read()
        self.config = dict(DENSITY_DISSIPATION=1, VELOCITY_DISSIPATION=0.2, PRESSURE=0.8,
                           PRESSURE_ITERATIONS=20, CURL=30, SPLAT_RADIUS=0.25)   # S:63-68
        self.config.update(config or {})
        self.linear = linear_filtering
        self.aspect = canvas_aspect if canvas_aspect is not None else sim_w / sim_h
        defs = () if linear_filtering else ("MANUAL_FILTERING",)
        P = lambda fs, d=(): Program(js, "baseVertexShader", fs, d)
        self.clearProgram = P("clearShader")
        self.splatProgram = P("splatShader")
        self.advectionProgram = P("advectionShader", defs)
        self.divergenceProgram = P("divergenceShader")
        self.curlProgram = P("curlShader")
        self.vorticityProgram = P("vorticityShader")
        self.pressureProgram = P("pressureShader")
        self.gradienSubtractProgram = P("gradientSubtractShader")
        lin = linear_filtering
        self.dye = DoubleTexture(dye_w, dye_h, 4, lin, half)
        self.velocity = DoubleTexture(sim_w, sim_h, 2, lin, half)
        self.divergence = Texture(sim_w, sim_h, 1, False, half)
        self.curl = Texture(sim_w, sim_h, 1, False, half)
        self.pressure = DoubleTexture(sim_w, sim_h, 1, False, half)

    # individual blits, usable by golden-vector generation on arbitrary input fields
    def run_curl(self):
        v = self.velocity
        self.curlProgram.set(texelSize=(v.texelSizeX, v.texelSizeY), uVelocity=v.read)
        self.curlProgram.blit(self.curl)

    def run_vorticity(self, dt):
        v = self.velocity
        self.vorticityProgram.set(texelSize=(v.texelSizeX, v.texelSizeY), uVelocity=v.read,
                                  uCurl=self.curl, curl=self.config["CURL"], dt=dt)
        self.vorticityProgram.blit(v.write); v.swap()

    def run_divergence(self):
        v = self.velocity
        self.divergenceProgram.set(texelSize=(v.texelSizeX, v.texelSizeY), uVelocity=v.read)
        self.divergenceProgram.blit(self.divergence)

    def run_clear(self):
        self.clearProgram.set(texelSize=(self.velocity.texelSizeX, self.velocity.texelSizeY),
                              uTexture=self.pressure.read, value=self.config["PRESSURE"])
        self.clearProgram.blit(self.pressure.write); self.pressure.swap()

    def run_pressure(self, iters):
        v = self.velocity
        self.pressureProgram.set(texelSize=(v.texelSizeX, v.texelSizeY), uDivergence=self.divergence)
        for _ in range(iters):
            self.pressureProgram.set(uPressure=self.pressure.read)
            self.pressureProgram.blit(self.pressure.write); self.pressure.swap()

    def run_gradient_subtract(self):
        v = self.velocity
        self.gradienSubtractProgram.set(texelSize=(v.texelSizeX, v.texelSizeY),
                                        uPressure=self.pressure.read, uVelocity=v.read)
        self.gradienSubtractProgram.blit(v.write); v.swap()

    def run_advect_velocity(self, dt):
        v = self.velocity
        self.advectionProgram.set(texelSize=(v.texelSizeX, v.texelSizeY))
        if not self.linear:
            self.advectionProgram.set(dyeTexelSize=(v.texelSizeX, v.texelSizeY))
        self.advectionProgram.set(uVelocity=v.read, uSource=v.read, dt=dt,
                                  dissipation=self.config["VELOCITY_DISSIPATION"])
        self.advectionProgram.blit(v.write); v.swap()

    def run_advect_dye(self, dt):
        v, d = self.velocity, self.dye
        # texelSize is NOT re-set between the two advection draws (S:1276 .. S:1292)
        self.advectionProgram.set(texelSize=(v.texelSizeX, v.texelSizeY))
        if not self.linear:
            self.advectionProgram.set(dyeTexelSize=(d.texelSizeX, d.texelSizeY))
        self.advectionProgram.set(uVelocity=v.read, uSource=d.read, dt=dt,
                                  dissipation=self.config["DENSITY_DISSIPATION"])
        self.advectionProgram.blit(d.write); d.swap()

    def step(self, dt):                                               # S:1231-1294
        self.run_curl()
        self.run_vorticity(dt)
        self.run_divergence()
        self.run_clear()
        self.run_pressure(self.config["PRESSURE_ITERATIONS"])
        self.run_gradient_subtract()
        self.run_advect_velocity(dt)
        self.run_advect_dye(dt)

    def splat(self, x, y, dx, dy, color):                              # S:1441-1455
        radius = self.config["SPLAT_RADIUS"] / 100.0                   # S:1447
        if self.aspect > 1:                                            # correctRadius S:1457-1462
            radius *= self.aspect
        sp = self.splatProgram
        sp.set(texelSize=(self.velocity.texelSizeX, self.velocity.texelSizeY),
               uTarget=self.velocity.read, aspectRatio=self.aspect, point=(x, y),
               color=(dx, dy, 0.0), radius=radius)
        sp.blit(self.velocity.write); self.velocity.swap()
        sp.set(uTarget=self.dye.read, color=tuple(color))
        sp.blit(self.dye.write); self.dye.swap()

    def render(self, width, height, shading=True, back_color=(0, 0, 0), transparent=False, to_screen=True):
        """render(target) with config.BLOOM = config.SUNRAYS = false (S:1296-1317).  TRANSPARENT false:
        drawColor(normalizeColor(BACK_COLOR)) then drawDisplay, blended ONE / ONE_MINUS_SRC_ALPHA
        (S:1305).  TRANSPARENT on the screen (target == null): drawCheckerboard instead of drawColor
        (S:1311-1312); TRANSPARENT into a capture target: no background, blending disabled (S:1307-1308).
        Returns the (height, width, 4) float target."""
        js = open(REFERENCE_JS).read()
        target = Texture(width, height, 4, True)
        if not transparent:
            color = Program(js, "baseVertexShader", "colorShader")
            color.set(texelSize=(1.0 / width, 1.0 / height),
                      color=(back_color[0] / 255, back_color[1] / 255, back_color[2] / 255, 1))   # S:1321, S:1599
            color.blit(target)
        elif to_screen:
            chk = Program(js, "baseVertexShader", "checkerboardShader")
            chk.set(texelSize=(1.0 / width, 1.0 / height), aspectRatio=width / height)            # S:1327
            chk.blit(target)
        dst = target.data.copy()
        disp = Program(js, "baseVertexShader", "displayShaderSource", ("SHADING",) if shading else ())
        # the dye texture is sampled through its LINEAR filter here whatever the advection path was
        dye = Texture(self.dye.read.w, self.dye.read.h, 4, True)
        dye.data = self.dye.read.data
        disp.set(texelSize=(1.0 / width, 1.0 / height), uTexture=dye)                      # S:1337-1338
        disp.blit(target)
        src = target.data
        if transparent and not to_screen:
            return src.astype(F)                                                            # gl.disable(BLEND), S:1308
        one_minus_a = (F(1.0) - src[..., 3:4]).astype(F)
        return (src + dst * one_minus_a).astype(F)                                          # S:1305

    def render_postfx(self, width, height, dither, cfg=None, back_color=(0, 0, 0)):
        """render(null) with BLOOM, SUNRAYS and SHADING on (the reference's desktop defaults,
        S:70-84): applyBloom (S:1350-1394), applySunrays + blur (S:1396-1419), drawColor,
        drawDisplay (S:1296-1348).  FBO sizes follow initBloomFramebuffers / initSunraysFramebuffers
        (S:1012-1043) for a width x height drawing buffer.  `dither` is the 64x64 LDR_LLL1_0.png
        as float RGB in [0,1].  Returns the target and the intermediate bloom / sunrays textures."""
        js = open(REFERENCE_JS).read()
        c = dict(BLOOM_ITERATIONS=8, BLOOM_RESOLUTION=256, BLOOM_INTENSITY=0.8, BLOOM_THRESHOLD=0.6,
                 BLOOM_SOFT_KNEE=0.7, SUNRAYS_RESOLUTION=196, SUNRAYS_WEIGHT=1.0)
        c.update(cfg or {})

        def get_resolution(res):                                                    # S:1612-1624
            ar = width / height
            if ar < 1: ar = 1.0 / ar
            mn, mx = int(np.floor(res + 0.5)), int(np.floor(res * ar + 0.5))
            return (mx, mn) if width > height else (mn, mx)

        P = lambda vs, fs, d=(): Program(js, vs, fs, d)
        dye = Texture(self.dye.read.w, self.dye.read.h, 4, True); dye.data = self.dye.read.data
        # ---- applyBloom(dye.read, bloom) ----------------------------------------------------------
        bw, bh = get_resolution(c["BLOOM_RESOLUTION"])
        bloom = Texture(bw, bh, 4, True)
        pyramid = []
        for i in range(c["BLOOM_ITERATIONS"]):
            pw, ph = bw >> (i + 1), bh >> (i + 1)
            if pw < 2 or ph < 2: break
            pyramid.append(Texture(pw, ph, 4, True))
        if len(pyramid) >= 2:
            knee = c["BLOOM_THRESHOLD"] * c["BLOOM_SOFT_KNEE"] + 0.0001
            pre = P("baseVertexShader", "bloomPrefilterShader")
            pre.set(texelSize=(0.0, 0.0), curve=(c["BLOOM_THRESHOLD"] - knee, knee * 2, 0.25 / knee),
                    threshold=c["BLOOM_THRESHOLD"], uTexture=dye)
            pre.blit(bloom)
            last = bloom
            blurp = P("baseVertexShader", "bloomBlurShader")
            for dest in pyramid:
                blurp.set(texelSize=(last.texelSizeX, last.texelSizeY), uTexture=last)
                blurp.blit(dest); last = dest
            for i in range(len(pyramid) - 2, -1, -1):
                base = pyramid[i]
                blurp.set(texelSize=(last.texelSizeX, last.texelSizeY), uTexture=last)
                blurp.blit(base, blend="add"); last = base
            fin = P("baseVertexShader", "bloomFinalShader")
            fin.set(texelSize=(last.texelSizeX, last.texelSizeY), uTexture=last, intensity=c["BLOOM_INTENSITY"])
            fin.blit(bloom)
        # ---- applySunrays(dye.read, dye.write, sunrays); blur(sunrays, sunraysTemp, 1) ---------------
        sw, sh = get_resolution(c["SUNRAYS_RESOLUTION"])
        mask = Texture(dye.w, dye.h, 4, True)
        sun, tmp = Texture(sw, sh, 1, True), Texture(sw, sh, 1, True)
        mp = P("baseVertexShader", "sunraysMaskShader"); mp.set(texelSize=(0.0, 0.0), uTexture=dye); mp.blit(mask)
        sp = P("baseVertexShader", "sunraysShader"); sp.set(texelSize=(0.0, 0.0), weight=c["SUNRAYS_WEIGHT"], uTexture=mask); sp.blit(sun)
        bp = P("blurVertexShader", "blurShader")
        bp.set(texelSize=(sun.texelSizeX, 0.0), uTexture=sun); bp.blit(tmp)
        bp.set(texelSize=(0.0, sun.texelSizeY), uTexture=tmp); bp.blit(sun)
        # ---- drawColor + drawDisplay -------------------------------------------------------------------
        target = Texture(width, height, 4, True)
        col = P("baseVertexShader", "colorShader")
        col.set(texelSize=(1.0 / width, 1.0 / height), color=(back_color[0] / 255, back_color[1] / 255, back_color[2] / 255, 1))
        col.blit(target)
        dith = Texture(dither.shape[1], dither.shape[0], 3, True, repeat=True)
        dith.data[..., :3] = dither
        disp = P("baseVertexShader", "displayShaderSource", ("SHADING", "BLOOM", "SUNRAYS"))
        disp.set(texelSize=(1.0 / width, 1.0 / height), uTexture=dye, uBloom=bloom, uDithering=dith,
                 ditherScale=(width / dith.w, height / dith.h), uSunrays=sun)          # S:1337-1346, S:1626-1631
        disp.blit(target, blend="premult")
        return dict(target=target.data.copy(), bloom=bloom.data.copy(), sunrays=sun.data[..., 0].copy(),
                    mask_alpha=mask.data[..., 3].copy(), pyramid=[t.data.copy() for t in pyramid])

    # numpy views in this repo's array conventions
    def fields(self):
        return dict(velocity=self.velocity.read.data[..., :2].copy(),
                    dye=self.dye.read.data.copy(),
                    pressure=self.pressure.read.data[..., 0].copy(),
                    divergence=self.divergence.data[..., 0].copy(),
                    curl=self.curl.data[..., 0].copy())

    def load(self, velocity=None, dye=None, pressure=None, divergence=None, curl=None):
        if velocity is not None: self.velocity.read.data[..., :2] = velocity
        if dye is not None: self.dye.read.data[...] = dye
        if pressure is not None: self.pressure.read.data[..., 0] = pressure
        if divergence is not None: self.divergence.data[..., 0] = divergence
        if curl is not None: self.curl.data[..., 0] = curl
