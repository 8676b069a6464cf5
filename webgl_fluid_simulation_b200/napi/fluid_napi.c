/*
 * fluid_napi.c — thin N-API shim: exposes the C ABI of include/fluid.h to Node as `fluid.node`.
 * js/fluid-sim.js builds the reference's step()/splat()/config surface on top of it.
 *
 * COMPILE-CHECKED ONLY here (gcc -fsyntax-only in build()): this image has no Node, so the addon
 * cannot be linked against node or loaded.  With a Node toolchain:
 *   gcc -shared -fPIC -DFLUID_USE_SYSTEM_NODE_API -I<node>/include/node -I../../include \
 *       fluid_napi.c -L.. -lfluid_b200 -Wl,-rpath,'$ORIGIN/..' -o fluid.node
 *
 * Exports: create(cfg) -> handle, destroy(h), step(h, dt), splat(h, x, y, dx, dy, r, g, b),
 *          setParam(h, key, value), resize(h, sw, sh, dw, dh), read(h, field) -> Float32Array,
 *          write(h, field, Float32Array), sync(h), dims(h, field) -> {width,height,channels,row0},
 *          getResolution(res, cw, ch) -> {width,height}.
 * Every non-zero status from the C ABI becomes a thrown JS Error carrying fluid_last_error().
 */
#ifdef FLUID_USE_SYSTEM_NODE_API
#include <node_api.h>
#else
#include "node_api_min.h"
#endif
#include <stdlib.h>
#include <string.h>

#include "../../include/fluid.h"

#define ARGS(n)                                                   \
    size_t argc = (n); napi_value argv[(n) > 0 ? (n) : 1];        \
    if (napi_get_cb_info(env, info, &argc, argv, NULL, NULL) != napi_ok || argc < (n)) { \
        napi_throw_error(env, NULL, "wrong number of arguments"); return NULL; }

static napi_value undefined(napi_env env) { napi_value u; napi_get_undefined(env, &u); return u; }

static fluid_t* handle_of(napi_env env, napi_value v) {
    void* p = NULL;
    if (napi_get_value_external(env, v, &p) != napi_ok || !p) napi_throw_error(env, NULL, "not a fluid handle");
    return (fluid_t*)p;
}

static int check(napi_env env, fluid_t* h, int rc) {
    if (rc == FLUID_OK) return 1;
    const char* m = fluid_last_error(h);
    napi_throw_error(env, NULL, (m && *m) ? m : "libfluid_b200 call failed");
    return 0;
}

static double num(napi_env env, napi_value v) { double d = 0; napi_get_value_double(env, v, &d); return d; }

static double prop(napi_env env, napi_value obj, const char* name, double dflt) {
    bool has = false; napi_value v;
    if (napi_has_named_property(env, obj, name, &has) != napi_ok || !has) return dflt;
    if (napi_get_named_property(env, obj, name, &v) != napi_ok) return dflt;
    return num(env, v);
}

static void finalize_handle(napi_env env, void* data, void* hint) { (void)env; (void)hint; fluid_destroy((fluid_t*)data); }

/* create({simWidth, simHeight, dyeWidth, dyeHeight, aspect, device, flags, jacobiBlock, ...config keys}) */
static napi_value Create(napi_env env, napi_callback_info info) {
    ARGS(1)
    fluid_config c; fluid_config_default(&c);
    c.sim_w = (int)prop(env, argv[0], "simWidth", c.sim_w);   c.sim_h = (int)prop(env, argv[0], "simHeight", c.sim_h);
    c.dye_w = (int)prop(env, argv[0], "dyeWidth", c.dye_w);   c.dye_h = (int)prop(env, argv[0], "dyeHeight", c.dye_h);
    c.density_dissipation = (float)prop(env, argv[0], "DENSITY_DISSIPATION", c.density_dissipation);
    c.velocity_dissipation = (float)prop(env, argv[0], "VELOCITY_DISSIPATION", c.velocity_dissipation);
    c.pressure = (float)prop(env, argv[0], "PRESSURE", c.pressure);
    c.pressure_iterations = (int)prop(env, argv[0], "PRESSURE_ITERATIONS", c.pressure_iterations);
    c.curl = (float)prop(env, argv[0], "CURL", c.curl);
    c.splat_radius = (float)prop(env, argv[0], "SPLAT_RADIUS", c.splat_radius);
    c.aspect = (float)prop(env, argv[0], "aspect", c.aspect);
    c.device = (int)prop(env, argv[0], "device", -1);
    c.flags = (uint32_t)prop(env, argv[0], "flags", 0);
    c.jacobi_block = (int)prop(env, argv[0], "jacobiBlock", 0);
    fluid_t* h = NULL;
    if (!check(env, NULL, fluid_create(&c, &h))) return NULL;
    napi_value ext;
    napi_create_external(env, h, finalize_handle, NULL, &ext);
    return ext;
}

static napi_value Step(napi_env env, napi_callback_info info) {
    ARGS(2)
    fluid_t* h = handle_of(env, argv[0]); if (!h) return NULL;
    check(env, h, fluid_step(h, (float)num(env, argv[1])));
    return undefined(env);
}

static napi_value Splat(napi_env env, napi_callback_info info) {
    ARGS(8)
    fluid_t* h = handle_of(env, argv[0]); if (!h) return NULL;
    float a[7]; for (int i = 0; i < 7; ++i) a[i] = (float)num(env, argv[i + 1]);
    check(env, h, fluid_splat(h, a[0], a[1], a[2], a[3], a[4], a[5], a[6]));
    return undefined(env);
}

static napi_value SetParam(napi_env env, napi_callback_info info) {
    ARGS(3)
    fluid_t* h = handle_of(env, argv[0]); if (!h) return NULL;
    int32_t key = 0; napi_get_value_int32(env, argv[1], &key);
    check(env, h, fluid_set_param(h, key, (float)num(env, argv[2])));
    return undefined(env);
}

static napi_value Resize(napi_env env, napi_callback_info info) {
    ARGS(5)
    fluid_t* h = handle_of(env, argv[0]); if (!h) return NULL;
    int32_t v[4]; for (int i = 0; i < 4; ++i) napi_get_value_int32(env, argv[i + 1], &v[i]);
    check(env, h, fluid_resize(h, v[0], v[1], v[2], v[3]));
    return undefined(env);
}

static napi_value Read(napi_env env, napi_callback_info info) {
    ARGS(2)
    fluid_t* h = handle_of(env, argv[0]); if (!h) return NULL;
    int32_t field = 0; napi_get_value_int32(env, argv[1], &field);
    const size_t n = fluid_field_elems(h, field);
    if (!n) { napi_throw_error(env, NULL, "bad field id"); return NULL; }
    void* data = NULL; napi_value ab, ta;
    if (napi_create_arraybuffer(env, n * sizeof(float), &data, &ab) != napi_ok) return NULL;
    if (!check(env, h, fluid_read(h, field, (float*)data, n))) return NULL;
    napi_create_typedarray(env, napi_float32_array, n, ab, 0, &ta);
    return ta;
}

static napi_value Write(napi_env env, napi_callback_info info) {
    ARGS(3)
    fluid_t* h = handle_of(env, argv[0]); if (!h) return NULL;
    int32_t field = 0; napi_get_value_int32(env, argv[1], &field);
    napi_typedarray_type t; size_t len = 0; void* data = NULL; napi_value ab; size_t off = 0;
    if (napi_get_typedarray_info(env, argv[2], &t, &len, &data, &ab, &off) != napi_ok || t != napi_float32_array) {
        napi_throw_error(env, NULL, "write() takes a Float32Array"); return NULL; }
    check(env, h, fluid_write(h, field, (const float*)data, len));
    return undefined(env);
}

static napi_value Sync(napi_env env, napi_callback_info info) {
    ARGS(1)
    fluid_t* h = handle_of(env, argv[0]); if (!h) return NULL;
    check(env, h, fluid_sync(h));
    return undefined(env);
}

static napi_value set_int(napi_env env, napi_value obj, const char* k, int v) {
    napi_value n; napi_create_int32(env, v, &n);
    napi_property_descriptor d = {k, NULL, NULL, NULL, NULL, n, napi_enumerable, NULL};
    napi_define_properties(env, obj, 1, &d);
    return obj;
}

static napi_value Dims(napi_env env, napi_callback_info info) {
    ARGS(2)
    fluid_t* h = handle_of(env, argv[0]); if (!h) return NULL;
    int32_t field = 0; napi_get_value_int32(env, argv[1], &field);
    int w, r, c, r0;
    if (!check(env, h, fluid_field_dims(h, field, &w, &r, &c, &r0))) return NULL;
    void* scratch; napi_value obj;                  /* any object will do as the property bag */
    napi_create_arraybuffer(env, 0, &scratch, &obj);
    set_int(env, obj, "width", w); set_int(env, obj, "height", r);
    set_int(env, obj, "channels", c); set_int(env, obj, "row0", r0);
    return obj;
}

static napi_value GetResolution(napi_env env, napi_callback_info info) {
    ARGS(3)
    int32_t v[3]; for (int i = 0; i < 3; ++i) napi_get_value_int32(env, argv[i], &v[i]);
    int w, hh; fluid_get_resolution(v[0], v[1], v[2], &w, &hh);
    void* scratch; napi_value obj;
    napi_create_arraybuffer(env, 0, &scratch, &obj);
    set_int(env, obj, "width", w); set_int(env, obj, "height", hh);
    return obj;
}

static napi_value Destroy(napi_env env, napi_callback_info info) {
    (void)info; /* lifetime is tied to the external's finalizer; explicit destroy is a no-op hint */
    return undefined(env);
}

NAPI_MODULE_INIT() {
    const napi_property_descriptor props[] = {
        {"create", NULL, Create, NULL, NULL, NULL, napi_enumerable, NULL},
        {"destroy", NULL, Destroy, NULL, NULL, NULL, napi_enumerable, NULL},
        {"step", NULL, Step, NULL, NULL, NULL, napi_enumerable, NULL},
        {"splat", NULL, Splat, NULL, NULL, NULL, napi_enumerable, NULL},
        {"setParam", NULL, SetParam, NULL, NULL, NULL, napi_enumerable, NULL},
        {"resize", NULL, Resize, NULL, NULL, NULL, napi_enumerable, NULL},
        {"read", NULL, Read, NULL, NULL, NULL, napi_enumerable, NULL},
        {"write", NULL, Write, NULL, NULL, NULL, napi_enumerable, NULL},
        {"sync", NULL, Sync, NULL, NULL, NULL, napi_enumerable, NULL},
        {"dims", NULL, Dims, NULL, NULL, NULL, napi_enumerable, NULL},
        {"getResolution", NULL, GetResolution, NULL, NULL, NULL, napi_enumerable, NULL},
    };
    napi_define_properties(env, exports, sizeof props / sizeof props[0], props);
    return exports;
}
