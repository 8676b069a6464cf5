/*
 * fluid_oracle.h — CPU restatement of the simulation path of WebGL-Fluid-Simulation.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this.  The product (libfluid_b200) never links or calls it.
 *
 * PARITY UNPINNED BY THE REFERENCE: the reference ships no tests, fixtures or golden vectors
 * and cannot be executed in this image (needs a browser DOM + WebGL; no node, no GL).  The
 * restatement is pinned instead against (a) hand-derived known-answer tests and (b) golden
 * vectors produced by oracle/glsl_exec.py, which EXECUTES the reference's GLSL shader source
 * text (read from /root/reference/script.js at generation time) under an emulated GL sampler.
 *
 * "S:n" = /root/reference/script.js line n.  All fields row-major, x fastest, row 0 = bottom.
 * Arithmetic is IEEE fp32 with no contraction (compile with -ffp-contract=off).
 */
#ifndef FLUID_ORACLE_H
#define FLUID_ORACLE_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

/* curlShader S:814-833 (geometry: baseVertexShader S:440-459, CLAMP_TO_EDGE S:1051-1052) */
void oracle_curl(const float* v, float* curl, int W, int H);
/* vorticityShader S:835-866 */
void oracle_vorticity(const float* v, const float* curl, float* vout, int W, int H, float curl_k,
                      float dt);
/* divergenceShader S:786-812 */
void oracle_divergence(const float* v, float* div, int W, int H);
/* clearShader S:508-519 as used at S:1253-1257 (value = config.PRESSURE), C channels */
void oracle_clear(const float* in, float* out, size_t n, float value);
/* pressureShader S:868-890: ONE Jacobi sweep */
void oracle_jacobi(const float* p, const float* div, float* pout, int W, int H);
/* the loop S:1259-1266: result is left in p (tmp is scratch of the same size) */
void oracle_copy_rows(float* dst, const float* src, int W, int H);
void oracle_jacobi_iters(float* p, float* tmp, const float* div, int W, int H, int iters);
/* gradientSubtractShader S:892-913 */
void oracle_gradient_subtract(const float* p, const float* v, float* vout, int W, int H);
/* advectionShader S:746-784 (the MANUAL_FILTERING bilerp S:758-770 is taken as the definition of
 * bilinear filtering).  vel is W x H float2 (texelSize = 1/W,1/H — always the SIM texel, S:1276);
 * src/out are Ws x Hs with C channels (dyeTexelSize = 1/Ws,1/Hs, S:1277-1278, S:1287-1288). */
void oracle_advect(const float* vel, int W, int H, const float* src, float* out, int Ws, int Hs,
                   int C, float dt, float dissipation);
/* splatShader S:726-744: out.xyz = base.xyz + exp(-dot(p,p)/radius)*color, 4th channel (if C==4)
 * := 1.  For the RG velocity target only .xy survive (C==2). */
void oracle_splat(const float* base, float* out, int W, int H, int C, float aspect, float px,
                  float py, const float* color3, float radius);
/* copyShader S:496-506 through a LINEAR sampler, as resizeFBO S:1108-1114 uses it */
void oracle_resample(const float* src, int Ws, int Hs, float* dst, int Wd, int Hd, int C);

/* render() without post-FX: drawColor + drawDisplay (S:1296-1348, displayShaderSource S:549-612) */
/* bg_mode: 0 drawColor(BACK_COLOR), 1 drawCheckerboard (aspect = canvas.width / canvas.height), 2 none / no blend */
void oracle_display(const float* dye, int Wd, int Hd, float* out, int w, int h, int shading,
                    const float* back_rgb, int bg_mode, float aspect);

/* post-FX chain: bloom S:614-674 + S:1350-1394, sunrays S:676-724 +
 * S:1396-1419, full display S:549-612 */
void oracle_bloom_prefilter(const float* dye, int Wd, int Hd, float* out, int w, int h, float curve0,
                            float curve1, float curve2, float threshold);
void oracle_box4(const float* src, int ws, int hs, float* dst, int w, int h, float scale, int add);
void oracle_sunrays_mask(const float* dye, float* mask, int Wd, int Hd);
void oracle_sunrays(const float* mask, int Wm, int Hm, float* out, int w, int h, float weight);
void oracle_blur3(const float* src, float* dst, int w, int h, float tsx, float tsy);
void oracle_display_full(const float* dye, int Wd, int Hd, const float* bloom, int bw, int bh,
                         const float* sun, int sw, int sh, const float* dither, int dw, int dh,
                         float* out, int w, int h, const float* back_rgb, int bg_mode, float aspect);

/* fp16 storage emulation (S:138-147, S:986-1006): round every element through IEEE half, RNE */
void oracle_round_half(float* a, size_t n);

/* ---- a whole simulation, mirroring include/fluid.h ------------------------------------------- */
typedef struct oracle_sim {
    int W, H, Wd, Hd;
    float density_dissipation, velocity_dissipation, pressure, curl, splat_radius, aspect;
    int pressure_iterations;
    int half_storage; /* 1: round every pass output through fp16 like the reference's textures */
    float *v, *v2, *dye, *dye2, *p, *p2, *div, *curl_f;
} oracle_sim;

oracle_sim* oracle_sim_create(int W, int H, int Wd, int Hd);
void oracle_sim_destroy(oracle_sim* s);
void oracle_sim_step(oracle_sim* s, float dt);                      /* S:1231-1294 */
void oracle_sim_splat(oracle_sim* s, float x, float y, float dx, float dy, float r, float g,
                      float b);                                    /* S:1441-1462 */
int oracle_num_threads(void);
void oracle_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
