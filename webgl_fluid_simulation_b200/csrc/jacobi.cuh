// jacobi.cuh — the pressure loop of step() (script.js S:1259-1266, pressureShader S:868-890).
//
//   p'[i,j] = ((((p[c(i-1),j] + p[c(i+1),j]) + p[i,c(j-1)]) + p[i,c(j+1)]) - div[i,j]) * 0.25
//
// Three kernels, all bit-identical per cell to the reference's one-blit-per-sweep loop:
//   jacobi_scalar_kernel  any W,H; one cell per thread (tiny / odd-sized grids).
//   jacobi_sweep_kernel   W%4==0; one float4 per thread, shuffles for x-neighbours; ONE sweep
//                         per launch == the literal S:1262 loop (12 B/update of HBM traffic).
//   jacobi_tb_kernel<K>   K sweeps per launch (temporal blocking) as a register-streaming stencil:
//     * a warp owns a 128-column window (lane l holds columns 4l..4l+3 as one float4) and streams
//       upward over a chunk of rows; no __syncthreads anywhere, warps are independent;
//     * time level t (t = 0..K-1) lives in a rotating 3-row register window; when level 0 loads
//       row r, level t produces row r-t, so the K sweeps form a software pipeline in y;
//     * x-neighbours across lanes come from __shfl_up/down; the two outermost lanes of a level
//       hold garbage that creeps inward one column per level: HX = roundup(K,4) halo columns on
//       each side of the window are recomputed by the neighbouring window (overlapped tiling);
//     * CLAMP_TO_EDGE in x is an even reflection p[-1-m] = p[m]; because L+R is the first add of
//       the expression (commutative) the reflected columns evolve bit-identically, so mirroring
//       the LOADS implements the boundary for all K levels.  In y the expression is not
//       symmetric in B/T, so rows 0 and H-1 select "below := centre" / "above := centre"
//       explicitly (EDGE instantiation, only on the few steps that touch those rows);
//     * div[r] is needed by level t when it produces row r, i.e. K times at K different steps:
//       each lane parks its float4 of the row in a shared-memory ring it alone reads back
//       (conflict-free LDS.128; smem is used as a software-managed per-lane spill ring).  Every
//       row is written twice, RING slots apart, so all K reads use one base register plus a
//       compile-time immediate offset;
//     * p / div rows are prefetched three steps ahead into registers (coalesced 512 B per warp
//       row segment, LDG.128).
//   The optional SCALE template fuses the clear pass (S:1253-1257, p <- PRESSURE*p) into the
//   level-0 load of the first launch: one fp32 multiply, same rounding as the separate blit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fk {

struct JacobiArgs {
    const float* pin;    // local row 0 of the source pressure buffer
    const float* div;    // local row 0 of divergence
    float* pout;         // local row 0 of the destination pressure buffer
    int W, H;            // GLOBAL grid size (clamp rules refer to it)
    int row_off;         // global row index of local row 0 (0 on a single GPU)
    int out_lo, out_hi;  // global rows [out_lo, out_hi) to produce
    int rows_per_chunk;  // tb kernel: output rows per warp stream
    float scale;         // SCALE: value of config.PRESSURE
};

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) jacobi_scalar_kernel(JacobiArgs a, bool do_scale) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = a.out_lo + blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= a.W || j >= a.out_hi) return;
    const int il = max(i - 1, 0), ir = min(i + 1, a.W - 1);
    const int jb = max(j - 1, 0) - a.row_off, jt = min(j + 1, a.H - 1) - a.row_off;
    const int jl = j - a.row_off;
    const float s = do_scale ? a.scale : 1.0f;
    float L = a.pin[(size_t)jl * a.W + il], R = a.pin[(size_t)jl * a.W + ir];
    float B = a.pin[(size_t)jb * a.W + i], T = a.pin[(size_t)jt * a.W + i];
    if (do_scale) { L = s * L; R = s * R; B = s * B; T = s * T; }
    a.pout[(size_t)jl * a.W + i] = ((((L + R) + B) + T) - a.div[(size_t)jl * a.W + i]) * 0.25f;
}

// ------------------------------------------------------------------------------------------------
// One sweep, one float4 per thread.  blockDim = (32 * WARPS_X, ROWS): a warp spans 128 columns.
template <bool SCALE>
__global__ void __launch_bounds__(256) jacobi_sweep_kernel(JacobiArgs a) {
    const int W4 = a.W >> 2;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;  // float4 column group
    const int j = a.out_lo + blockIdx.y * blockDim.y + threadIdx.y;
    const bool live = (g < W4) && (j < a.out_hi);
    const int gc = min(g, W4 - 1), jc = min(j, a.out_hi - 1);
    const int jl = jc - a.row_off;
    const int jb = max(jc - 1, 0) - a.row_off, jt = min(jc + 1, a.H - 1) - a.row_off;
    const float4* P = reinterpret_cast<const float4*>(a.pin);
    float4 c = __ldg(P + (size_t)jl * W4 + gc);
    float4 b = __ldg(P + (size_t)jb * W4 + gc);
    float4 t = __ldg(P + (size_t)jt * W4 + gc);
    const float4 d = __ldg(reinterpret_cast<const float4*>(a.div) + (size_t)jl * W4 + gc);
    if (SCALE) {
        const float s = a.scale;
        c.x = s * c.x; c.y = s * c.y; c.z = s * c.z; c.w = s * c.w;
        b.x = s * b.x; b.y = s * b.y; b.z = s * b.z; b.w = s * b.w;
        t.x = s * t.x; t.y = s * t.y; t.z = s * t.z; t.w = s * t.w;
    }
    float l = __shfl_up_sync(0xffffffffu, c.w, 1);
    float r = __shfl_down_sync(0xffffffffu, c.x, 1);
    const int lane = threadIdx.x & 31;
    if (lane == 0) {   // warp edge: fetch the neighbour (or clamp at the wall)
        float e = (gc > 0) ? __ldg(a.pin + (size_t)jl * a.W + 4 * gc - 1) : 0.0f;
        l = (gc > 0) ? (SCALE ? a.scale * e : e) : c.x;
    }
    if (lane == 31 || gc == W4 - 1) {
        float e = (gc < W4 - 1) ? __ldg(a.pin + (size_t)jl * a.W + 4 * gc + 4) : 0.0f;
        r = (gc < W4 - 1) ? (SCALE ? a.scale * e : e) : c.w;
    }
    float4 o;
    o.x = ((((l + c.y) + b.x) + t.x) - d.x) * 0.25f;
    o.y = ((((c.x + c.z) + b.y) + t.y) - d.y) * 0.25f;
    o.z = ((((c.y + c.w) + b.z) + t.z) - d.z) * 0.25f;
    o.w = ((((c.z + r) + b.w) + t.w) - d.w) * 0.25f;
    if (live) reinterpret_cast<float4*>(a.pout)[(size_t)jl * W4 + gc] = o;
}

// ------------------------------------------------------------------------------------------------
template <int K>
struct TB {
    static constexpr int HX = (K + 3) / 4 * 4;       // x halo (columns) on each side of a window
    static constexpr int VALID = 128 - 2 * HX;       // columns a window produces
    static constexpr int RING = K + 1;               // div ring slots (each row stored twice)
    static constexpr int WARPS = 1;                  // one warp per CTA (see jacobi_tb_kernel)
    static constexpr int SMEM_PER_WARP = 2 * RING * 32 * (int)sizeof(float4);
    static constexpr int SMEM = WARPS * SMEM_PER_WARP;
};

__device__ __forceinline__ float4 rev4(float4 v, bool rev) {
    return rev ? make_float4(v.w, v.z, v.y, v.x) : v;
}

// one Jacobi update of a float4 of row `c`, with rows below/above and the shuffled neighbours
__device__ __forceinline__ float4 jacobi4(const float4 below, const float4 c, const float4 above,
                                          const float4 d) {
    const float l = __shfl_up_sync(0xffffffffu, c.w, 1);
    const float r = __shfl_down_sync(0xffffffffu, c.x, 1);
    float4 o;
    o.x = ((((l + c.y) + below.x) + above.x) - d.x) * 0.25f;
    o.y = ((((c.x + c.z) + below.y) + above.y) - d.y) * 0.25f;
    o.z = ((((c.y + c.w) + below.z) + above.z) - d.z) * 0.25f;
    o.w = ((((c.z + r) + below.w) + above.w) - d.w) * 0.25f;
    return o;
}

// One triple of pipeline steps (phases 0,1,2 of the 3-slot window rotation).  EDGE instantiates
// the wall selects; the caller picks the instantiation with a warp-uniform branch, so the
// steady-state code carries no select at all.
template <int K, bool SCALE, bool EDGE>
__device__ __forceinline__ void tb_triple(float4 (&w)[K][3], float4 (&pf)[3], float4 (&df)[3],
                                          float4* __restrict__ ring, int& slot,
                                          const float4* __restrict__ Pg, const float4* __restrict__ Dg,
                                          float4* __restrict__ Og, const int W4, const int ys,
                                          const int ye, const int y0, const int y1, const int H,
                                          const int s0, const bool rev, const bool lane_out,
                                          const float scale) {
    using T = TB<K>;
#pragma unroll
    for (int ph = 0; ph < 3; ++ph) {
        const int s = s0 + ph;
        // ---- level 0: take the prefetched row, start the prefetch three rows ahead ---------------
        float4 in = rev4(pf[ph], rev);
        if (SCALE) {
            in.x = scale * in.x; in.y = scale * in.y; in.z = scale * in.z; in.w = scale * in.w;
        }
        const float4 dv = rev4(df[ph], rev);
        {
            const int r = min(ys + s + 3, ye);
            pf[ph] = __ldg(Pg + (size_t)r * W4);
            df[ph] = __ldg(Dg + (size_t)r * W4);
        }
        w[0][(ph + 2) % 3] = in;
        // park div row (ys+s) in the ring, twice (slot and slot+RING)
        ring[slot * 32] = dv;
        ring[(slot + T::RING) * 32] = dv;
        const float4* rbase = ring + (slot + T::RING) * 32;   // row (ys+s-t) is at rbase[-t*32]
        // ---- levels 1..K ------------------------------------------------------------------------
#pragma unroll
        for (int t = 1; t <= K; ++t) {
            const int r = ys + s - t;             // row produced by level t at this step
            const float4 c = w[t - 1][(ph + 1) % 3];
            float4 below = w[t - 1][(ph + 0) % 3];
            float4 above = w[t - 1][(ph + 2) % 3];
            if (EDGE) {                           // warp-uniform conditions
                if (r == 0) below = c;            // CLAMP_TO_EDGE: p[i,-1] = p[i,0]
                if (r == H - 1) above = c;        //                p[i,H]  = p[i,H-1]
            }
            const float4 d = rbase[-t * 32];
            const float4 o = jacobi4(below, c, above, d);
            if (t < K) {
                w[t][(ph + 2) % 3] = o;
            } else if (lane_out && r >= y0 && r < y1) {
                Og[(size_t)r * W4] = o;
            }
        }
        slot = (slot + 1 == T::RING) ? 0 : slot + 1;
    }
}

// One warp per CTA: every quantity that steers control flow derives from blockIdx and kernel
// arguments only, so the compiler can prove the warp converged at each shuffle (no WARPSYNC /
// BSSY scaffolding) and keeps loop state in uniform registers.
template <int K, bool SCALE>
__global__ void __launch_bounds__(32) jacobi_tb_kernel(JacobiArgs a) {
    using T = TB<K>;
    extern __shared__ float4 smem4[];
    const int lane = threadIdx.x;
    const int nxw = (a.W + T::VALID - 1) / T::VALID;
    const int wid = blockIdx.x;
    const int wx = wid % nxw, cy = wid / nxw;

    // ---- x geometry of this lane: mirrored / clamped float4 column group ------------------------
    const int W = a.W, H = a.H;
    const int gx = wx * T::VALID - T::HX + 4 * lane;  // first global column of this lane
    int lc = gx;
    bool rev = false;
    if (gx < 0) { lc = -gx - 4; rev = true; }                 // p[-1-m] = p[m]
    else if (gx >= W) { lc = 2 * W - 4 - gx; rev = true; }    // p[W+m]  = p[W-1-m]
    lc = min(max(lc, 0), W - 4);
    const bool lane_out = (lane >= T::HX / 4) && (lane < 32 - T::HX / 4) && (gx >= 0) && (gx < W);

    // ---- y geometry of this warp's stream ----------------------------------------------------------
    const int y0 = a.out_lo + cy * a.rows_per_chunk;
    const int y1 = min(y0 + a.rows_per_chunk, a.out_hi);
    const int ys = max(y0 - K, 0);                    // first input row
    const int ye = min(y1 - 1 + K, H - 1);            // last input row (loads clamp to it)
    const int nsteps = y1 - ys + K;                   // level K emits row ys+s-K at step s

    const int W4 = W >> 2;
    const ptrdiff_t base = -(ptrdiff_t)a.row_off * W4;
    const float4* Pg = reinterpret_cast<const float4*>(a.pin) + base + (lc >> 2);
    const float4* Dg = reinterpret_cast<const float4*>(a.div) + base + (lc >> 2);
    float4* Og = reinterpret_cast<float4*>(a.pout) + base + (lane_out ? (gx >> 2) : 0);

    float4* ring = smem4 + lane;                      // slot k at ring[k*32]

    // rotating windows: w[t][(ph+0)%3] = row r-1, [(ph+1)%3] = row r, [(ph+2)%3] = fresh row r+1
    float4 w[K][3];
#pragma unroll
    for (int t = 0; t < K; ++t)
#pragma unroll
        for (int q = 0; q < 3; ++q) w[t][q] = make_float4(0.f, 0.f, 0.f, 0.f);

    // prefetch buffers, three rows ahead
    float4 pf[3], df[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const int r = min(ys + q, ye);
        pf[q] = __ldg(Pg + (size_t)r * W4);
        df[q] = __ldg(Dg + (size_t)r * W4);
    }

    int slot = 0;                                     // ring slot of the row loaded at this step
    for (int s0 = 0; s0 < nsteps; s0 += 3) {
        // does any row this triple touches lie on the bottom / top wall?
        const int lo = ys + s0 - K, hi = ys + s0 + 2;
        if ((lo <= 0) || (hi >= H - 1))
            tb_triple<K, SCALE, true>(w, pf, df, ring, slot, Pg, Dg, Og, W4, ys, ye, y0, y1, H, s0,
                                      rev, lane_out, a.scale);
        else
            tb_triple<K, SCALE, false>(w, pf, df, ring, slot, Pg, Dg, Og, W4, ys, ye, y0, y1, H, s0,
                                       rev, lane_out, a.scale);
    }
}

}  // namespace fk
