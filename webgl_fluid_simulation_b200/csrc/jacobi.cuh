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
//       each lane keeps its float4 of the last K+3 rows in a register ring (see tb_block);
//     * p / div rows are staged D rows ahead into a per-warp shared-memory ring, so the few
//       resident warps still keep enough bytes in flight for HBM.  Two interchangeable fills:
//       TMA — one elected lane issues cp.async.bulk (UBLKCP) of the window's 512 B row segment
//       with mbarrier complete_tx; or LDGSTS — every lane cp.asyncs its own 16 B.
//   The optional SCALE template fuses the clear pass (S:1253-1257, p <- PRESSURE*p) into the
//   level-0 load of the first launch: one fp32 multiply, same rounding as the separate blit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fk {

// ---- cross-GPU hand-offs of a slab solve (peer-memory transport; everything null on one GPU) ------
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t;
}
// bounded spin: a neighbour that never arrives costs 4 s and an error flag, never a hung GPU
__device__ __forceinline__ void spin_until(const unsigned* flag, unsigned seq, int* err) {
    const unsigned long long t0 = global_ns();
    while ((int)(ld_acquire_sys(flag) - seq) < 0) {
        if (global_ns() - t0 > 4000000000ull) { *err = 2; break; }
        __nanosleep(100);
    }
}

// The LAST blocked launch of a solve stores the rows its neighbours will need as ghost rows of the
// NEXT solve straight into their arenas (NVLink stores from the compute kernel: no separate message,
// no exchange kernel between two solves), guarded by two flag hand-offs:
//   * "done reading": the neighbour's PENULTIMATE launch was the last reader of the ghost rows about
//     to be overwritten (same ping-pong buffer); its last CTA to finish release-stores the solve's
//     sequence number into my flag word, and my mirroring warps acquire it first;
//   * "mirror ready": my last CTA to finish release-stores the sequence number into the neighbour's
//     flag word; the streams of its next first launch that read ghost rows acquire it first.
// The same wait for a whole (converged, one-warp-CTA) warp: lane 0 polls, a vote carries the outcome — one
// request per poll instead of 32, and no lane-dependent branch (see tb_mirror for why that matters here).
// Each lane then performs its own acquire load of the flag before it touches what the flag guards.
__device__ __forceinline__ void warp_spin_until(const unsigned* flag, unsigned seq, int* err) {
    const unsigned long long t0 = global_ns();
    const unsigned l0 = (threadIdx.x == 0) ? 1u : 0u;
    for (;;) {
        unsigned v = seq;
        asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q ld.acquire.sys.global.u32 %0, [%1]; }" : "+r"(v) : "l"(flag), "r"(l0) : "memory");
        if (!__any_sync(0xffffffffu, (int)(v - seq) < 0)) break;
        if (global_ns() - t0 > 4000000000ull) { *err = 2; break; }      // all lanes agree within a poll or two; the store is idempotent
        __nanosleep(200);
    }
    // every lane's own acquire of the (now satisfied) flag — not a fence: MEMBAR.SYS would first drain the rows
    // this warp has just stored (measured: 6-22 us per mirroring warp)
    (void)ld_acquire_sys(flag);
}

struct TbSync {
    const unsigned* pre_wait[2];   // [0] below, [1] above: streams reading rows outside [own_lo, own_hi) wait for *flag >= pre_seq
    unsigned pre_seq;
    int own_lo, own_hi;
    float* mirror[2];              // neighbour's copy of pout, offset so that GLOBAL row j starts at mirror + j*W
    int mir_lo[2], mir_hi[2];      // global rows [lo, hi) of mine the neighbour on that side keeps as ghost rows
    const unsigned* mir_wait[2];   // "done reading" words (local), value mir_seq
    unsigned mir_seq;
    unsigned* ticket;              // CTA completion counter (local); null: this launch signals nothing
    unsigned* done_flag[2];        // words in the neighbours' arenas that receive done_seq when the whole grid is done
    unsigned done_seq;
    // FLUID_DEBUG_HALO_TIMING: u64 sums — [0] ns streams waited for "mirror ready", [1] such streams, [2] ns mirroring
    // warps waited for "done reading", [3] ns they spent copying, [4] such warps, [5] / [6] ns and count of streams
    // touching ghost rows (whole life), [7] / [8] of the other streams
    unsigned long long* dbg;
};
__device__ __forceinline__ void dbg_add(unsigned long long* p, unsigned long long v, bool on) {   // predicated, no branch (see tb_mirror)
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q red.global.add.u64 [%0], %1; }" ::"l"(p), "l"(v), "r"((unsigned)on) : "memory");
}

struct JacobiArgs {
    const float* pin;    // local row 0 of the source pressure buffer
    const float* div;    // local row 0 of divergence
    float* pout;         // local row 0 of the destination pressure buffer
    int W, H;            // GLOBAL grid size (clamp rules refer to it)
    int row_off;         // global row index of local row 0 (0 on a single GPU)
    int out_lo, out_hi;  // global rows [out_lo, out_hi) to produce
    int rows_per_chunk;  // tb kernel: output rows per warp stream
    float scale;         // SCALE: value of config.PRESSURE
    int* err;            // device error word (3 = a bounded mbarrier wait gave up)
    const unsigned char* tiny_map;   // tb kernel: where divergence defeats the fma contraction (may be null)
    // tb kernel, optional SECOND row range of the same launch (the two boundary strips of a slab, which
    // wait for the halo while the interior runs): chunks 0..nch1-1 tile [out_lo, out_hi), the rest
    // tile [seg2_lo, seg2_hi).  nch1 <= 0: single range.
    int nch1, seg2_lo, seg2_hi;
};
// kernel argument block: the hand-offs exist only in the SYNC instantiation.  The single-GPU kernel must
// not pay for slab plumbing: with TbSync inside JacobiArgs and the null checks compiled into the one
// kernel, the 4096^2 solve measured 3 % slower (profiles/r02_scaling.md).  (A derived struct, not a
// separate kernel parameter: with the latter ptxas gave up the converged-warp assumption of the hot
// loops — shuffles with BRA.DIV scaffolding, a fifth copy of the block body.)
template <bool SYNC> struct TbArgs : JacobiArgs {};
template <> struct TbArgs<true> : JacobiArgs { TbSync sy; };

// the two tensor maps a temporally blocked launch reads through (TMA staging); 64-byte aligned
// CUtensorMap images, passed by value as a __grid_constant__ kernel parameter
struct __align__(64) TmapPair {
    unsigned char p[128];
    unsigned char d[128];
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
// Temporally blocked kernel, generation 7.
//
// Arithmetic.  The fp32 pipes deliver 32 lane-operations per cycle per SM sub-partition whether
// they are issued scalar (FADD/FFMA, 1 cycle) or packed (FADD2/FFMA2, 2 cycles) — measured,
// tools/ubench/fp32_pipe.cu — so the kernel's floor is LANE-OPERATIONS per update, and the packed
// forms only save issue slots.  The reference expression costs 5:  ((((L+R)+B)+T)-d)*0.25.  The
// last two contract into ONE fma with a pre-scaled divergence,
//        fma(S, 0.25, -0.25*d)  ==  (S - d) * 0.25      bit for bit,
// whenever d == 0 or |d| >= 2^-123 (and the result does not overflow):  -0.25*d is then exact;
// for |S-d| >= 2^-124 scaling by 2^-2 commutes with the rounding of S-d (normal results); for
// |S-d| < 2^-124 <= |d|/2 Sterbenz makes S-d exact, so both sides round the same real number once.
// A divergence value with 0 < |d| < 2^-123 (subnormal-range far fields of a Gaussian splat do
// produce them) can double-round differently, so the producers of `divergence` keep a coarse map —
// one byte per 128-column x 32-row cell — of where such values occur, and a warp whose footprint
// touches a flagged cell runs the EXACT instantiation (subtract, then multiply).  4 lane-ops per
// update instead of 5 everywhere else.
//
// Staging.  p / div rows reach a warp through a shared-memory ring, filled either
//   * TMA: by 2-D tensor-map TMA (cp.async.bulk.tensor.2d, SASS UTMALDG): one elected lane arms an
//     mbarrier and fetches a 128-column x 3-row box of p and of div per pipeline block; columns
//     outside the grid and rows outside the buffer are zero-filled by the hardware (never used
//     by a valid output); or
//   * LDGSTS: by per-lane 16-byte cp.async (the generation-5 path, kept selectable).
template <int K>
struct TB {
    static constexpr int HX = (K + 3) / 4 * 4;       // x halo (columns) on each side of a window
    static constexpr int VALID = 128 - 2 * HX;       // columns a window produces
    static constexpr int RD = K + 3;                 // div register-ring slots
    static constexpr int U = 3;                      // pipeline steps per unrolled block
    // LDGSTS ring: D rows of p + D rows of div, one float4 per lane per row
    static constexpr int D = 8;
    static constexpr int SMEM_LDGSTS = D * 2 * 32 * (int)sizeof(float4);
    // TMA ring: NS stages, each a 3-row box of p followed by a 3-row box of div (row = 512 B)
    static constexpr int NS = 4;
    static constexpr int STAGE4 = 2 * U * 32;        // float4 per stage
    static constexpr int SMEM_TMA = NS * STAGE4 * (int)sizeof(float4) + NS * 8;
};

// coarse map of "divergence has a value with 0 < |d| < 2^-123 here": one byte per cell
constexpr int TINY_CW = 128, TINY_CH = 32;
constexpr float TINY_DIV = 9.4039548e-38f;           // 2^-123
__host__ __device__ inline int tiny_map_w(int W) { return (W + TINY_CW - 1) / TINY_CW; }
__host__ __device__ inline int tiny_map_h(int H) { return (H + TINY_CH - 1) / TINY_CH; }
__device__ __forceinline__ bool is_tiny_div(float d) { return d != 0.0f && fabsf(d) < TINY_DIV; }

// ---- cp.async (LDGSTS) staging: global -> shared without passing through registers ------------
// Each lane copies, and later reads back, ONLY its own 16 bytes of a row, so the per-thread
// completion wait (cp.async.wait_group) is all the synchronisation the ring needs.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- TMA (2-D tensor map) + mbarrier ----------------------------------------------------------------
// Waits are bounded: a mis-armed barrier costs an error word, never a hung kernel.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(void* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, int x, int y, void* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(void* bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity, int* err) {
    if (mbar_try_wait(bar, parity)) return;              // the common case: the box landed long ago
#pragma unroll 1
    for (int tries = 0; tries < (1 << 16); ++tries)      // each try_wait already blocks for a HW time slice
        if (mbar_try_wait(bar, parity)) return;
    if (err) *err = 3;
}

__device__ __forceinline__ float4 lds128(const float4* p) {
    float4 v;
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}

__device__ __forceinline__ float4 rev4(float4 v, bool rev) {
    return rev ? make_float4(v.w, v.z, v.y, v.x) : v;
}

// Blackwell packed fp32 arithmetic (FADD2 / FMUL2 / FFMA2): two IEEE round-to-nearest results per
// issue slot, element-wise on the (x,y) and (z,w) halves of a float4.  The horizontal sums need the
// misaligned pair (y,z) and stay scalar.  Rounding is per element and identical to the scalar form.
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float a, float b) {
    u64 v; asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(a), "f"(b)); return v;
}
__device__ __forceinline__ void unpack2(u64 v, float& a, float& b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 c; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b)); return c;
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
    u64 c; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b)); return c;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 c; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b)); return c;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}

// one Jacobi update of a float4 of row `c`, with rows below/above and the shuffled neighbours.
//   EXACT:  o = ((((L + R) + below) + above) - d) * 0.25          (d = divergence)
//   else :  o = fma(((L + R) + below) + above, 0.25, d)           (d = -0.25 * divergence)
template <bool EXACT>
__device__ __forceinline__ float4 jacobi4(const float4 below, const float4 c, const float4 above,
                                          const float4 d) {
    const float l = __shfl_up_sync(0xffffffffu, c.w, 1);
    const float r = __shfl_down_sync(0xffffffffu, c.x, 1);
    const u64 q = pack2(0.25f, 0.25f);
    u64 lo = pack2(l + c.y, c.x + c.z);
    u64 hi = pack2(c.y + c.w, c.z + r);
    lo = add2(lo, pack2(below.x, below.y));
    hi = add2(hi, pack2(below.z, below.w));
    lo = add2(lo, pack2(above.x, above.y));
    hi = add2(hi, pack2(above.z, above.w));
    if (EXACT) {
        lo = sub2(lo, pack2(d.x, d.y));
        hi = sub2(hi, pack2(d.z, d.w));
        lo = mul2(lo, q);
        hi = mul2(hi, q);
    } else {
        lo = fma2(lo, q, pack2(d.x, d.y));
        hi = fma2(hi, q, pack2(d.z, d.w));
    }
    float4 o;
    unpack2(lo, o.x, o.y);
    unpack2(hi, o.z, o.w);
    return o;
}

// Per-warp stream state that survives across unrolled blocks.
struct TBStream {
    float4* op;          // where level K's row of THIS step goes (advances one row per step)
    int rout;            // global row index op points at  (= ys + s - K)
    // LDGSTS variant
    const float4* pl;    // next p row to stage (this lane's float4 column group)
    const float4* dl;    // next div row to stage
    const float4* stage; // this lane's cell of the staging slot consumed at this step
    int rload;           // global row index pl/dl point at
    int slot;            // staging slot consumed at this step (0..D-1)
    // TMA variant
    int stg;             // stage consumed by this block (0..NS-1)
    unsigned phase;      // mbarrier parity of that stage
    int yfill;           // local row coordinate of the next box to fetch
};

struct TBTma {           // warp-uniform TMA bookkeeping
    const void* tm_p;    // tensor map of the source pressure buffer
    const void* tm_d;    // tensor map of divergence
    unsigned long long* bars;   // NS mbarriers
    float4* slots;       // stage q at slots + q * STAGE4
    int x0;              // first (possibly negative) column of the window
    int lane;
    int* err;
};

// One unrolled block of 3 pipeline steps.
//   * p windows: 3 register slots per level, rotating with period 3 == the unroll, so they are
//     addressed with compile-time constants and never moved;
//   * div: level t needs div[r] when it produces row r, i.e. at K different steps.  The last K+3
//     rows live in the register ring dr[]: the row staged at phase ph sits in dr[K+ph], level t
//     reads dr[K+ph-t], and the ring is shifted down by 3 at the end of the block (MOVs issue on
//     the ALU pipe, beside the fp32 pipe that bounds the kernel);
//   * EDGE instantiates the wall selects in y, REV the mirrored-lane reversal in x, EXACT the
//     un-contracted tail; all are chosen by warp-uniform branches OUTSIDE the steady-state loop.
template <int K, bool SCALE, bool EDGE, bool REV, bool EXACT, bool TMA>
__device__ __forceinline__ void tb_block(float4 (&w)[K][3], float4 (&dr)[TB<K>::RD], TBStream& st,
                                         const float4* __restrict__ ring, const TBTma& tm, const int W4,
                                         const int ye, const int y0, const int y1, const int H,
                                         const bool rev, const bool lane_out, const float scale) {
    using T = TB<K>;
    const float4* srow = ring;
    if (TMA) {
        mbar_wait(tm.bars + st.stg, st.phase, tm.err);
        srow = ring + st.stg * T::STAGE4;
    }
    // The K levels x 3 steps of a block form a dependency lattice: node (ph, t) needs node
    // (ph, t-1) of the same step (the fresh row above) and rows of level t-1 from the two previous
    // steps.  Source order is step by step; ptxas interleaves the chains itself.  (Emitting the
    // lattice anti-diagonal by anti-diagonal — FLUID_TB_DIAG, legal because ph ascends inside a
    // diagonal — was measured 3 % SLOWER on B200: profiles/r02_jacobi_layout.md.)
#ifdef FLUID_TB_DIAG
#pragma unroll
    for (int dg = 0; dg < K + 3; ++dg) {
#pragma unroll
        for (int ph = 0; ph < 3; ++ph) {
            const int t = dg - ph;
            if (t < 0 || t > K) continue;
#else
#pragma unroll
    for (int ph = 0; ph < 3; ++ph) {
#pragma unroll
        for (int t = 0; t <= K; ++t) {
#endif
            if (t == 0) {
                // ---- level 0: the oldest staged row ---------------------------------------------------
                float4 in, dv;
                if (TMA) {
                    in = lds128(srow + ph * 32);
                    dv = lds128(srow + (3 + ph) * 32);
                } else {
                    cp_async_wait<T::D - 1>();
                    in = lds128(st.stage);
                    dv = lds128(st.stage + T::D * 32);
                    cp_async16(const_cast<float4*>(st.stage), st.pl);
                    cp_async16(const_cast<float4*>(st.stage) + T::D * 32, st.dl);
                    cp_async_commit();
                    if (st.rload < ye) { ++st.rload; st.pl += W4; st.dl += W4; }   // loads clamp to row ye
                    st.slot = (st.slot + 1 == T::D) ? 0 : st.slot + 1;
                    st.stage = ring + st.slot * 32;
                }
                if (REV) { in = rev4(in, rev); dv = rev4(dv, rev); }
                if (SCALE) {
                    in.x = scale * in.x; in.y = scale * in.y; in.z = scale * in.z; in.w = scale * in.w;
                }
                if (!EXACT) {                     // d' = -0.25 * d, exact outside the flagged cells
                    const u64 nq = pack2(-0.25f, -0.25f);
                    u64 a = mul2(pack2(dv.x, dv.y), nq), b = mul2(pack2(dv.z, dv.w), nq);
                    unpack2(a, dv.x, dv.y); unpack2(b, dv.z, dv.w);
                }
                w[0][(ph + 2) % 3] = in;
                dr[K + ph] = dv;
            } else {
                // ---- level t of step ph ---------------------------------------------------------------
                const float4 c = w[t - 1][(ph + 1) % 3];
                float4 below = w[t - 1][(ph + 0) % 3];
                float4 above = w[t - 1][(ph + 2) % 3];
                if (EDGE) {                       // warp-uniform conditions
                    const int r = st.rout + ph + (K - t);   // row produced by level t at this step
                    if (r == 0) below = c;        // CLAMP_TO_EDGE: p[i,-1] = p[i,0]
                    if (r == H - 1) above = c;    //                p[i,H]  = p[i,H-1]
                }
                const float4 d = dr[K + ph - t];  // div row staged t steps ago
                const float4 o = jacobi4<EXACT>(below, c, above, d);
                if (t < K) {
                    w[t][(ph + 2) % 3] = o;
                } else if (lane_out && (unsigned)(st.rout + ph - y0) < (unsigned)(y1 - y0)) {
                    st.op[ph * W4] = o;
                }
            }
        }
    }
    st.rout += 3;
    st.op += 3 * W4;
#pragma unroll
    for (int j = 0; j < K; ++j) dr[j] = dr[j + 3];
    if (TMA) {
        // Refill the stage consumed ONE BLOCK AGO with the box NS-1 blocks ahead: every value read
        // from it has been used by now, so all lanes' reads of it have completed.
        __syncwarp();
        if (tm.lane == 0) {
            const int q = (st.stg == 0) ? T::NS - 1 : st.stg - 1;
            float4* dst = tm.slots + q * T::STAGE4;
            mbar_expect_tx(tm.bars + q, 2u * 3u * 512u);
            tma_load_2d(dst, tm.tm_p, tm.x0, st.yfill, tm.bars + q);
            tma_load_2d(dst + 3 * 32, tm.tm_d, tm.x0, st.yfill, tm.bars + q);
        }
        st.yfill += 3;
        if (st.stg + 1 == T::NS) { st.stg = 0; st.phase ^= 1u; } else { ++st.stg; }
    }
}

template <int K, bool SCALE, bool REV, bool EXACT, bool TMA>
__device__ __forceinline__ void tb_stream(const JacobiArgs& a, const void* tm_p, const void* tm_d,
                                          float4* __restrict__ smem4, const int lane, const int x0,
                                          const int lc, const int gx, const bool rev,
                                          const bool lane_out, const int y0, const int y1,
                                          const int ys, const int ye) {
    using T = TB<K>;
    const int W = a.W, H = a.H, W4 = W >> 2;
    const int nsteps = y1 - ys + K;                   // level K emits row ys+s-K at step s

    const ptrdiff_t base = -(ptrdiff_t)a.row_off * W4;
    float4* Og = reinterpret_cast<float4*>(a.pout) + base + (lane_out ? (gx >> 2) : 0);

    // rotating windows: w[t][(ph+0)%3] = row r-1, [(ph+1)%3] = row r, [(ph+2)%3] = fresh row r+1
    float4 w[K][3];
#pragma unroll
    for (int t = 0; t < K; ++t)
#pragma unroll
        for (int q = 0; q < 3; ++q) w[t][q] = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 dr[T::RD];
#pragma unroll
    for (int q = 0; q < T::RD; ++q) dr[q] = make_float4(0.f, 0.f, 0.f, 0.f);

    TBStream st{};
    TBTma tm{};
    const float4* ring;
    if (TMA) {
        // every lane reads its (mirrored / clamped) column group out of the staged 128-column box
        ring = smem4 + ((lc - x0) >> 2);
        tm.tm_p = tm_p; tm.tm_d = tm_d;
        tm.bars = reinterpret_cast<unsigned long long*>(smem4 + T::NS * T::STAGE4);
        tm.slots = smem4;
        tm.x0 = x0; tm.lane = lane; tm.err = a.err;
        st.stg = 0; st.phase = 0;
        st.yfill = ys - a.row_off;
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < T::NS; ++q) mbar_init(tm.bars + q, 1);
            mbar_fence_init();
        }
        __syncwarp();
        // boxes of blocks 0 .. NS-2 into stages 0 .. NS-2; stage NS-1 is filled at the end of block 0
#pragma unroll
        for (int q = 0; q < T::NS - 1; ++q) {
            if (lane == 0) {
                float4* dst = tm.slots + q * T::STAGE4;
                mbar_expect_tx(tm.bars + q, 2u * 3u * 512u);
                tma_load_2d(dst, tm_p, x0, st.yfill, tm.bars + q);
                tma_load_2d(dst + 3 * 32, tm_d, x0, st.yfill, tm.bars + q);
            }
            st.yfill += 3;
        }
    } else {
        ring = smem4 + lane;                          // p slot q at ring[q*32], div at ring[(D+q)*32]
        const float4* Pg = reinterpret_cast<const float4*>(a.pin) + base + (lc >> 2);
        const float4* Dg = reinterpret_cast<const float4*>(a.div) + base + (lc >> 2);
        st.pl = Pg + (ptrdiff_t)ys * W4;
        st.dl = Dg + (ptrdiff_t)ys * W4;
        st.rload = ys; st.slot = 0;
        // fill the staging ring: rows ys .. ys+D-1 (clamped to ye), one cp.async group per row
#pragma unroll
        for (int q = 0; q < T::D; ++q) {
            cp_async16(smem4 + lane + q * 32, st.pl);
            cp_async16(smem4 + lane + (T::D + q) * 32, st.dl);
            cp_async_commit();
            if (st.rload < ye) { ++st.rload; st.pl += W4; st.dl += W4; }
        }
        st.stage = ring;
    }
    st.rout = ys - K;
    st.op = Og + (ptrdiff_t)st.rout * W4;             // only dereferenced for rows in [y0, y1)

    // A block touches a wall row when its rows [ys+s0-K, ys+s0+2] reach row 0 or row H-1: the
    // first blocks of a bottom chunk and the last ones of a top chunk.  Wall and steady-state
    // blocks run in SEPARATE loops (not one loop with a branch) so that the steady-state loop
    // has its own register assignment and its window rotation closes without moves.
    int s0 = 0;
#pragma unroll 1
    for (int part = 0; part < 2; ++part) {
#pragma unroll 1
        for (; s0 < nsteps && ((ys + s0 - K <= 0) || (ys + s0 + 2 >= H - 1)); s0 += 3)
            tb_block<K, SCALE, true, REV, EXACT, TMA>(w, dr, st, ring, tm, W4, ye, y0, y1, H, rev, lane_out, a.scale);
#pragma unroll 1
        for (; s0 < nsteps && !((ys + s0 - K <= 0) || (ys + s0 + 2 >= H - 1)); s0 += 3)
            tb_block<K, SCALE, false, REV, EXACT, TMA>(w, dr, st, ring, tm, W4, ye, y0, y1, H, rev, lane_out, a.scale);
    }
    // drain the over-fetched tail before the CTA (and its shared memory) goes away
    if (TMA) {
#pragma unroll 1
        for (int q = 0; q < T::NS - 1; ++q) {
            mbar_wait(tm.bars + st.stg, st.phase, tm.err);
            if (st.stg + 1 == T::NS) { st.stg = 0; st.phase ^= 1u; } else { ++st.stg; }
        }
    } else {
        cp_async_wait<0>();
    }
}

#ifdef FLUID_TB_EXACT_INLINE
#define FLUID_TB_EXACT_ATTR __forceinline__
#else
#define FLUID_TB_EXACT_ATTR __noinline__
#endif
template <int K, bool SCALE, bool TMA>
__device__ FLUID_TB_EXACT_ATTR void tb_stream_exact(const JacobiArgs& a, const void* tm_p, const void* tm_d,
                                             float4* __restrict__ smem4, const int lane, const int x0,
                                             const int lc, const int gx, const bool rev, const bool lane_out,
                                             const int y0, const int y1, const int ys, const int ye,
                                             const bool any_rev) {
    if (any_rev) tb_stream<K, SCALE, true, true, TMA>(a, tm_p, tm_d, smem4, lane, x0, lc, gx, rev, lane_out, y0, y1, ys, ye);
    else tb_stream<K, SCALE, false, true, TMA>(a, tm_p, tm_d, smem4, lane, x0, lc, gx, rev, lane_out, y0, y1, ys, ye);
}

// Slab epilogue of a stream (last launch of a solve): copy the rows this warp just produced that lie in
// a neighbour's ghost zone into the neighbour's buffer.  The warp re-reads its own stores (same
// thread, same address: program order), waits for the neighbour's "done reading" first.
// Written WITHOUT lane-dependent control flow (predicated PTX instead of if (lane ...)): code that may
// diverge after the stream makes the compiler drop its "this one-warp CTA is always converged"
// assumption for the whole kernel, and the shuffles of the hot loops get divergence scaffolding.
__device__ __forceinline__ void st_v4_if(float4* p, const float4 v, const bool on) {
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %5, 0; @q st.global.v4.f32 [%0], {%1, %2, %3, %4}; }"
                 ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"((unsigned)on) : "memory");
}
__device__ __forceinline__ void tb_mirror(const JacobiArgs& a, const TbSync& sy, const int gx, const bool lane_out, const int y0, const int y1) {
    const int W4 = a.W >> 2;
    const int col = lane_out ? (gx >> 2) : 0;
    const float4* own = reinterpret_cast<const float4*>(a.pout) - (ptrdiff_t)a.row_off * W4 + col;
#pragma unroll 1
    for (int s = 0; s < 2; ++s) {
        if (sy.mirror[s] == nullptr) continue;
        const int lo = max(y0, sy.mir_lo[s]), hi = min(y1, sy.mir_hi[s]);
        if (lo >= hi) continue;
        const unsigned long long t1 = sy.dbg ? global_ns() : 0ull;
        warp_spin_until(sy.mir_wait[s], sy.mir_seq, a.err);
        const unsigned long long t2 = sy.dbg ? global_ns() : 0ull;
        float4* dst = reinterpret_cast<float4*>(sy.mirror[s]) + col;
        // 16 rows at a time: all the loads first (the stores are volatile asm, nothing moves across them —
        // one load -> store pair per row was a chain of ~50 L2 round trips, 35 us per solve)
#pragma unroll 1
        for (int r = lo; r < hi; r += 16) {
            float4 v[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = own[(ptrdiff_t)min(r + q, hi - 1) * W4];
#pragma unroll
            for (int q = 0; q < 16; ++q) st_v4_if(dst + (ptrdiff_t)min(r + q, hi - 1) * W4, v[q], lane_out);   // the last row may be stored twice
        }
        // this lane's stores are performed at the neighbour before the lane goes on to the completion ticket.
        // (fence.acq_rel, not __threadfence_system(): that one is fence.sc — MEMBAR.SC.SYS, totally ordered
        // among ALL the grid's warps — and measured 35 us per solve when every warp issued it)
        asm volatile("fence.acq_rel.sys;" ::: "memory");
        if (sy.dbg) {
            const unsigned long long t3 = global_ns();
            const bool l0 = (threadIdx.x == 0);
            dbg_add(sy.dbg + 2, t2 - t1, l0); dbg_add(sy.dbg + 3, t3 - t2, l0); dbg_add(sy.dbg + 4, 1ull, l0);
        }
    }
}

// Grid-completion signal: the last CTA to get here release-stores done_seq into the neighbours' words.
__device__ __forceinline__ void tb_signal_done(const TbSync& sy, const int lane) {
    // every lane is past its stores (and, in tb_mirror, its fence) before lane 0 takes the ticket — a vote,
    // not __syncwarp: see above.  The ticket is a release at GPU scope; the last CTA's fence + release-store
    // at system scope then publishes everything that happened before any ticket.
    const unsigned first = (__all_sync(0xffffffffu, 1) && lane == 0) ? 1u : 0u;
    unsigned t;
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %1, 0; mov.u32 %0, 0xffffffff; @q atom.acq_rel.gpu.global.add.u32 %0, [%2], 1; }"
                 : "=r"(t) : "r"(first), "l"(sy.ticket) : "memory");
    const unsigned last = (t == gridDim.x - 1) ? 1u : 0u;              // lane 0 of the last CTA only
    if (__any_sync(0xffffffffu, last)) {                               // warp-uniform branch: one CTA of the grid enters
        asm volatile("{ .reg .pred q; setp.ne.u32 q, %0, 0; @q st.global.u32 [%1], 0; @q fence.acq_rel.sys; }"
                     ::"r"(last), "l"(sy.ticket) : "memory");
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            if (sy.done_flag[s] == nullptr) continue;
            asm volatile("{ .reg .pred q; setp.ne.u32 q, %0, 0; @q st.release.sys.global.u32 [%1], %2; }"
                         ::"r"(last), "l"(sy.done_flag[s]), "r"(sy.done_seq) : "memory");
        }
    }
}

// One warp per CTA: every quantity that steers control flow derives from blockIdx and kernel
// arguments only, so the compiler can prove the warp converged at each shuffle (no WARPSYNC /
// BSSY scaffolding) and keeps loop state in uniform registers.
#ifndef FLUID_TB_MINBLOCKS
#define FLUID_TB_MINBLOCKS 1     // tuning builds: resident CTAs per SM the register allocation must allow
#endif
#ifdef FLUID_TB_MAXNREG            // tuning builds: cap registers directly (192 -> 10 one-warp CTAs per SM)
#define FLUID_TB_BOUNDS __maxnreg__(FLUID_TB_MAXNREG)
#else
#define FLUID_TB_BOUNDS __launch_bounds__(32, FLUID_TB_MINBLOCKS)
#endif
template <int K, bool SCALE, bool TMA, bool SYNC>
__global__ void FLUID_TB_BOUNDS jacobi_tb_kernel(TbArgs<SYNC> a, const __grid_constant__ TmapPair maps) {
    using T = TB<K>;
    extern __shared__ __align__(128) float4 smem4[];
    // Programmatic dependent launch: back-to-back blocked launches of one solve are chained with
    // cudaLaunchAttributeProgrammaticStreamSerialization, so the NEXT launch's CTAs are already
    // resident (index math done, parked at griddepcontrol.wait) while this grid drains — the
    // single-wave tail and the launch gap overlap instead of adding up.  Both instructions are
    // no-ops for a launch without the attribute.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int lane = threadIdx.x;
    const int nxw = (a.W + T::VALID - 1) / T::VALID;
    const int wid = blockIdx.x;
    const int wx = wid % nxw;
    int cy = wid / nxw, seg_lo = a.out_lo, seg_hi = a.out_hi;
    if (a.nch1 > 0 && cy >= a.nch1) { cy -= a.nch1; seg_lo = a.seg2_lo; seg_hi = a.seg2_hi; }

    // ---- x geometry of this lane: mirrored / clamped float4 column group ------------------------
    const int W = a.W, H = a.H;
    const int gx = wx * T::VALID - T::HX + 4 * lane;  // first global column of this lane
    int lc = gx;
    bool rev = false;
    if (gx < 0) { lc = -gx - 4; rev = true; }                 // p[-1-m] = p[m]
    else if (gx >= W) { lc = 2 * W - 4 - gx; rev = true; }    // p[W+m]  = p[W-1-m]
    lc = min(max(lc, 0), W - 4);
    const bool any_rev = (wx == 0) || ((wx + 1) * T::VALID + T::HX > W);   // warp-uniform
    const bool lane_out = (lane >= T::HX / 4) && (lane < 32 - T::HX / 4) && (gx >= 0) && (gx < W);
    const int x0 = wx * T::VALID - T::HX;

    // ---- y geometry of this warp's stream ----------------------------------------------------------
    const int y0 = seg_lo + cy * a.rows_per_chunk;
    const int y1 = min(y0 + a.rows_per_chunk, seg_hi);
    const int ys = max(y0 - K, 0);                    // first input row
    const int ye = min(y1 - 1 + K, H - 1);            // last input row

    asm volatile("griddepcontrol.wait;" ::: "memory");      // everything below reads what the previous launch wrote

    // slab, first launch of a solve whose pressure ghost rows were stored by the neighbours' previous
    // last launch: the streams that read those rows acquire "mirror ready" first
    // (every lane polls the same word — warp-uniform control flow, see tb_mirror)
    unsigned long long t_start = 0;
    if constexpr (SYNC) {
        const TbSync& sy = a.sy;
        if (sy.dbg) t_start = global_ns();
        const bool w0 = sy.pre_wait[0] != nullptr && ys < sy.own_lo, w1 = sy.pre_wait[1] != nullptr && ye >= sy.own_hi;
        if (w0) warp_spin_until(sy.pre_wait[0], sy.pre_seq, a.err);
        if (w1) warp_spin_until(sy.pre_wait[1], sy.pre_seq, a.err);
        if (sy.dbg && (w0 || w1)) { dbg_add(sy.dbg + 0, global_ns() - t_start, lane == 0); dbg_add(sy.dbg + 1, 1ull, lane == 0); }
    }

    // ---- does the divergence this stream reads hold a value that defeats the fma contraction? ----
    bool exact = false;
#ifndef FLUID_TB_NO_EXACT          // tuning builds only: measures what the second instantiation costs in code size
    if (a.tiny_map) {
        const int mw = tiny_map_w(W);
        const int cx0 = max(x0, 0) / TINY_CW, cx1 = min(x0 + 127, W - 1) / TINY_CW;
        const int cy0 = ys / TINY_CH, cy1 = ye / TINY_CH;
        const int ncx = cx1 - cx0 + 1, n = ncx * (cy1 - cy0 + 1);
        unsigned f = 0;
        for (int k = lane; k < n; k += 32) f |= a.tiny_map[(cy0 + k / ncx) * mw + cx0 + k % ncx];
        exact = __any_sync(0xffffffffu, f != 0);
    }
#endif
    const void* tp = &maps.p;
    const void* td = &maps.d;
    // The common case first, so that it is laid out first: code placement matters to this kernel (it
    // runs one or two warps per scheduler, straight out of the instruction caches).  The un-contracted
    // instantiations are kept out of line for the same reason.
    if (!exact) {
        if (!any_rev) tb_stream<K, SCALE, false, false, TMA>(a, tp, td, smem4, lane, x0, lc, gx, rev, lane_out, y0, y1, ys, ye);
        else tb_stream<K, SCALE, true, false, TMA>(a, tp, td, smem4, lane, x0, lc, gx, rev, lane_out, y0, y1, ys, ye);
    } else {
        tb_stream_exact<K, SCALE, TMA>(a, tp, td, smem4, lane, x0, lc, gx, rev, lane_out, y0, y1, ys, ye, any_rev);
    }
    if constexpr (SYNC) {
        const TbSync& sy = a.sy;
        if (sy.dbg) {
            const bool edge = (ys < sy.own_lo) || (ye >= sy.own_hi);
            dbg_add(sy.dbg + (edge ? 5 : 7), global_ns() - t_start, lane == 0); dbg_add(sy.dbg + (edge ? 6 : 8), 1ull, lane == 0);
        }
        if (sy.mirror[0] != nullptr || sy.mirror[1] != nullptr) tb_mirror(a, sy, gx, lane_out, y0, y1);
        if (sy.ticket != nullptr) tb_signal_done(sy, lane);
    }
}

// ---- producers of the tiny-divergence map ----------------------------------------------------------
// Host-written divergence (fluid_write, fluid_pressure_solve_host) and ghost rows received from a
// neighbour are scanned by this kernel; the kernels that COMPUTE divergence flag cells themselves.
// The map must have been zeroed for the rows being scanned (cudaMemsetAsync by the caller).
__global__ void __launch_bounds__(256) tiny_scan_kernel(const float* __restrict__ div, unsigned char* __restrict__ map,
                                                        int W, int row_off, int j_lo, int j_hi) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = j_lo + blockIdx.y;
    if (i >= W || j >= j_hi) return;
    if (is_tiny_div(__ldg(div + (size_t)(j - row_off) * W + i)))
        map[(j / TINY_CH) * tiny_map_w(W) + i / TINY_CW] = 1;
}

}  // namespace fk
