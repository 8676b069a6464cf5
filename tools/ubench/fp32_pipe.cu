// fp32_pipe.cu — micro-benchmark: issue / pipe throughput of the fp32 instructions the temporally
// blocked Jacobi kernel is built from (scalar FADD / FFMA vs the packed f32x2 forms, SHFL, MOV mixes)
// on one B200, as warp-instructions per cycle per SM sub-partition (SMSP).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp32_pipe fp32_pipe.cu && ./fp32_pipe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

typedef unsigned long long u64;
#ifndef NCH
#define NCH 8           // independent dependency chains per thread (compile with -DNCH=1 for dependent-issue latency)
#endif
#define BODY 4          // chain updates per loop iteration (per chain)

enum Kind { K_FADD, K_FADD2, K_FFMA, K_FFMA2, K_FMUL2, K_MIX_1A2P, K_SHFL, K_FADD2_MOV, K_FADD_IMM, K_FFMA_IMM, K_FADD2_DIST, K_FFMA2_DIST, K_FADD_DIST, K_NKIND };
static const char* NAMES[] = {"FADD r,r,r", "FADD2 (add.f32x2)", "FFMA r,r,r,r", "FFMA2 (fma.f32x2)", "FMUL2 (mul.f32x2)",
                              "mix 1 FADD : 2 FADD2", "SHFL.UP", "mix 1 FADD2 : 1 MOV", "FADD r,r,imm", "FFMA r,r,imm,r",
                              "FADD2 distinct operands", "FFMA2 distinct operands", "FADD distinct operands"};
// warp-instructions per loop iteration per thread
static const int INSTR[] = {NCH * BODY, NCH * BODY, NCH * BODY, NCH * BODY, NCH * BODY, NCH * BODY, NCH * BODY, NCH * BODY, NCH * BODY, NCH * BODY, NCH * BODY, NCH * BODY, NCH * BODY};

template <int KIND>
__global__ void bench(float* out, const float* in, int iters, u64* cycles) {
    float a[NCH], b = in[1], c = in[2];
    u64 p[NCH], pb, pc;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        a[k] = in[3 + k] + threadIdx.x;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p[k]) : "f"(a[k]), "f"(a[k] + 1.0f));
    }
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b), "f"(b));
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(pc) : "f"(c), "f"(c));
    __syncthreads();
    const u64 t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < BODY; ++r) {
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                if (KIND == K_FADD) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[k]) : "f"(b));
                if (KIND == K_FADD_IMM) asm volatile("add.rn.f32 %0, %0, 0f3F800000;" : "+f"(a[k]));
                if (KIND == K_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[k]) : "f"(b), "f"(c));
                if (KIND == K_FFMA_IMM) asm volatile("fma.rn.f32 %0, %0, 0f3E800000, %1;" : "+f"(a[k]) : "f"(c));
                if (KIND == K_FADD2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[k]) : "l"(pb));
                if (KIND == K_FMUL2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[k]) : "l"(pb));
                if (KIND == K_FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[k]) : "l"(pb), "l"(pc));
                if (KIND == K_MIX_1A2P) {
                    if (k % 3 == 0) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[k]) : "f"(b));
                    else asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[k]) : "l"(pb));
                }
                // every operand a different register (pair): no operand-reuse cache hits, like real code
                if (KIND == K_FADD2_DIST) asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(p[k]) : "l"(p[(k + 3) % NCH]), "l"(p[(k + 5) % NCH]));
                if (KIND == K_FFMA2_DIST) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p[k]) : "l"(p[(k + 3) % NCH]), "l"(p[(k + 5) % NCH]), "l"(p[(k + 6) % NCH]));
                if (KIND == K_FADD_DIST) asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(a[k]) : "f"(a[(k + 3) % NCH]), "f"(a[(k + 5) % NCH]));
                if (KIND == K_SHFL) asm volatile("shfl.sync.up.b32 %0, %0, 1, 0, 0xffffffff;" : "+f"(a[k]));
                if (KIND == K_FADD2_MOV) {
                    if (k & 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[k]) : "l"(pb));
                    else asm volatile("mov.b32 %0, %1;" : "=f"(a[k]) : "f"(a[(k + 2) % NCH]));
                }
            }
        }
    }
    const u64 t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        float lo, hi;
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[k]));
        s += a[k] + lo + hi;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int KIND>
void run(int warps_per_smsp, float* out, const float* in, u64* cyc, int nsm) {
    const int iters = 2000;
    const int threads = 32 * 4 * warps_per_smsp;     // one CTA per SM, 4 SMSPs
    const int tpb = threads > 1024 ? 1024 : threads;
    const int ctas_per_sm = threads / tpb;
    bench<KIND><<<nsm * ctas_per_sm, tpb>>>(out, in, 10, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<KIND><<<nsm * ctas_per_sm, tpb>>>(out, in, iters, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    u64 h[4096]; cudaMemcpy(h, cyc, sizeof(u64) * nsm * ctas_per_sm, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < nsm * ctas_per_sm; ++i) avg += (double)h[i]; avg /= nsm * ctas_per_sm;
    const double winstr_per_smsp = (double)iters * INSTR[KIND] * warps_per_smsp;
    printf("%-24s warps/SMSP %2d : %.3f warp-instr/clk/SMSP  (%.0f cycles, %.3f ms)\n", NAMES[KIND], warps_per_smsp,
           winstr_per_smsp / avg, avg, ms);
}

int main() {
    int dev = 0; cudaSetDevice(dev);
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev);
    const int nsm = pr.multiProcessorCount;
    printf("# %s, %d SMs, sm_%d%d\n", pr.name, nsm, pr.major, pr.minor);
    float *out, *in; u64* cyc;
    cudaMalloc(&out, sizeof(float) * nsm * 4096); cudaMalloc(&in, 64 * sizeof(float)); cudaMalloc(&cyc, sizeof(u64) * 4096);
    float hin[64]; for (int i = 0; i < 64; ++i) hin[i] = 1.0f + i * 1e-3f; hin[1] = 1e-7f; hin[2] = 1e-9f;
    cudaMemcpy(in, hin, sizeof hin, cudaMemcpyHostToDevice);
    for (int w : {1, 2, 4, 8}) {
        run<K_FADD>(w, out, in, cyc, nsm); run<K_FADD_IMM>(w, out, in, cyc, nsm);
        run<K_FADD2>(w, out, in, cyc, nsm); run<K_FFMA>(w, out, in, cyc, nsm); run<K_FFMA_IMM>(w, out, in, cyc, nsm);
        run<K_FFMA2>(w, out, in, cyc, nsm); run<K_FMUL2>(w, out, in, cyc, nsm); run<K_MIX_1A2P>(w, out, in, cyc, nsm);
        run<K_SHFL>(w, out, in, cyc, nsm); run<K_FADD2_MOV>(w, out, in, cyc, nsm);
        run<K_FADD_DIST>(w, out, in, cyc, nsm); run<K_FADD2_DIST>(w, out, in, cyc, nsm); run<K_FFMA2_DIST>(w, out, in, cyc, nsm);
        printf("\n");
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("# %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
