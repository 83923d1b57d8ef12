// Issue-rate micro-benchmark for the instruction mix of the attention softmax warps (B200, sm_100a).
// For each op: one CTA per SM, W warps per scheduler, 8 independent dependency chains per thread, reports
// cycles per warp-instruction per SM sub-partition (rt_SMSP).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

enum Op { FFMA3, FFMA_IMM, FADD_, FMUL_, FFMA2_, FADD2_, FADD2RM, FMNMX2, FMNMX3_, EX2, EX2_F16X2, F2FP_, IADD_, SHL_ADD, MIX_SOFT, MIX_POLY };
static const char* kNames[] = {"FFMA(3reg)", "FFMA(imm)", "FADD", "FMUL", "FFMA2", "FADD2", "FADD2.RM", "FMNMX", "FMNMX3", "MUFU.EX2", "MUFU.EX2.F16x2", "F2FP.pack", "IADD3", "SHL+IADD", "mix: ffma2+ex2x2+fadd2+f2fp (per 2 elems)", "mix: poly exp2 x2 (per 2 elems)"};
static const int kInstrPerIter[] = {8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 16, 5 * 4, 9 * 4};

template <int OP>
__global__ void __launch_bounds__(1024, 1) bench(float* out, long long* cyc, int iters, float seed) {
    float a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = seed * (i + 1) + threadIdx.x * 1e-3f; b[i] = seed * 0.5f + i; }
    const float c0 = seed * 0.999f, c1 = seed * 0.25f;
    unsigned long long pa[8], pb[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(pa[i]) : "f"(a[i]), "f"(b[i]));
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(pb[i]) : "f"(c0), "f"(c1));
    }
    uint32_t u[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) u[i] = __float_as_uint(a[i]);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if constexpr (OP == FFMA3) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c0), "f"(b[i]));
        } else if constexpr (OP == FFMA_IMM) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32 %0, %0, %1, 0f3F000000;" : "+f"(a[i]) : "f"(c0));
        } else if constexpr (OP == FADD_) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c1));
        } else if constexpr (OP == FMUL_) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c0));
        } else if constexpr (OP == FFMA2_) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(pa[i]) : "l"(pb[i]));
        } else if constexpr (OP == FADD2_) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(pa[i]) : "l"(pb[i]));
        } else if constexpr (OP == FADD2RM) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("add.rm.ftz.f32x2 %0, %0, %1;" : "+l"(pa[i]) : "l"(pb[i]));
        } else if constexpr (OP == FMNMX2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
        } else if constexpr (OP == FMNMX3_) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(c0));
        } else if constexpr (OP == EX2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        } else if constexpr (OP == EX2_F16X2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i]));
        } else if constexpr (OP == F2FP_) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(__uint_as_float(u[i])));
        } else if constexpr (OP == IADD_) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
        } else if constexpr (OP == SHL_ADD) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { uint32_t t; asm volatile("shl.b32 %0, %1, 23;" : "=r"(t) : "r"(u[i])); asm volatile("add.u32 %0, %1, %2;" : "=r"(u[i]) : "r"(t), "r"(u[i])); }
        } else if constexpr (OP == MIX_SOFT) {
            // per PAIR of elements: FFMA2 (scale, -max) ; 2 x MUFU.EX2 ; FADD2 (row sum) ; F2FP pack  = 5 instr
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                unsigned long long x;
                asm volatile("fma.rn.f32x2 %0, %1, %2, %2;" : "=l"(x) : "l"(pa[i]), "l"(pb[i]));
                float x0, x1;
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(x));
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x0));
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x1));
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(x0), "f"(x1));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(pa[4 + i]) : "l"(x));
                asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(x1), "f"(x0));
            }
        } else if constexpr (OP == MIX_POLY) {
            // per PAIR: FFMA2 (scale,-max); FADD2.RM (magic floor); FADD2 (-magic) ; FADD2 (frac) ; 3 x FFMA2 (poly); 2 x (SHL+IADD) -> counted 9
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                unsigned long long x, f, r, pl;
                asm volatile("fma.rn.f32x2 %0, %1, %2, %2;" : "=l"(x) : "l"(pa[i]), "l"(pb[i]));
                asm volatile("add.rm.ftz.f32x2 %0, %1, %2;" : "=l"(f) : "l"(x), "l"(pb[4 + i]));
                asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f), "l"(pb[i]));
                asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(x), "l"(r));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %2;" : "=l"(pl) : "l"(r), "l"(pb[i]));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(pl) : "l"(r), "l"(pb[i]));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(pl) : "l"(r), "l"(pb[i]));
                uint32_t f0, f1, p0, p1;
                asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(f0), "=r"(f1) : "l"(f));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(p0), "=r"(p1) : "l"(pl));
                asm volatile("shl.b32 %0, %0, 23;" : "+r"(f0));
                asm volatile("shl.b32 %0, %0, 23;" : "+r"(f1));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(p0) : "r"(f0));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(p1) : "r"(f1));
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(pl) : "r"(p0), "r"(p1));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(pa[4 + i]) : "l"(pl));
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float x, y;
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(pa[i]));
        s += a[i] + x + y + __uint_as_float(u[i]);
    }
    if (s == 123.456f) out[0] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(float* out, long long* cyc, int nsm) {
    const int iters = 2048;
    for (int warps_per_smsp : {1, 2, 4, 8}) {
        const int threads = warps_per_smsp * 128;
        bench<OP><<<nsm, threads>>>(out, cyc, iters, 1.0001f);
        cudaDeviceSynchronize();
        bench<OP><<<nsm, threads>>>(out, cyc, iters, 1.0001f);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", kNames[OP], cudaGetErrorString(e)); return; }
        long long h[256];
        cudaMemcpy(h, cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < nsm; ++i) avg += (double)h[i];
        avg /= nsm;
        const double winstr = (double)iters * kInstrPerIter[OP] * warps_per_smsp;   // warp-instructions per SMSP
        printf("%-46s warps/SMSP=%d  cycles/warp-instr/SMSP = %.3f\n", kNames[OP], warps_per_smsp, avg / winstr);
    }
}

int main() {
    int nsm = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    float* out; long long* cyc;
    cudaMalloc(&out, 1024); cudaMalloc(&cyc, sizeof(long long) * 256);
    printf("SMs: %d\n", nsm);
    run<FFMA3>(out, cyc, nsm); run<FFMA_IMM>(out, cyc, nsm); run<FADD_>(out, cyc, nsm); run<FMUL_>(out, cyc, nsm);
    run<FFMA2_>(out, cyc, nsm); run<FADD2_>(out, cyc, nsm); run<FADD2RM>(out, cyc, nsm); run<FMNMX2>(out, cyc, nsm);
    run<FMNMX3_>(out, cyc, nsm); run<EX2>(out, cyc, nsm); run<EX2_F16X2>(out, cyc, nsm); run<F2FP_>(out, cyc, nsm);
    run<IADD_>(out, cyc, nsm); run<SHL_ADD>(out, cyc, nsm); run<MIX_SOFT>(out, cyc, nsm); run<MIX_POLY>(out, cyc, nsm);
    return 0;
}
