// microbench.cu -- per-SM throughput of the instruction classes the force kernel is made of
// (B200 / sm_100a).  Prints ops per clock per SM.  Build: see tools/build_microbench.sh
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define ITERS 4096
#define ILP 8

template <int OP>
__global__ void k(double *out, double seed, float fseed) {
    double a[ILP];
    float f[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { a[i] = seed + i * 0.37 + threadIdx.x * 1e-3; f[i] = fseed + i * 0.37f + threadIdx.x * 1e-3f; }
    const double c = seed * 0.999, d = seed * 1e-3;
    const float g0 = fseed * 0.999f, g1 = fseed * 1e-3f;
    float2 qa = make_float2(g0, g0), qb = make_float2(g1, g1);
    const unsigned long long q0 = *reinterpret_cast<unsigned long long *>(&qa), q1 = *reinterpret_cast<unsigned long long *>(&qb);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == 0) a[i] = fma(a[i], c, d);
            else if (OP == 1) a[i] = __dadd_rn(a[i], d);
            else if (OP == 2) a[i] = __dmul_rn(a[i], c);
            else if (OP == 3) a[i] = rint(a[i]) + d;          // FRND.F64 + DADD
            else if (OP == 4) a[i] = round(a[i]) + d;         // DADD.RZ + FRND.TRUNC + DADD
            else if (OP == 5) a[i] = __ddiv_rn(1.0, a[i]);    // MUFU.RCP64H + 5 DFMA (+ guard)
            else if (OP == 6) { f[i] = (float)a[i]; a[i] = a[i] + d + (double)0.0; a[i] = __dadd_rn(a[i], (double)f[i] * 0.0); }
            else if (OP == 7) f[i] = fmaf(f[i], 0.999f, 0.001f);
            else if (OP == 8) a[i] = floor(a[i]) + d;
            else if (OP == 9) { a[i] = (a[i] + 6755399441055744.0) - 6755399441055744.0 + d; }  // magic rint: 3 DADD
            else if (OP == 10) f[i] = rintf(f[i]) + 0.001f;
            else if (OP == 11) f[i] = fmaf(f[i], g0, g1);      // FFMA, three register operands
            else if (OP == 12) {                               // FFMA2 (fma.rn.f32x2): two FP32 FMAs per instruction
                unsigned long long &p2 = reinterpret_cast<unsigned long long *>(a)[i];
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(q0), "l"(q1));
            } else if (OP == 13) {                             // FADD2
                unsigned long long &p2 = reinterpret_cast<unsigned long long *>(a)[i];
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p2) : "l"(q1));
            }
        }
    }
    double s = 0; float fs = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { s += a[i]; fs += f[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + fs;
}

template <int OP>
void run(const char *name, double ops_per_iter) {
    int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    int blocks = p.multiProcessorCount * 4, tpb = 512;
    double *out; cudaMalloc(&out, sizeof(double) * blocks * tpb);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<blocks, tpb>>>(out, 1.0001, 1.0001f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<OP><<<blocks, tpb>>>(out, 1.0001, 1.0001f);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev);
    double total = (double)blocks * tpb * ITERS * ILP * ops_per_iter;
    double per_s = total / (ms * 1e-3);
    printf("%-28s %8.3f ms  %10.3e ops/s  %7.2f ops/clk/SM @%d MHz(nominal max)\n", name, ms, per_s,
           per_s / p.multiProcessorCount / (clk_khz * 1e3), clk_khz / 1000);
    cudaFree(out);
}

int main() {
    run<0>("DFMA", 1);
    run<1>("DADD", 1);
    run<2>("DMUL", 1);
    run<3>("rint(FRND.F64)+DADD", 1);
    run<4>("round(DADD.RZ+FRND.TRUNC)+DADD", 1);
    run<5>("1.0/x (RCP64H+5DFMA)", 1);
    run<6>("F2F f64->f32->f64 mix", 1);
    run<7>("FFMA", 1);
    run<8>("floor(FRND.FLOOR)+DADD", 1);
    run<9>("magic-rint (3 DADD)", 1);
    run<10>("rintf(FRND)+FADD", 1);
    run<11>("FFMA reg,reg,reg", 1);
    run<12>("FFMA2 (f32x2), FMAs", 2);
    run<13>("FADD2 (f32x2), adds", 2);
    return 0;
}
