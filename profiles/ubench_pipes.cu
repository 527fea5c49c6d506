// ubench_pipes.cu -- how the FP64, XU (MUFU / F2F / FRND), ALU and FMA pipes of one B200 SM share
// issue bandwidth.  Each kernel runs ITER iterations of a fixed instruction mix on independent
// register chains (8 chains per thread), 148*k CTAs of 256 threads; we report warp-instructions
// per clock per SM for every mix.  Build: nvcc -arch=sm_100a -O3 -o ubench_pipes ubench_pipes.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CHAINS 8
constexpr int ITER = 4096;

template <int ND, int NX, int NA, int NF, int XKIND>
__global__ void __launch_bounds__(256) mix(double* out, double seed, int iters) {
    double d[CHAINS];
    float f[CHAINS];
    int a[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { d[c] = seed + c + threadIdx.x; f[c] = (float)d[c]; a[c] = threadIdx.x + c; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < ND; ++r)
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) d[c] = fma(d[c], 1.0000001, 1e-9);
#pragma unroll
        for (int r = 0; r < NX; ++r)
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) {
                if (XKIND == 0) { double t; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(t) : "f"(f[c])); f[c] = __int_as_float(__double2loint(t) ^ __float_as_int(f[c])); }
                if (XKIND == 1) { double t; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(t) : "d"(d[c])); a[c] ^= __double2hiint(t); }
                if (XKIND == 2) { double t; asm volatile("cvt.rmi.f64.f64 %0, %1;" : "=d"(t) : "d"(d[c])); a[c] ^= __double2loint(t); }
                if (XKIND == 3) { float t; asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(t) : "d"(d[c])); a[c] ^= __float_as_int(t); }
                if (XKIND == 4) { float t; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(f[c])); f[c] = t; }
            }
#pragma unroll
        for (int r = 0; r < NA; ++r)
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) asm volatile("lop3.b32 %0, %0, %1, 0x55aa, 0x96;" : "+r"(a[c]) : "r"(it));
#pragma unroll
        for (int r = 0; r < NF; ++r)
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) asm volatile("fma.rn.f32 %0, %0, 0f3F800001, 0f2F800000;" : "+f"(f[c]));
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += d[c] + f[c] + a[c];
    if (s == 123.456) out[threadIdx.x] = s;
}

template <int ND, int NX, int NA, int NF, int XKIND>
void run(const char* name, int sms, double mhz) {
    double* out;
    cudaMalloc(&out, 4096);
    const int grid = sms * 4;   // 4 CTAs x 8 warps = 32 warps per SM
    mix<ND, NX, NA, NF, XKIND><<<grid, 256>>>(out, 1.0, 16);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    mix<ND, NX, NA, NF, XKIND><<<grid, 256>>>(out, 1.0, ITER);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double per_it = (double)(ND + NX + NA + NF) * CHAINS;            // warp instr per warp per iteration
    const double winst = per_it * ITER * (double)grid * 8;                 // total warp instructions
    const double clk = ms * 1e-3 * mhz * 1e6;
    printf("%-34s D=%d X=%d A=%d F=%d : %8.3f ms  %6.3f winst/clk/SM   cycles per (iteration,chain) per SMSP-warp-slot: %6.2f\n",
           name, ND, NX, NA, NF, ms, winst / clk / sms, clk / (ITER * (double)CHAINS * 8 /*warps per SMSP*/));
    cudaFree(out);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1000.0;
    printf("%s  SMs=%d  clock attr=%.0f MHz (rates below assume this clock)\n", p.name, p.multiProcessorCount, mhz);
    const int sms = p.multiProcessorCount;
    run<4, 0, 0, 0, 0>("DFMA only", sms, mhz);
    run<0, 4, 0, 0, 0>("F2F.F64.F32 only", sms, mhz);
    run<0, 4, 0, 0, 1>("MUFU.RCP64H only", sms, mhz);
    run<0, 4, 0, 0, 2>("FRND.F64.FLOOR only", sms, mhz);
    run<0, 4, 0, 0, 3>("F2F.F32.F64 only", sms, mhz);
    run<0, 4, 0, 0, 4>("MUFU.RCP (f32) only", sms, mhz);
    run<0, 0, 4, 0, 0>("LOP3 only", sms, mhz);
    run<0, 0, 0, 4, 0>("FFMA only", sms, mhz);
    run<4, 0, 4, 0, 0>("DFMA + LOP3 1:1", sms, mhz);
    run<4, 0, 8, 0, 0>("DFMA + LOP3 1:2", sms, mhz);
    run<4, 0, 0, 4, 0>("DFMA + FFMA 1:1", sms, mhz);
    run<4, 0, 4, 4, 0>("DFMA + LOP3 + FFMA 1:1:1", sms, mhz);
    run<4, 1, 0, 0, 0>("DFMA + F2F64 4:1", sms, mhz);
    run<4, 2, 0, 0, 0>("DFMA + F2F64 4:2", sms, mhz);
    run<8, 2, 0, 0, 0>("DFMA + F2F64 8:2", sms, mhz);
    run<4, 1, 0, 0, 1>("DFMA + RCP64H 4:1", sms, mhz);
    run<4, 1, 0, 0, 2>("DFMA + FRND 4:1", sms, mhz);
    run<4, 1, 4, 0, 0>("DFMA + F2F64 + LOP3 4:1:4", sms, mhz);
    run<8, 2, 8, 4, 0>("DFMA+F2F64+LOP3+FFMA 8:2:8:4 (kernel-like)", sms, mhz);
    run<0, 2, 8, 0, 0>("F2F64 + LOP3 2:8", sms, mhz);
    run<0, 2, 0, 8, 0>("F2F64 + FFMA 2:8", sms, mhz);
    return 0;
}
