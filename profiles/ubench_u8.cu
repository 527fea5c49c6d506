// ubench_u8.cu -- round 2: issue rates of the instructions the u8 RGB rectification kernel is made of
// (PRMT, FFMA2, IMAD, I2F.U8 with byte select, FMNMX3, SHFL, LDS.32/64/128/U8, STS.128) alone and in
// kernel-like mixes, plus a check that a u8 TMA box may start at ANY byte coordinate.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_u8 ubench_u8.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

constexpr int ITER = 2048;
constexpr int CH = 8;

// KIND: which instruction the slot emits
enum { K_PRMT, K_FFMA2, K_IMAD, K_I2FU8, K_FMNMX3, K_SHFL, K_LDS32, K_LDS64, K_LDS128, K_LDSU8, K_STS128, K_FFMA, K_IADD3, K_FSETP, K_HADD2F32 };

template <int KIND>
__device__ __forceinline__ void emit(uint32_t (&a)[CH], float2 (&f)[CH], uint32_t sa, int c, int it, int r = 0) {
    if (KIND == K_PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x4321;" : "+r"(a[c]) : "r"(it));
    if (KIND == K_FFMA2) { unsigned long long v = *reinterpret_cast<unsigned long long*>(&f[c]);
        asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(v) : "l"(0x3f8000013f800001ull)); *reinterpret_cast<unsigned long long*>(&f[c]) = v; }
    if (KIND == K_FFMA) asm volatile("fma.rn.f32 %0, %0, 0f3F800001, 0f2F800000;" : "+f"(f[c].x));
    if (KIND == K_IMAD) asm volatile("mad.lo.u32 %0, %0, 257, %1;" : "+r"(a[c]) : "r"(it));
    if (KIND == K_I2FU8) { float t; asm volatile("{ .reg .b8 b0,b1,b2,b3; mov.b32 {b0,b1,b2,b3}, %1; cvt.rn.f32.u8 %0, b2; }" : "=f"(t) : "r"(a[c])); a[c] ^= __float_as_uint(t); }
    if (KIND == K_FMNMX3) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(f[c].x) : "f"(f[c].y), "f"(f[(c + 1) % CH].y));
    if (KIND == K_SHFL) a[c] = __shfl_down_sync(0xffffffffu, a[c], 1);
    if (KIND == K_LDS32) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(sa + 4u * c + 128u * r)); a[c] += v; }
    if (KIND == K_LDS64) { uint32_t v, w; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v), "=r"(w) : "r"(sa + (sa & 127u) + 8u * (c & 3) + 256u * (r & 3))); a[c] += v + w; }
    if (KIND == K_LDS128) { uint32_t v, w, x, y; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v), "=r"(w), "=r"(x), "=r"(y) : "r"(sa + 3u * (sa & 127u) + 16u * (c & 1) + 512u * (r & 1))); a[c] += v + w + x + y; }
    if (KIND == K_LDSU8) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(sa + 3u * c + 128u * r)); a[c] += v; }
    if (KIND == K_STS128) asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" :: "r"(sa + 3u * (sa & 127u) + 16u * (c & 1) + 512u * (r & 1)), "r"(a[c]));
    if (KIND == K_IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[c]) : "r"(it));
    if (KIND == K_FSETP) { asm volatile("{ .reg .pred p; setp.gt.f32 p, %1, 0f3F000000; @p add.u32 %0, %0, 1; }" : "+r"(a[c]) : "f"(f[c].x)); }
    if (KIND == K_HADD2F32) { asm volatile("{ .reg .f16 h0, h1; mov.b32 {h0,h1}, %1; cvt.f32.f16 %0, h1; }" : "=f"(f[c].y) : "r"(a[c])); }
}

// N0 x KIND0 + N1 x KIND1 + N2 x KIND2 + N3 x KIND3 per chain per iteration
template <int K0, int N0, int K1, int N1, int K2, int N2, int K3, int N3>
__global__ void __launch_bounds__(256) mix(float* out, int iters) {
    __shared__ __align__(128) uint32_t sm[2048];
    uint32_t a[CH];
    float2 f[CH];
    for (int i = threadIdx.x; i < 2048; i += 256) sm[i] = i * 2654435761u;
    __syncthreads();
    const uint32_t sa = (threadIdx.x & 31) * 4u;   // lane-consecutive words
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm);
#pragma unroll
    for (int c = 0; c < CH; ++c) { a[c] = threadIdx.x * 77u + c; f[c] = make_float2(1.0f + c, 0.5f + threadIdx.x); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < N0; ++r)
#pragma unroll
            for (int c = 0; c < CH; ++c) emit<K0>(a, f, sa + sbase, c, it, r);
#pragma unroll
        for (int r = 0; r < N1; ++r)
#pragma unroll
            for (int c = 0; c < CH; ++c) emit<K1>(a, f, sa + sbase, c, it, r);
#pragma unroll
        for (int r = 0; r < N2; ++r)
#pragma unroll
            for (int c = 0; c < CH; ++c) emit<K2>(a, f, sa + sbase, c, it, r);
#pragma unroll
        for (int r = 0; r < N3; ++r)
#pragma unroll
            for (int c = 0; c < CH; ++c) emit<K3>(a, f, sa + sbase, c, it, r);
    }
    float s = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += f[c].x + f[c].y + (float)a[c];
    if (s == 123.456f) out[threadIdx.x] = s;
}

static double g_mhz;
static int g_sms;

template <int K0, int N0, int K1, int N1, int K2, int N2, int K3, int N3>
void run(const char* name) {
    float* out;
    cudaMalloc(&out, 4096);
    const int grid = g_sms * 4;
    mix<K0, N0, K1, N1, K2, N2, K3, N3><<<grid, 256>>>(out, 16);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    mix<K0, N0, K1, N1, K2, N2, K3, N3><<<grid, 256>>>(out, ITER);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double per_it = (double)(N0 + N1 + N2 + N3) * CH;
    const double winst = per_it * ITER * (double)grid * 8;
    const double clk = ms * 1e-3 * g_mhz * 1e6;
    printf("%-44s %d:%d:%d:%d  %8.3f ms  %6.3f winst/clk/SM  (err %s)\n", name, N0, N1, N2, N3, ms, winst / clk / g_sms,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

// ---- TMA: may a u8 box start at any byte coordinate? -------------------------------------------
__global__ void tma_probe(const __grid_constant__ CUtensorMap tmap, uint8_t* out, int c0, int c1, int bytes) {
    extern __shared__ __align__(128) uint8_t buf[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"((uint32_t)__cvta_generic_to_shared(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&bar)), "r"(bytes));
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     :: "r"((uint32_t)__cvta_generic_to_shared(buf)), "l"(reinterpret_cast<uint64_t>(&tmap)),
                        "r"((uint32_t)__cvta_generic_to_shared(&bar)), "r"(c0), "r"(c1) : "memory");
    }
    uint32_t done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"((uint32_t)__cvta_generic_to_shared(&bar)) : "memory");
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = buf[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static void tma_test() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { printf("no encoder\n"); return; }
    const int W = 1536, H = 64, BW = 112, BH = 8;      // bytes
    uint8_t* src; uint8_t* out;
    cudaMalloc(&src, W * H); cudaMalloc(&out, BW * BH);
    uint8_t* h = (uint8_t*)malloc(W * H);
    for (int i = 0; i < W * H; ++i) h[i] = (uint8_t)(i * 131 + (i >> 8) * 7);
    cudaMemcpy(src, h, W * H, cudaMemcpyHostToDevice);
    CUtensorMap tm;
    cuuint64_t dims[2] = {W, H}, strides[1] = {W};
    cuuint32_t box[2] = {BW, BH}, es[2] = {1, 1};
    CUresult r = ((EncodeTiledFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode u8 box %dx%d: %d\n", BW, BH, (int)r);
    uint8_t* ho = (uint8_t*)malloc(BW * BH);
    const int c0s[] = {0, 16, 12, 3, 5, 7, 1, -3, W - 50};
    for (int t = 0; t < 9; ++t) {
        const int c0 = c0s[t], c1 = 3;
        cudaMemset(out, 0xEE, BW * BH);
        tma_probe<<<1, 128, BW * BH>>>(tm, out, c0, c1, BW * BH);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(ho, out, BW * BH, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int y = 0; y < BH; ++y)
            for (int x = 0; x < BW; ++x) {
                const int gx = c0 + x, gy = c1 + y;
                const uint8_t want = (gx >= 0 && gx < W && gy < H) ? h[gy * W + gx] : 0;
                bad += ho[y * BW + x] != want;
            }
        printf("tma u8 box at byte coordinate %5d: %s, %d mismatches\n", c0, cudaGetErrorString(e), bad);
    }
}

int main(int argc, char** argv) {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    g_mhz = khz / 1000.0; g_sms = p.multiProcessorCount;
    setvbuf(stdout, NULL, _IONBF, 0);
    cudaFree(0);
    printf("%s  SMs=%d  clock attr=%.0f MHz\n", p.name, g_sms, g_mhz);
    if (argc > 1 && !strcmp(argv[1], "tma")) { tma_test(); return 0; }
    if (argc > 1 && !strcmp(argv[1], "lsu")) {
        run<K_LDS32, 4, 0, 0, 0, 0, 0, 0>("LDS.32");
        run<K_LDS64, 4, 0, 0, 0, 0, 0, 0>("LDS.64");
        run<K_LDS128, 4, 0, 0, 0, 0, 0, 0>("LDS.128");
        run<K_LDSU8, 4, 0, 0, 0, 0, 0, 0>("LDS.U8");
        run<K_STS128, 4, 0, 0, 0, 0, 0, 0>("STS.128");
        run<K_SHFL, 4, 0, 0, 0, 0, 0, 0>("SHFL");
        run<K_LDS32, 4, K_SHFL, 4, 0, 0, 0, 0>("LDS.32 + SHFL");
        run<K_LDS64, 4, K_SHFL, 4, 0, 0, 0, 0>("LDS.64 + SHFL");
        run<K_LDS32, 2, K_PRMT, 6, 0, 0, 0, 0>("LDS.32 + PRMT");
        run<K_LDS32, 3, K_PRMT, 4, K_FFMA2, 3, 0, 0>("LDS.32 + PRMT + FFMA2");
        run<K_LDS32, 3, K_PRMT, 8, K_FFMA2, 4, K_SHFL, 1>("LDS.32 + PRMT + FFMA2 + SHFL (6:16:8:2 per 2px)");
        run<K_LDS64, 1, K_PRMT, 6, K_FFMA2, 4, K_SHFL, 1>("LDS.64 + PRMT + FFMA2 + SHFL");
        run<K_LDS64, 2, K_PRMT, 9, K_FFMA2, 8, K_I2FU8, 3>("LDS.64 + PRMT + FFMA2 + I2F (2:9:8:3)");
        run<K_LDS64, 2, K_PRMT, 12, K_FFMA2, 8, 0, 0>("LDS.64 + PRMT + FFMA2 (2:12:8)");
        run<K_LDS64, 2, K_PRMT, 12, K_FFMA, 12, 0, 0>("LDS.64 + PRMT + FFMA (2:12:12)");
        run<K_LDS32, 6, K_PRMT, 16, K_FFMA2, 8, 0, 0>("LDS.32 + PRMT + FFMA2 (6:16:8)");
        return 0;
    }
    run<K_PRMT, 4, 0, 0, 0, 0, 0, 0>("PRMT");
    run<K_IADD3, 4, 0, 0, 0, 0, 0, 0>("IADD");
    run<K_FFMA, 4, 0, 0, 0, 0, 0, 0>("FFMA");
    run<K_FFMA2, 4, 0, 0, 0, 0, 0, 0>("FFMA2");
    run<K_IMAD, 4, 0, 0, 0, 0, 0, 0>("IMAD");
    run<K_I2FU8, 4, 0, 0, 0, 0, 0, 0>("I2F.U8.Bk (+LOP3)");
    run<K_HADD2F32, 4, 0, 0, 0, 0, 0, 0>("cvt.f32.f16 (H1)");
    run<K_FMNMX3, 4, 0, 0, 0, 0, 0, 0>("FMNMX3");
    run<K_FSETP, 4, 0, 0, 0, 0, 0, 0>("FSETP + @p IADD");
    run<K_SHFL, 4, 0, 0, 0, 0, 0, 0>("SHFL.DOWN");
    run<K_LDS32, 4, 0, 0, 0, 0, 0, 0>("LDS.32 (+LOP3)");
    run<K_LDS64, 4, 0, 0, 0, 0, 0, 0>("LDS.64 (+2 LOP3)");
    run<K_LDS128, 4, 0, 0, 0, 0, 0, 0>("LDS.128 (+LOP3s)");
    run<K_LDSU8, 4, 0, 0, 0, 0, 0, 0>("LDS.U8 (+LOP3)");
    run<K_STS128, 4, 0, 0, 0, 0, 0, 0>("STS.128");
    run<K_PRMT, 4, K_FFMA2, 4, 0, 0, 0, 0>("PRMT + FFMA2");
    run<K_PRMT, 6, K_FFMA2, 4, 0, 0, 0, 0>("PRMT + FFMA2");
    run<K_PRMT, 4, K_IMAD, 4, 0, 0, 0, 0>("PRMT + IMAD");
    run<K_FFMA2, 4, K_IMAD, 4, 0, 0, 0, 0>("FFMA2 + IMAD");
    run<K_PRMT, 4, K_I2FU8, 1, 0, 0, 0, 0>("PRMT + I2F.U8");
    run<K_PRMT, 3, K_I2FU8, 1, K_FFMA2, 3, 0, 0>("PRMT + I2F.U8 + FFMA2");
    run<K_PRMT, 6, K_FFMA2, 4, K_LDS64, 1, 0, 0>("PRMT + FFMA2 + LDS.64");
    run<K_PRMT, 6, K_FFMA2, 4, K_LDS32, 2, 0, 0>("PRMT + FFMA2 + LDS.32");
    run<K_PRMT, 6, K_FFMA2, 4, K_LDS32, 3, 0, 0>("PRMT + FFMA2 + LDS.32");
    run<K_PRMT, 6, K_FFMA2, 4, K_LDS64, 1, K_SHFL, 1>("PRMT + FFMA2 + LDS.64 + SHFL (kernel-like)");
    run<K_PRMT, 5, K_FFMA2, 4, K_LDS64, 1, K_IMAD, 1>("PRMT + FFMA2 + LDS.64 + IMAD");
    run<K_LDS32, 4, K_SHFL, 4, 0, 0, 0, 0>("LDS.32 + SHFL");
    run<K_LDS64, 4, K_STS128, 1, 0, 0, 0, 0>("LDS.64 + STS.128");
    return 0;
}
