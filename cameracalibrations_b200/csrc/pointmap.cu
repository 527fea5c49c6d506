// pointmap.cu -- batched pixel<->world maps over SoA point sets.
//
// Replaces the Julia broadcasts c.(imgpoints, i) / c.(objpoints, i)
// (src/buildcalibrations.jl:29,46) over the callables of src/meta.jl:82,88.
//
// HBM-bound streaming kernels: every array is read or written exactly once with
// 128-bit accesses (double2 / float4), L1 no-allocate loads and evict-first stores;
// a persistent grid of (SM count x resident CTAs) strides over the vectors with two
// independent vectors in flight per thread.  Algorithmic bytes per point:
// img2world 40 B (f64) / 20 B (f32), 32 / 16 B without z; world2img 40 / 20 B
// (32 / 16 B when z == NULL).
#include "chain_device.cuh"

namespace cc {

template <typename T> struct Vec;
template <> struct Vec<double> { using type = double2; static constexpr int N = 2; };
template <> struct Vec<float>  { using type = float4;  static constexpr int N = 4; };

template <typename T> __device__ __forceinline__ T& lane(typename Vec<T>::type& v, int i);
template <> __device__ __forceinline__ double& lane<double>(double2& v, int i) { return i == 0 ? v.x : v.y; }
template <> __device__ __forceinline__ float& lane<float>(float4& v, int i) {
    return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w));
}

constexpr int kThreads = 256;
constexpr int kUnroll = 2;

template <typename T, bool HAS_Z>
__global__ void __launch_bounds__(kThreads)
img2world_kernel(const Chain<T> ch, const T* __restrict__ row, const T* __restrict__ col,
                 T* __restrict__ x, T* __restrict__ y, T* __restrict__ z, size_t n, bool vec_ok) {
    using V = typename Vec<T>::type;
    constexpr int N = Vec<T>::N;
    const size_t tid = (size_t)blockIdx.x * kThreads + threadIdx.x;
    const size_t nthreads = (size_t)gridDim.x * kThreads;
    size_t done = 0;
    if (vec_ok) {
        const size_t nvec = n / N;
        const V* rv = reinterpret_cast<const V*>(row);
        const V* cv = reinterpret_cast<const V*>(col);
        V* xv = reinterpret_cast<V*>(x);
        V* yv = reinterpret_cast<V*>(y);
        V* zv = reinterpret_cast<V*>(z);
        for (size_t i = tid; i < nvec; i += nthreads * kUnroll) {
            V r[kUnroll], c[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const size_t j = i + u * nthreads;
                if (j < nvec) { r[u] = ldg_stream(rv + j); c[u] = ldg_stream(cv + j); }
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const size_t j = i + u * nthreads;
                if (j < nvec) {
                    V ox, oy, oz;
#pragma unroll
                    for (int e = 0; e < N; ++e)
                        img2world(ch, lane<T>(r[u], e), lane<T>(c[u], e), lane<T>(ox, e),
                                  lane<T>(oy, e), lane<T>(oz, e));
                    stg_stream(xv + j, ox);
                    stg_stream(yv + j, oy);
                    if (HAS_Z) stg_stream(zv + j, oz);
                }
            }
        }
        done = nvec * N;
    }
    for (size_t i = done + tid; i < n; i += nthreads) {   // tail / unaligned pointers
        T ox, oy, oz;
        img2world(ch, row[i], col[i], ox, oy, oz);
        x[i] = ox; y[i] = oy;
        if (HAS_Z) z[i] = oz;
    }
}

template <typename T, bool HAS_Z>
__global__ void __launch_bounds__(kThreads)
world2img_kernel(const Chain<T> ch, const T* __restrict__ x, const T* __restrict__ y,
                 const T* __restrict__ z, T* __restrict__ row, T* __restrict__ col, size_t n,
                 bool vec_ok) {
    using V = typename Vec<T>::type;
    constexpr int N = Vec<T>::N;
    const size_t tid = (size_t)blockIdx.x * kThreads + threadIdx.x;
    const size_t nthreads = (size_t)gridDim.x * kThreads;
    size_t done = 0;
    if (vec_ok) {
        const size_t nvec = n / N;
        const V* xv = reinterpret_cast<const V*>(x);
        const V* yv = reinterpret_cast<const V*>(y);
        const V* zv = reinterpret_cast<const V*>(z);
        V* rv = reinterpret_cast<V*>(row);
        V* cv = reinterpret_cast<V*>(col);
        for (size_t i = tid; i < nvec; i += nthreads * kUnroll) {
            V a[kUnroll], b[kUnroll], c[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const size_t j = i + u * nthreads;
                if (j < nvec) {
                    a[u] = ldg_stream(xv + j);
                    b[u] = ldg_stream(yv + j);
                    if (HAS_Z) c[u] = ldg_stream(zv + j);
                }
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const size_t j = i + u * nthreads;
                if (j < nvec) {
                    V orow, ocol;
#pragma unroll
                    for (int e = 0; e < N; ++e)
                        world2img(ch, lane<T>(a[u], e), lane<T>(b[u], e),
                                  HAS_Z ? lane<T>(c[u], e) : T(0), lane<T>(orow, e),
                                  lane<T>(ocol, e));
                    stg_stream(rv + j, orow);
                    stg_stream(cv + j, ocol);
                }
            }
        }
        done = nvec * N;
    }
    for (size_t i = done + tid; i < n; i += nthreads) {
        T orow, ocol;
        world2img(ch, x[i], y[i], HAS_Z ? z[i] : T(0), orow, ocol);
        row[i] = orow; col[i] = ocol;
    }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename K>
static int grid_for(cc_ctx* ctx, K kernel, size_t work_items) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0) != cudaSuccess ||
        per_sm < 1)
        per_sm = 1;
    size_t want = (work_items + (size_t)kThreads * kUnroll - 1) / ((size_t)kThreads * kUnroll);
    size_t cap = (size_t)ctx->sm_count * per_sm;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

template <typename T>
int launch_img2world(cc_ctx* ctx, const ChainD& chd, const T* row, const T* col, T* x, T* y, T* z,
                     size_t n, cudaStream_t st);
template <typename T>
int launch_world2img(cc_ctx* ctx, const ChainD& chd, const T* x, const T* y, const T* z, T* row,
                     T* col, size_t n, cudaStream_t st);

template <typename T> static Chain<T> pick_chain(const ChainD& d);
template <> Chain<double> pick_chain<double>(const ChainD& d) { return d; }
template <> Chain<float> pick_chain<float>(const ChainD& d) { ChainF f; narrow_chain(d, &f); return f; }

template <typename T>
int launch_img2world(cc_ctx* ctx, const ChainD& chd, const T* row, const T* col, T* x, T* y, T* z,
                     size_t n, cudaStream_t st) {
    if (n == 0) return CC_OK;
    const Chain<T> ch = pick_chain<T>(chd);
    const bool vec = aligned16(row) && aligned16(col) && aligned16(x) && aligned16(y) &&
                     (z == nullptr || aligned16(z));
    const size_t items = vec ? n / Vec<T>::N + 1 : n;
    if (z) {
        auto k = img2world_kernel<T, true>;
        k<<<grid_for(ctx, k, items), kThreads, 0, st>>>(ch, row, col, x, y, z, n, vec);
    } else {
        auto k = img2world_kernel<T, false>;
        k<<<grid_for(ctx, k, items), kThreads, 0, st>>>(ch, row, col, x, y, z, n, vec);
    }
    ctx->launches++;
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

template <typename T>
int launch_world2img(cc_ctx* ctx, const ChainD& chd, const T* x, const T* y, const T* z, T* row,
                     T* col, size_t n, cudaStream_t st) {
    if (n == 0) return CC_OK;
    const Chain<T> ch = pick_chain<T>(chd);
    const bool vec = aligned16(x) && aligned16(y) && aligned16(row) && aligned16(col) &&
                     (z == nullptr || aligned16(z));
    const size_t items = vec ? n / Vec<T>::N + 1 : n;
    if (z) {
        auto k = world2img_kernel<T, true>;
        k<<<grid_for(ctx, k, items), kThreads, 0, st>>>(ch, x, y, z, row, col, n, vec);
    } else {
        auto k = world2img_kernel<T, false>;
        k<<<grid_for(ctx, k, items), kThreads, 0, st>>>(ch, x, y, z, row, col, n, vec);
    }
    ctx->launches++;
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

template int launch_img2world<double>(cc_ctx*, const ChainD&, const double*, const double*, double*, double*, double*, size_t, cudaStream_t);
template int launch_img2world<float>(cc_ctx*, const ChainD&, const float*, const float*, float*, float*, float*, size_t, cudaStream_t);
template int launch_world2img<double>(cc_ctx*, const ChainD&, const double*, const double*, const double*, double*, double*, size_t, cudaStream_t);
template int launch_world2img<float>(cc_ctx*, const ChainD&, const float*, const float*, const float*, float*, float*, size_t, cudaStream_t);

}  // namespace cc
