// lm.cu -- one Levenberg-Marquardt step on the normal equations that reproj_jtj_kernel
// accumulates (the solve that OpenCV.calibrateCamera runs inside the
// reference's fit, src/detect_fit.jl:47, flags :40, criteria src/CameraCalibrations.jl:16).
//
// The system is an arrowhead: per view a 6x6 block A_v (rvec, tvec), a 6x4 coupling B_v to the
// shared intrinsics (f, crow, ccol, k) and the shared 4x4 block C:
//     [ A  B ] [de]     [ge]          A_v' = A_v + lambda*diag(A_v)   (Marquardt damping, as
//     [ B' C ] [di] = - [gi]          C'   = C   + lambda*diag(C)      CvLevMarq scales the diagonal)
// Phase 1 (lm_schur_kernel, one thread per view): Cholesky of A_v', Y_v = A_v'^-1 B_v,
//   z_v = A_v'^-1 ge_v, and the view's share of the Schur complement S = sum B_v' Y_v,
//   s = sum B_v' z_v (reduced over views in a fixed order; all-reduced over ranks by the host
//   layer when views are sharded across GPUs -- 20 doubles).
// Phase 2 (lm_update_kernel): (C' - S) di = -(gi - s);  de_v = -(z_v + Y_v di);  candidate
//   parameters = parameters + delta; step and parameter norms for the stopping rule.
// Fixed shared parameters (CALIB_FIX_K1 when with_distortion == false) are masked out.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace cc {

constexpr int kLmThreads = 128;

// in-place Cholesky of a symmetric positive definite N x N matrix (lower triangle); false if a
// pivot is not positive
template <int N>
__device__ __forceinline__ bool cholesky(double (&a)[N][N]) {
    bool ok = true;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        double d = a[j][j];
#pragma unroll
        for (int k = 0; k < j; ++k) d -= a[j][k] * a[j][k];
        ok &= d > 0.0;
        const double l = sqrt(d), il = 1.0 / l;
        a[j][j] = l;
#pragma unroll
        for (int i = j + 1; i < N; ++i) {
            double s = a[i][j];
#pragma unroll
            for (int k = 0; k < j; ++k) s -= a[i][k] * a[j][k];
            a[i][j] = s * il;
        }
    }
    return ok;
}

// solve L L' x = b in place
template <int N>
__device__ __forceinline__ void chol_solve(const double (&a)[N][N], double (&x)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double s = x[i];
#pragma unroll
        for (int k = 0; k < i; ++k) s -= a[i][k] * x[k];
        x[i] = s / a[i][i];
    }
#pragma unroll
    for (int i = N - 1; i >= 0; --i) {
        double s = x[i];
#pragma unroll
        for (int k = i + 1; k < N; ++k) s -= a[k][i] * x[k];
        x[i] = s / a[i][i];
    }
}

// scratch layout (component-major, like reproj_jtj_kernel): [0,16) S, [16,20) s, [20] failed views
constexpr int kSchurComponents = 21;

__global__ void __launch_bounds__(kLmThreads)
lm_schur_kernel(const double* __restrict__ per_view, int nviews, double lambda, double* __restrict__ yz,
                double* __restrict__ scratch) {
    const int v = blockIdx.x * kLmThreads + threadIdx.x;
    if (v >= nviews) return;
    const double* pv = per_view + (size_t)v * CC_PER_VIEW;
    double A[6][6], B[6][4], g[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
        for (int j = 0; j < 6; ++j) A[i][j] = pv[6 * i + j];
#pragma unroll
        for (int j = 0; j < 4; ++j) B[i][j] = pv[36 + 4 * i + j];
        g[i] = pv[60 + i];
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) A[i][i] = fma(lambda, A[i][i], A[i][i]);
    const bool ok = cholesky<6>(A);
    double Y[6][4], z[6];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        double col[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) col[i] = B[i][j];
        chol_solve<6>(A, col);
#pragma unroll
        for (int i = 0; i < 6; ++i) Y[i][j] = col[i];
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) z[i] = g[i];
    chol_solve<6>(A, z);
    double* o = yz + (size_t)v * 30;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) o[4 * i + j] = ok ? Y[i][j] : 0.0;
        o[24 + i] = ok ? z[i] : 0.0;
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < 6; ++i) s = fma(B[i][a], Y[i][b], s);
            scratch[(size_t)(4 * a + b) * nviews + v] = ok ? s : 0.0;
        }
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < 6; ++i) s = fma(B[i][a], z[i], s);
        scratch[(size_t)(16 + a) * nviews + v] = ok ? s : 0.0;
    }
    scratch[(size_t)20 * nviews + v] = ok ? 0.0 : 1.0;
}

// out[c] = sum_v scratch[c][v] in a fixed order (same scheme as residual.cu)
__global__ void __launch_bounds__(256)
lm_reduce_kernel(const double* __restrict__ scratch, int nviews, double* __restrict__ out) {
    __shared__ double sm[256];
    const double* col = scratch + (size_t)blockIdx.x * nviews;
    double s = 0.0;
    for (int v = threadIdx.x; v < nviews; v += 256) s += col[v];
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = sm[0];
}

// delta layout: [0,4) di | [4] sum |de|^2 | [5] sum |pe|^2 | [6] 1 if the 4x4 solve succeeded | [7] spare
__global__ void __launch_bounds__(kLmThreads)
lm_update_kernel(const double* __restrict__ shared, const double* __restrict__ schur, double lambda,
                 unsigned free_mask, const double* __restrict__ yz, const cc_view* __restrict__ vin,
                 int nviews, cc_view* __restrict__ vout, double* __restrict__ delta,
                 double* __restrict__ scratch) {
    __shared__ double di_sm[4];
    __shared__ int ok_sm;
    if (threadIdx.x == 0) {
        double M[4][4], rhs[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
#pragma unroll
            for (int b = 0; b < 4; ++b) M[a][b] = shared[4 * a + b] - schur[4 * a + b];
            M[a][a] = fma(lambda, shared[5 * a], M[a][a]);
            rhs[a] = -(shared[16 + a] - schur[16 + a]);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
            if (!((free_mask >> a) & 1u)) {
#pragma unroll
                for (int b = 0; b < 4; ++b) { M[a][b] = 0.0; M[b][a] = 0.0; }
                M[a][a] = 1.0;
                rhs[a] = 0.0;
            }
        const bool ok = cholesky<4>(M);
        chol_solve<4>(M, rhs);
#pragma unroll
        for (int a = 0; a < 4; ++a) di_sm[a] = ok ? rhs[a] : 0.0;
        ok_sm = ok ? 1 : 0;
        if (blockIdx.x == 0) {
#pragma unroll
            for (int a = 0; a < 4; ++a) delta[a] = di_sm[a];
            delta[6] = ok ? 1.0 : 0.0;
            delta[7] = 0.0;
        }
    }
    __syncthreads();
    const int v = blockIdx.x * kLmThreads + threadIdx.x;
    if (v >= nviews) return;
    const double* o = yz + (size_t)v * 30;
    const double p[6] = {vin[v].rvec[0], vin[v].rvec[1], vin[v].rvec[2],
                         vin[v].tvec[0], vin[v].tvec[1], vin[v].tvec[2]};
    double de[6], n_de = 0.0, n_p = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double s = o[24 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) s = fma(o[4 * i + j], di_sm[j], s);
        de[i] = ok_sm ? -s : 0.0;
        n_de = fma(de[i], de[i], n_de);
        n_p = fma(p[i], p[i], n_p);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        vout[v].rvec[i] = p[i] + de[i];
        vout[v].tvec[i] = p[3 + i] + de[3 + i];
    }
    scratch[v] = n_de;
    scratch[(size_t)nviews + v] = n_p;
}

static int ensure_lm_scratch(cc_ctx* ctx, size_t elems) {
    if (ctx->jtj_scratch_elems >= elems) return CC_OK;
    if (ctx->jtj_scratch) { CC_CUDA(cudaFree(ctx->jtj_scratch)); ctx->jtj_scratch = nullptr; }
    ctx->jtj_scratch_elems = 0;
    CC_CUDA(cudaMalloc(&ctx->jtj_scratch, elems * sizeof(double)));
    ctx->jtj_scratch_elems = elems;
    return CC_OK;
}

int launch_lm_schur(cc_ctx* ctx, const double* per_view, int nviews, double lambda, double* yz,
                    double* schur, cudaStream_t st) {
    const int nv = nviews > 0 ? nviews : 1;
    int rc = ensure_lm_scratch(ctx, (size_t)kSchurComponents * nv);
    if (rc) return rc;
    if (nviews > 0) {
        lm_schur_kernel<<<(nviews + kLmThreads - 1) / kLmThreads, kLmThreads, 0, st>>>(per_view, nviews, lambda, yz,
                                                                                   ctx->jtj_scratch);
        ctx->launches++;
        CC_CUDA(cudaGetLastError());
    }
    lm_reduce_kernel<<<kSchurComponents, 256, 0, st>>>(ctx->jtj_scratch, nviews, schur);
    ctx->launches++;
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

int launch_lm_update(cc_ctx* ctx, const double* shared, const double* schur, double lambda,
                     unsigned free_mask, const double* yz, const cc_view* views_in, int nviews,
                     cc_view* views_out, double* delta, cudaStream_t st) {
    const int nv = nviews > 0 ? nviews : 1;
    int rc = ensure_lm_scratch(ctx, (size_t)kSchurComponents * nv);
    if (rc) return rc;
    // at least one block: block 0 publishes the shared-parameter step even without local views
    const int blocks = std::max(1, (nviews + kLmThreads - 1) / kLmThreads);
    lm_update_kernel<<<blocks, kLmThreads, 0, st>>>(shared, schur, lambda, free_mask, yz, views_in, nviews,
                                                    views_out, delta, ctx->jtj_scratch);
    ctx->launches++;
    CC_CUDA(cudaGetLastError());
    lm_reduce_kernel<<<2, 256, 0, st>>>(ctx->jtj_scratch, nviews, delta + 4);
    ctx->launches++;
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

// launchers of residual.cu
int launch_reproj_jtj(cc_ctx*, const cc_intr*, double, const cc_view*, int, const double*, const double*, int,
                      double*, double*, cudaStream_t);

// The whole fit on one device, host arrays in and out: what a binding calls instead of
// OpenCV.calibrateCamera.  Same loop as cameracalibrations_b200/lm.py::lm_fit (which adds the
// all-reduces for views sharded over ranks): lambda starts at 1e-3, /10 on an accepted step, *10 on
// a rejected one (CvLevMarq's schedule); stops after max_iter steps or when
// |delta| < eps * |parameters| (the reference's CRITERIA: 30, 1e-3).
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 8); }
    template <typename T> T* as() { return static_cast<T*>(p); }
};

int lm_fit_host(cc_ctx* ctx, cc_intr* intr, double aspect, unsigned free_mask, cc_view* views, int nviews,
                const double* obj, const double* img, int ncorners, int max_iter, double eps, double* rms,
                int* iterations) {
    cudaStream_t st = nullptr;
    CC_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } guard{st};
    const size_t nv = (size_t)nviews;
    DevBuf d_views[2], d_pv[2], d_sh[2], d_obj, d_img, d_yz, d_schur, d_delta;
    for (int i = 0; i < 2; ++i) {
        CC_CUDA(d_views[i].alloc(nv * sizeof(cc_view)));
        CC_CUDA(d_pv[i].alloc(nv * CC_PER_VIEW * sizeof(double)));
        CC_CUDA(d_sh[i].alloc(CC_SHARED * sizeof(double)));
    }
    CC_CUDA(d_obj.alloc((size_t)ncorners * 3 * sizeof(double)));
    CC_CUDA(d_img.alloc(nv * ncorners * 2 * sizeof(double)));
    CC_CUDA(d_yz.alloc(nv * CC_LM_YZ * sizeof(double)));
    CC_CUDA(d_schur.alloc(CC_LM_SCHUR * sizeof(double)));
    CC_CUDA(d_delta.alloc(CC_LM_DELTA * sizeof(double)));
    CC_CUDA(cudaMemcpyAsync(d_views[0].p, views, nv * sizeof(cc_view), cudaMemcpyHostToDevice, st));
    CC_CUDA(cudaMemcpyAsync(d_obj.p, obj, (size_t)ncorners * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
    CC_CUDA(cudaMemcpyAsync(d_img.p, img, nv * ncorners * 2 * sizeof(double), cudaMemcpyHostToDevice, st));

    double f = intr->fcol, crow = intr->crow, ccol = intr->ccol, k = (free_mask & 8u) ? intr->k : 0.0;
    auto blocks = [&](int slot, double f_, double crow_, double ccol_, double k_, double* sse) -> int {
        cc_intr in = *intr;
        in.frow = aspect * f_; in.fcol = f_; in.crow = crow_; in.ccol = ccol_; in.k = k_;
        int rc = launch_reproj_jtj(ctx, &in, aspect, d_views[slot].as<cc_view>(), nviews, d_obj.as<double>(),
                                   d_img.as<double>(), ncorners, d_pv[slot].as<double>(), d_sh[slot].as<double>(), st);
        if (rc) return rc;
        CC_CUDA(cudaMemcpyAsync(sse, d_sh[slot].as<double>() + 20, sizeof(double), cudaMemcpyDeviceToHost, st));
        CC_CUDA(cudaStreamSynchronize(st));
        return CC_OK;
    };
    int cur = 0, rc = CC_OK, it = 0;
    double sse = 0.0, lambda = 1e-3;
    if ((rc = blocks(cur, f, crow, ccol, k, &sse))) return rc;
    while (it < max_iter) {
        ++it;
        const int nxt = cur ^ 1;
        if ((rc = launch_lm_schur(ctx, d_pv[cur].as<double>(), nviews, lambda, d_yz.as<double>(),
                                  d_schur.as<double>(), st))) return rc;
        if ((rc = launch_lm_update(ctx, d_sh[cur].as<double>(), d_schur.as<double>(), lambda, free_mask,
                                   d_yz.as<double>(), d_views[cur].as<cc_view>(), nviews,
                                   d_views[nxt].as<cc_view>(), d_delta.as<double>(), st))) return rc;
        double d[CC_LM_DELTA], failed = 0.0;
        CC_CUDA(cudaMemcpyAsync(d, d_delta.p, sizeof(d), cudaMemcpyDeviceToHost, st));
        CC_CUDA(cudaMemcpyAsync(&failed, d_schur.as<double>() + 20, sizeof(double), cudaMemcpyDeviceToHost, st));
        CC_CUDA(cudaStreamSynchronize(st));
        bool bad = failed > 0.0 || d[6] == 0.0;
        for (int i = 0; i < 6; ++i) bad = bad || !std::isfinite(d[i]);
        double sse_c = 0.0;
        const double fc = f + d[0], crc = crow + d[1], ccc = ccol + d[2], kc = k + d[3];
        if (!bad) {
            if ((rc = blocks(nxt, fc, crc, ccc, kc, &sse_c))) return rc;
            bad = !std::isfinite(sse_c) || sse_c >= sse;
        }
        if (bad) { lambda = std::min(lambda * 10.0, 1e16); continue; }
        const double step = std::sqrt(d[4] + d[0] * d[0] + d[1] * d[1] + d[2] * d[2] + d[3] * d[3]);
        const double size = std::sqrt(d[5] + f * f + crow * crow + ccol * ccol + k * k);
        f = fc; crow = crc; ccol = ccc; k = kc; sse = sse_c; cur = nxt;
        lambda = std::max(lambda / 10.0, 1e-16);
        if (step < eps * size) break;
    }
    CC_CUDA(cudaMemcpyAsync(views, d_views[cur].p, nv * sizeof(cc_view), cudaMemcpyDeviceToHost, st));
    CC_CUDA(cudaStreamSynchronize(st));
    intr->frow = aspect * f; intr->fcol = f; intr->crow = crow; intr->ccol = ccol; intr->k = k;
    if (rms) *rms = std::sqrt(sse / (double)std::max<size_t>(1, nv * (size_t)ncorners));
    if (iterations) *iterations = it;
    return CC_OK;
}

}  // namespace cc
