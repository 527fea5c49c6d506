// rectify.cu -- full-frame rectification (backward warp + bilinear gather).
//
// Replaces warp(img, tform, axs) of src/plot_calibration.jl:40 (tform/axs from
// image_transformations :15-22 and get_axes :1-6) for batches of frames that share
// one view.  The map is computed in-kernel and never stored: algorithmic traffic is
// one read + one write of every pixel (8 B/px fp32 gray, 6 B/px u8 RGB).
//
// Thread layout: the first RowCol axis is contiguous in memory, so a warp covers
// 128 consecutive first-axis pixels (4 per lane -> one 128-bit store per lane for
// fp32), a CTA of 8 warps covers 8 consecutive second-axis lines: a 128 x 8 output
// tile whose source footprint is a compact patch that stays in L1/L2.
#include "rectify_device.cuh"

namespace cc {

constexpr int kPx = 4;            // output pixels per thread along the contiguous axis
constexpr int kTile1 = 32 * kPx;  // 128
constexpr int kTile2 = 8;         // warps per CTA
constexpr int kRectThreads = 32 * kTile2;

// ---------------------------------------------------------------------------------
// fp32 single channel, direct gather through L1/L2
// ---------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kRectThreads)
rectify_f32c1_direct(const RectParams<T> p, const float* __restrict__ src, float* __restrict__ dst,
                     float fill, bool vec_ok) {
    const int lane_id = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int a0 = blockIdx.x * kTile1 + lane_id * kPx;
    const int b = blockIdx.y * kTile2 + warp;
    if (b >= p.sz2 || a0 >= p.sz1) return;
    const float* s = src + (long long)blockIdx.z * p.frame_stride;
    float* o = dst + (long long)blockIdx.z * p.frame_stride + (long long)b * p.pitch + a0;

    const ColTerm<T> ct = rect_col_term(p, p.axs1 + b);
    float out[kPx];
    // phase 1: coordinates and addresses for all kPx pixels, phase 2: 16 loads in flight
    int idx[kPx];
    T d1[kPx], d2[kPx];
#pragma unroll
    for (int e = 0; e < kPx; ++e) {
        T row, col;
        rect_coord(p, ct, p.axs0 + a0 + e, row, col);
        int i1, i2;
        const bool ok = lin_pos(row, p.sz1, i1, d1[e]) & lin_pos(col, p.sz2, i2, d2[e]);
        idx[e] = ok ? (int)(i2 * p.pitch + i1) : -1;
    }
    float a00[kPx], a10[kPx], a01[kPx], a11[kPx];
#pragma unroll
    for (int e = 0; e < kPx; ++e) {
        if (idx[e] >= 0) {
            const float* q = s + idx[e];
            a00[e] = __ldg(q);
            a10[e] = __ldg(q + 1);
            a01[e] = __ldg(q + p.pitch);
            a11[e] = __ldg(q + p.pitch + 1);
        }
    }
#pragma unroll
    for (int e = 0; e < kPx; ++e)
        out[e] = idx[e] >= 0 ? (float)bilerp<T>((T)a00[e], (T)a10[e], (T)a01[e], (T)a11[e], d1[e], d2[e])
                             : fill;
    if (vec_ok && a0 + kPx <= p.sz1) {
        stg_stream(reinterpret_cast<float4*>(o), make_float4(out[0], out[1], out[2], out[3]));
    } else {
#pragma unroll
        for (int e = 0; e < kPx; ++e)
            if (a0 + e < p.sz1) o[e] = out[e];
    }
}

// ---------------------------------------------------------------------------------
// u8 x 3 interleaved (RGB{N0f8}), direct gather.  Each tap is 3 bytes at an arbitrary
// byte offset; the two taps of one source line are 6 contiguous bytes.  The output of
// a warp (128 px = 384 B) is transposed through shared memory so that 24 lanes issue
// one aligned 128-bit store each.
// ---------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T u8_round(T v);
template <> __device__ __forceinline__ double u8_round<double>(double v) { return rint(v); }
template <> __device__ __forceinline__ float u8_round<float>(float v) { return rintf(v); }

template <typename T>
__global__ void __launch_bounds__(kRectThreads)
rectify_u8c3_direct(const RectParams<T> p, const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                    uchar3 fill, int store_mode /*16, 4 or 1 byte stores*/) {
    __shared__ __align__(16) uint32_t stage[kTile2][32 * 3];
    const int lane_id = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int a0 = blockIdx.x * kTile1 + lane_id * kPx;
    const int b = blockIdx.y * kTile2 + warp;
    if (b >= p.sz2) return;                       // warp-uniform
    const uint8_t* s = src + (long long)blockIdx.z * p.frame_stride * 3;
    uint8_t* orow = dst + ((long long)blockIdx.z * p.frame_stride + (long long)b * p.pitch) * 3;
    const long long pitch3 = p.pitch * 3;

    uint8_t px[kPx][3];
    const ColTerm<T> ct = rect_col_term(p, p.axs1 + b);
#pragma unroll
    for (int e = 0; e < kPx; ++e) {
        px[e][0] = fill.x; px[e][1] = fill.y; px[e][2] = fill.z;
        if (a0 + e < p.sz1) {
            T row, col, d1, d2;
            rect_coord(p, ct, p.axs0 + a0 + e, row, col);
            int i1, i2;
            if (lin_pos(row, p.sz1, i1, d1) & lin_pos(col, p.sz2, i2, d2)) {
                const uint8_t* q = s + (i2 * p.pitch + i1) * 3;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const T v = bilerp<T>((T)__ldg(q + c), (T)__ldg(q + 3 + c),
                                          (T)__ldg(q + pitch3 + c), (T)__ldg(q + pitch3 + 3 + c),
                                          d1, d2);
                    px[e][c] = (uint8_t)(int)u8_round<T>(v);   // weights in [0,1]: never leaves [0,255]
                }
            }
        }
    }
    // 12 bytes per lane -> 3 little-endian words
    const uint32_t w0 = px[0][0] | (px[0][1] << 8) | (px[0][2] << 16) | ((uint32_t)px[1][0] << 24);
    const uint32_t w1 = px[1][1] | (px[1][2] << 8) | (px[2][0] << 16) | ((uint32_t)px[2][1] << 24);
    const uint32_t w2 = px[2][2] | (px[3][0] << 8) | (px[3][1] << 16) | ((uint32_t)px[3][2] << 24);
    const int tile_a = blockIdx.x * kTile1;       // first pixel of this warp's 128-px run
    const int valid_px = min(kTile1, p.sz1 - tile_a);
    const int valid_bytes = valid_px * 3;
    uint8_t* obase = orow + (long long)tile_a * 3;
    if (store_mode == 16) {
        stage[warp][lane_id * 3 + 0] = w0;
        stage[warp][lane_id * 3 + 1] = w1;
        stage[warp][lane_id * 3 + 2] = w2;
        __syncwarp();
        if (lane_id < 24) {
            const uint4 v = *reinterpret_cast<const uint4*>(&stage[warp][lane_id * 4]);
            const int off = lane_id * 16;
            if (off + 16 <= valid_bytes) {
                stg_stream(reinterpret_cast<uint4*>(obase + off), v);
            } else if (off < valid_bytes) {
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                for (int j = 0; off + j < valid_bytes; ++j)
                    obase[off + j] = (uint8_t)(w[j >> 2] >> (8 * (j & 3)));
            }
        }
    } else if (store_mode == 4 && a0 + kPx <= p.sz1) {
        uint32_t* ow = reinterpret_cast<uint32_t*>(obase + lane_id * 12);
        ow[0] = w0; ow[1] = w1; ow[2] = w2;
    } else {
#pragma unroll
        for (int e = 0; e < kPx; ++e)
            if (a0 + e < p.sz1) {
                uint8_t* q = obase + (lane_id * kPx + e) * 3;
                q[0] = px[e][0]; q[1] = px[e][1]; q[2] = px[e][2];
            }
    }
}

// ---------------------------------------------------------------------------------
// the map alone (FP64): the source coordinate each output pixel samples
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRectThreads)
rectify_map_kernel(const RectParams<double> p, double* __restrict__ map_row,
                   double* __restrict__ map_col) {
    const int lane_id = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int a0 = blockIdx.x * kTile1 + lane_id * kPx;
    const int b = blockIdx.y * kTile2 + warp;
    if (b >= p.sz2) return;
    const ColTerm<double> ct = rect_col_term(p, p.axs1 + b);
#pragma unroll
    for (int e = 0; e < kPx; ++e)
        if (a0 + e < p.sz1) {
            double row, col;
            rect_coord(p, ct, p.axs0 + a0 + e, row, col);
            map_row[(long long)b * p.pitch + a0 + e] = row;
            map_col[(long long)b * p.pitch + a0 + e] = col;
        }
}

// ---------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------
template <typename T>
static RectParams<T> make_params(const ChainD& chd, double ratio, const int64_t axs_min[2], int sz1,
                                 int sz2, size_t pitch, size_t frame_stride, int nframes);
template <>
RectParams<double> make_params<double>(const ChainD& chd, double ratio, const int64_t axs_min[2],
                                       int sz1, int sz2, size_t pitch, size_t frame_stride,
                                       int nframes) {
    RectParams<double> p;
    p.ch = chd;
    p.inv_ratio = 1.0 / ratio;     // inv(LinearMap(ratio*I)), src/plot_calibration.jl:16-17
    p.axs0 = (int)axs_min[0]; p.axs1 = (int)axs_min[1];
    p.sz1 = sz1; p.sz2 = sz2; p.pitch = (long long)pitch; p.frame_stride = (long long)frame_stride;
    p.nframes = nframes;
    return p;
}
template <>
RectParams<float> make_params<float>(const ChainD& chd, double ratio, const int64_t axs_min[2],
                                     int sz1, int sz2, size_t pitch, size_t frame_stride,
                                     int nframes) {
    RectParams<float> p;
    narrow_chain(chd, &p.ch);
    p.inv_ratio = (float)(1.0 / ratio);
    p.axs0 = (int)axs_min[0]; p.axs1 = (int)axs_min[1];
    p.sz1 = sz1; p.sz2 = sz2; p.pitch = (long long)pitch; p.frame_stride = (long long)frame_stride;
    p.nframes = nframes;
    return p;
}

int check_rect_args(const int64_t axs_min[2], int sz1, int sz2, size_t pitch, size_t frame_stride,
                    int nframes, double ratio) {
    CC_REQUIRE(axs_min != nullptr, "axs_min is NULL");
    CC_REQUIRE(sz1 > 0 && sz2 > 0 && nframes >= 0, "frame size must be positive");
    CC_REQUIRE(pitch >= (size_t)sz1, "pitch smaller than sz1");
    CC_REQUIRE(nframes <= 1 || frame_stride >= pitch * (size_t)(sz2 - 1) + sz1, "frames overlap");
    CC_REQUIRE(pitch * (size_t)sz2 < (size_t)1 << 30, "frame too large (pitch*sz2 must be < 2^30)");
    CC_REQUIRE(nframes <= 65535, "at most 65535 frames per call");
    CC_REQUIRE(ratio > 0.0 && ratio == ratio, "ratio must be positive");
    const int64_t lim = (int64_t)1 << 30;
    CC_REQUIRE(axs_min[0] > -lim && axs_min[0] < lim && axs_min[1] > -lim && axs_min[1] < lim,
               "axs_min out of range");
    return CC_OK;
}

static dim3 rect_grid(int sz1, int sz2, int nframes) {
    return dim3((sz1 + kTile1 - 1) / kTile1, (sz2 + kTile2 - 1) / kTile2, nframes);
}

int launch_rectify_f32c1(cc_ctx* ctx, const ChainD& chd, double ratio, const int64_t axs_min[2],
                         const float* src, float* dst, int sz1, int sz2, size_t pitch,
                         size_t frame_stride, int nframes, float fill, unsigned flags,
                         cudaStream_t st) {
    if (nframes == 0) return CC_OK;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) && (pitch % 4 == 0) &&
                        (frame_stride % 4 == 0);
    const dim3 grid = rect_grid(sz1, sz2, nframes);
    if (flags & CC_COORD_F32) {
        rectify_f32c1_direct<float><<<grid, kRectThreads, 0, st>>>(
            make_params<float>(chd, ratio, axs_min, sz1, sz2, pitch, frame_stride, nframes), src, dst,
            fill, vec_ok);
    } else {
        rectify_f32c1_direct<double><<<grid, kRectThreads, 0, st>>>(
            make_params<double>(chd, ratio, axs_min, sz1, sz2, pitch, frame_stride, nframes), src, dst,
            fill, vec_ok);
    }
    ctx->launches++;
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

int launch_rectify_u8c3(cc_ctx* ctx, const ChainD& chd, double ratio, const int64_t axs_min[2],
                        const uint8_t* src, uint8_t* dst, int sz1, int sz2, size_t pitch,
                        size_t frame_stride, int nframes, const uint8_t fill[3], unsigned flags,
                        cudaStream_t st) {
    if (nframes == 0) return CC_OK;
    const uintptr_t d = reinterpret_cast<uintptr_t>(dst);
    int store_mode = 1;
    if ((d & 15u) == 0 && (pitch * 3) % 16 == 0 && (frame_stride * 3) % 16 == 0) store_mode = 16;
    else if ((d & 3u) == 0 && (pitch * 3) % 4 == 0 && (frame_stride * 3) % 4 == 0) store_mode = 4;
    const dim3 grid = rect_grid(sz1, sz2, nframes);
    const uchar3 f = make_uchar3(fill[0], fill[1], fill[2]);
    if (flags & CC_COORD_F32) {
        rectify_u8c3_direct<float><<<grid, kRectThreads, 0, st>>>(
            make_params<float>(chd, ratio, axs_min, sz1, sz2, pitch, frame_stride, nframes), src, dst,
            f, store_mode);
    } else {
        rectify_u8c3_direct<double><<<grid, kRectThreads, 0, st>>>(
            make_params<double>(chd, ratio, axs_min, sz1, sz2, pitch, frame_stride, nframes), src, dst,
            f, store_mode);
    }
    ctx->launches++;
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

int launch_rectify_map(cc_ctx* ctx, const ChainD& chd, double ratio, const int64_t axs_min[2],
                       double* map_row, double* map_col, int sz1, int sz2, size_t pitch,
                       cudaStream_t st) {
    rectify_map_kernel<<<rect_grid(sz1, sz2, 1), kRectThreads, 0, st>>>(
        make_params<double>(chd, ratio, axs_min, sz1, sz2, pitch, pitch * (size_t)sz2, 1), map_row,
        map_col);
    ctx->launches++;
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

}  // namespace cc
