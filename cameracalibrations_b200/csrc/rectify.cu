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
#include <cmath>

#include "rectify_device.cuh"

namespace cc {

constexpr int kChunks = 4;              // 32-pixel chunks per warp along the contiguous axis
constexpr int kTile1 = 32 * kChunks;    // 128
constexpr int kTile2 = 8;               // warps per CTA = second-axis lines per tile
constexpr int kRectThreads = 32 * kTile2;

// ---------------------------------------------------------------------------------
// fp32 single channel, direct gather through L1/L2.
// One warp per output line (fixed I2): everything that depends on I2 only is computed
// once per line; the warp then walks the line in batches of 4 x 32 pixels.  Lane l owns
// pixels a = 128*batch + 32*e + l (e = 0..3): consecutive lanes sample consecutive source
// texels, so one gather request touches ~5 sectors (the first layout, 4 consecutive pixels
// per lane, touched 16.7: profiles/r1_rectify.md) and every store instruction writes one
// full 128-byte line.  Addresses are 32-bit element offsets from a per-frame base.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ const float* elem_ptr(const float* base, unsigned idx) {
    const float* q;
    asm("mad.wide.u32 %0, %1, 4, %2;" : "=l"(q) : "r"(idx), "l"(base));
    return q;
}

__global__ void __launch_bounds__(kRectThreads)
rectify_f32c1_exact(const RectExact p, const RectGeom g, const float* __restrict__ src,
                    float* __restrict__ dst, float fill) {
    const int lane_id = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x * kTile2 + warp;
    if (b >= g.sz2) return;
    const float* s = src + (long long)blockIdx.y * g.frame_stride;
    float* o = dst + (long long)blockIdx.y * g.frame_stride + (long long)b * g.pitch;
    const ColTermD ct = rect_col_term(p, g.axs1 + b);
    const unsigned pitch = (unsigned)g.pitch;

    for (int a0 = lane_id; a0 < g.sz1; a0 += kTile1) {
        unsigned idx[kChunks];
        bool ok[kChunks];
        double d1[kChunks], d2[kChunks];
#pragma unroll
        for (int e = 0; e < kChunks; ++e) {
            double row, col;
            rect_coord(p, ct, g.axs0 + a0 + 32 * e, row, col);
            int i1, i2;
            const bool v1 = lin_pos(row, g.sz1, i1, d1[e]);
            const bool v2 = lin_pos(col, g.sz2, i2, d2[e]);
            ok[e] = v1 & v2;
            idx[e] = (unsigned)i2 * pitch + (unsigned)i1;
        }
        float a00[kChunks], a10[kChunks], a01[kChunks], a11[kChunks];
#pragma unroll
        for (int e = 0; e < kChunks; ++e) {
            const float* q = elem_ptr(s, idx[e]);
            const float* q2 = elem_ptr(s, idx[e] + pitch);
            if (ok[e]) {
                a00[e] = __ldg(q);
                a10[e] = __ldg(q + 1);
                a01[e] = __ldg(q2);
                a11[e] = __ldg(q2 + 1);
            }
        }
#pragma unroll
        for (int e = 0; e < kChunks; ++e) {
            const int a = a0 + 32 * e;
            float v = fill;
            if (ok[e])
                v = (float)bilerp((double)a00[e], (double)a10[e], (double)a01[e], (double)a11[e],
                                  d1[e], d2[e]);
            if (a < g.sz1) __stcs(o + a, v);
        }
    }
}

__global__ void __launch_bounds__(kRectThreads)
rectify_f32c1_fast(const RectFast p, const RectGeom g, const float* __restrict__ src,
                   float* __restrict__ dst, float fill) {
    const int lane_id = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x * kTile2 + warp;
    if (b >= g.sz2) return;
    const float* s = src + (long long)blockIdx.y * g.frame_stride;
    float* o = dst + (long long)blockIdx.y * g.frame_stride + (long long)b * g.pitch;
    const ColTermF ct = rect_col_term(p, g.axs1 + b);
    const unsigned pitch = (unsigned)g.pitch;
    float i1f = (float)(g.axs0 + lane_id) - p.c1;

    for (int a0 = lane_id; a0 < g.sz1; a0 += kTile1, i1f += (float)kTile1) {
        unsigned idx[kChunks];
        bool ok[kChunks];
        float d1[kChunks], d2[kChunks];
#pragma unroll
        for (int e = 0; e < kChunks; ++e) {
            float row, col;
            rect_coord(p, ct, i1f + 32.0f * e, row, col);
            int i1, i2;
            const bool v1 = lin_pos_fast(row, g.sz1, i1, d1[e]);
            const bool v2 = lin_pos_fast(col, g.sz2, i2, d2[e]);
            ok[e] = v1 & v2;
            idx[e] = (unsigned)i2 * pitch + (unsigned)i1;
        }
        float a00[kChunks], a10[kChunks], a01[kChunks], a11[kChunks];
#pragma unroll
        for (int e = 0; e < kChunks; ++e) {
            const float* q = elem_ptr(s, idx[e]);
            const float* q2 = elem_ptr(s, idx[e] + pitch);
            if (ok[e]) {
                a00[e] = __ldg(q);
                a10[e] = __ldg(q + 1);
                a01[e] = __ldg(q2);
                a11[e] = __ldg(q2 + 1);
            }
        }
#pragma unroll
        for (int e = 0; e < kChunks; ++e) {
            const int a = a0 + 32 * e;
            float v = fill;
            if (ok[e]) v = bilerp_fast(a00[e], a10[e], a01[e], a11[e], d1[e], d2[e]);
            if (a < g.sz1) __stcs(o + a, v);
        }
    }
}

// ---------------------------------------------------------------------------------
// u8 x 3 interleaved (RGB{N0f8}), direct gather.  Lane l owns pixels tile + 32*e + l.
// The two taps of one source line are 6 contiguous bytes at byte offset 3*i: fetched as
// three aligned 32-bit words and funnelled with PRMT (2 per line) instead of 6 byte loads.
// A warp's 32 output pixels of one chunk (96 B) are packed through shuffles into
// 24 aligned 32-bit stores... kept simple here: each lane stores its 3 bytes; the L2
// merges them (write traffic is still one pass, see profiles/).
// ---------------------------------------------------------------------------------
struct Taps6 { uint32_t lo, hi; };   // bytes [o, o+4) and [o+4, o+8) of a byte stream

__device__ __forceinline__ Taps6 load6(const uint8_t* __restrict__ base, unsigned o) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(base) + (o >> 2);
    const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
    const unsigned sh = o & 3u;
    const unsigned sel = 0x3210u + 0x1111u * sh;     // bytes sh..sh+3 of the pair
    Taps6 t;
    t.lo = __byte_perm(w0, w1, sel);
    t.hi = __byte_perm(w1, w2, sel);
    return t;
}

__device__ __forceinline__ float byte_f(uint32_t w, int k) { return (float)((w >> (8 * k)) & 0xffu); }
__device__ __forceinline__ double byte_d(uint32_t w, int k) { return (double)((w >> (8 * k)) & 0xffu); }

template <bool EXACT, typename P>
__global__ void __launch_bounds__(kRectThreads)
rectify_u8c3_kernel(const P p, const RectGeom g, const uint8_t* __restrict__ src,
                    uint8_t* __restrict__ dst, uchar3 fill, unsigned src_bytes) {
    const int lane_id = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y * kTile2 + warp;
    if (b >= g.sz2) return;
    const int a_base = blockIdx.x * kTile1 + lane_id;
    const uint8_t* s = src + (long long)blockIdx.z * g.frame_stride * 3;
    uint8_t* o = dst + ((long long)blockIdx.z * g.frame_stride + (long long)b * g.pitch) * 3;
    const auto ct = rect_col_term(p, g.axs1 + b);
    const unsigned pitch3 = (unsigned)g.pitch * 3u;
    // the word-granular gather may read up to 5 bytes past the last needed byte: stay
    // inside the caller's buffer by taking the byte path for the few taps at its very end
    const unsigned frame_off = (unsigned)((reinterpret_cast<uintptr_t>(s)) & 3u);
    const uint8_t* s4 = s - frame_off;                       // 4-byte aligned base
    const unsigned safe_end = src_bytes;                     // bytes of this frame reachable from s

    float i1f0 = 0.f;
    if constexpr (!EXACT) i1f0 = (float)(g.axs0 + a_base) - p.c1;
#pragma unroll
    for (int e = 0; e < kChunks; ++e) {
        const int a = a_base + 32 * e;
        uint32_t r = fill.x, gg = fill.y, bb = fill.z;
        int i1, i2;
        bool ok;
        [[maybe_unused]] double d1d, d2d;
        [[maybe_unused]] float d1f, d2f;
        if constexpr (EXACT) {
            double row, col;
            rect_coord(p, ct, g.axs0 + a, row, col);
            const bool v1 = lin_pos(row, g.sz1, i1, d1d);
            const bool v2 = lin_pos(col, g.sz2, i2, d2d);
            ok = v1 & v2;
        } else {
            float row, col;
            rect_coord(p, ct, i1f0 + 32.0f * e, row, col);
            const bool v1 = lin_pos_fast(row, g.sz1, i1, d1f);
            const bool v2 = lin_pos_fast(col, g.sz2, i2, d2f);
            ok = v1 & v2;
        }
        if (ok) {
            const unsigned off = (unsigned)i2 * pitch3 + (unsigned)i1 * 3u + frame_off;
            Taps6 t0, t1;
            if (off + pitch3 + 12u <= safe_end + frame_off) {
                t0 = load6(s4, off);
                t1 = load6(s4, off + pitch3);
            } else {                                          // last bytes of the buffer
                const uint8_t* q = s4 + off;
                t0.lo = q[0] | (q[1] << 8) | (q[2] << 16) | ((uint32_t)q[3] << 24);
                t0.hi = q[4] | (q[5] << 8);
                q += pitch3;
                t1.lo = q[0] | (q[1] << 8) | (q[2] << 16) | ((uint32_t)q[3] << 24);
                t1.hi = q[4] | (q[5] << 8);
            }
            // t.lo = [a00.r a00.g a00.b a10.r], t.hi = [a10.g a10.b . .]
            if constexpr (EXACT) {
                r  = (uint32_t)(int)rint(bilerp(byte_d(t0.lo, 0), byte_d(t0.lo, 3), byte_d(t1.lo, 0), byte_d(t1.lo, 3), d1d, d2d));
                gg = (uint32_t)(int)rint(bilerp(byte_d(t0.lo, 1), byte_d(t0.hi, 0), byte_d(t1.lo, 1), byte_d(t1.hi, 0), d1d, d2d));
                bb = (uint32_t)(int)rint(bilerp(byte_d(t0.lo, 2), byte_d(t0.hi, 1), byte_d(t1.lo, 2), byte_d(t1.hi, 1), d1d, d2d));
            } else {
                r  = (uint32_t)__float2int_rn(bilerp_fast(byte_f(t0.lo, 0), byte_f(t0.lo, 3), byte_f(t1.lo, 0), byte_f(t1.lo, 3), d1f, d2f));
                gg = (uint32_t)__float2int_rn(bilerp_fast(byte_f(t0.lo, 1), byte_f(t0.hi, 0), byte_f(t1.lo, 1), byte_f(t1.hi, 0), d1f, d2f));
                bb = (uint32_t)__float2int_rn(bilerp_fast(byte_f(t0.lo, 2), byte_f(t0.hi, 1), byte_f(t1.lo, 2), byte_f(t1.hi, 1), d1f, d2f));
                r = min(r, 255u); gg = min(gg, 255u); bb = min(bb, 255u);
            }
        }
        if (a < g.sz1) {
            uint8_t* q = o + 3 * a;
            q[0] = (uint8_t)r; q[1] = (uint8_t)gg; q[2] = (uint8_t)bb;
        }
    }
}

// ---------------------------------------------------------------------------------
// the map alone (FP64): the source coordinate each output pixel samples
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRectThreads)
rectify_map_kernel(const RectExact p, const RectGeom g, double* __restrict__ map_row,
                   double* __restrict__ map_col) {
    const int lane_id = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y * kTile2 + warp;
    if (b >= g.sz2) return;
    const ColTermD ct = rect_col_term(p, g.axs1 + b);
#pragma unroll
    for (int e = 0; e < kChunks; ++e) {
        const int a = blockIdx.x * kTile1 + 32 * e + lane_id;
        if (a < g.sz1) {
            double row, col;
            rect_coord(p, ct, g.axs0 + a, row, col);
            map_row[(long long)b * g.pitch + a] = row;
            map_col[(long long)b * g.pitch + a] = col;
        }
    }
}

// ---------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------
static RectGeom make_geom(const int64_t axs_min[2], int sz1, int sz2, size_t pitch,
                          size_t frame_stride, int nframes) {
    RectGeom g;
    g.axs0 = (int)axs_min[0]; g.axs1 = (int)axs_min[1];
    g.sz1 = sz1; g.sz2 = sz2; g.pitch = (long long)pitch; g.frame_stride = (long long)frame_stride;
    g.nframes = nframes;
    return g;
}

static RectExact make_exact(const ChainD& ch, double ratio) {
    RectExact p;
    for (int i = 0; i < 3; ++i) { p.R0[i] = ch.R[3 * i]; p.R1[i] = ch.R[3 * i + 1]; p.t[i] = ch.t[i]; }
    p.inv_ratio = 1.0 / ratio;     // inv(LinearMap(ratio*I)), src/plot_calibration.jl:16-17
    p.inv_cs = ch.inv_cs; p.k = ch.k;
    p.frow = ch.frow; p.fcol = ch.fcol; p.crow = ch.crow; p.ccol = ch.ccol;
    return p;
}

// fast path: fold scale, 1/ratio and the rotation columns in double, shift the origin to
// the centre of the output so the FP32 products stay small
static RectFast make_fast(const ChainD& ch, double ratio, const RectGeom& g) {
    RectFast p;
    const double sc = (1.0 / ratio) * ch.inv_cs;
    const double c1 = (double)g.axs0 + 0.5 * (g.sz1 - 1), c2 = (double)g.axs1 + 0.5 * (g.sz2 - 1);
    const double c1r = std::nearbyint(c1), c2r = std::nearbyint(c2);
    for (int i = 0; i < 3; ++i) {
        const double A = ch.R[3 * i] * sc, C = ch.R[3 * i + 1] * sc;
        p.A[i] = (float)A; p.Cc[i] = (float)C;
        p.T[i] = (float)(ch.t[i] + A * c1r + C * c2r);
    }
    p.c1 = (float)c1r; p.c2 = (float)c2r;
    p.k = (float)ch.k; p.frow = (float)ch.frow; p.fcol = (float)ch.fcol;
    p.crow = (float)ch.crow; p.ccol = (float)ch.ccol;
    return p;
}

int check_rect_args(const int64_t axs_min[2], int sz1, int sz2, size_t pitch, size_t frame_stride,
                    int nframes, double ratio) {
    CC_REQUIRE(axs_min != nullptr, "axs_min is NULL");
    CC_REQUIRE(sz1 > 0 && sz2 > 0 && nframes >= 0, "frame size must be positive");
    CC_REQUIRE(sz1 < (1 << 22) && sz2 < (1 << 22), "frame extent must be < 2^22");
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
    CC_REQUIRE((flags & CC_GATHER_TMA) == 0, "TMA gather not available for this layout");
    const dim3 grid((sz2 + kTile2 - 1) / kTile2, nframes);
    const RectGeom g = make_geom(axs_min, sz1, sz2, pitch, frame_stride, nframes);
    if (flags & CC_COORD_F32)
        rectify_f32c1_fast<<<grid, kRectThreads, 0, st>>>(make_fast(chd, ratio, g), g, src, dst, fill);
    else
        rectify_f32c1_exact<<<grid, kRectThreads, 0, st>>>(make_exact(chd, ratio), g, src, dst, fill);
    ctx->launches++;
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

int launch_rectify_u8c3(cc_ctx* ctx, const ChainD& chd, double ratio, const int64_t axs_min[2],
                        const uint8_t* src, uint8_t* dst, int sz1, int sz2, size_t pitch,
                        size_t frame_stride, int nframes, const uint8_t fill[3], unsigned flags,
                        cudaStream_t st) {
    if (nframes == 0) return CC_OK;
    CC_REQUIRE((flags & CC_GATHER_TMA) == 0, "TMA gather not available for this layout");
    const dim3 grid = rect_grid(sz1, sz2, nframes);
    const RectGeom g = make_geom(axs_min, sz1, sz2, pitch, frame_stride, nframes);
    const uchar3 f = make_uchar3(fill[0], fill[1], fill[2]);
    // bytes of one frame that belong to the caller: the last line is only sz1 pixels long
    const unsigned frame_bytes = (unsigned)((pitch * (size_t)(sz2 - 1) + (size_t)sz1) * 3);
    if (flags & CC_COORD_F32)
        rectify_u8c3_kernel<false, RectFast><<<grid, kRectThreads, 0, st>>>(make_fast(chd, ratio, g), g, src, dst, f, frame_bytes);
    else
        rectify_u8c3_kernel<true, RectExact><<<grid, kRectThreads, 0, st>>>(make_exact(chd, ratio), g, src, dst, f, frame_bytes);
    ctx->launches++;
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

int launch_rectify_map(cc_ctx* ctx, const ChainD& chd, double ratio, const int64_t axs_min[2],
                       double* map_row, double* map_col, int sz1, int sz2, size_t pitch,
                       cudaStream_t st) {
    const RectGeom g = make_geom(axs_min, sz1, sz2, pitch, pitch * (size_t)sz2, 1);
    rectify_map_kernel<<<rect_grid(sz1, sz2, 1), kRectThreads, 0, st>>>(make_exact(chd, ratio), g,
                                                                        map_row, map_col);
    ctx->launches++;
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

}  // namespace cc
