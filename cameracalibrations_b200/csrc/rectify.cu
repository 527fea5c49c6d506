// rectify.cu -- full-frame rectification (backward warp + bilinear gather): host side.
//
// Replaces warp(img, tform, axs) of src/plot_calibration.jl:40 (tform/axs from
// image_transformations :15-22 and get_axes :1-6) for batches of frames that share one
// view.  The map is never stored in memory: algorithmic traffic is one read and one write of
// every pixel (8 B/px fp32 gray, 6 B/px u8 RGB).
//
// Decomposition.  The first RowCol axis is contiguous in memory.  A tile is 32 consecutive
// first-axis pixels (lane = pixel, so every store instruction writes one full 128-byte line
// and consecutive lanes sample consecutive source texels) x 32 second-axis lines, 8 lines per
// consumer warp.  A work unit is one tile of a GROUP of consecutive frames: the tile's map
// (tap offsets, weights, pixel classes) is built once per unit and reused for every frame of
// the group (rectify_f32c1.cuh, rectify_u8c3.cuh).
//
// This file: the tile plan (per-tile source box, floor constants, valid ranges; built on the
// host once per calibration/geometry, cached in the context, uploaded once), the tensor map of
// the source frames, the persistent grid and its ticket counter (rectify_ring.cuh), and the
// choice between the TMA-staged kernels and the direct ones (pointers/pitches TMA cannot
// address, footprints too large to stage).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "rectify_device.cuh"
#include "tma.cuh"

namespace cc {

constexpr int kT = 32;                 // strip width along the first axis = lanes of a warp
constexpr int kWarps = 4;              // consumer warps per CTA
#ifndef CAMCAL_TL_F32
#define CAMCAL_TL_F32 32
#endif
constexpr int kTLf = CAMCAL_TL_F32;    // f32c1: lines per tile (a warp owns kTLf / kWarps of them)
constexpr int kTLmax = 32;             // most lines of a tile (sizes the per-stage q2 slots)
// most frames one unit rectifies with one map (measured, profiles/r1_rectify.md: the costlier the
// map, the larger the group; the cheaper, the finer the units for the tail of the ticket queue)
#ifndef CAMCAL_FG_F32_EXACT
#define CAMCAL_FG_F32_EXACT 12
#endif
#ifndef CAMCAL_FG_F32_FAST
#define CAMCAL_FG_F32_FAST 4
#endif
#ifndef CAMCAL_FG_U8_EXACT
#define CAMCAL_FG_U8_EXACT 16
#endif
#ifndef CAMCAL_FG_U8_FAST
#define CAMCAL_FG_U8_FAST 16
#endif
constexpr int kFGf32Exact = CAMCAL_FG_F32_EXACT, kFGf32Fast = CAMCAL_FG_F32_FAST, kFGu8Exact = CAMCAL_FG_U8_EXACT, kFGu8Fast = CAMCAL_FG_U8_FAST;
#ifndef CAMCAL_TL_U8
#define CAMCAL_TL_U8 32
#endif
constexpr int kTLu = CAMCAL_TL_U8;     // u8c3: lines per tile
// floor of the exact path, per axis: 0 = FRND.F64.FLOOR (XU pipe), 1 = DADD.RM (FP64 pipe)
#ifndef CAMCAL_FLOOR1
#define CAMCAL_FLOOR1 0
#endif
#ifndef CAMCAL_FLOOR2
#define CAMCAL_FLOOR2 1
#endif
constexpr int kFloorMode1 = CAMCAL_FLOOR1, kFloorMode2 = CAMCAL_FLOOR2;
#ifndef CAMCAL_MAX_STAGES
#define CAMCAL_MAX_STAGES 8
#endif
constexpr int kMaxStages = CAMCAL_MAX_STAGES;
// dynamic shared memory one CTA may spend on its ring (4 resident CTAs of 48 KB + static slots fit in 227 KB)
#ifndef CAMCAL_RING_BYTES
#define CAMCAL_RING_BYTES (48 * 1024)
#endif
// staged line pitch granularity in bytes, per pixel format (16: dense boxes; 128: all 32 banks, see plan_boxes)
#ifndef CAMCAL_PITCH_ALIGN_U8
#define CAMCAL_PITCH_ALIGN_U8 0        // 0: the bank-disjoint pitch of plan_boxes; else round the line up to this many bytes
#endif
#ifndef CAMCAL_PITCH_ALIGN_F32
#define CAMCAL_PITCH_ALIGN_F32 16
#endif
#ifndef CAMCAL_PRODUCER_SLEEP
#define CAMCAL_PRODUCER_SLEEP 256
#endif
constexpr int kProducerSleep = CAMCAL_PRODUCER_SLEEP;   // ns between the producer's probes of a full ring
#ifndef CAMCAL_MINB
#define CAMCAL_MINB 1
#endif
#ifndef CAMCAL_MINB_EXACT
#define CAMCAL_MINB_EXACT 4
#endif
// __launch_bounds__ min CTAs/SM of the staged f32c1 kernels (fast / exact coordinates)
constexpr int kMinBlocks = CAMCAL_MINB, kMinBlocksExact = CAMCAL_MINB_EXACT;
#ifndef CAMCAL_MINB_U8
#define CAMCAL_MINB_U8 1
#endif
#ifndef CAMCAL_MINB_U8_EXACT
#define CAMCAL_MINB_U8_EXACT 3
#endif
constexpr int kMinBlocksU8 = CAMCAL_MINB_U8, kMinBlocksU8Exact = CAMCAL_MINB_U8_EXACT;
constexpr int kConsumerThreads = 32 * kWarps;

struct TileCfg {
    int box1, box2;        // staged box, in pixels (box1 along the contiguous axis)
    int pitch_b;           // bytes between consecutive lines of a staged box (see plan_boxes)
    int stages;
    int ntiles2;           // tiles along the second axis
    int box_bytes;         // bytes one TMA load delivers = stage stride (multiple of 128)
    // persistent f32c1 kernels: work units (strip x, tile y, frame z), x fastest
    int strips;            // tiles along the first axis
    uint32_t units;        // strips * ntiles2 * frame groups
    int fg;                // frames per group: the tile's map is built once per group and reused
    uint32_t widen_mul;    // f32c1 exact: 2^29, the multiplier of the integer float->double widening (a kernel
                           // parameter so that ptxas keeps ONE IMAD.WIDE instead of two shifts, rectify_f32c1.cuh)
    // *_views_kernel only (frames with different views in ONE launch): units are (strip x, tile y, group) with
    // group = view * gpv + frame group of the view; the tile headers / q2 terms of view v start at v * tiles / v * sz2
    int fpv;               // frames per view
    int gpv;               // frame groups per view
    int tiles;             // strips * ntiles2
};

// Frames with different views in one launch (cc_rectify_*_views): what the single-view kernels take as
// launch parameters, per view, in the kernel's parameter space (constant bank: a warp-uniform index).
constexpr int kMaxViews = 64;
struct ViewParams { RectExact pe; RectFast pf; RectGeom g; };
struct ViewTable { ViewParams v[kMaxViews]; };

// f32c1: per-tile header precomputed on the host (RectPlan), read straight from global memory
struct __align__(16) TileHdr {
    double Mk1, Mk2;       // exact: 2^52 - K   (K = global 1-based index of local tap 0)
    float mk1, mk2;        // fast:  1.5*2^23 - K
    uint32_t R1, R2;       // number of valid local first-tap indices per axis (0: nothing staged)
    int x0, y0;            // box origin: TMA coordinates (x0: f32 texels / u8 BYTES of the line, multiple of 16 bytes; y0: lines; may be negative)
    uint32_t base_off;     // byte offset of local tap (0, 0) inside the stage (pixels are 4 or 3 bytes)
    uint32_t view;         // *_views_kernel: index of the unit's view, written by the producer (0 in the plan tables)
};
static_assert(sizeof(TileHdr) == 48, "TileHdr is read as three 16-byte words");

// per-stage slot of the persistent f32c1 kernels, published by the producer warp
struct SmemRing {
    uint64_t full[kMaxStages];
    uint64_t empty[kMaxStages];
    int4 pos[kMaxStages];            // (strip x, tile y, frame z, -); z < 0: no more work
    TileHdr hdr[kMaxStages];
    double q2[kMaxStages][kTLmax];   // exact path: second-axis world term of the lines of each staged tile
};

// global ticket counter of one launch (self-resetting: the last producer zeroes it)
struct RectSched { uint32_t next, done; };

}  // namespace cc
#include "rectify_ring.cuh"
#include "rectify_f32c1.cuh"
#include "rectify_u8c3.cuh"
namespace cc {

// --------------------------------------------------------------------------------------
// the map alone (FP64): the source coordinate each output pixel samples
// --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kConsumerThreads)
rectify_map_kernel(const RectExact p, const RectGeom g, double* __restrict__ map_row,
                   double* __restrict__ map_col) {
    const int lane_id = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int a = blockIdx.x * kT + lane_id;
    if (a >= g.sz1) return;
    const RowTermD rt = rect_row_term(p, g.axs0 + a);
#pragma unroll
    for (int e = 0; e < kT / kWarps; ++e) {
        const int b = blockIdx.y * kT + warp * (kT / kWarps) + e;
        if (b < g.sz2) {
            double row, col;
            rect_coord(p, rt, rect_q2(p, g.axs1 + b), row, col);
            map_row[(long long)b * g.pitch + a] = row;
            map_col[(long long)b * g.pitch + a] = col;
        }
    }
}

// the FP32 fast path's map: the same arithmetic as the fast staged kernels (rect_row_term /
// rect_coord on RectFast), written out as floats
__global__ void __launch_bounds__(kConsumerThreads)
rectify_map_f32_kernel(const RectFast p, const RectGeom g, float* __restrict__ map_row,
                       float* __restrict__ map_col) {
    const int lane_id = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int a = blockIdx.x * kT + lane_id;
    if (a >= g.sz1) return;
    const RowTermF rt = rect_row_term(p, g.axs0 + a);
#pragma unroll
    for (int e = 0; e < kT / kWarps; ++e) {
        const int b = blockIdx.y * kT + warp * (kT / kWarps) + e;
        if (b < g.sz2) {
            float row, col;
            rect_coord(p, rt, (float)(g.axs1 + b) - p.c2, row, col);
            map_row[(long long)b * g.pitch + a] = row;
            map_col[(long long)b * g.pitch + a] = col;
        }
    }
}

// --------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------
static RectGeom make_geom(const int64_t axs_min[2], int sz1, int sz2, size_t pitch,
                          size_t frame_stride, int nframes) {
    RectGeom g;
    memset(&g, 0, sizeof(g));
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
    const double c1r = std::nearbyint((double)g.axs0 + 0.5 * (g.sz1 - 1));
    const double c2r = std::nearbyint((double)g.axs1 + 0.5 * (g.sz2 - 1));
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

// host evaluation of the map (double) for the box-size estimate only
static void host_coord(const ChainD& ch, double inv_ratio, long long I1, long long I2, double* row,
                       double* col) {
    const double q1 = ((double)I1 * inv_ratio) * ch.inv_cs, q2 = ((double)I2 * inv_ratio) * ch.inv_cs;
    const double P1 = ch.R[1] * q2 + (ch.R[0] * q1 + ch.t[0]);
    const double P2 = ch.R[4] * q2 + (ch.R[3] * q1 + ch.t[1]);
    const double P3 = ch.R[7] * q2 + (ch.R[6] * q1 + ch.t[2]);
    const double u = P1 / P3, v = P2 / P3;
    const double radial = 1.0 + ch.k * (u * u + v * v);
    *row = ch.frow * radial * u + ch.crow;
    *col = ch.fcol * radial * v + ch.ccol;
}

// Tile plan.  For every output tile: the bounding box of its source footprint (32 samples on the
// tile perimeter, evaluated in double) and everything the kernels derive from it.  A plan
// depends on the calibration, ratio/axes, frame size and tile shape only -- not on the pixels --
// so it is built once on the host, uploaded once and reused by every later call with the same
// parameters (a video stream): the kernels' producer warp then only issues TMA loads.
struct PlanKey { ChainD ch; double ratio; RectGeom g; int tw, tl, pxb; };

struct RectPlan {
    PlanKey key;
    int n1, n2;                    // tiles along the first / second axis
    int need1, need2;              // largest footprint (pixels), incl. taps and slack
    int box1, box2, box_bytes;     // staged box (pixels) and its size; box_bytes == 0: not stageable
    int pitch_b;                   // bytes per staged line
    int tilt;                      // sign of d(source line)/d(first output index): which way a warp's taps change lines
    std::vector<int> origin;       // per tile: floor(min row), floor(min col) of the perimeter samples
    std::vector<unsigned char> p3_ok;
    std::vector<TileHdr> hdr;      // tile headers
    std::vector<double> q2;        // exact path: second-axis world term of every output line
    TileHdr* d_hdr;
    double* d_q2;
    int per_sm[2];                 // resident CTAs per SM of the staged kernel, [exact], cached (0: not queried yet)
    size_t per_sm_smem[2];
};

static void plan_free(RectPlan* p) {
    if (!p) return;
    if (p->d_hdr) cudaFree(p->d_hdr);
    if (p->d_q2) cudaFree(p->d_q2);
    delete p;
}

static void multi_free(void* mp);

void rectify_free_plans(cc_ctx* ctx) {
    for (int i = 0; i < cc_ctx::NPLAN; ++i) {
        plan_free(static_cast<RectPlan*>(ctx->rect_plans[i]));
        ctx->rect_plans[i] = nullptr;
    }
    for (int i = 0; i < cc_ctx::NMULTI; ++i) {
        multi_free(ctx->multi_plans[i]);
        ctx->multi_plans[i] = nullptr;
    }
}

static void plan_footprints(RectPlan* p, bool threads) {
    const ChainD& ch = p->key.ch;
    const RectGeom& g = p->key.g;
    const int tw = p->key.tw, tl = p->key.tl;
    const double inv_ratio = 1.0 / p->key.ratio;
    p->n1 = (g.sz1 + tw - 1) / tw;
    p->n2 = (g.sz2 + tl - 1) / tl;
    p->origin.assign((size_t)p->n1 * p->n2 * 2, -4);
    p->p3_ok.assign((size_t)p->n1 * p->n2, 1);
    // strips [s0, s1) of the plan; partial maxima / tilt votes of the range go to out[3]
    auto strips = [&](int s0, int s1, long long* out) {
    int m1 = 0, m2 = 0;
    long long tilt = 0;
    for (int t1 = s0; t1 < s1; ++t1) {
        const int a_lo = t1 * tw, a_hi = std::min(a_lo + tw - 1, g.sz1 - 1);
        for (int t2 = 0; t2 < p->n2; ++t2) {
            const int b_lo = t2 * tl, b_hi = std::min(b_lo + tl - 1, g.sz2 - 1);
            double rmin = 0, rmax = 0, cmin = 0, cmax = 0;
            bool finite = true;
            double c_first = 0, c_last = 0;         // source line at the two ends of the tile's first output line
            for (int j = 0; j < 32; ++j) {          // 8 samples on each of the four edges
                const int k = j & 7;
                const bool far = (j & 8) != 0;
                int ca, cb;
                if (j < 16) { ca = far ? a_hi : a_lo; cb = b_lo + ((b_hi - b_lo) * k) / 7; }
                else        { cb = far ? b_hi : b_lo; ca = a_lo + ((a_hi - a_lo) * k) / 7; }
                double r, c;
                host_coord(ch, inv_ratio, g.axs0 + ca, g.axs1 + cb, &r, &c);
                finite = finite && std::isfinite(r) && std::isfinite(c);
                if (j == 0) { rmin = rmax = r; cmin = cmax = c; }
                if (j == 16) c_first = c;
                if (j == 23) c_last = c;
                rmin = std::min(rmin, r); rmax = std::max(rmax, r);
                cmin = std::min(cmin, c); cmax = std::max(cmax, c);
            }
            // P3 is monotone in each output index: the four corners bound it over the tile.  The
            // branch-free reciprocal of the exact kernels needs a sane exponent and one sign.
            double p3min = 0, p3max = 0;
            for (int j = 0; j < 4; ++j) {
                const long long I1 = g.axs0 + ((j & 1) ? a_hi : a_lo), I2 = g.axs1 + ((j & 2) ? b_hi : b_lo);
                const double q1 = ((double)I1 * inv_ratio) * ch.inv_cs, q2 = ((double)I2 * inv_ratio) * ch.inv_cs;
                const double P3 = std::fma(ch.R[7], q2, std::fma(ch.R[6], q1, ch.t[2]));
                if (j == 0) p3min = p3max = P3;
                p3min = std::min(p3min, P3); p3max = std::max(p3max, P3);
            }
            const bool p3_ok = std::isfinite(p3min) && std::isfinite(p3max) &&
                               ((p3min > 1e-270 && p3max < 1e270) || (p3max < -1e-270 && p3min > -1e270));
            const size_t t = (size_t)t1 * p->n2 + t2;
            p->p3_ok[t] = p3_ok ? 1 : 0;
            if (!finite) continue;                  // origin stays at (-4, -4): the range tests fail
            tilt += (c_last > c_first) - (c_last < c_first);
            // clamp so that far-away footprints still give a legal (fully out-of-frame) box
            const double rc = std::min(std::max(rmin, -4.0), (double)g.sz1 + 4.0);
            const double cc_ = std::min(std::max(cmin, -4.0), (double)g.sz2 + 4.0);
            p->origin[2 * t] = (int)std::floor(rc);
            p->origin[2 * t + 1] = (int)std::floor(cc_);
            // tiles entirely outside the frame never gather
            if (rmax < 1.0 || cmax < 1.0 || rmin > g.sz1 || cmin > g.sz2) continue;
            m1 = std::max(m1, (int)(std::floor(rmax) - std::floor(rmin)));
            m2 = std::max(m2, (int)(std::floor(cmax) - std::floor(cmin)));
        }
    }
    out[0] = m1; out[1] = m2; out[2] = tilt;
    };
    // large plans (a 4K frame has 8160 tiles x 36 coordinate evaluations) are split over host threads; every
    // tile writes its own slots of origin / p3_ok, the partials are combined in strip order
    const long long tiles = (long long)p->n1 * p->n2;
    const int nt = !threads ? 1 : (int)std::max(1ll, std::min<long long>({tiles / 512, 16ll, (long long)std::max(1u, std::thread::hardware_concurrency()), (long long)p->n1}));
    std::vector<long long> part((size_t)nt * 3, 0);
    if (nt == 1) {
        strips(0, p->n1, part.data());
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) {
            const int s0 = (int)((long long)p->n1 * t / nt), s1 = (int)((long long)p->n1 * (t + 1) / nt);
            bool spawned = false;
            if (t + 1 < nt) {
                try { th.emplace_back(strips, s0, s1, part.data() + 3 * t); spawned = true; }
                catch (...) {}                  // no thread to be had: this range runs here (nothing throws across the C ABI)
            }
            if (!spawned) strips(s0, s1, part.data() + 3 * t);
        }
        for (auto& x : th) x.join();
    }
    int m1 = 0, m2 = 0;
    long long tilt = 0;
    for (int t = 0; t < nt; ++t) { m1 = std::max(m1, (int)part[3 * t]); m2 = std::max(m2, (int)part[3 * t + 1]); tilt += part[3 * t + 2]; }
    // taps floor-1 .. floor; 2 texels of slack below (origin) and 1 above
    p->need1 = m1 + 2 + 3;
    p->need2 = m2 + 2 + 3;
    p->tilt = tilt >= 0 ? 1 : -1;
}

// box size from the footprint; tile headers
//
// f32c1 (4-byte pixels): the box origin along the first axis is a texel index rounded down to 4
// texels (16 bytes, what TMA can address); dense lines (pitch = box width).
//
// u8c3 (3-byte pixels): the TMA tensor is the line of BYTES, so the origin is a byte offset rounded
// down to 16 bytes -- at most 15 bytes (5 texels) of slack instead of the 15 texels a texel-granular
// origin needs -- and the tile header carries the byte offset of local tap 0.  Line pitch: the lanes
// of a warp read consecutive 3-byte texels (32 lanes span ~24-26 words); on a rotated map the upper
// part of the warp samples the neighbouring source line, `pitch` bytes away.  Both parts hit disjoint
// banks iff the pitch in words is +4 (mod 32) in the direction the source line changes: with a
// multiple of 128 bytes the two lanes at the change share a bank (ncu, round 2: 1.2 wavefronts per
// LDS), with anything else whole groups of lanes collide (1.7 wavefronts with 192-byte lines).
static int floor_to(int v, int m) { return v >= 0 ? (v / m) * m : -(((-v) + m - 1) / m) * m; }

// staged box of a footprint of need1 x need2 texels (tilt: RectPlan.tilt); box_bytes == 0: not worth staging
struct BoxDims { int box1, box2, pitch_b, box_bytes; };

static BoxDims box_dims(int need1, int need2, int tilt, int pxb) {
    BoxDims d;
    if (pxb == 4) {
        const int unit = 4;
        d.box1 = (need1 + unit - 1 + unit - 1) / unit * unit;          // + unit - 1: the origin is rounded down
        const int align = CAMCAL_PITCH_ALIGN_F32;
        d.pitch_b = (d.box1 * pxb + align - 1) / align * align;
        d.box1 = d.pitch_b / pxb;                                      // usable pixels per line
    } else {
        const int need_b = need1 * 3 + 15;                             // + 15: the byte origin is rounded down
        int pitch = (need_b + 15) / 16 * 16;
#if CAMCAL_PITCH_ALIGN_U8 == 0
        const int want = tilt >= 0 ? 4 : 28;                           // pitch in words, mod 32
        const int dense = pitch;
        while ((pitch / 4) % 32 != want) pitch += 16;
        // a TMA box line holds at most 256 elements (bytes here): a footprint whose bank-disjoint pitch does
        // not fit keeps the dense pitch (some bank conflicts) rather than losing the staged path altogether
        if (pitch > 256 && dense <= 256) pitch = dense;
#else
        pitch = (pitch + CAMCAL_PITCH_ALIGN_U8 - 1) / CAMCAL_PITCH_ALIGN_U8 * CAMCAL_PITCH_ALIGN_U8;
#endif
        d.pitch_b = pitch;
        d.box1 = pitch / 3;
    }
    d.box2 = need2;
    while (((size_t)d.pitch_b * d.box2) % 128) ++d.box2;              // every stage base stays 128-byte aligned
    const int box1_elems = (pxb == 4) ? d.pitch_b / 4 : d.pitch_b;
    d.box_bytes = d.pitch_b * d.box2;
    if (box1_elems > 256 || d.box2 > 256 || d.box_bytes > 40 * 1024) d.box_bytes = 0;   // not worth staging
    return d;
}

// tile headers of a plan's footprints for a staged box of the given size (the plan's own, or the common box of
// a group of views: any box at least as large as the plan's footprint gives the same pixels)
static void fill_headers(const RectPlan* p, const BoxDims& d, TileHdr* hdr) {
    const RectGeom& g = p->key.g;
    const int pxb = p->key.pxb;
    const size_t ntiles = (size_t)p->n1 * p->n2;
    for (size_t t = 0; t < ntiles; ++t) {
        TileHdr& h = hdr[t];
        memset(&h, 0, sizeof(h));
        const int y0 = p->origin[2 * t + 1] - 2;
        // first tap g1 (0-based texel index): inside the box (both taps) and inside the frame
        int g_lo, g_hi;
        if (pxb == 4) {
            const int x0 = floor_to(p->origin[2 * t] - 2, 4);
            h.x0 = x0;
            g_lo = std::max(0, x0);
            g_hi = std::min(x0 + d.box1 - 2, g.sz1 - 2);
            h.base_off = (uint32_t)(g_lo - x0) * 4u;
        } else {
            const int xb0 = floor_to((p->origin[2 * t] - 2) * 3, 16);
            h.x0 = xb0;
            g_lo = std::max(0, xb0 >= 0 ? (xb0 + 2) / 3 : 0);                       // 3 g1 >= xb0
            g_hi = std::min((xb0 + d.pitch_b - 6) >= 0 ? (xb0 + d.pitch_b - 6) / 3 : -1, g.sz1 - 2);   // 3 g1 + 6 <= xb0 + pitch
            h.base_off = (uint32_t)(3 * g_lo - xb0);
        }
        const int lo2 = std::max(0, -y0), hi2 = std::min(d.box2 - 2, g.sz2 - 2 - y0);
        h.y0 = y0;
        h.R1 = p->p3_ok[t] ? (uint32_t)std::max(0, g_hi - g_lo + 1) : 0u;
        h.R2 = (uint32_t)std::max(0, hi2 - lo2 + 1);
        const int k1 = 1 + g_lo, k2 = 1 + y0 + lo2;       // global 1-based index of local tap 0
        h.Mk1 = 4503599627370496.0 - (double)k1;
        h.Mk2 = 4503599627370496.0 - (double)k2;
        h.mk1 = 12582912.0f - (float)k1;
        h.mk2 = 12582912.0f - (float)k2;
        h.base_off += (uint32_t)lo2 * (uint32_t)d.pitch_b;
    }
}

static void plan_boxes(RectPlan* p) {
    const RectGeom& g = p->key.g;
    const BoxDims d = box_dims(p->need1, p->need2, p->tilt, p->key.pxb);
    p->box1 = d.box1; p->box2 = d.box2; p->pitch_b = d.pitch_b; p->box_bytes = d.box_bytes;
    // the second-axis world terms do not depend on the box (the views call needs them for unstageable plans too)
    p->q2.resize((size_t)g.sz2);
    const double inv_ratio = 1.0 / p->key.ratio;
    for (int b = 0; b < g.sz2; ++b)       // rect_q2() of rectify_device.cuh, same two products
        p->q2[b] = ((double)((long long)g.axs1 + b) * inv_ratio) * p->key.ch.inv_cs;
    if (!p->box_bytes) return;
    p->hdr.resize((size_t)p->n1 * p->n2);
    fill_headers(p, d, p->hdr.data());
}

// tile headers and q2 terms of a plan to the device (once); false: out of device memory
static bool plan_upload(RectPlan* p, cudaStream_t st) {
    if (p->d_hdr || p->hdr.empty()) return true;
    const size_t hb = p->hdr.size() * sizeof(TileHdr), qb = p->q2.size() * sizeof(double);
    if (cudaMalloc(&p->d_hdr, hb) != cudaSuccess || cudaMalloc(&p->d_q2, qb) != cudaSuccess ||
        cudaMemcpyAsync(p->d_hdr, p->hdr.data(), hb, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(p->d_q2, p->q2.data(), qb, cudaMemcpyHostToDevice, st) != cudaSuccess) {
        cudaGetLastError();
        if (p->d_hdr) cudaFree(p->d_hdr);
        if (p->d_q2) cudaFree(p->d_q2);
        p->d_hdr = nullptr; p->d_q2 = nullptr;
        return false;
    }
    // later calls may come on other streams: make the tables visible to all of them
    cudaStreamSynchronize(st);
    return true;
}

static RectPlan* plan_find(cc_ctx* ctx, const PlanKey& key) {
    for (int i = 0; i < cc_ctx::NPLAN; ++i) {
        RectPlan* p = static_cast<RectPlan*>(ctx->rect_plans[i]);
        if (p && memcmp(&p->key, &key, sizeof(key)) == 0) return p;
    }
    return nullptr;
}

// footprints, box and headers of one parameter set: host arithmetic only (safe to run on several threads;
// threads: split a large plan over host threads itself -- off when the caller already runs one plan per thread)
static RectPlan* plan_build(const PlanKey& key, bool threads) {
    RectPlan* p = new (std::nothrow) RectPlan();
    if (!p) return nullptr;
    p->key = key; p->d_hdr = nullptr; p->d_q2 = nullptr;
    p->per_sm[0] = p->per_sm[1] = 0; p->per_sm_smem[0] = p->per_sm_smem[1] = 0;
    plan_footprints(p, threads);
    plan_boxes(p);
    return p;
}

static void plan_insert(cc_ctx* ctx, RectPlan* p) {
    const int slot = ctx->rect_plan_next++ % cc_ctx::NPLAN;
    if (ctx->rect_plans[slot]) {
        cudaDeviceSynchronize();       // an evicted plan may still be read by a running kernel
        plan_free(static_cast<RectPlan*>(ctx->rect_plans[slot]));
    }
    ctx->rect_plans[slot] = p;
}

static PlanKey plan_key(const ChainD& ch, double ratio, const RectGeom& g, int tw, int tl, int pxb) {
    PlanKey key;
    memset(&key, 0, sizeof(key));
    key.ch = ch; key.ratio = ratio; key.g = g; key.g.nframes = 0; key.g.frame_stride = 0;
    key.tw = tw; key.tl = tl; key.pxb = pxb;
    return key;
}

// find or build the plan of this (calibration, geometry, tile shape); its tables go to the device on first use
static RectPlan* plan_get(cc_ctx* ctx, const ChainD& ch, double ratio, const RectGeom& g, int tw, int tl,
                          int pxb, cudaStream_t st) {
    const PlanKey key = plan_key(ch, ratio, g, tw, tl, pxb);
    if (RectPlan* p = plan_find(ctx, key)) return plan_upload(p, st) ? p : nullptr;
    RectPlan* p = plan_build(key, true);
    if (!p) return nullptr;
    if (!plan_upload(p, st)) {
        plan_free(p);
        return nullptr;
    }
    plan_insert(ctx, p);
    return p;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn encoder(cc_ctx* ctx) {
    if (!ctx->encode_tiled) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            ctx->encode_tiled = fn;
    }
    return reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled);
}

// Decide whether the TMA-staged variant applies and build its tensor map + tile config.
// pxb: bytes per pixel (4: one float element; 3: three u8 elements).
static bool plan_tma(cc_ctx* ctx, const ChainD& ch, double ratio, const RectGeom& g, const void* src,
                     int pxb, int tw, int tl, cudaStream_t st, CUtensorMap* tmap, TileCfg* cfg,
                     RectPlan** plan_out) {
    const size_t pitch_b = (size_t)g.pitch * pxb, frame_b = (size_t)g.frame_stride * pxb;
    if ((reinterpret_cast<uintptr_t>(src) & 15u) || (pitch_b & 15u) || (g.nframes > 1 && (frame_b & 15u)))
        return false;
    EncodeTiledFn enc = encoder(ctx);
    if (!enc) return false;
    RectPlan* plan = plan_get(ctx, ch, ratio, g, tw, tl, pxb, st);
    if (!plan || !plan->box_bytes) return false;
    // measured (profiles/r1_rectify.md): occupancy beats ring depth; small boxes afford more stages
    // the ring hides the TMA round trip: per SM, (stages - 1) x resident CTAs x the USEFUL bytes of a box
    // must cover bandwidth x latency (~30 KB of reads at the HBM roofline, profiles/r2_rectify.md)
    int stages = std::min(kMaxStages, std::max(2, CAMCAL_RING_BYTES / plan->box_bytes));
    if (const char* e = getenv("CAMCAL_STAGES")) stages = std::min(kMaxStages, std::max(1, atoi(e)));   // tuning knob
    while (stages > 2 && stages * plan->box_bytes > 56 * 1024) --stages;

    const int box1_elems = (pxb == 4) ? plan->pitch_b / 4 : plan->pitch_b;
    cuuint64_t dims[3] = {(cuuint64_t)g.sz1 * (pxb == 4 ? 1 : 3), (cuuint64_t)g.sz2,
                          (cuuint64_t)std::max(g.nframes, 1)};
    cuuint64_t strides[2] = {(cuuint64_t)pitch_b, (cuuint64_t)(g.nframes > 1 ? frame_b : pitch_b * g.sz2)};
    cuuint32_t box[3] = {(cuuint32_t)box1_elems, (cuuint32_t)plan->box2, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUtensorMapL2promotion l2promo = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    if (const char* e = getenv("CAMCAL_L2PROMO")) {     // tuning knob: 0 none, 64, 128, 256
        const int v = atoi(e);
        l2promo = v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                : v == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    }
    const CUresult r = enc(tmap, pxb == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 3,
                           const_cast<void*>(src), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, l2promo,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    cfg->box1 = plan->box1; cfg->box2 = plan->box2; cfg->stages = stages; cfg->box_bytes = plan->box_bytes;
    cfg->pitch_b = plan->pitch_b;
    *plan_out = plan;
    return true;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the KERNEL (per device), not to a plan:
// the library remembers the largest value it has set per staged kernel and device and only ever
// raises it (a smaller plan must not lower the limit under a cached larger one, whichever context
// of the process made either call).
static std::mutex g_smem_mu;
static size_t g_smem_attr[64][12];
template <typename K>
static int set_smem(cc_ctx* ctx, int kslot, K kernel, size_t bytes) {
    if (bytes <= 32 * 1024) return CC_OK;     // static smem (barriers + headers) counts against the 48 KB default too
    std::lock_guard<std::mutex> lk(g_smem_mu);
    size_t& cur = g_smem_attr[ctx->device & 63][kslot];
    if (bytes > cur) {
        CC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        cur = bytes;
    }
    return CC_OK;
}

// Ticket counters: a small ring of self-resetting counters per context.  Launches on different
// streams may overlap, so consecutive launches take different counters, and a counter is only
// handed out again after the launch that used it last has finished (event wait on the new stream).
static int sched_acquire(cc_ctx* ctx, cudaStream_t st, RectSched** out) {
    if (!ctx->sched_pool) {
        CC_CUDA(cudaMalloc(&ctx->sched_pool, cc_ctx::NSCHED * sizeof(RectSched)));
        CC_CUDA(cudaMemset(ctx->sched_pool, 0, cc_ctx::NSCHED * sizeof(RectSched)));
        for (int i = 0; i < cc_ctx::NSCHED; ++i)
            CC_CUDA(cudaEventCreateWithFlags(&ctx->sched_event[i], cudaEventDisableTiming));
    }
    const int slot = ctx->sched_next % cc_ctx::NSCHED;
    if (ctx->sched_used[slot] && ctx->sched_stream[slot] != st)
        CC_CUDA(cudaStreamWaitEvent(st, ctx->sched_event[slot], 0));
    *out = static_cast<RectSched*>(ctx->sched_pool) + slot;
    return CC_OK;
}

static void sched_release(cc_ctx* ctx, cudaStream_t st) {
    const int slot = ctx->sched_next++ % cc_ctx::NSCHED;
    cudaEventRecord(ctx->sched_event[slot], st);
    ctx->sched_used[slot] = 1;
    ctx->sched_stream[slot] = st;
}

void rectify_free_sched(cc_ctx* ctx) {
    if (!ctx->sched_pool) return;
    for (int i = 0; i < cc_ctx::NSCHED; ++i) cudaEventDestroy(ctx->sched_event[i]);
    cudaFree(ctx->sched_pool);
    ctx->sched_pool = nullptr;
}

// persistent grid of a staged kernel: every CTA slot of the device (tickets do the balancing)
template <typename K, typename PL>
static int persistent_grid(cc_ctx* ctx, int kslot, K kernel, size_t smem, const TileCfg& cfg, PL* plan, bool exact,
                           uint32_t* gsz) {
    // the occupancy query and the shared-memory attribute cost microseconds per call: once per plan
    int rc = set_smem(ctx, kslot, kernel, smem);
    if (rc) return rc;
    int per_sm = plan->per_sm_smem[exact] == smem ? plan->per_sm[exact] : 0;
    if (per_sm == 0) {
        CC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kConsumerThreads + 32, smem));
        plan->per_sm[exact] = per_sm;
        plan->per_sm_smem[exact] = smem;
    }
    if (const char* e = getenv("CAMCAL_CTAS_PER_SM")) per_sm = std::min(per_sm, std::max(1, atoi(e)));   // tuning knob
    *gsz = std::min<uint32_t>(cfg.units, (uint32_t)ctx->sm_count * (uint32_t)std::max(per_sm, 1));
    if (getenv("CAMCAL_DEBUG"))
        fprintf(stderr, "[camcal] staged: box %dx%d pitch %d (%d B) stages %d units %u (fg %d) grid %u (%d/SM) smem %zu\n",
                cfg.box1, cfg.box2, cfg.pitch_b, cfg.box_bytes, cfg.stages, cfg.units, cfg.fg, *gsz, per_sm, smem);
    return CC_OK;
}

// Frames per group.  The map of a tile (indices, weights) depends on the calibration only, so a
// unit rectifies the same tile of `fg` consecutive frames and computes the map once.  Larger
// groups amortise better; smaller ones leave more units for the tail of the ticket queue: keep
// at least ~8 units per resident CTA slot.
static int unit_cfg(cc_ctx* ctx, TileCfg* cfg, int sz1, int sz2, int nframes, int tw, int tl, int fg_max) {
    cfg->strips = (sz1 + tw - 1) / tw;
    cfg->ntiles2 = (sz2 + tl - 1) / tl;
    const long long tiles = (long long)cfg->strips * cfg->ntiles2;
    const long long want = (long long)ctx->sm_count * 6 * 8;
    int fg = std::min(fg_max, nframes);
    while (fg > 1 && tiles * ((nframes + fg - 1) / fg) < want) fg = (fg + 1) / 2;
    if (const char* e = getenv("CAMCAL_FG")) fg = std::max(1, std::min(nframes, atoi(e)));   // tuning knob
    cfg->fg = fg;
    const long long groups = (nframes + fg - 1) / fg;
    CC_REQUIRE(tiles * groups < (1ll << 31), "too many tiles in one call: split the batch");
    cfg->units = (uint32_t)(tiles * groups);
    return CC_OK;
}

// grid of the direct (unstaged) kernels: ~64 CTAs per SM worth of work, at least 32 lines per CTA
static dim3 direct_grid(cc_ctx* ctx, int sz1, int sz2, int nframes, int* lines) {
    const int strips = (sz1 + kT - 1) / kT;
    const long long per_line_ctas = (long long)strips * nframes;
    const long long segs = std::max<long long>(1, ((long long)ctx->sm_count * 64 + per_line_ctas - 1) / per_line_ctas);
    *lines = (int)std::max<long long>(32, (sz2 + segs - 1) / segs);
    return dim3(strips, (sz2 + *lines - 1) / *lines, nframes);
}

int launch_rectify_f32c1(cc_ctx* ctx, const ChainD& chd, double ratio, const int64_t axs_min[2],
                         const float* src, float* dst, int sz1, int sz2, size_t pitch,
                         size_t frame_stride, int nframes, float fill, unsigned flags,
                         cudaStream_t st) {
    if (nframes == 0) return CC_OK;
    const RectGeom g = make_geom(axs_min, sz1, sz2, pitch, frame_stride, nframes);
    const RectExact pe = make_exact(chd, ratio);
    const RectFast pf = make_fast(chd, ratio, g);
    const bool exact = !(flags & CC_COORD_F32);
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    TileCfg cfg;
    memset(&cfg, 0, sizeof(cfg));
    RectPlan* plan = nullptr;
    const bool tma = !(flags & CC_GATHER_DIRECT) && plan_tma(ctx, chd, ratio, g, src, 4, kT, kTLf, st, &tmap, &cfg, &plan);
    if ((flags & CC_GATHER_TMA) && !tma)
        return set_error(CC_ERR_INVALID_ARG, "TMA gather not available for this layout / footprint");
    int rc = CC_OK;
    if (tma) {
        if ((rc = unit_cfg(ctx, &cfg, sz1, sz2, nframes, kT, kTLf, exact ? kFGf32Exact : kFGf32Fast))) return rc;
        cfg.widen_mul = 0x20000000u;
        // two frames per ring stage when the ring holds at least four boxes (large boxes keep one) and the
        // units have at least two frames (single frames, e.g. the per-view launches of cc_rectify_*_views, keep one)
        const int nf = (kF32FramesPerStage == 2 && cfg.stages >= 4 && cfg.fg >= 2) ? 2 : 1;
        cfg.stages /= nf;
        const size_t smem = (size_t)cfg.stages * nf * cfg.box_bytes;
        uint32_t gsz = 0;
        if (nf == 2) rc = exact ? persistent_grid(ctx, 0, rectify_f32c1_kernel<true>, smem, cfg, plan, true, &gsz)
                                : persistent_grid(ctx, 1, rectify_f32c1_kernel<false>, smem, cfg, plan, false, &gsz);
        else         rc = exact ? persistent_grid(ctx, 4, rectify_f32c1_single_kernel<true>, smem, cfg, plan, true, &gsz)
                                : persistent_grid(ctx, 5, rectify_f32c1_single_kernel<false>, smem, cfg, plan, false, &gsz);
        if (rc) return rc;
        RectSched* sched = nullptr;
        if ((rc = sched_acquire(ctx, st, &sched))) return rc;
        if (nf == 2) {
            if (exact) rectify_f32c1_kernel<true><<<gsz, kConsumerThreads + 32, smem, st>>>(tmap, pe, pf, g, cfg, plan->d_hdr, plan->d_q2, sched, src, dst, fill);
            else       rectify_f32c1_kernel<false><<<gsz, kConsumerThreads + 32, smem, st>>>(tmap, pe, pf, g, cfg, plan->d_hdr, plan->d_q2, sched, src, dst, fill);
        } else {
            if (exact) rectify_f32c1_single_kernel<true><<<gsz, kConsumerThreads + 32, smem, st>>>(tmap, pe, pf, g, cfg, plan->d_hdr, plan->d_q2, sched, src, dst, fill);
            else       rectify_f32c1_single_kernel<false><<<gsz, kConsumerThreads + 32, smem, st>>>(tmap, pe, pf, g, cfg, plan->d_hdr, plan->d_q2, sched, src, dst, fill);
        }
        sched_release(ctx, st);
    } else {
        int lines = 0;
        const dim3 grid = direct_grid(ctx, sz1, sz2, nframes, &lines);
        if (exact) rectify_f32c1_direct_kernel<true><<<grid, kConsumerThreads, 0, st>>>(pe, pf, g, lines, src, dst, fill);
        else       rectify_f32c1_direct_kernel<false><<<grid, kConsumerThreads, 0, st>>>(pe, pf, g, lines, src, dst, fill);
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
    const RectGeom g = make_geom(axs_min, sz1, sz2, pitch, frame_stride, nframes);
    const RectExact pe = make_exact(chd, ratio);
    const RectFast pf = make_fast(chd, ratio, g);
    const bool exact = !(flags & CC_COORD_F32);
    const uchar3 f = make_uchar3(fill[0], fill[1], fill[2]);
    // bytes of one frame that belong to the caller: the last line is only sz1 pixels long
    const unsigned frame_bytes = (unsigned)((pitch * (size_t)(sz2 - 1) + (size_t)sz1) * 3);
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    TileCfg cfg;
    memset(&cfg, 0, sizeof(cfg));
    RectPlan* plan = nullptr;
    // the staged kernel writes whole 32-bit words: 4-byte aligned output lines
    const bool dst_ok = (reinterpret_cast<uintptr_t>(dst) & 3u) == 0 && ((pitch * 3) & 3u) == 0 &&
                        (nframes <= 1 || ((frame_stride * 3) & 3u) == 0);
    const bool tma = !(flags & CC_GATHER_DIRECT) && dst_ok &&
                     plan_tma(ctx, chd, ratio, g, src, 3, kT, kTLu, st, &tmap, &cfg, &plan);
    if ((flags & CC_GATHER_TMA) && !tma)
        return set_error(CC_ERR_INVALID_ARG, "TMA gather not available for this layout / footprint");
    int rc = CC_OK;
    if (tma) {
        if ((rc = unit_cfg(ctx, &cfg, sz1, sz2, nframes, kT, kTLu, exact ? kFGu8Exact : kFGu8Fast))) return rc;
        // + 16: the word-granular gather may read the aligned words that hold the last tap bytes
        cfg.stages = std::max(2, cfg.stages / kU8FramesPerStage);        // a stage holds kU8FramesPerStage boxes
        const size_t smem = (size_t)cfg.stages * kU8FramesPerStage * cfg.box_bytes + 16;
        uint32_t gsz = 0;
        if ((rc = exact ? persistent_grid(ctx, 2, rectify_u8c3_kernel<true>, smem, cfg, plan, true, &gsz)
                        : persistent_grid(ctx, 3, rectify_u8c3_kernel<false>, smem, cfg, plan, false, &gsz))) return rc;
        RectSched* sched = nullptr;
        if ((rc = sched_acquire(ctx, st, &sched))) return rc;
        if (exact)
            rectify_u8c3_kernel<true><<<gsz, kConsumerThreads + 32, smem, st>>>(tmap, pe, pf, g, cfg, plan->d_hdr, plan->d_q2, sched, src, dst, f, frame_bytes);
        else
            rectify_u8c3_kernel<false><<<gsz, kConsumerThreads + 32, smem, st>>>(tmap, pe, pf, g, cfg, plan->d_hdr, plan->d_q2, sched, src, dst, f, frame_bytes);
        sched_release(ctx, st);
    } else {
        int lines = 0;
        const dim3 grid = direct_grid(ctx, sz1, sz2, nframes, &lines);
        if (exact) rectify_u8c3_direct_kernel<true><<<grid, kConsumerThreads, 0, st>>>(pe, pf, g, lines, src, dst, f, frame_bytes);
        else       rectify_u8c3_direct_kernel<false><<<grid, kConsumerThreads, 0, st>>>(pe, pf, g, lines, src, dst, f, frame_bytes);
    }
    ctx->launches++;
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

// --------------------------------------------------------------------------------------
// Frames with different views in ONE launch (cc_rectify_*_views; the reference's plot loop,
// src/plot_calibration.jl:36-42).  The views of a group share the frame layout, so they share the tensor map
// and -- with the largest footprint of the group as the common staged box -- the ring; what differs per view
// (coordinate parameters, axes, tile headers, q2 terms) is indexed by the unit's view.  One launch per view
// (round 2's first version) spent more time in 64 ramp-ups, tails and host launches than in the work.
// --------------------------------------------------------------------------------------
struct MultiPlan {
    std::vector<PlanKey> keys;     // the views of the group, in order
    BoxDims d;                     // the common staged box
    int n1, n2;
    TileHdr* d_hdr;                // [view][tile]
    double* d_q2;                  // [view][sz2]
    int per_sm[2];
    size_t per_sm_smem[2];
};

static void multi_free(void* p) {
    MultiPlan* mp = static_cast<MultiPlan*>(p);
    if (!mp) return;
    if (mp->d_hdr) cudaFree(mp->d_hdr);
    if (mp->d_q2) cudaFree(mp->d_q2);
    delete mp;
}

// find or build the plan of a group of views; nullptr: not stageable (or out of memory) -- the caller falls
// back to one launch per view
static MultiPlan* multi_get(cc_ctx* ctx, const ChainD* chs, const double* ratios, const RectGeom* gs, int nviews,
                            int tw, int tl, int pxb, cudaStream_t st) {
    std::vector<PlanKey> keys((size_t)nviews);
    for (int v = 0; v < nviews; ++v) keys[v] = plan_key(chs[v], ratios[v], gs[v], tw, tl, pxb);
    for (int i = 0; i < cc_ctx::NMULTI; ++i) {
        MultiPlan* mp = static_cast<MultiPlan*>(ctx->multi_plans[i]);
        if (mp && mp->keys.size() == keys.size() && memcmp(mp->keys.data(), keys.data(), keys.size() * sizeof(PlanKey)) == 0)
            return mp;
    }
    // footprints per view: the cached RectPlans, the missing ones built on up to 16 host threads (the plot loop
    // is a one-shot call: 64 cold views cost 64 x ~0.5 ms of host arithmetic on one thread) and cached without
    // their device tables (the group uploads its own).  Copied at once: inserting may evict an earlier plan.
    std::vector<RectPlan*> pl((size_t)nviews, nullptr);
    std::vector<int> missing;
    for (int v = 0; v < nviews; ++v) {
        pl[v] = plan_find(ctx, keys[v]);
        bool dup = false;                                  // the same view twice: build it once
        for (int w = 0; w < v && !pl[v] && !dup; ++w) dup = memcmp(&keys[w], &keys[v], sizeof(PlanKey)) == 0;
        if (!pl[v] && !dup) missing.push_back(v);
    }
    if (!missing.empty()) {
        const int nt = (int)std::min<size_t>(missing.size(), std::min(16u, std::max(1u, std::thread::hardware_concurrency())));
        std::vector<std::thread> th;
        auto work = [&](int t) { for (size_t i = (size_t)t; i < missing.size(); i += (size_t)nt) pl[missing[i]] = plan_build(keys[missing[i]], nt == 1); };
        for (int t = 1; t < nt; ++t) {
            try { th.emplace_back(work, t); }
            catch (...) { work(t); }            // no thread to be had: this share runs here
        }
        work(0);
        for (auto& x : th) x.join();
    }
    std::vector<RectPlan> fp;
    fp.reserve((size_t)nviews);
    int need1 = 0, need2 = 0;
    long long tilt = 0;
    bool oom = false;
    for (int v = 0; v < nviews; ++v) {
        if (!pl[v])                                        // a duplicate of an earlier view (or out of memory)
            for (int w = 0; w < v && !pl[v]; ++w)
                if (memcmp(&keys[w], &keys[v], sizeof(PlanKey)) == 0) pl[v] = pl[w];
        if (!pl[v]) { oom = true; continue; }
        fp.push_back(*pl[v]);
        need1 = std::max(need1, pl[v]->need1); need2 = std::max(need2, pl[v]->need2);
        tilt += pl[v]->tilt;
    }
    for (int v : missing) if (pl[v]) plan_insert(ctx, pl[v]);
    if (oom) return nullptr;
    MultiPlan* mp = new (std::nothrow) MultiPlan();
    if (!mp) return nullptr;
    mp->keys = keys;
    mp->d = box_dims(need1, need2, tilt >= 0 ? 1 : -1, pxb);
    mp->n1 = fp[0].n1; mp->n2 = fp[0].n2;
    mp->d_hdr = nullptr; mp->d_q2 = nullptr;
    mp->per_sm[0] = mp->per_sm[1] = 0; mp->per_sm_smem[0] = mp->per_sm_smem[1] = 0;
    if (getenv("CAMCAL_DEBUG"))
        fprintf(stderr, "[camcal] views: common footprint %d x %d -> box %d x %d pitch %d (%d B)\n", need1, need2, mp->d.box1,
                mp->d.box2, mp->d.pitch_b, mp->d.box_bytes);
    if (!mp->d.box_bytes) { delete mp; return nullptr; }
    const size_t ntiles = (size_t)mp->n1 * mp->n2, sz2 = (size_t)gs[0].sz2;
    std::vector<TileHdr> hdr(ntiles * nviews);
    std::vector<double> q2(sz2 * nviews);
    for (int v = 0; v < nviews; ++v) {
        fill_headers(&fp[v], mp->d, hdr.data() + (size_t)v * ntiles);
        memcpy(q2.data() + (size_t)v * sz2, fp[v].q2.data(), sz2 * sizeof(double));
    }
    const size_t hb = hdr.size() * sizeof(TileHdr), qb = q2.size() * sizeof(double);
    if (cudaMalloc(&mp->d_hdr, hb) != cudaSuccess || cudaMalloc(&mp->d_q2, qb) != cudaSuccess ||
        cudaMemcpyAsync(mp->d_hdr, hdr.data(), hb, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(mp->d_q2, q2.data(), qb, cudaMemcpyHostToDevice, st) != cudaSuccess) {
        cudaGetLastError();
        multi_free(mp);
        return nullptr;
    }
    cudaStreamSynchronize(st);         // the host vectors go away; later calls may come on other streams
    const int slot = ctx->multi_plan_next++ % cc_ctx::NMULTI;
    if (ctx->multi_plans[slot]) {
        cudaDeviceSynchronize();       // an evicted plan may still be read by a running kernel
        multi_free(ctx->multi_plans[slot]);
    }
    ctx->multi_plans[slot] = mp;
    return mp;
}

struct ViewsLaunch {
    CUtensorMap tmap;
    TileCfg cfg;
    ViewTable vt;
    RectGeom g;
    MultiPlan* mp;
};

// tensor map, unit configuration and view table of one group; *ok = false: the single launch does not apply
static int views_prepare(cc_ctx* ctx, const ChainD* chs, const double* ratios, const int64_t* axs_mins, int nviews,
                         const void* src, int pxb, int tl, int sz1, int sz2, size_t pitch, size_t frame_stride, int fpv,
                         int fg_max, cudaStream_t st, ViewsLaunch* L, bool* ok) {
    *ok = false;
    if (nviews < 1 || nviews > kMaxViews || fpv < 1) return CC_OK;
    const int nframes = nviews * fpv;
    const size_t pitch_b = pitch * pxb, frame_b = frame_stride * pxb;
    if ((reinterpret_cast<uintptr_t>(src) & 15u) || (pitch_b & 15u) || (nframes > 1 && (frame_b & 15u))) return CC_OK;
    EncodeTiledFn enc = encoder(ctx);
    if (!enc) return CC_OK;
    std::vector<RectGeom> gs((size_t)nviews);
    for (int v = 0; v < nviews; ++v) gs[v] = make_geom(axs_mins + 2 * v, sz1, sz2, pitch, frame_stride, nframes);
    MultiPlan* mp = multi_get(ctx, chs, ratios, gs.data(), nviews, kT, tl, pxb, st);
    if (!mp) {
        if (getenv("CAMCAL_DEBUG")) fprintf(stderr, "[camcal] views: group of %d not stageable (pxb %d)\n", nviews, pxb);
        return CC_OK;
    }
    int stages = std::min(kMaxStages, std::max(2, CAMCAL_RING_BYTES / mp->d.box_bytes));
    if (const char* e = getenv("CAMCAL_STAGES")) stages = std::min(kMaxStages, std::max(1, atoi(e)));   // tuning knob
    while (stages > 2 && stages * mp->d.box_bytes > 56 * 1024) --stages;
    memset(&L->tmap, 0, sizeof(L->tmap));
    const int box1_elems = (pxb == 4) ? mp->d.pitch_b / 4 : mp->d.pitch_b;
    cuuint64_t dims[3] = {(cuuint64_t)sz1 * (pxb == 4 ? 1 : 3), (cuuint64_t)sz2, (cuuint64_t)nframes};
    cuuint64_t strides[2] = {(cuuint64_t)pitch_b, (cuuint64_t)(nframes > 1 ? frame_b : pitch_b * sz2)};
    cuuint32_t box[3] = {(cuuint32_t)box1_elems, (cuuint32_t)mp->d.box2, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if (enc(&L->tmap, pxb == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 3,
            const_cast<void*>(src), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return CC_OK;
    TileCfg& cfg = L->cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.box1 = mp->d.box1; cfg.box2 = mp->d.box2; cfg.pitch_b = mp->d.pitch_b; cfg.box_bytes = mp->d.box_bytes;
    cfg.stages = stages;
    cfg.strips = (sz1 + kT - 1) / kT;
    cfg.ntiles2 = (sz2 + tl - 1) / tl;
    cfg.tiles = cfg.strips * cfg.ntiles2;
    // frames per group: as unit_cfg (keep ~8 units per resident CTA slot), within a view
    const long long want = (long long)ctx->sm_count * 6 * 8;
    int fg = std::min(fg_max, fpv);
    while (fg > 1 && (long long)cfg.tiles * nviews * ((fpv + fg - 1) / fg) < want) fg = (fg + 1) / 2;
    cfg.fg = fg;
    cfg.fpv = fpv;
    cfg.gpv = (fpv + fg - 1) / fg;
    const long long units = (long long)cfg.tiles * nviews * cfg.gpv;
    CC_REQUIRE(units < (1ll << 31), "too many tiles in one call: split the batch");
    cfg.units = (uint32_t)units;
    cfg.widen_mul = 0x20000000u;
    for (int v = 0; v < nviews; ++v) {
        L->vt.v[v].pe = make_exact(chs[v], ratios[v]);
        L->vt.v[v].pf = make_fast(chs[v], ratios[v], gs[v]);
        L->vt.v[v].g = gs[v];
    }
    for (int v = nviews; v < kMaxViews; ++v) L->vt.v[v] = L->vt.v[0];
    L->g = gs[0];
    L->mp = mp;
    *ok = true;
    return CC_OK;
}

int launch_rectify_f32c1_views(cc_ctx* ctx, const ChainD* chs, const double* ratios, const int64_t* axs_mins, int nviews,
                               const float* src, float* dst, int sz1, int sz2, size_t pitch, size_t frame_stride,
                               int fpv, float fill, unsigned flags, cudaStream_t st, bool* ok) {
    const bool exact = !(flags & CC_COORD_F32);
    *ok = false;
    ViewsLaunch L;                     // 15 KB: the view table goes to the kernel by value
    ViewsLaunch* Lp = &L;
    int rc = views_prepare(ctx, chs, ratios, axs_mins, nviews, src, 4, kTLf, sz1, sz2, pitch, frame_stride, fpv,
                           exact ? kFGf32Exact : kFGf32Fast, st, Lp, ok);
    if (rc || !*ok) return rc;
    const size_t smem = (size_t)Lp->cfg.stages * Lp->cfg.box_bytes;
    uint32_t gsz = 0;
    rc = exact ? persistent_grid(ctx, 6, rectify_f32c1_views_kernel<true>, smem, Lp->cfg, Lp->mp, true, &gsz)
               : persistent_grid(ctx, 7, rectify_f32c1_views_kernel<false>, smem, Lp->cfg, Lp->mp, false, &gsz);
    RectSched* sched = nullptr;
    if (!rc) rc = sched_acquire(ctx, st, &sched);
    if (rc) return rc;
    if (exact) rectify_f32c1_views_kernel<true><<<gsz, kConsumerThreads + 32, smem, st>>>(Lp->tmap, Lp->vt, Lp->g, Lp->cfg, Lp->mp->d_hdr, Lp->mp->d_q2, sched, src, dst, fill);
    else       rectify_f32c1_views_kernel<false><<<gsz, kConsumerThreads + 32, smem, st>>>(Lp->tmap, Lp->vt, Lp->g, Lp->cfg, Lp->mp->d_hdr, Lp->mp->d_q2, sched, src, dst, fill);
    sched_release(ctx, st);
    ctx->launches++;
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

int launch_rectify_u8c3_views(cc_ctx* ctx, const ChainD* chs, const double* ratios, const int64_t* axs_mins, int nviews,
                              const uint8_t* src, uint8_t* dst, int sz1, int sz2, size_t pitch, size_t frame_stride,
                              int fpv, const uint8_t fill[3], unsigned flags, cudaStream_t st, bool* ok) {
    const bool exact = !(flags & CC_COORD_F32);
    *ok = false;
    const int nframes = nviews * fpv;
    // the staged kernel writes whole 32-bit words: 4-byte aligned output lines
    const bool dst_ok = (reinterpret_cast<uintptr_t>(dst) & 3u) == 0 && ((pitch * 3) & 3u) == 0 &&
                        (nframes <= 1 || ((frame_stride * 3) & 3u) == 0);
    if (!dst_ok) return CC_OK;
    ViewsLaunch L;
    ViewsLaunch* Lp = &L;
    int rc = views_prepare(ctx, chs, ratios, axs_mins, nviews, src, 3, kTLu, sz1, sz2, pitch, frame_stride, fpv,
                           exact ? kFGu8Exact : kFGu8Fast, st, Lp, ok);
    if (rc || !*ok) return rc;
    const uchar3 f = make_uchar3(fill[0], fill[1], fill[2]);
    const unsigned frame_bytes = (unsigned)((pitch * (size_t)(sz2 - 1) + (size_t)sz1) * 3);
    Lp->cfg.stages = std::max(2, Lp->cfg.stages / kU8FramesPerStage);
    const size_t smem = (size_t)Lp->cfg.stages * kU8FramesPerStage * Lp->cfg.box_bytes + 16;
    uint32_t gsz = 0;
    rc = exact ? persistent_grid(ctx, 8, rectify_u8c3_views_kernel<true>, smem, Lp->cfg, Lp->mp, true, &gsz)
               : persistent_grid(ctx, 9, rectify_u8c3_views_kernel<false>, smem, Lp->cfg, Lp->mp, false, &gsz);
    RectSched* sched = nullptr;
    if (!rc) rc = sched_acquire(ctx, st, &sched);
    if (rc) return rc;
    if (exact) rectify_u8c3_views_kernel<true><<<gsz, kConsumerThreads + 32, smem, st>>>(Lp->tmap, Lp->vt, Lp->g, Lp->cfg, Lp->mp->d_hdr, Lp->mp->d_q2, sched, src, dst, f, frame_bytes);
    else       rectify_u8c3_views_kernel<false><<<gsz, kConsumerThreads + 32, smem, st>>>(Lp->tmap, Lp->vt, Lp->g, Lp->cfg, Lp->mp->d_hdr, Lp->mp->d_q2, sched, src, dst, f, frame_bytes);
    sched_release(ctx, st);
    ctx->launches++;
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

int rectify_views_per_launch() { return kMaxViews; }

int launch_rectify_map(cc_ctx* ctx, const ChainD& chd, double ratio, const int64_t axs_min[2],
                       double* map_row, double* map_col, int sz1, int sz2, size_t pitch,
                       cudaStream_t st) {
    const RectGeom g = make_geom(axs_min, sz1, sz2, pitch, pitch * (size_t)sz2, 1);
    const dim3 grid((sz1 + kT - 1) / kT, (sz2 + kT - 1) / kT);
    rectify_map_kernel<<<grid, kConsumerThreads, 0, st>>>(make_exact(chd, ratio), g, map_row, map_col);
    ctx->launches++;
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

int launch_rectify_map_f32(cc_ctx* ctx, const ChainD& chd, double ratio, const int64_t axs_min[2],
                           float* map_row, float* map_col, int sz1, int sz2, size_t pitch, cudaStream_t st) {
    const RectGeom g = make_geom(axs_min, sz1, sz2, pitch, pitch * (size_t)sz2, 1);
    const dim3 grid((sz1 + kT - 1) / kT, (sz2 + kT - 1) / kT);
    rectify_map_f32_kernel<<<grid, kConsumerThreads, 0, st>>>(make_fast(chd, ratio, g), g, map_row, map_col);
    ctx->launches++;
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

}  // namespace cc
