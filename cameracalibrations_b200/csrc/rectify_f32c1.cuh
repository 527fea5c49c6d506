// rectify_f32c1.cuh -- fp32 single-channel rectification kernels (included by rectify.cu).
//
// The per-pixel instruction budget bounds these kernels (profiles/r1_rectify.md), so the staged
// path is written to the instruction:
//  * the producer folds the tile's index origin into the floor constant: the exact path adds
//    Mk = 2^52 - K to the coordinate with round-down (DADD.RM is the floor) or to
//    FRND.F64.FLOOR(x), the fast path adds mk = 1.5*2^23 - K with FADD2.RM; the low word of the
//    sum is magic + tile-local tap index.  "high word == 0x43300000" and
//    "max over the batch of (bits - magic) < R" are the complete range tests (negative, NaN,
//    huge and out-of-box coordinates all fail them), and the smem address takes the raw bits
//    (the magic is folded into the tile's base address);
//  * 1/P3 is NVIDIA's own correctly rounded sequence (MUFU.RCP64H + 5 DFMA) inlined without
//    its per-pixel branch: its validity test (exponent of P3 not extreme) is made once per
//    tile by the producer on the tile corners -- P3 is affine in the output index;
//  * one vote per batch of lines decides between the staged gather and the generic path,
//    which lives out of line so that the hot loop keeps its registers.
#pragma once

namespace cc {

// ---- generic per-pixel path: every check, direct global taps ---------------------------
template <bool EXACT>
__device__ __forceinline__ float sample_direct_f32(const RectExact& pe, const RectFast& pf,
                                                const RowTermD& rtd, const RowTermF& rtf,
                                                const RectGeom& g, const float* __restrict__ sframe,
                                                unsigned pitch, int b, float fill) {
    if (EXACT) {
        double row, col, d1, d2;
        int i1, i2;
        rect_coord(pe, rtd, rect_q2(pe, g.axs1 + b), row, col);
        if (!(lin_ok(row, g.sz1) & lin_ok(col, g.sz2))) return fill;
        lin_floor(row, i1, d1);
        lin_floor(col, i2, d2);
        lin_fix_edge(g.sz1, i1, d1);
        lin_fix_edge(g.sz2, i2, d2);
        const float* q = sframe + ((unsigned)(i2 - 1) * pitch + (unsigned)(i1 - 1));
        return (float)bilerp((double)__ldg(q), (double)__ldg(q + 1), (double)__ldg(q + pitch),
                             (double)__ldg(q + pitch + 1), d1, d2);
    } else {
        float row, col, d1, d2;
        int t1, t2;
        rect_coord(pf, rtf, (float)(g.axs1 + b) - pf.c2, row, col);
        lin_floor_fast(row, t1, d1);
        lin_floor_fast(col, t2, d2);
        const int g1 = t1 - (kMagicBits + 1), g2 = t2 - (kMagicBits + 1);
        if (!(((unsigned)g1 <= (unsigned)(g.sz1 - 2)) & ((unsigned)g2 <= (unsigned)(g.sz2 - 2)))) return fill;
        const float* q = sframe + ((unsigned)g2 * pitch + (unsigned)g1);
        return bilerp_fast(__ldg(q), __ldg(q + 1), __ldg(q + pitch), __ldg(q + pitch + 1), d1, d2);
    }
}

// `n` consecutive lines of one lane's pixel column through the generic path (cold)
#ifndef CAMCAL_GENERIC_INLINE
#define CAMCAL_GENERIC_INLINE 1
#endif
#if CAMCAL_GENERIC_INLINE
#define CC_GENERIC_ATTR __forceinline__
#else
#define CC_GENERIC_ATTR __noinline__
#endif
template <bool EXACT>
__device__ CC_GENERIC_ATTR void generic_lines_f32(const RectExact* pe, const RectFast* pf, const RectGeom* g,
                                               const float* __restrict__ sframe, float* __restrict__ o,
                                               int a, int b, int n, float fill) {
    if (a >= g->sz1) return;
    RowTermD rtd;
    RowTermF rtf;
    if (EXACT) rtd = rect_row_term(*pe, g->axs0 + a); else rtf = rect_row_term(*pf, g->axs0 + a);
    const unsigned pitch = (unsigned)g->pitch;
    n = min(n, g->sz2 - b);
#pragma unroll 1
    for (int e = 0; e < n; ++e, o += pitch)
        __stcs(o, sample_direct_f32<EXACT>(*pe, *pf, rtd, rtf, *g, sframe, pitch, b + e, fill));
}

__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
template <int OFF>
__device__ __forceinline__ float lds_f32_off(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF));
    return v;
}

// ---- direct kernel: no staging (unaligned layouts, footprints too large for a box) -------
template <bool EXACT>
__global__ void __launch_bounds__(kConsumerThreads)
rectify_f32c1_direct_kernel(const __grid_constant__ RectExact pe, const __grid_constant__ RectFast pf,
                            const __grid_constant__ RectGeom g, const int lines_per_cta,
                            const float* __restrict__ src, float* __restrict__ dst, float fill) {
    const int lane_id = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int a = blockIdx.x * kT + lane_id;
    if (a >= g.sz1) return;
    const int frame = blockIdx.z;
    const float* sframe = src + (long long)frame * g.frame_stride;
    const unsigned pitch = (unsigned)g.pitch;
    RowTermD rtd;
    RowTermF rtf;
    if (EXACT) rtd = rect_row_term(pe, g.axs0 + a); else rtf = rect_row_term(pf, g.axs0 + a);
    const int b_begin = blockIdx.y * lines_per_cta;
    const int b_end = min(b_begin + lines_per_cta, g.sz2);
    float* o = dst + (long long)frame * g.frame_stride + (long long)(b_begin + warp) * g.pitch + a;
    for (int b = b_begin + warp; b < b_end; b += kWarps, o += (long long)kWarps * g.pitch)
        __stcs(o, sample_direct_f32<EXACT>(pe, pf, rtd, rtf, g, sframe, pitch, b, fill));
}

// ---- staged kernel (persistent; scheduling and producer: rectify_ring.cuh) ------------------
template <bool EXACT>
__global__ void __launch_bounds__(kConsumerThreads + 32, EXACT ? kMinBlocksExact : kMinBlocks)
rectify_f32c1_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ RectExact pe,
                     const __grid_constant__ RectFast pf, const __grid_constant__ RectGeom g,
                     const __grid_constant__ TileCfg cfg, const TileHdr* __restrict__ plan,
                     const double* __restrict__ q2tab, RectSched* __restrict__ sched,
                     const float* __restrict__ src, float* __restrict__ dst, float fill) {
    constexpr int KB = EXACT ? kBatchExact : kBatchFast;
    constexpr int TL = kTLf;                      // lines per tile
    constexpr int WX = kWXf, WY = kWarps / WX;    // consumer warps across / down the tile
    constexpr int LPW = TL / WY;                  // lines per warp per tile
    static_assert(LPW % KB == 0, "batch must divide the lines of a warp");
    extern __shared__ __align__(128) uint8_t stage_mem[];
    __shared__ SmemRing ring;
    const int lane_id = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform for the compiler
    ring_init(&ring, cfg.stages);

    if (warp == kWarps) {                              // ---- producer warp
        producer_loop<EXACT, TL, 1>(&tmap, g, cfg, plan, q2tab, sched, &ring, stage_mem, lane_id);
        return;
    }

    // ---- consumer warps
    const int wx = warp % WX, wy = warp / WX;
    const unsigned pitch = (unsigned)g.pitch;
    const uint32_t box_pitch_b = (uint32_t)cfg.box1 * 4u;
    const uint32_t stage0 = smem_u32(stage_mem);

    int s = 0;
    uint32_t phase = 0;
    for (;;) {
        mbar_wait(&ring.full[s], phase);
        const int4 pos = ring.pos[s];
        if (pos.z < 0) break;
        const TileHdr* h = &ring.hdr[s];
        const int a_w = pos.x * (kT * WX) + wx * kT;           // first pixel of this warp's lanes
        const int a = a_w + lane_id;
        const int b0 = pos.y * TL + wy * LPW;
        const float* sframe = src + (long long)pos.z * g.frame_stride;
        float* optr = dst + (long long)pos.z * g.frame_stride + (long long)b0 * g.pitch + a;
        [[maybe_unused]] RowTermD rtd;
        [[maybe_unused]] RowTermF rtf;
        const int a_c = min(a, g.sz1 - 1);             // out-of-frame lanes shadow the last pixel
        if (EXACT) rtd = rect_row_term(pe, g.axs0 + a_c); else rtf = rect_row_term(pf, g.axs0 + a_c);
        [[maybe_unused]] double Mk1 = 0, Mk2 = 0;
        [[maybe_unused]] float mk1 = 0, mk2 = 0;
        if (EXACT) { Mk1 = h->Mk1; Mk2 = h->Mk2; } else { mk1 = h->mk1; mk2 = h->mk2; }
        uint32_t R1 = h->R1;
        const uint32_t R2 = h->R2;
        if (!(a_w + kT <= g.sz1 && b0 + LPW <= g.sz2)) R1 = 0;   // partial lines: everything generic
        // raw magic-biased bits index the box directly: fold the bias into the base
        const uint32_t magic = EXACT ? 0u : (uint32_t)kMagicBits;
        const uint32_t base = stage0 + (uint32_t)s * (uint32_t)cfg.box_bytes + h->base_off - magic * (box_pitch_b + 4u);
        [[maybe_unused]] float2 ip;
        ip.x = (float)(g.axs1 + b0) - pf.c2;
        ip.y = ip.x + 1.0f;
        [[maybe_unused]] const double* q2p = &ring.q2[s][wy * LPW];
#pragma unroll 1
        for (int batch = 0; batch < LPW / KB; ++batch) {
            uint32_t t1[KB], t2[KB];
            uint32_t m1 = 0, m2 = 0;
            [[maybe_unused]] uint32_t hi_bad = 0;
            [[maybe_unused]] double d1d[KB], d2d[KB];
            [[maybe_unused]] float2 d1p[KB / 2 + 1], d2p[KB / 2 + 1];
            if (EXACT) {
#pragma unroll
                for (int e = 0; e < KB; ++e) {
                    double row, col;
                    rect_coord_nobranch(pe, rtd, q2p[e], row, col);
                    uint32_t h1, h2;
                    floor_index<kFloorMode1>(row, Mk1, t1[e], h1, d1d[e]);
                    floor_index<kFloorMode2>(col, Mk2, t2[e], h2, d2d[e]);
                    hi_bad |= (h1 ^ 0x43300000u) | (h2 ^ 0x43300000u);
                    m1 = max(m1, t1[e]);
                    m2 = max(m2, t2[e]);
                }
                q2p += KB;
            } else {
#pragma unroll
                for (int hh = 0; hh < KB / 2; ++hh) {
                    float2 row, col;
                    rect_coord2(pf, rtf, ip, row, col);
                    ip = add2(ip, bc2(2.0f));
                    floor_bits_fast2(row, mk1, t1[2 * hh], t1[2 * hh + 1], d1p[hh]);
                    floor_bits_fast2(col, mk2, t2[2 * hh], t2[2 * hh + 1], d2p[hh]);
                }
#pragma unroll
                for (int e = 0; e < KB; ++e) {
                    m1 = max(m1, t1[e] - (uint32_t)kMagicBits);
                    m2 = max(m2, t2[e] - (uint32_t)kMagicBits);
                }
            }
            bool ok = (m1 < R1) & (m2 < R2);
            if (EXACT) ok &= hi_bad == 0u;
            if (__all_sync(0xffffffffu, ok)) {
                float a00[KB], a10[KB], a01[KB], a11[KB];
#pragma unroll
                for (int e = 0; e < KB; ++e) {
                    const uint32_t q = base + t2[e] * box_pitch_b + t1[e] * 4u;
                    const uint32_t q1 = q + box_pitch_b;
#ifdef CAMCAL_CHECK_BOUNDS      // debug builds (profiles/mkvariant.sh chk "-DCAMCAL_CHECK_BOUNDS"): taps inside the stage
                    {
                        const uint32_t lo = stage0 + (uint32_t)s * (uint32_t)cfg.box_bytes;
                        if (q < lo || q1 + 8u > lo + (uint32_t)cfg.box_bytes || (q & 3u)) __trap();
                    }
#endif
                    a00[e] = lds_f32(q); a10[e] = lds_f32_off<4>(q);
                    a01[e] = lds_f32(q1); a11[e] = lds_f32_off<4>(q1);
                }
                float* o = optr;
                if (EXACT) {
#pragma unroll
                    for (int e = 0; e < KB; ++e) {
                        const float v = (float)bilerp((double)a00[e], (double)a10[e], (double)a01[e],
                                                      (double)a11[e], d1d[e], d2d[e]);
                        __stcs(o, v);
                        o += pitch;
                    }
                } else {
#pragma unroll
                    for (int hh = 0; hh < KB / 2; ++hh) {
                        const float2 v = bilerp_fast2(make_float2(a00[2 * hh], a00[2 * hh + 1]),
                                                      make_float2(a10[2 * hh], a10[2 * hh + 1]),
                                                      make_float2(a01[2 * hh], a01[2 * hh + 1]),
                                                      make_float2(a11[2 * hh], a11[2 * hh + 1]),
                                                      d1p[hh], d2p[hh]);
                        __stcs(o, v.x); __stcs(o + pitch, v.y);
                        o += 2 * pitch;
                    }
                }
            } else {
                generic_lines_f32<EXACT>(&pe, &pf, &g, sframe, optr, a, b0 + batch * KB, KB, fill);
            }
            optr += (long long)KB * pitch;
        }
        __syncwarp();
        if (lane_id == 0) mbar_arrive(&ring.empty[s]);
        if (++s == cfg.stages) { s = 0; phase ^= 1; }
    }
}

}  // namespace cc
