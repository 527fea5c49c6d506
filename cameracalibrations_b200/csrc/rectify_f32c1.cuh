// rectify_f32c1.cuh -- fp32 single-channel rectification kernels (included by rectify.cu).
//
// Staged kernel: persistent CTAs, ticket scheduler and TMA producer in rectify_ring.cuh; the
// consumers build a tile's map once per group of frames and reuse it (see the kernel's header).
// Details that keep the map build short:
//  * the host plan folds the tile's index origin into the floor constant: the exact path adds
//    Mk = 2^52 - K to the coordinate with round-down (DADD.RM is the floor) or to
//    FRND.F64.FLOOR(x), the fast path adds mk = 1.5*2^23 - K with FADD2.RM; the low word of the
//    sum is (magic +) the tile-local tap index, "high word == 0x43300000" and "index < R" are the
//    complete range tests (negative, NaN, huge and out-of-box coordinates all fail them);
//  * 1/P3 is NVIDIA's own correctly rounded sequence (MUFU.RCP64H + 5 DFMA) inlined without its
//    per-pixel branch: its validity test (exponent of P3 not extreme) is made once per tile by
//    the plan on the tile corners -- P3 is affine in the output index.
// Direct kernel: no staging, every pixel through sample_direct_f32 (layouts TMA cannot address).
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

// output store flavour (tuning knob): 0 st.global.cs, 1 .wt, 2 .cg, 3 default (.wb), 4 .cs + L2 evict_first hint
#ifndef CAMCAL_F32_STORE
#define CAMCAL_F32_STORE 0
#endif
__device__ __forceinline__ void st_out(float* p, float v) {
    if (CAMCAL_F32_STORE == 1) __stwt(p, v);
    else if (CAMCAL_F32_STORE == 2) __stcg(p, v);
    else if (CAMCAL_F32_STORE == 3) *p = v;
    else if (CAMCAL_F32_STORE == 4) {
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" :: "l"(p), "f"(v), "l"(pol) : "memory");
    } else __stcs(p, v);
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

// CAMCAL_F32_PAIR 1: two frames per ring stage.  The consumers pay the hand-over (barrier wait, stage
// addresses, arrive) once per PAIR of frames, and the exact blend computes 1 - d1, 1 - d2 once for both
// (7 instead of 8 FP64 instructions per pixel and frame).
#ifndef CAMCAL_F32_PAIR
#define CAMCAL_F32_PAIR 1
#endif
constexpr int kF32FramesPerStage = CAMCAL_F32_PAIR != 0 ? 2 : 1;
static_assert(kF32FramesPerStage == 1 || kPosTrack, "pairs of frames need the frame counters (CAMCAL_POS_TRACK)");

#ifndef CAMCAL_F32_BORDER
#define CAMCAL_F32_BORDER 1
#endif
constexpr bool kBorderUnrolled = CAMCAL_F32_BORDER != 0;

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
// The map of a tile -- tap address, weights, and whether the pixel is staged / fill / generic --
// depends on the calibration only.  A unit is one tile of a GROUP of frames: the consumers build
// the map once (pos.w set on the group's first frame), keep it in registers (8 pixels per lane),
// and for every further frame only gather, blend and store.  Per frame and pixel that is 4 LDS,
// the blend and one store; the coordinate chain (23 of the 31 FP64 instructions of the exact
// variant) is paid once per group.
//
// Pixel classes (bit e of the lane's masks, e = line of the warp's slab):
//   staged  : all four taps inside the staged box and the frame        -> LDS gather
//   fill    : the coordinate is outside the frame ([1, n] exact, [1, n) fast) -> fill value
//   skip    : the output pixel itself is outside the frame (partial strip / last tile)
//   neither : generic path (per-pixel checks, direct global taps): x == n exactly, or a
//             footprint that leaves the box.  Results never depend on the class.
// 1: keep e1 = 1 - d1, e2 = 1 - d2 with the map (two DADDs fewer per pixel and frame, 163 instead of
// 177 instructions per 8 pixels) -- measured SLOWER (0.233 vs 0.221 ms on c2): 127 registers instead
// of 96 drop the kernel from 4 to 3 resident CTAs per SM.  Kept as a knob.
#ifndef CAMCAL_F32_KEEP_E
#define CAMCAL_F32_KEEP_E 0
#endif
constexpr bool kKeepE = CAMCAL_F32_KEEP_E != 0;

// Widening the four taps float -> double is what bounds the exact kernel: F2F.F64.F32 runs on the
// XU pipe (4 lanes per clock and sub-partition; ncu: 75 % busy, profiles/r2).  For a finite float
// f >= +0 with bit pattern b, the 64-bit integer b * 2^29 read as a double is f * 2^-896 EXACTLY
// (exponent field e instead of e + 896; zero stays zero; a float denormal becomes the double
// denormal with the same significand) -- one IMAD.WIDE.U32 on the idle FMA pipe.  The scale is
// undone for free: the map keeps d2 * 2^896 instead of d2, so e2' = 2^896 - d2' and the two inner
// products e2'*a', d2'*a' are the unscaled products rounded once -- bit for bit the oracle's
// blend (rounding commutes with a power-of-two scale; nothing under- or overflows: e2*a >= 2^-202).
// Negative, -0, Inf and NaN taps do not fit the rule: the unsigned maximum of a warp's tap bits
// tells (>= 0x7f800000), and such a warp takes the F2F blend for that frame.
#ifndef CAMCAL_F32_WIDEN
#define CAMCAL_F32_WIDEN 0
#endif
constexpr bool kWiden = CAMCAL_F32_WIDEN != 0;
constexpr double kWidenScale = 0x1p896, kWidenUnscale = 0x1p-896;

__device__ __forceinline__ double widen_scaled(float a, uint32_t mul) {
    unsigned long long r;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(__float_as_uint(a)), "r"(mul));
    return __longlong_as_double((long long)r);
}
__device__ __forceinline__ uint32_t umax3(uint32_t a, uint32_t b, uint32_t c) { return max(max(a, b), c); }

// bilerp() on taps scaled by 2^-896 with the second-axis weight scaled by 2^896
__device__ __forceinline__ double bilerp_scaled(double a00, double a10, double a01, double a11, double d1,
                                                double d2s) {
    const double e1 = 1.0 - d1, e2s = kWidenScale - d2s;
    const double lo = fma(d2s, a01, e2s * a00);
    const double hi = fma(d2s, a11, e2s * a10);
    return fma(d1, hi, e1 * lo);
}

// CAMCAL_F32_MAXNREG_FAST > 0: cap the FP32-coordinate variant with __maxnreg__ (5 CTAs of 5 warps put 7
// warps on one scheduler: 7 * 32 * 72 registers is what its 16 K registers hold, 80 is not)
#ifndef CAMCAL_F32_MAXNREG_FAST
#define CAMCAL_F32_MAXNREG_FAST 0
#endif
// the kernel body; NF = frames per ring stage (the two __global__ wrappers are below)
// VIEWS: frames with different views in one launch (cc_rectify_f32c1_views) -- coordinate parameters and axes of
// the unit's view come from the view table in the kernel's parameter space (vt; index = TileHdr.view)
template <bool EXACT, int NF, bool VIEWS = false>
__device__ __forceinline__ void rectify_f32c1_body(const CUtensorMap& tmap, const RectExact& pe0, const RectFast& pf0,
                                                   const RectGeom& g, const TileCfg& cfg, const TileHdr* __restrict__ plan,
                                                   const double* __restrict__ q2tab, RectSched* __restrict__ sched,
                                                   const float* __restrict__ src, float* __restrict__ dst, float fill,
                                                   const ViewTable* vt = nullptr) {
    constexpr int TL = kTLf;                      // lines per tile
    constexpr int LPW = TL / kWarps;              // lines per warp per tile = pixels per lane
    static_assert(LPW % 2 == 0 && LPW <= 16, "pairs of lines; masks are 16 bits");
    extern __shared__ __align__(128) uint8_t stage_mem[];
    __shared__ SmemRing ring;
    const int lane_id = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform for the compiler
    ring_init(&ring, cfg.stages);

    if (warp == kWarps) {                              // ---- producer warp
        producer_loop<EXACT, TL, 1, NF, VIEWS>(&tmap, g, cfg, plan, q2tab, sched, &ring, stage_mem, lane_id);
        return;
    }

    // ---- consumer warps
    const unsigned pitch = (unsigned)g.pitch;
    const uint32_t box_pitch_b = (uint32_t)cfg.pitch_b;
    const uint32_t stage0 = smem_u32(stage_mem);

    // the map of the current unit
    uint32_t rel[LPW];                                 // tap (0,0) byte offset inside a stage
    [[maybe_unused]] double wd1[LPW], wd2[LPW];        // exact weights
    [[maybe_unused]] double we1[LPW], we2[LPW];        // kKeepE: 1 - wd1, 1 - wd2 (the oracle's e1, e2) kept instead of two DADDs per pixel and frame
    [[maybe_unused]] float2 wf1[LPW / 2], wf2[LPW / 2];   // fast weights, pairs of lines
    uint32_t m_staged = 0, m_fill = 0, m_skip = 0;
    bool all_staged = false;                           // warp-uniform: every pixel of every lane staged
    int a = 0, b0 = 0;
    long long off0 = 0;

    int s = 0;
    uint32_t phase = 0;
    int4 pos = make_int4(0, 0, 0, 0);
    int frames_left = 0, frame_z = 0;
    [[maybe_unused]] uint32_t view = 0;
    float* o_cur = dst;
    for (;;) {
        mbar_wait(&ring.full[s], phase);
        // CAMCAL_POS_TRACK: the slot is read on a unit's first frame only; frame index and frames left
        // are carried in warp-uniform registers (declared uniform through a shuffle) -- no LDS + two
        // dependent branches in front of every frame
        if (!kPosTrack || frames_left == 0) {
            pos = ring.pos[s];
            if (pos.z < 0) break;
            if (kPosTrack) {
                frame_z = __shfl_sync(0xffffffffu, pos.z, 0);
                frames_left = __shfl_sync(0xffffffffu, pos.w, 0);
            } else {
                frame_z = pos.z;
            }
        } else {
            pos.w = 0;
        }
        if (VIEWS && pos.w) view = __shfl_sync(0xffffffffu, ring.hdr[s].view, 0);
        const RectExact& pe = VIEWS ? vt->v[view].pe : pe0;
        const RectFast& pf = VIEWS ? vt->v[view].pf : pf0;
        const RectGeom& gv = VIEWS ? vt->v[view].g : g;     // axs0 / axs1 of the view (everything else is the launch's)
        if (pos.w) {                                   // ---- first frame of a unit: build the map
            const TileHdr* h = &ring.hdr[s];
            const int a_w = pos.x * kT;
            a = a_w + lane_id;
            b0 = pos.y * TL + warp * LPW;
            off0 = (long long)b0 * g.pitch + a;
            const int a_c = min(a, g.sz1 - 1);         // out-of-frame lanes shadow the last pixel
            const uint32_t R1 = h->R1, R2 = h->R2;
            const uint32_t rel0 = h->base_off;
            m_staged = m_fill = m_skip = 0;
            if ((!EXACT && !VIEWS) || a_w + kT > g.sz1 || b0 + LPW > g.sz2) {   // partial tile (warp-uniform test; measured: pays only in the exact kernel and with a map per frame)
#pragma unroll
                for (int e = 0; e < LPW; ++e)
                    if (a >= g.sz1 || b0 + e >= g.sz2) m_skip |= 1u << e;
            }
            if (EXACT) {
                const RowTermD rtd = rect_row_term(pe, gv.axs0 + a_c);
                const double Mk1 = h->Mk1, Mk2 = h->Mk2;
                const double* q2p = &ring.q2[s][warp * LPW];
#pragma unroll
                for (int e = 0; e < LPW; ++e) {
                    double row, col;
                    rect_coord_nobranch(pe, rtd, q2p[e], row, col);
                    uint32_t t1, t2, h1, h2;
                    floor_index<kFloorMode1>(row, Mk1, t1, h1, wd1[e]);
                    floor_index<kFloorMode2>(col, Mk2, t2, h2, wd2[e]);
                    if (kKeepE) { we1[e] = 1.0 - wd1[e]; we2[e] = 1.0 - wd2[e]; }
                    if (kWiden) wd2[e] *= kWidenScale;     // exact
                    const bool st = (((h1 ^ 0x43300000u) | (h2 ^ 0x43300000u)) == 0u) & (t1 < R1) & (t2 < R2);
                    rel[e] = rel0 + t2 * box_pitch_b + t1 * 4u;
                    if (st) m_staged |= 1u << e;
                    else if (!(lin_ok(row, g.sz1) & lin_ok(col, g.sz2))) m_fill |= 1u << e;   // rare: border tiles
                }
            } else {
                const RowTermF rtf = rect_row_term(pf, gv.axs0 + a_c);
                const float mk1 = h->mk1, mk2 = h->mk2;
                float2 ip;
                ip.x = (float)(gv.axs1 + b0) - pf.c2;
                ip.y = ip.x + 1.0f;
#pragma unroll
                for (int hh = 0; hh < LPW / 2; ++hh) {
                    float2 row, col;
                    rect_coord2(pf, rtf, ip, row, col);
                    ip = add2(ip, bc2(2.0f));
                    uint32_t t1[2], t2[2];
                    floor_bits_fast2(row, mk1, t1[0], t1[1], wf1[hh]);
                    floor_bits_fast2(col, mk2, t2[0], t2[1], wf2[hh]);
                    const float rr[2] = {row.x, row.y}, cc_[2] = {col.x, col.y};
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int e = 2 * hh + j;
                        const uint32_t l1 = t1[j] - (uint32_t)kMagicBits, l2 = t2[j] - (uint32_t)kMagicBits;
                        const bool st = (l1 < R1) & (l2 < R2);
                        rel[e] = rel0 + l2 * box_pitch_b + l1 * 4u;
                        if (st) m_staged |= 1u << e;
                        if (!VIEWS) {
                            // (unconditional: four FSETPs cost less than a divergent branch here; measured)
                            const bool inframe = (rr[j] >= 1.0f) & (rr[j] < (float)g.sz1) & (cc_[j] >= 1.0f) & (cc_[j] < (float)g.sz2);
                            if (!inframe) m_fill |= 1u << e;
                        }
                    }
                }
                // VIEWS: the map is built per frame, so the fill class (nine instructions per pixel) is worked out
                // for the warps that need it only -- border tiles, 4 % -- from the same coordinates once more
                if (VIEWS && !__all_sync(0xffffffffu, m_staged == (1u << LPW) - 1u)) {
                    ip.x = (float)(gv.axs1 + b0) - pf.c2;
                    ip.y = ip.x + 1.0f;
#pragma unroll
                    for (int hh = 0; hh < LPW / 2; ++hh) {
                        float2 row, col;
                        rect_coord2(pf, rtf, ip, row, col);
                        ip = add2(ip, bc2(2.0f));
                        if (!((row.x >= 1.0f) & (row.x < (float)g.sz1) & (col.x >= 1.0f) & (col.x < (float)g.sz2))) m_fill |= 1u << (2 * hh);
                        if (!((row.y >= 1.0f) & (row.y < (float)g.sz1) & (col.y >= 1.0f) & (col.y < (float)g.sz2))) m_fill |= 1u << (2 * hh + 1);
                    }
                }
            }
            // a tile the plan marked unusable (P3 not sane, empty box): everything generic
            if (R1 == 0u || R2 == 0u) { m_staged = 0; m_fill = 0; }
            constexpr uint32_t kAll = (1u << LPW) - 1u;
            all_staged = __all_sync(0xffffffffu, (m_staged == kAll) & (m_skip == 0u));
        }

        // ---- every frame of the unit: gather, blend, store
        const float* sframe = src + (long long)frame_z * g.frame_stride;
        // NF == 2: this stage holds frames frame_z and frame_z + 1 (only frame_z at the odd end of a unit)
        const bool two = NF == 2 && frames_left > 1;
        constexpr bool kPairBlock = NF == 2 && !(EXACT && (kWiden || kKeepE));   // the two-frame hot block exists
        // kPosTrack: a running output pointer (set on the unit's first frame, one 64-bit add per frame)
        if (!kPosTrack || pos.w) o_cur = dst + (long long)frame_z * g.frame_stride + off0;
        float* o = o_cur;
        uint32_t sbase = stage0 + (uint32_t)(s * NF) * (uint32_t)cfg.box_bytes;
        uint32_t sbase1 = sbase + box_pitch_b;    // one uniform base per source line: LDS [R + UR + imm], no per-pixel add
        auto border = [&](float* o) {
            // border tiles (and frames whose taps the integer widening cannot take): per-pixel class.
            // CAMCAL_F32_BORDER 1: staged and fill pixels unrolled (static indices into the map), then
            // one rolled pass over what is left for the generic sampler (it needs no map entry);
            // 0: round 1's rolled loop that selects the map entry through a compare chain -- 98
            // instructions per pixel, 9 % of all executed instructions of c2 (profiles/r2_rectify.md)
            if (kBorderUnrolled) {
                float* ob = o;
#pragma unroll
                for (int e = 0; e < LPW; ++e, ob += pitch) {
                    const uint32_t bit = 1u << e;
                    if (m_skip & bit) continue;              // the output pixel itself is outside the frame
                    if (m_staged & bit) {
                        const uint32_t q = sbase + rel[e], q1 = sbase1 + rel[e];
                        const float t00 = lds_f32(q), t10 = lds_f32_off<4>(q), t01 = lds_f32(q1), t11 = lds_f32_off<4>(q1);
                        float v;
                        if (EXACT) v = (float)bilerp((double)t00, (double)t10, (double)t01, (double)t11, wd1[e],
                                                     kWiden ? wd2[e] * kWidenUnscale : wd2[e]);
                        else v = bilerp_fast(t00, t10, t01, t11, (e & 1) ? wf1[e / 2].y : wf1[e / 2].x,
                                             (e & 1) ? wf2[e / 2].y : wf2[e / 2].x);
                        __stcs(ob, v);
                    } else if (m_fill & bit) {
                        __stcs(ob, fill);
                    }
                }
                // what is left: x == n exactly, footprints that leave the box (needs no map entry: rolled)
                const uint32_t m_gen = ~(m_staged | m_fill | m_skip) & ((1u << LPW) - 1u);
                if (m_gen) {
                    RowTermD rtd;
                    RowTermF rtf;
                    if (EXACT) rtd = rect_row_term(pe, gv.axs0 + a); else rtf = rect_row_term(pf, gv.axs0 + a);
#pragma unroll 1
                    for (int e = 0; e < LPW; ++e)
                        if ((m_gen >> e) & 1u)
                            __stcs(o + (long long)e * pitch, sample_direct_f32<EXACT>(pe, pf, rtd, rtf, gv, sframe, pitch, b0 + e, fill));
                }
            } else {
#pragma unroll 1
            for (int e = 0; e < LPW; ++e, o += pitch) {
                if ((m_skip >> e) & 1u) continue;
                float v;
                if ((m_staged >> e) & 1u) {
                    // (rel[], weights indexed dynamically would spill: select through a switch-free copy)
                    uint32_t r = 0;
                    [[maybe_unused]] double d1 = 0, d2 = 0;
                    [[maybe_unused]] float f1 = 0, f2 = 0;
#pragma unroll
                    for (int j = 0; j < LPW; ++j)
                        if (j == e) {
                            r = rel[j];
                            if (EXACT) { d1 = wd1[j]; d2 = kWiden ? wd2[j] * kWidenUnscale : wd2[j]; }
                            else { f1 = (j & 1) ? wf1[j / 2].y : wf1[j / 2].x; f2 = (j & 1) ? wf2[j / 2].y : wf2[j / 2].x; }
                        }
                    const uint32_t q = sbase + r, q1 = q + box_pitch_b;
                    const float t00 = lds_f32(q), t10 = lds_f32_off<4>(q), t01 = lds_f32(q1), t11 = lds_f32_off<4>(q1);
                    v = EXACT ? (float)bilerp((double)t00, (double)t10, (double)t01, (double)t11, d1, d2)
                              : bilerp_fast(t00, t10, t01, t11, f1, f2);
                } else if ((m_fill >> e) & 1u) {
                    v = fill;
                } else {
                    RowTermD rtd;
                    RowTermF rtf;
                    if (EXACT) rtd = rect_row_term(pe, gv.axs0 + a); else rtf = rect_row_term(pf, gv.axs0 + a);
                    v = sample_direct_f32<EXACT>(pe, pf, rtd, rtf, gv, sframe, pitch, b0 + e, fill);
                }
                __stcs(o, v);
            }
            }
        };
        if (kPairBlock && all_staged && two) {
            // two frames with one map: taps of frame A at sbase, of frame B one box further
            const uint32_t tbase = sbase + (uint32_t)cfg.box_bytes, tbase1 = sbase1 + (uint32_t)cfg.box_bytes;
            float* oA = o;
            float* oB = o + g.frame_stride;
            if (EXACT) {
#pragma unroll
                for (int e = 0; e < LPW; ++e) {
                    const uint32_t r = rel[e];
#ifdef CAMCAL_CHECK_BOUNDS
                    if (r + box_pitch_b + 8u > (uint32_t)cfg.box_bytes || (r & 3u)) __trap();
#endif
                    const float a00 = lds_f32(sbase + r), a10 = lds_f32_off<4>(sbase + r);
                    const float a01 = lds_f32(sbase1 + r), a11 = lds_f32_off<4>(sbase1 + r);
                    const float c00 = lds_f32(tbase + r), c10 = lds_f32_off<4>(tbase + r);
                    const float c01 = lds_f32(tbase1 + r), c11 = lds_f32_off<4>(tbase1 + r);
                    const double d1 = wd1[e], d2 = wd2[e], e1 = 1.0 - d1, e2 = 1.0 - d2;
                    st_out(oA, (float)bilerp_e((double)a00, (double)a10, (double)a01, (double)a11, d1, e1, d2, e2));
                    st_out(oB, (float)bilerp_e((double)c00, (double)c10, (double)c01, (double)c11, d1, e1, d2, e2));
                    oA += pitch; oB += pitch;
                }
            } else {
#pragma unroll
                for (int hh = 0; hh < LPW / 2; ++hh) {
                    const uint32_t r0 = rel[2 * hh], r1 = rel[2 * hh + 1];
                    const float2 vA = bilerp_fast2(make_float2(lds_f32(sbase + r0), lds_f32(sbase + r1)),
                                                   make_float2(lds_f32_off<4>(sbase + r0), lds_f32_off<4>(sbase + r1)),
                                                   make_float2(lds_f32(sbase1 + r0), lds_f32(sbase1 + r1)),
                                                   make_float2(lds_f32_off<4>(sbase1 + r0), lds_f32_off<4>(sbase1 + r1)), wf1[hh], wf2[hh]);
                    const float2 vB = bilerp_fast2(make_float2(lds_f32(tbase + r0), lds_f32(tbase + r1)),
                                                   make_float2(lds_f32_off<4>(tbase + r0), lds_f32_off<4>(tbase + r1)),
                                                   make_float2(lds_f32(tbase1 + r0), lds_f32(tbase1 + r1)),
                                                   make_float2(lds_f32_off<4>(tbase1 + r0), lds_f32_off<4>(tbase1 + r1)), wf1[hh], wf2[hh]);
                    st_out(oA, vA.x); st_out(oA + pitch, vA.y);
                    st_out(oB, vB.x); st_out(oB + pitch, vB.y);
                    oA += 2 * pitch; oB += 2 * pitch;
                }
            }
        } else if (all_staged) {
            float a00[LPW], a10[LPW], a01[LPW], a11[LPW];
#pragma unroll
            for (int e = 0; e < LPW; ++e) {
                const uint32_t q = sbase + rel[e];
                const uint32_t q1 = sbase1 + rel[e];
#ifdef CAMCAL_CHECK_BOUNDS      // debug builds (profiles/mkvariant.sh chk "-DCAMCAL_CHECK_BOUNDS"): taps inside the stage
                if (q < sbase || q1 + 8u > sbase + (uint32_t)cfg.box_bytes || (q & 3u)) __trap();
#endif
                a00[e] = lds_f32(q); a10[e] = lds_f32_off<4>(q);
                a01[e] = lds_f32(q1); a11[e] = lds_f32_off<4>(q1);
            }
            if (EXACT && kWiden) {
                // blend as if every tap fits the integer widening, keep the maximum of the tap bits on
                // the side; a warp that saw a negative / Inf / NaN tap redoes the frame below (F2F blend)
                const uint32_t mul = cfg.widen_mul;
                uint32_t mx = 0;
                float* ow = o;
#pragma unroll
                for (int e = 0; e < LPW; ++e) {
                    if (CAMCAL_F32_WIDEN != 2) {       // 2: tuning build without the check (wrong for negative taps)
                        mx = umax3(mx, __float_as_uint(a00[e]), __float_as_uint(a10[e]));
                        mx = umax3(mx, __float_as_uint(a01[e]), __float_as_uint(a11[e]));
                    }
                    __stcs(ow, (float)bilerp_scaled(widen_scaled(a00[e], mul), widen_scaled(a10[e], mul),
                                                    widen_scaled(a01[e], mul), widen_scaled(a11[e], mul), wd1[e], wd2[e]));
                    ow += pitch;
                }
                if (__any_sync(0xffffffffu, mx >= 0x7f800000u)) border(o);
            } else if (EXACT) {
#pragma unroll
                for (int e = 0; e < LPW; ++e) {
                    if (kKeepE) {
                        // opaque to the compiler: otherwise it rematerialises 1 - d per frame to save registers
                        asm volatile("" : "+d"(we1[e]), "+d"(we2[e]));
                        __stcs(o, (float)bilerp_e((double)a00[e], (double)a10[e], (double)a01[e], (double)a11[e],
                                                  wd1[e], we1[e], wd2[e], we2[e]));
                    } else {
                        st_out(o, (float)bilerp((double)a00[e], (double)a10[e], (double)a01[e], (double)a11[e],
                                                wd1[e], wd2[e]));
                    }
                    o += pitch;
                }
            } else {
#pragma unroll
                for (int hh = 0; hh < LPW / 2; ++hh) {
                    const float2 v = bilerp_fast2(make_float2(a00[2 * hh], a00[2 * hh + 1]),
                                                  make_float2(a10[2 * hh], a10[2 * hh + 1]),
                                                  make_float2(a01[2 * hh], a01[2 * hh + 1]),
                                                  make_float2(a11[2 * hh], a11[2 * hh + 1]), wf1[hh], wf2[hh]);
                    st_out(o, v.x); st_out(o + pitch, v.y);
                    o += 2 * pitch;
                }
            }
        } else {
            border(o);
        }
        if (two && !(kPairBlock && all_staged)) {
            // second frame of the stage through the per-frame path (border tiles; the knob variants)
            sbase += (uint32_t)cfg.box_bytes; sbase1 += (uint32_t)cfg.box_bytes;
            sframe += g.frame_stride;
            border(o_cur + g.frame_stride);
        }
        __syncwarp();
        if (kElectArrive ? elect_one() : lane_id == 0) mbar_arrive(&ring.empty[s]);
        if (kPosTrack) { frames_left = max(frames_left - NF, 0); frame_z += NF; o_cur += NF * g.frame_stride; }
        if (++s == cfg.stages) { s = 0; phase ^= 1; }
    }
}

#if CAMCAL_F32_MAXNREG_FAST > 0
#define CAMCAL_F32_KERNEL_ATTR(EXACT) __launch_bounds__(kConsumerThreads + 32) __maxnreg__(EXACT ? 65536 / ((kConsumerThreads + 32) * kMinBlocksExact) / 8 * 8 : CAMCAL_F32_MAXNREG_FAST)
#else
#define CAMCAL_F32_KERNEL_ATTR(EXACT) __launch_bounds__(kConsumerThreads + 32, EXACT ? kMinBlocksExact : kMinBlocks)
#endif
// two frames per ring stage (the default: rings of at least four boxes)
template <bool EXACT>
__global__ void CAMCAL_F32_KERNEL_ATTR(EXACT)
rectify_f32c1_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ RectExact pe,
                     const __grid_constant__ RectFast pf, const __grid_constant__ RectGeom g,
                     const __grid_constant__ TileCfg cfg, const TileHdr* __restrict__ plan,
                     const double* __restrict__ q2tab, RectSched* __restrict__ sched,
                     const float* __restrict__ src, float* __restrict__ dst, float fill) {
    rectify_f32c1_body<EXACT, kF32FramesPerStage>(tmap, pe, pf, g, cfg, plan, q2tab, sched, src, dst, fill);
}
// one frame per stage: large boxes (strongly tilted or magnifying views), whose ring holds two or three
template <bool EXACT>
__global__ void CAMCAL_F32_KERNEL_ATTR(EXACT)
rectify_f32c1_single_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ RectExact pe,
                            const __grid_constant__ RectFast pf, const __grid_constant__ RectGeom g,
                            const __grid_constant__ TileCfg cfg, const TileHdr* __restrict__ plan,
                            const double* __restrict__ q2tab, RectSched* __restrict__ sched,
                            const float* __restrict__ src, float* __restrict__ dst, float fill) {
    rectify_f32c1_body<EXACT, 1>(tmap, pe, pf, g, cfg, plan, q2tab, sched, src, dst, fill);
}

// frames with different views in one launch, one frame per stage (cc_rectify_f32c1_views)
template <bool EXACT>
__global__ void CAMCAL_F32_KERNEL_ATTR(EXACT)
rectify_f32c1_views_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ ViewTable vt,
                           const __grid_constant__ RectGeom g, const __grid_constant__ TileCfg cfg,
                           const TileHdr* __restrict__ plan, const double* __restrict__ q2tab,
                           RectSched* __restrict__ sched, const float* __restrict__ src, float* __restrict__ dst,
                           float fill) {
    rectify_f32c1_body<EXACT, 1, true>(tmap, vt.v[0].pe, vt.v[0].pf, g, cfg, plan, q2tab, sched, src, dst, fill, &vt);
}

}  // namespace cc
