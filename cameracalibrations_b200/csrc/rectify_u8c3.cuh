// rectify_u8c3.cuh -- 8-bit RGB (RGB{N0f8}, 3 interleaved bytes) rectification kernels
// (included by rectify.cu).  Same skeleton as rectify_f32c1.cuh: persistent CTAs, ticket
// scheduler, host-built tile plan, one TMA box per tile, raw-bit addressing.
//
// Per pixel the two taps of one source line are 6 contiguous bytes at byte offset 3*i: they are
// fetched as three aligned 32-bit words and funnelled with PRMT.  Bytes become floats without a
// conversion instruction (an isolated byte IS the denormal float b * 2^-149; a packed FMUL2 by
// 2^100 makes it normal and exact), the blend runs on PAIRS of lines in FFMA2/FADD2 in that scaled
// domain, and the final FFMA2 by 2^49 onto a rounding magic undoes the scale.
//
// Exact variant (FP64 coordinates): indices and weights come from the FP64 chain, bit for bit
// like the oracle.  The blend of 8-bit taps is then done in FP32 with a CERTIFIED rounding:
//   |v32 - v64| <= 5.3e-5 LSB  (three roundings at ulp(255)/2 = 7.6e-6 each, accumulated:
//   3.8e-5, plus the weights' FP32 representation error 2 * 255 * 2^-25 = 1.5e-5),
// so whenever v32 is farther than 6.5e-5 from a rounding boundary, rint(v32) == rint(v64).
// The pixels that are closer (1.3e-4 of the values; exact .5 ties included) are re-blended in
// FP64 with the oracle's operation order.  The output is therefore bit-identical to the FP64
// blend while ~99.9 % of the pixels never touch the FP64 pipe after the coordinate chain.
//
// Stores: the 32 pixels of a warp's line are 96 contiguous bytes; two shuffles and one PRMT per
// line turn the per-lane 0x00BBGGRR into 24 aligned words (full 32-byte sectors).
#pragma once

namespace cc {

struct Taps6 { uint32_t lo, hi; };   // bytes [o, o+4) and [o+4, o+8) of a byte stream

template <typename LD>
__device__ __forceinline__ Taps6 load6(const uint32_t* words, unsigned o, unsigned sel, LD ld) {
    const uint32_t* w = words + (o >> 2);
    const uint32_t w0 = ld(w), w1 = ld(w + 1), w2 = ld(w + 2);
    Taps6 t;
    t.lo = __byte_perm(w0, w1, sel);
    t.hi = __byte_perm(w1, w2, sel);
    return t;
}
__device__ __forceinline__ unsigned sel6(unsigned o) { return 0x3210u + 0x1111u * (o & 3u); }

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
template <int OFF>
__device__ __forceinline__ uint32_t lds_u32_off(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF));
    return v;
}
// six bytes out of the three words at word-aligned shared address wa
__device__ __forceinline__ Taps6 lds6w(uint32_t wa, unsigned sel) {
    const uint32_t w0 = lds_u32(wa), w1 = lds_u32_off<4>(wa), w2 = lds_u32_off<8>(wa);
    Taps6 t;
    t.lo = __byte_perm(w0, w1, sel);
    t.hi = __byte_perm(w1, w2, sel);
    return t;
}
// byte k of w as a float without a conversion instruction: 0x4B000000 | b  ==  2^23 + b
__device__ __forceinline__ float byte_f(uint32_t w, int k) {
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440u + (unsigned)k)) - 8388608.0f;
}
__device__ __forceinline__ double byte_d(uint32_t w, int k) { return (double)((w >> (8 * k)) & 0xffu); }

// t0 = source line i2, t1 = line i2+1;  t.lo = [a00.r a00.g a00.b a10.r], t.hi = [a10.g a10.b . .]
template <bool EXACT>
__device__ __forceinline__ uint32_t blend_rgb(const Taps6& t0, const Taps6& t1, double d1d, double d2d,
                                              float d1f, float d2f) {
    uint32_t r, g, b;
    if (EXACT) {
        r = (uint32_t)(int)rint(bilerp(byte_d(t0.lo, 0), byte_d(t0.lo, 3), byte_d(t1.lo, 0), byte_d(t1.lo, 3), d1d, d2d));
        g = (uint32_t)(int)rint(bilerp(byte_d(t0.lo, 1), byte_d(t0.hi, 0), byte_d(t1.lo, 1), byte_d(t1.hi, 0), d1d, d2d));
        b = (uint32_t)(int)rint(bilerp(byte_d(t0.lo, 2), byte_d(t0.hi, 1), byte_d(t1.lo, 2), byte_d(t1.hi, 1), d1d, d2d));
    } else {
        // round-to-nearest by magic add; weights in [0,1] keep the result inside [0,255]
        const float m = 12582912.0f;
        const float fr = bilerp_fast(byte_f(t0.lo, 0), byte_f(t0.lo, 3), byte_f(t1.lo, 0), byte_f(t1.lo, 3), d1f, d2f) + m;
        const float fg = bilerp_fast(byte_f(t0.lo, 1), byte_f(t0.hi, 0), byte_f(t1.lo, 1), byte_f(t1.hi, 0), d1f, d2f) + m;
        const float fb = bilerp_fast(byte_f(t0.lo, 2), byte_f(t0.hi, 1), byte_f(t1.lo, 2), byte_f(t1.hi, 1), d1f, d2f) + m;
        r = __float_as_uint(fr); g = __float_as_uint(fg); b = __float_as_uint(fb);
    }
    // low bytes of r, g, b -> 0x00BBGGRR
    return __byte_perm(__byte_perm(r, g, 0x0040), b, 0x0410);
}

// Two pixels at once (FMUL2/FFMA2).  A byte's bit pattern IS a float already -- the denormal
// b * 2^-149 -- so one packed FMUL2 by 2^100 per pair of bytes makes it the normal float b * 2^-49
// (exact); the blend runs in that scaled domain (all normal numbers, full relative precision) and
// the final FFMA2 by 2^49 onto the rounding magic undoes the scale.
constexpr float kTwo100 = 1.2676506002282294e30f;      // 2^100
constexpr float kTwo49 = 562949953421312.0f;           // 2^49
__device__ __forceinline__ float2 byte_f2(uint32_t wp, uint32_t wq, int k) {
    return mul2(make_float2(__uint_as_float(__byte_perm(wp, 0u, 0x4440u + (unsigned)k)),
                            __uint_as_float(__byte_perm(wq, 0u, 0x4440u + (unsigned)k))), bc2(kTwo100));
}

// distance of a blended value from the integer it rounds to, against the certification bound
constexpr float kCertThr = 0.5f - 6.5e-5f;

// scaled blended values -> packed 0x00BBGGRR per pixel (+ certification of the roundings)
template <bool CERT>
__device__ __forceinline__ void round_pack2(float2 vr, float2 vg, float2 vb, uint32_t& rgb_p, uint32_t& rgb_q,
                                            bool& amb_p, bool& amb_q) {
    const float2 m = bc2(12582912.0f), up = bc2(kTwo49);
    const float2 fr = fma2(vr, up, m), fg = fma2(vg, up, m), fb = fma2(vb, up, m);     // rint(v) in the low byte
    rgb_p = __byte_perm(__byte_perm(__float_as_uint(fr.x), __float_as_uint(fg.x), 0x0040), __float_as_uint(fb.x), 0x0410);
    rgb_q = __byte_perm(__byte_perm(__float_as_uint(fr.y), __float_as_uint(fg.y), 0x0040), __float_as_uint(fb.y), 0x0410);
    if (CERT) {
        // v - rint(v): the scaling by 2^49 is exact, the FMA rounds once (|e| <= 0.5, far above ulp)
        const float2 er = fma2(vr, up, sub2(m, fr)), eg = fma2(vg, up, sub2(m, fg)), eb = fma2(vb, up, sub2(m, fb));
        amb_p = (fabsf(er.x) > kCertThr) | (fabsf(eg.x) > kCertThr) | (fabsf(eb.x) > kCertThr);
        amb_q = (fabsf(er.y) > kCertThr) | (fabsf(eg.y) > kCertThr) | (fabsf(eb.y) > kCertThr);
    }
}

// CERT: also report (per pixel) whether any channel is too close to a rounding boundary
template <bool CERT>
__device__ __forceinline__ void blend_rgb2(const Taps6& p0, const Taps6& p1, const Taps6& q0,
                                           const Taps6& q1, float2 d1, float2 d2, uint32_t& rgb_p,
                                           uint32_t& rgb_q, bool& amb_p, bool& amb_q) {
    const float2 vr = bilerp_fast2(byte_f2(p0.lo, q0.lo, 0), byte_f2(p0.lo, q0.lo, 3),
                                   byte_f2(p1.lo, q1.lo, 0), byte_f2(p1.lo, q1.lo, 3), d1, d2);
    const float2 vg = bilerp_fast2(byte_f2(p0.lo, q0.lo, 1), byte_f2(p0.hi, q0.hi, 0),
                                   byte_f2(p1.lo, q1.lo, 1), byte_f2(p1.hi, q1.hi, 0), d1, d2);
    const float2 vb = bilerp_fast2(byte_f2(p0.lo, q0.lo, 2), byte_f2(p0.hi, q0.hi, 1),
                                   byte_f2(p1.lo, q1.lo, 2), byte_f2(p1.hi, q1.hi, 1), d1, d2);
    round_pack2<CERT>(vr, vg, vb, rgb_p, rgb_q, amb_p, amb_q);
}

__device__ __forceinline__ void store_rgb(uint8_t* q, uint32_t rgb) {
    q[0] = (uint8_t)rgb; q[1] = (uint8_t)(rgb >> 8); q[2] = (uint8_t)(rgb >> 16);
}

// generic per-pixel path with every check and direct global taps
template <bool EXACT>
__device__ __forceinline__ uint32_t sample_direct_u8(const RectExact& pe, const RectFast& pf,
                                                     const RowTermD& rtd, const RowTermF& rtf,
                                                     const RectGeom& g, const uint8_t* __restrict__ sframe,
                                                     unsigned pitch3, unsigned frame_bytes, int b,
                                                     uint32_t fill) {
    int g1, g2;
    double d1d = 0, d2d = 0;
    float d1f = 0, d2f = 0;
    if (EXACT) {
        double row, col;
        rect_coord(pe, rtd, rect_q2(pe, g.axs1 + b), row, col);
        if (!(lin_ok(row, g.sz1) & lin_ok(col, g.sz2))) return fill;
        lin_floor(row, g1, d1d);
        lin_floor(col, g2, d2d);
        lin_fix_edge(g.sz1, g1, d1d);
        lin_fix_edge(g.sz2, g2, d2d);
        g1 -= 1; g2 -= 1;
    } else {
        float row, col;
        int t1, t2;
        rect_coord(pf, rtf, (float)(g.axs1 + b) - pf.c2, row, col);
        lin_floor_fast(row, t1, d1f);
        lin_floor_fast(col, t2, d2f);
        g1 = t1 - (kMagicBits + 1); g2 = t2 - (kMagicBits + 1);
        if (!(((unsigned)g1 <= (unsigned)(g.sz1 - 2)) & ((unsigned)g2 <= (unsigned)(g.sz2 - 2)))) return fill;
    }
    const unsigned off = (unsigned)g2 * pitch3 + (unsigned)g1 * 3u;
    Taps6 t0, t1;
    // word-granular gather from a 4-byte aligned base; the byte path for the last few taps of
    // the frame so nothing outside the caller's buffer is touched
    if (off + pitch3 + 12u <= frame_bytes) {
        const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(sframe) & 3u);
        const uint32_t* gwords = reinterpret_cast<const uint32_t*>(sframe - mis);
        auto ld = [](const uint32_t* p) { return __ldg(p); };
        t0 = load6(gwords, off + mis, sel6(off + mis), ld);
        t1 = load6(gwords, off + pitch3 + mis, sel6(off + pitch3 + mis), ld);
    } else {
        const uint8_t* q = sframe + off;
        t0.lo = q[0] | (q[1] << 8) | (q[2] << 16) | ((uint32_t)q[3] << 24);
        t0.hi = q[4] | (q[5] << 8);
        q += pitch3;
        t1.lo = q[0] | (q[1] << 8) | (q[2] << 16) | ((uint32_t)q[3] << 24);
        t1.hi = q[4] | (q[5] << 8);
    }
    return blend_rgb<EXACT>(t0, t1, d1d, d2d, d1f, d2f);
}

// ---- direct kernel: no staging -------------------------------------------------------------
template <bool EXACT>
__global__ void __launch_bounds__(kConsumerThreads)
rectify_u8c3_direct_kernel(const __grid_constant__ RectExact pe, const __grid_constant__ RectFast pf,
                           const __grid_constant__ RectGeom g, const int lines_per_cta,
                           const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uchar3 fill3,
                           unsigned frame_bytes) {
    const int lane_id = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int a = blockIdx.x * kT + lane_id;
    if (a >= g.sz1) return;
    const int frame = blockIdx.z;
    const uint8_t* sframe = src + (long long)frame * g.frame_stride * 3;
    const unsigned pitch3 = (unsigned)g.pitch * 3u;
    const uint32_t fill = fill3.x | (fill3.y << 8) | (fill3.z << 16);
    RowTermD rtd;
    RowTermF rtf;
    if (EXACT) rtd = rect_row_term(pe, g.axs0 + a); else rtf = rect_row_term(pf, g.axs0 + a);
    const int b_begin = blockIdx.y * lines_per_cta;
    const int b_end = min(b_begin + lines_per_cta, g.sz2);
    uint8_t* o = dst + ((long long)frame * g.frame_stride + (long long)(b_begin + warp) * g.pitch + a) * 3;
    for (int b = b_begin + warp; b < b_end; b += kWarps, o += (long long)kWarps * pitch3)
        store_rgb(o, sample_direct_u8<EXACT>(pe, pf, rtd, rtf, g, sframe, pitch3, frame_bytes, b, fill));
}

// ---- staged kernel (persistent; scheduling and producer: rectify_ring.cuh) ------------------
// Same unit structure as rectify_f32c1_kernel: the map of a tile (tap byte offset, weights, pixel
// class) is built once per group of frames and kept in registers (8 pixels per lane); every
// frame then only gathers, blends and stores.  Both coordinate variants blend in FP32 from FP32
// copies of the weights; the exact variant certifies each rounding and, for the rare pixel that
// fails, recomputes the FP64 weights and blends in FP64 (oracle order).
template <bool EXACT>
__device__ __noinline__ uint32_t reblend_exact_u8(const RectExact* pe, const RectGeom* g, int a, int b,
                                                  const Taps6 t0, const Taps6 t1) {
    const RowTermD rtd = rect_row_term(*pe, g->axs0 + a);
    double row, col, d1, d2;
    int i1, i2;
    rect_coord(*pe, rtd, rect_q2(*pe, g->axs1 + b), row, col);
    lin_floor(row, i1, d1);
    lin_floor(col, i2, d2);
    return blend_rgb<true>(t0, t1, d1, d2, 0.f, 0.f);
}

template <bool EXACT>
__global__ void __launch_bounds__(kConsumerThreads + 32, EXACT ? kMinBlocksU8Exact : kMinBlocksU8)
rectify_u8c3_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ RectExact pe,
                    const __grid_constant__ RectFast pf, const __grid_constant__ RectGeom g,
                    const __grid_constant__ TileCfg cfg, const TileHdr* __restrict__ plan,
                    const double* __restrict__ q2tab, RectSched* __restrict__ sched,
                    const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uchar3 fill3,
                    unsigned frame_bytes) {
    constexpr int KB = 4;                         // pixels blended together (two packed pairs)
    constexpr int TL = kTLu;                      // lines per tile
    constexpr int LPW = TL / kWarps;              // lines per warp per tile = pixels per lane
    static_assert(LPW % KB == 0 && LPW <= 16, "batches of four lines; masks are 16 bits");
    extern __shared__ __align__(128) uint8_t stage_mem[];
    __shared__ SmemRing ring;
    const int lane_id = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform for the compiler
    ring_init(&ring, cfg.stages);

    if (warp == kWarps) {                              // ---- producer warp
        producer_loop<EXACT, TL, 3>(&tmap, g, cfg, plan, q2tab, sched, &ring, stage_mem, lane_id);
        return;
    }

    // ---- consumer warps
    const unsigned pitch3 = (unsigned)g.pitch * 3u;
    const uint32_t box_pitch_b = (uint32_t)cfg.box1 * 3u;      // bytes per box line (multiple of 48)
    const uint32_t stage0 = smem_u32(stage_mem);
    const uint32_t fill = fill3.x | (fill3.y << 8) | (fill3.z << 16);
    // store transpose: output word L (< 24) of a line holds bytes 4L..4L+3 = pixels p0, p0+1
    const int wp0 = min((4 * lane_id) / 3, 31), wo = (4 * lane_id) % 3;
    const unsigned wsel = wo == 0 ? 0x4210u : (wo == 1 ? 0x5421u : 0x6542u);
    const int wp1 = min(wp0 + 1, 31);

    // the map of the current unit
    uint32_t rel[LPW];                                 // first tap's byte offset inside a stage, rounded down to a word
    uint32_t selv[LPW];                                // PRMT selector that funnels the 6 tap bytes out of 3 words
    float2 wf1[LPW / 2], wf2[LPW / 2];                 // weights, pairs of lines (FP32 copies when EXACT)
    uint32_t m_staged = 0, m_fill = 0, m_skip = 0;
    bool all_staged = false;
    int a = 0, b0 = 0;
    long long off0 = 0;

    int s = 0;
    uint32_t phase = 0;
    for (;;) {
        mbar_wait(&ring.full[s], phase);
        const int4 pos = ring.pos[s];
        if (pos.z < 0) break;
        if (pos.w) {                                   // ---- first frame of a unit: build the map
            const TileHdr* h = &ring.hdr[s];
            const int a_w = pos.x * kT;
            a = a_w + lane_id;
            b0 = pos.y * TL + warp * LPW;
            off0 = ((long long)b0 * g.pitch + a_w) * 3;        // the warp's first output byte of line b0
            const int a_c = min(a, g.sz1 - 1);         // out-of-frame lanes shadow the last pixel
            const uint32_t R1 = h->R1, R2 = h->R2;
            const uint32_t rel0 = h->base_off;
            m_staged = m_fill = m_skip = 0;
            if (!EXACT || a_w + kT > g.sz1 || b0 + LPW > g.sz2) {   // partial tile (warp-uniform test; measured: pays only in the exact kernel)
#pragma unroll
                for (int e = 0; e < LPW; ++e)
                    if (a >= g.sz1 || b0 + e >= g.sz2) m_skip |= 1u << e;
            }
            if (EXACT) {
                const RowTermD rtd = rect_row_term(pe, g.axs0 + a_c);
                const double Mk1 = h->Mk1, Mk2 = h->Mk2;
                const double* q2p = &ring.q2[s][warp * LPW];
#pragma unroll
                for (int e = 0; e < LPW; ++e) {
                    double row, col, d1, d2;
                    rect_coord_nobranch(pe, rtd, q2p[e], row, col);
                    uint32_t t1, t2, h1, h2;
                    floor_index<kFloorMode1>(row, Mk1, t1, h1, d1);
                    floor_index<kFloorMode2>(col, Mk2, t2, h2, d2);
                    const bool st = (((h1 ^ 0x43300000u) | (h2 ^ 0x43300000u)) == 0u) & (t1 < R1) & (t2 < R2);
                    rel[e] = rel0 + t2 * box_pitch_b + t1 * 3u;
                    selv[e] = sel6(rel[e]);            // stages are 128-byte aligned: (address & 3) == (rel & 3)
                    rel[e] &= ~3u;
                    if (e & 1) { wf1[e / 2].y = (float)d1; wf2[e / 2].y = (float)d2; }
                    else       { wf1[e / 2].x = (float)d1; wf2[e / 2].x = (float)d2; }
                    if (st) m_staged |= 1u << e;
                    else if (!(lin_ok(row, g.sz1) & lin_ok(col, g.sz2))) m_fill |= 1u << e;   // rare: border tiles
                }
            } else {
                const RowTermF rtf = rect_row_term(pf, g.axs0 + a_c);
                const float mk1 = h->mk1, mk2 = h->mk2;
                float2 ip;
                ip.x = (float)(g.axs1 + b0) - pf.c2;
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
                        rel[e] = rel0 + l2 * box_pitch_b + l1 * 3u;
                        selv[e] = sel6(rel[e]);
                        rel[e] &= ~3u;
                        // (unconditional: four FSETPs cost less than a divergent branch here; measured)
                        const bool inframe = (rr[j] >= 1.0f) & (rr[j] < (float)g.sz1) & (cc_[j] >= 1.0f) & (cc_[j] < (float)g.sz2);
                        if (st) m_staged |= 1u << e;
                        if (!inframe) m_fill |= 1u << e;
                    }
                }
            }
            if (R1 == 0u || R2 == 0u) { m_staged = 0; m_fill = 0; }   // tile the plan marked unusable
            constexpr uint32_t kAll = (1u << LPW) - 1u;
            all_staged = __all_sync(0xffffffffu, (m_staged == kAll) & (m_skip == 0u));
        }

        // ---- every frame of the unit: gather, blend, store
        const uint8_t* sframe = src + (long long)pos.z * g.frame_stride * 3;
        uint8_t* oline = dst + (long long)pos.z * g.frame_stride * 3 + off0;
        const uint32_t sbase = stage0 + (uint32_t)s * (uint32_t)cfg.box_bytes;
        if (all_staged) {
            uint32_t* ow = reinterpret_cast<uint32_t*>(oline) + lane_id;
#pragma unroll
            for (int bt = 0; bt < LPW / KB; ++bt) {
                uint32_t rgb[KB];
                [[maybe_unused]] bool amb[KB];
                Taps6 ta[KB], tb[KB];
#pragma unroll
                for (int j = 0; j < KB; ++j) {
                    const uint32_t o = sbase + rel[bt * KB + j];       // word aligned
#ifdef CAMCAL_CHECK_BOUNDS      // debug builds: the three words of both lines inside the stage
                    if (o < sbase || o + box_pitch_b + 12u > sbase + (uint32_t)cfg.box_bytes + 4u || (o & 3u)) __trap();
#endif
                    const unsigned sel = selv[bt * KB + j];    // box_pitch_b % 4 == 0: same for both lines
                    ta[j] = lds6w(o, sel);
                    tb[j] = lds6w(o + box_pitch_b, sel);
                }
#pragma unroll
                for (int hh = 0; hh < KB / 2; ++hh)
                    blend_rgb2<EXACT>(ta[2 * hh], tb[2 * hh], ta[2 * hh + 1], tb[2 * hh + 1],
                                      wf1[bt * (KB / 2) + hh], wf2[bt * (KB / 2) + hh],
                                      rgb[2 * hh], rgb[2 * hh + 1], amb[2 * hh], amb[2 * hh + 1]);
                if (EXACT) {
                    bool any = false;
#pragma unroll
                    for (int j = 0; j < KB; ++j) any |= amb[j];
                    if (__any_sync(0xffffffffu, any)) {        // rare: certify by the FP64 blend
#pragma unroll
                        for (int j = 0; j < KB; ++j)
                            if (amb[j]) rgb[j] = reblend_exact_u8<true>(&pe, &g, a, b0 + bt * KB + j, ta[j], tb[j]);
                    }
                }
#pragma unroll
                for (int j = 0; j < KB; ++j) {
                    const uint32_t v0 = __shfl_sync(0xffffffffu, rgb[j], wp0);
                    const uint32_t v1 = __shfl_sync(0xffffffffu, rgb[j], wp1);
                    if (lane_id < 24) __stcs(ow, __byte_perm(v0, v1, wsel));
                    ow = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(ow) + pitch3);
                }
            }
        } else {
            // border tiles: per-pixel class, byte stores
            uint8_t* o = oline + lane_id * 3;
#pragma unroll 1
            for (int e = 0; e < LPW; ++e, o += pitch3) {
                if ((m_skip >> e) & 1u) continue;
                uint32_t v;
                if ((m_staged >> e) & 1u) {
                    uint32_t r = 0, sel = 0;
                    float f1 = 0, f2 = 0;
#pragma unroll
                    for (int j = 0; j < LPW; ++j)
                        if (j == e) {
                            r = rel[j];
                            sel = selv[j];
                            f1 = (j & 1) ? wf1[j / 2].y : wf1[j / 2].x;
                            f2 = (j & 1) ? wf2[j / 2].y : wf2[j / 2].x;
                        }
                    const uint32_t q = sbase + r;
                    const Taps6 t0 = lds6w(q, sel), t1 = lds6w(q + box_pitch_b, sel);
                    uint32_t vq;
                    bool am, amq;
                    blend_rgb2<EXACT>(t0, t1, t0, t1, make_float2(f1, f1), make_float2(f2, f2), v, vq, am, amq);
                    if (EXACT && am) v = reblend_exact_u8<true>(&pe, &g, a, b0 + e, t0, t1);
                } else if ((m_fill >> e) & 1u) {
                    v = fill;
                } else {
                    RowTermD rtd;
                    RowTermF rtf;
                    if (EXACT) rtd = rect_row_term(pe, g.axs0 + a); else rtf = rect_row_term(pf, g.axs0 + a);
                    v = sample_direct_u8<EXACT>(pe, pf, rtd, rtf, g, sframe, pitch3, frame_bytes, b0 + e, fill);
                }
                store_rgb(o, v);
            }
        }
        __syncwarp();
        if (lane_id == 0) mbar_arrive(&ring.empty[s]);
        if (++s == cfg.stages) { s = 0; phase ^= 1; }
    }
}

}  // namespace cc
