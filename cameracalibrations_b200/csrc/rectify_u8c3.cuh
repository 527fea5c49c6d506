// rectify_u8c3.cuh -- 8-bit RGB (RGB{N0f8}, 3 interleaved bytes) rectification kernels
// (included by rectify.cu).  Same skeleton as rectify_f32c1.cuh: persistent CTAs, ticket
// scheduler, host-built tile plan, one TMA box per tile, raw-bit addressing.
//
// Per pixel the two taps of one source line are 6 contiguous bytes at byte offset 3*i: they are
// fetched as three aligned 32-bit words and funnelled with PRMT.  Bytes become floats without a
// conversion instruction (PRMT into 0x4B0000bb = 2^23 + b, minus 2^23 in a packed FADD2), the
// blend runs on PAIRS of lines in FFMA2/FADD2, and the result is rounded by a magic add.
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
// six bytes at shared byte address o (any alignment)
__device__ __forceinline__ Taps6 lds6(uint32_t o, unsigned sel) {
    const uint32_t wa = o & ~3u;
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

// two pixels at once (FADD2/FFMA2): same arithmetic as blend_rgb<false>
__device__ __forceinline__ float2 byte_f2(uint32_t wp, uint32_t wq, int k) {
    return add2(make_float2(__uint_as_float(__byte_perm(wp, 0x4B000000u, 0x7440u + (unsigned)k)),
                            __uint_as_float(__byte_perm(wq, 0x4B000000u, 0x7440u + (unsigned)k))),
                bc2(-8388608.0f));
}

// distance of a blended value from the integer it rounds to, against the certification bound
constexpr float kCertThr = 0.5f - 6.5e-5f;

// CERT: also report (per pixel) whether any channel is too close to a rounding boundary
template <bool CERT>
__device__ __forceinline__ void blend_rgb2(const Taps6& p0, const Taps6& p1, const Taps6& q0,
                                           const Taps6& q1, float2 d1, float2 d2, uint32_t& rgb_p,
                                           uint32_t& rgb_q, bool& amb_p, bool& amb_q) {
    const float2 m = bc2(12582912.0f);
    const float2 vr = bilerp_fast2(byte_f2(p0.lo, q0.lo, 0), byte_f2(p0.lo, q0.lo, 3),
                                   byte_f2(p1.lo, q1.lo, 0), byte_f2(p1.lo, q1.lo, 3), d1, d2);
    const float2 vg = bilerp_fast2(byte_f2(p0.lo, q0.lo, 1), byte_f2(p0.hi, q0.hi, 0),
                                   byte_f2(p1.lo, q1.lo, 1), byte_f2(p1.hi, q1.hi, 0), d1, d2);
    const float2 vb = bilerp_fast2(byte_f2(p0.lo, q0.lo, 2), byte_f2(p0.hi, q0.hi, 1),
                                   byte_f2(p1.lo, q1.lo, 2), byte_f2(p1.hi, q1.hi, 1), d1, d2);
    const float2 fr = add2(vr, m), fg = add2(vg, m), fb = add2(vb, m);
    rgb_p = __byte_perm(__byte_perm(__float_as_uint(fr.x), __float_as_uint(fg.x), 0x0040), __float_as_uint(fb.x), 0x0410);
    rgb_q = __byte_perm(__byte_perm(__float_as_uint(fr.y), __float_as_uint(fg.y), 0x0040), __float_as_uint(fb.y), 0x0410);
    if (CERT) {
        const float2 er = sub2(vr, sub2(fr, m)), eg = sub2(vg, sub2(fg, m)), eb = sub2(vb, sub2(fb, m));
        amb_p = (fabsf(er.x) > kCertThr) | (fabsf(eg.x) > kCertThr) | (fabsf(eb.x) > kCertThr);
        amb_q = (fabsf(er.y) > kCertThr) | (fabsf(eg.y) > kCertThr) | (fabsf(eb.y) > kCertThr);
    }
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

// `n` consecutive lines of one lane's pixel column through the generic path (cold)
template <bool EXACT>
__device__ __forceinline__ void generic_lines_u8(const RectExact& pe, const RectFast& pf, const RectGeom& g,
                                                 const uint8_t* __restrict__ sframe, uint8_t* __restrict__ o,
                                                 unsigned frame_bytes, int a, int b, int n, uint32_t fill) {
    if (a >= g.sz1) return;
    RowTermD rtd;
    RowTermF rtf;
    if (EXACT) rtd = rect_row_term(pe, g.axs0 + a); else rtf = rect_row_term(pf, g.axs0 + a);
    const unsigned pitch3 = (unsigned)g.pitch * 3u;
    n = min(n, g.sz2 - b);
#pragma unroll 1
    for (int e = 0; e < n; ++e, o += pitch3)
        store_rgb(o, sample_direct_u8<EXACT>(pe, pf, rtd, rtf, g, sframe, pitch3, frame_bytes, b + e, fill));
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
template <bool EXACT>
__global__ void __launch_bounds__(kConsumerThreads + 32, EXACT ? kMinBlocksU8Exact : kMinBlocksU8)
rectify_u8c3_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ RectExact pe,
                    const __grid_constant__ RectFast pf, const __grid_constant__ RectGeom g,
                    const __grid_constant__ TileCfg cfg, const TileHdr* __restrict__ plan,
                    const double* __restrict__ q2tab, RectSched* __restrict__ sched,
                    const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uchar3 fill3,
                    unsigned frame_bytes) {
    constexpr int KB = 4;                         // lines in flight per lane (two packed pairs)
    constexpr int TL = kTLu;                      // lines per tile
    constexpr int LPW = TL / kWarps;              // lines per warp per tile
    static_assert(LPW % KB == 0, "batch must divide the lines of a warp");
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

    int s = 0;
    uint32_t phase = 0;
    for (;;) {
        mbar_wait(&ring.full[s], phase);
        const int4 pos = ring.pos[s];
        if (pos.z < 0) break;
        const TileHdr* h = &ring.hdr[s];
        const int a_w = pos.x * kT;                            // first pixel of this warp's lanes
        const int a = a_w + lane_id;
        const int b0 = pos.y * TL + warp * LPW;
        const uint8_t* sframe = src + (long long)pos.z * g.frame_stride * 3;
        // the warp's first output byte of line b0 (4-byte aligned: checked on the host)
        uint8_t* oline = dst + ((long long)pos.z * g.frame_stride + (long long)b0 * g.pitch + a_w) * 3;
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
        const uint32_t base = stage0 + (uint32_t)s * (uint32_t)cfg.box_bytes + h->base_off - magic * (box_pitch_b + 3u);
        [[maybe_unused]] float2 ip;
        ip.x = (float)(g.axs1 + b0) - pf.c2;
        ip.y = ip.x + 1.0f;
        [[maybe_unused]] const double* q2p = &ring.q2[s][warp * LPW];
#pragma unroll 1
        for (int batch = 0; batch < LPW / KB; ++batch) {
            uint32_t t1[KB], t2[KB];
            uint32_t m1 = 0, m2 = 0;
            [[maybe_unused]] uint32_t hi_bad = 0;
            [[maybe_unused]] double d1d[KB], d2d[KB];
            float2 d1p[KB / 2], d2p[KB / 2];
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
#pragma unroll
                for (int hh = 0; hh < KB / 2; ++hh) {
                    d1p[hh] = make_float2((float)d1d[2 * hh], (float)d1d[2 * hh + 1]);
                    d2p[hh] = make_float2((float)d2d[2 * hh], (float)d2d[2 * hh + 1]);
                }
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
                Taps6 ta[KB], tb[KB];
#pragma unroll
                for (int e = 0; e < KB; ++e) {
                    const uint32_t o = base + t2[e] * box_pitch_b + t1[e] * 3u;
#ifdef CAMCAL_CHECK_BOUNDS      // debug builds: the six tap bytes of both lines inside the stage
                    {
                        const uint32_t lo = stage0 + (uint32_t)s * (uint32_t)cfg.box_bytes;
                        if (o < lo || o + box_pitch_b + 6u > lo + (uint32_t)cfg.box_bytes) __trap();
                    }
#endif
                    const unsigned sel = sel6(o);              // box_pitch_b % 4 == 0: same for both lines
                    ta[e] = lds6(o, sel);
                    tb[e] = lds6(o + box_pitch_b, sel);
                }
                uint32_t rgb[KB];
                [[maybe_unused]] bool amb[KB];
#pragma unroll
                for (int hh = 0; hh < KB / 2; ++hh)
                    blend_rgb2<EXACT>(ta[2 * hh], tb[2 * hh], ta[2 * hh + 1], tb[2 * hh + 1], d1p[hh], d2p[hh],
                                      rgb[2 * hh], rgb[2 * hh + 1], amb[2 * hh], amb[2 * hh + 1]);
                if (EXACT) {
                    bool any = false;
#pragma unroll
                    for (int e = 0; e < KB; ++e) any |= amb[e];
                    if (__any_sync(0xffffffffu, any)) {        // rare: certify by the FP64 blend
#pragma unroll
                        for (int e = 0; e < KB; ++e)
                            if (amb[e]) rgb[e] = blend_rgb<true>(ta[e], tb[e], d1d[e], d2d[e], 0.f, 0.f);
                    }
                }
                uint32_t* ow = reinterpret_cast<uint32_t*>(oline) + lane_id;
#pragma unroll
                for (int e = 0; e < KB; ++e) {
                    const uint32_t v0 = __shfl_sync(0xffffffffu, rgb[e], wp0);
                    const uint32_t v1 = __shfl_sync(0xffffffffu, rgb[e], wp1);
                    if (lane_id < 24) __stcs(ow, __byte_perm(v0, v1, wsel));
                    ow = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(ow) + pitch3);
                }
            } else {
                generic_lines_u8<EXACT>(pe, pf, g, sframe, oline + lane_id * 3, frame_bytes, a, b0 + batch * KB, KB, fill);
            }
            oline += (long long)KB * pitch3;
        }
        __syncwarp();
        if (lane_id == 0) mbar_arrive(&ring.empty[s]);
        if (++s == cfg.stages) { s = 0; phase ^= 1; }
    }
}

}  // namespace cc
