// rectify_f32c1.cuh -- fp32 single-channel rectification kernel (included by rectify.cu).
//
// Per-pixel instruction budget is what bounds this kernel (profiles/r1_rectify.md), so the
// staged path is written to the instruction:
//  * the producer folds the tile's index origin into the floor constant: the exact path adds
//    Mk = 2^52 - K to floor(x) (or adds it with round-down, DADD.RM, which IS the floor), the
//    fast path adds mk = 1.5*2^23 - K with FADD.RM; the low word of the sum is the tile-local
//    tap index, and "high word == 0x43300000" / "(bits - 0x4B400000) < R" are the complete
//    range tests (negative, NaN, huge and out-of-box coordinates all fail them);
//  * 1/P3 is NVIDIA's own correctly rounded sequence (MUFU.RCP64H + 5 DFMA) inlined without
//    its per-pixel branch: the exponent test that guards it joins the range predicate, and a
//    failing warp takes the generic path, which divides;
//  * one vote per batch of lines decides between the staged gather and the generic path.
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

template <bool EXACT, bool TMA>
__global__ void __launch_bounds__(TMA ? kConsumerThreads + 32 : kConsumerThreads)
rectify_f32c1_kernel(const __grid_constant__ CUtensorMap tmap, const RectExact pe, const RectFast pf,
                     const RectGeom g, const TileCfg cfg, const float* __restrict__ src,
                     float* __restrict__ dst, float fill) {
    constexpr int KB = EXACT ? kBatchExact : kBatchFast;
    extern __shared__ __align__(128) uint8_t stage_mem[];
    __shared__ SmemCtl ctl;
    const int lane_id = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int a_lo = blockIdx.x * kT;
    const int frame = blockIdx.z;
    const int t_begin = blockIdx.y * cfg.tiles_per_seg;
    const int t_end = min(t_begin + cfg.tiles_per_seg, cfg.ntiles2);
    pipeline_init(&ctl, cfg.stages, TMA);

    if (TMA && warp == kWarps) {                       // ---- producer warp
        if (lane_id == 0) tma_prefetch_desc(&tmap);
        int s = 0;
        uint32_t phase = 1;                            // a fresh barrier passes a parity-1 wait
        for (int tile = t_begin; tile < t_end; ++tile) {
            mbar_wait(&ctl.empty[s], phase);
            producer_tile<EXACT, 1>(&tmap, pf, pe, g, cfg, &ctl, stage_mem + (size_t)s * cfg.box_bytes,
                                    s, a_lo, tile, frame, lane_id);
            if (++s == cfg.stages) { s = 0; phase ^= 1; }
        }
        return;
    }

    // ---- consumer warps
    const int a = a_lo + lane_id;
    const bool a_in = a < g.sz1;
    const float* sframe = src + (long long)frame * g.frame_stride;
    const unsigned pitch = (unsigned)g.pitch;
    RowTermD rtd;
    RowTermF rtf;
    const int a_c = min(a, g.sz1 - 1);                 // out-of-frame lanes shadow the last pixel
    if (EXACT) rtd = rect_row_term(pe, g.axs0 + a_c); else rtf = rect_row_term(pf, g.axs0 + a_c);
    const int line0 = t_begin * kT + warp * kLines;
    float* optr = dst + (long long)frame * g.frame_stride + (long long)line0 * g.pitch + a;
    const long long tile_step = (long long)(kT - kLines) * g.pitch;
    const uint32_t box_pitch_b = (uint32_t)cfg.box1 * 4u;
    [[maybe_unused]] const float2 fc0 = bc2(pf.Cc[0]), fc1 = bc2(pf.Cc[1]), fc2 = bc2(pf.Cc[2]);

    int s = 0;
    uint32_t phase = 0;
    for (int tile = t_begin; tile < t_end; ++tile) {
        const int b0 = tile * kT + warp * kLines;
        const bool full_lines = (b0 + kLines <= g.sz2) && a_lo + kT <= g.sz1;
        const StageHdr* h = &ctl.hdr[s];
        [[maybe_unused]] double Mk1 = 0, Mk2 = 0;
        [[maybe_unused]] float mk1 = 0, mk2 = 0;
        uint32_t R1 = 0, R2 = 0, base = 0;
        if (TMA) {
            mbar_wait(&ctl.full[s], phase);
            if (EXACT) { Mk1 = h->Mk1; Mk2 = h->Mk2; } else { mk1 = h->mk1; mk2 = h->mk2; }
            R1 = (uint32_t)h->R1; R2 = (uint32_t)h->R2; base = h->base;
        }
        // fast path: second-axis term of the first line of this warp, relative to the output centre
        [[maybe_unused]] const float i2f = (float)(g.axs1 + b0) - pf.c2;
#pragma unroll
        for (int batch = 0; batch < kLines / KB; ++batch) {
            bool fast = false;
            if (TMA) {
                uint32_t l1[KB], l2[KB];
                bool ok = full_lines;
                [[maybe_unused]] double d1d[KB], d2d[KB];
                [[maybe_unused]] float2 d1p[KB / 2 + 1], d2p[KB / 2 + 1];
                if (EXACT) {
#pragma unroll
                    for (int e = 0; e < KB; ++e) {
                        double row, col;
                        ok &= rect_coord_nobranch(pe, rtd, h->q2[warp * kLines + batch * KB + e], row, col);
                        uint32_t h1, h2;
                        floor_index<kFloorMode1>(row, Mk1, l1[e], h1, d1d[e]);
                        floor_index<kFloorMode2>(col, Mk2, l2[e], h2, d2d[e]);
                        ok &= ((h1 ^ 0x43300000u) | (h2 ^ 0x43300000u)) == 0u;
                        ok &= (l1[e] < R1) & (l2[e] < R2);
                    }
                } else {
#pragma unroll
                    for (int hh = 0; hh < KB / 2; ++hh) {
                        float2 row, col;
                        const float e0 = (float)(batch * KB + 2 * hh);
                        rect_coord2(pf, rtf, fc0, fc1, fc2, make_float2(i2f + e0, i2f + (e0 + 1.0f)), row, col);
                        floor_index_fast2(row, mk1, l1[2 * hh], l1[2 * hh + 1], d1p[hh]);
                        floor_index_fast2(col, mk2, l2[2 * hh], l2[2 * hh + 1], d2p[hh]);
                        ok &= (l1[2 * hh] < R1) & (l2[2 * hh] < R2) & (l1[2 * hh + 1] < R1) & (l2[2 * hh + 1] < R2);
                    }
                }
                fast = __all_sync(0xffffffffu, ok);
                if (fast) {
                    float a00[KB], a10[KB], a01[KB], a11[KB];
#pragma unroll
                    for (int e = 0; e < KB; ++e) {
                        const uint32_t q = base + l2[e] * box_pitch_b + l1[e] * 4u;
                        const uint32_t q1 = q + box_pitch_b;
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
                }
            }
            if (!fast) {
                float* o = optr;
                const int bb = b0 + batch * KB;
                if (TMA) {
#pragma unroll 1
                    for (int e = 0; e < KB; ++e) {
                        const int b = bb + e;
                        if (b < g.sz2 && a_in)
                            __stcs(o, sample_direct_f32<EXACT>(pe, pf, rtd, rtf, g, sframe, pitch, b, fill));
                        o += pitch;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < KB; ++e) {
                        const int b = bb + e;
                        if (b < g.sz2 && a_in)
                            __stcs(o, sample_direct_f32<EXACT>(pe, pf, rtd, rtf, g, sframe, pitch, b, fill));
                        o += pitch;
                    }
                }
            }
            optr += (long long)KB * pitch;
        }
        optr += tile_step;
        if (TMA) {
            __syncwarp();
            if (lane_id == 0) mbar_arrive(&ctl.empty[s]);
            if (++s == cfg.stages) { s = 0; phase ^= 1; }
        }
    }
}

}  // namespace cc
