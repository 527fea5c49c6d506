// rectify_ring.cuh -- scheduling and producer side shared by the staged rectification kernels.
//
// Persistent CTAs fed from a global ticket counter.  A work unit is one tile (strip x, tile y)
// of a GROUP of consecutive frames, numbered x fastest, and units are handed out in that order to
// whichever CTA is free: at any instant the device works on a window of consecutive units, so tiles that
// share source lines (the halos of neighbouring strips) are fetched within microseconds of each
// other and the second fetch hits L2.  Static schedules (long private walks, round-robin) let
// the CTAs drift apart and were measured to read 1.5x the frame from DRAM
// (profiles/r1_rectify.md).
//
// Producer warp, per unit: take a ticket (one ticket ahead, so the atomic's latency is hidden),
// decode it, read the tile header and the q2 terms from the plan, wait for a free stage, publish
// {position, header, q2} in the stage's slot and issue ONE cp.async.bulk.tensor.  A ticket past
// the last unit is published as a stop marker.  The last producer to leave resets the counter.
#pragma once

namespace cc {

__device__ __forceinline__ void ring_init(SmemRing* ring, int stages) {
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(&ring->full[s], 1);
            mbar_init(&ring->empty[s], kWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();
}

// CAMCAL_PRODUCER_PARK > 0: wait with the barrier's own suspend-time hint (the polling loop of
// try_wait + nanosleep was 8 % of all executed warp instructions of the u8 kernel, ncu round 2)
#ifndef CAMCAL_PRODUCER_PARK
#define CAMCAL_PRODUCER_PARK 2000
#endif
__device__ __forceinline__ void producer_wait(uint64_t* bar, uint32_t parity) {
    if (CAMCAL_PRODUCER_PARK > 0) mbar_wait_parked<CAMCAL_PRODUCER_PARK>(bar, parity);
    else mbar_wait<kProducerSleep>(bar, parity);
}

__device__ __forceinline__ uint32_t take_ticket(RectSched* sched, int lane_id) {
    uint32_t u = 0;
    if (lane_id == 0) u = atomicAdd(&sched->next, 1u);
    return __shfl_sync(0xffffffffu, u, 0);
}

// PXB: multiplier from TileHdr.x0 to the tensor-map coordinate (1: f32c1 texels, and u8c3 whose x0 already is a byte offset).
// A unit is (strip x, tile y, frame group): the producer publishes one slot per FRAME of the
// group -- pos = (x, y, frame, 1 on the first frame of a unit) -- so the consumers rebuild the
// tile's map only when pos.w is set and otherwise just gather.
template <bool EXACT, int TL, int PXB>
__device__ __forceinline__ void producer_loop(const CUtensorMap* tmap, const RectGeom& g, const TileCfg& cfg,
                                              const TileHdr* __restrict__ plan,
                                              const double* __restrict__ q2tab, RectSched* sched,
                                              SmemRing* ring, uint8_t* stage_mem, int lane_id) {
    if (lane_id == 0) tma_prefetch_desc(tmap);
    int s = 0;
    uint32_t phase = 1;                            // a fresh barrier passes a parity-1 wait
    uint32_t u_next = take_ticket(sched, lane_id);
    for (;;) {
        const uint32_t u = u_next;
        if (u >= cfg.units) break;
        const int x = (int)(u % (uint32_t)cfg.strips);
        const uint32_t r = u / (uint32_t)cfg.strips;
        const int y = (int)(r % (uint32_t)cfg.ntiles2);
        const int f0 = (int)(r / (uint32_t)cfg.ntiles2) * cfg.fg;
        const int f1 = min(f0 + cfg.fg, g.nframes);
        const uint32_t* hp = reinterpret_cast<const uint32_t*>(plan + x * cfg.ntiles2 + y);
        uint32_t hword = 0;
        if (lane_id < 12) hword = __ldg(hp + lane_id);            // the 48-byte header, one word per lane
        [[maybe_unused]] double q2a = 0, q2b = 0;
        if (EXACT) {
            const int b = y * TL + lane_id;
            q2a = __ldg(q2tab + min(b, g.sz2 - 1));
            if (TL > 32) q2b = __ldg(q2tab + min(b + 32, g.sz2 - 1));
        }
        u_next = take_ticket(sched, lane_id);
        const int x0 = __shfl_sync(0xffffffffu, (int)hword, 8), y0 = __shfl_sync(0xffffffffu, (int)hword, 9);
        for (int f = f0; f < f1; ++f) {
            producer_wait(&ring->empty[s], phase);
            if (f == f0) {                                        // the map inputs travel with the first frame
                if (lane_id < 12) reinterpret_cast<uint32_t*>(&ring->hdr[s])[lane_id] = hword;
                if (EXACT) {
                    if (lane_id < TL) ring->q2[s][lane_id] = q2a;
                    if (TL > 32) ring->q2[s][lane_id + 32] = q2b;
                }
            }
            if (lane_id == 12) ring->pos[s] = make_int4(x, y, f, f == f0 ? 1 : 0);
            __syncwarp();
            if (lane_id == 0) {
                mbar_arrive_expect_tx(&ring->full[s], (uint32_t)cfg.box_bytes);
                tma_load_3d(stage_mem + (size_t)s * cfg.box_bytes, tmap, &ring->full[s], x0 * PXB, y0, f);
            }
            if (++s == cfg.stages) { s = 0; phase ^= 1; }
        }
    }
    // stop marker
    producer_wait(&ring->empty[s], phase);
    if (lane_id == 0) { ring->pos[s] = make_int4(0, 0, -1, 0); mbar_arrive(&ring->full[s]); }
    // every producer has taken its last ticket before it counts itself out
    if (lane_id == 0 && atomicAdd(&sched->done, 1u) == gridDim.x - 1) {
        sched->next = 0;
        sched->done = 0;
    }
}

}  // namespace cc
