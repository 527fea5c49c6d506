// lm_state.cuh -- device-resident state of the Levenberg-Marquardt loop (lm.cu, residual.cu).
//
// The loop of cc_lm_fit_f64 never returns to the host between iterations: damping, the
// accept / reject decision and the stopping rule live in this block in device memory; every kernel
// of an iteration reads what it needs from it (which of the two view / block buffers is current,
// lambda, the candidate intrinsics) and returns at once when `done` is set.
#pragma once

#include "common.cuh"

namespace cc {

struct LmState {
    double par[4];        // current shared parameters (f, crow, ccol, k); frow = aspect * f
    double cand[4];       // candidate = par + di of the last update
    double lambda;        // Marquardt damping (CvLevMarq: starts at 1e-3, /10 accepted, *10 rejected)
    double sse;           // sum |residual|^2 at `par` over ALL ranks
    double npoints;       // residual points over all ranks (views * corners)
    double step2, size2;  // |step|^2, |parameters|^2 of the last accepted step
    int cur;              // index of the current view / per-view-block buffers
    int done;             // stopping rule met, or max_iter reached
    int iterations, accepted;
    int step_ok;          // the 4x4 solve of the last update succeeded
    int pad[3];
};

// the two generations of per-rank arrays the loop ping-pongs between
struct LmBufs {
    cc_view* views[2];
    double* pv[2];        // per-view blocks, nviews x CC_PER_VIEW
};

// residual.cu
int launch_reproj_jtj_state(cc_ctx* ctx, const LmState* st, int which, const LmBufs& b, double aspect,
                            double checker_size, int nviews, const double* obj, const double* img,
                            int ncorners, double* shared_out, cudaStream_t stream);
// comm.cu
int comm_allreduce_sum(cc_ctx* ctx, double* buf, size_t count, cudaStream_t st);

}  // namespace cc
