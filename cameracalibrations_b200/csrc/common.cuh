// common.cuh -- shared declarations of libcamcal_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include "../../include/camcal_b200.h"

struct cc_ctx;
namespace cc {

// What `Calibration(...)` / `img2obj` derive once per view (src/meta.jl:27-33, 71-76),
// expanded on the host and passed to kernels by value.  Row-major 3x3.
template <typename T>
struct Chain {
    T R[9], t[3], Rinv[9], tinv[3];
    T a_row, b_row, a_col, b_col;   // inv(intrinsic): u = a*rc + b
    T frow, fcol, crow, ccol, k;
    T inv_cs, cs_back;              // scale and inv(scale)
};
using ChainD = Chain<double>;
using ChainF = Chain<float>;

// host: rotation vector -> matrix and derived maps (chain_host.cu)
void rodrigues_host(const double r[3], double R[9]);
void build_chain(const cc_intr* in, const cc_view* vw, ChainD* out);
void narrow_chain(const ChainD& d, ChainF* f);
void rectify_free_plans(struct ::cc_ctx* ctx);
void rectify_free_sched(struct ::cc_ctx* ctx);
void ingest_free(struct ::cc_ctx* ctx);
int scratch_acquire(struct ::cc_ctx* ctx, size_t elems, cudaStream_t st);
int scratch_release(struct ::cc_ctx* ctx, cudaStream_t st);
int jpeg_info(const uint8_t* data, size_t length, int* sz1, int* sz2, int* channels);
int jpeg_decode_u8c3(struct ::cc_ctx* ctx, const uint8_t* const* jpegs, const size_t* lengths, int n, uint8_t* dst,
                     int sz1, int sz2, size_t pitch, size_t frame_stride, cudaStream_t st);
void comm_free(struct ::cc_ctx* ctx);
void lm_free_workspace(struct ::cc_ctx* ctx);

// error plumbing (abi.cu)
int set_error(int status, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

}  // namespace cc

struct cc_ctx {
    int device;
    int sm_count;
    int cc_major, cc_minor;
    unsigned long long launches;
    // reprojection scratch: [21][nviews] component-major partials
    double* jtj_scratch;
    size_t jtj_scratch_elems;
    // the scratch is shared by reproj_jtj / calc_errors / lm_*: calls on DIFFERENT streams are ordered
    // through an event recorded on the previous user's stream (scratch_acquire)
    cudaStream_t scratch_stream;
    cudaEvent_t scratch_event;
    int scratch_used;
    // host pipeline: NSLOT streams with device staging buffers
    static const int NSLOT = 4;
    cudaStream_t pipe_stream[NSLOT];
    void* pipe_in[NSLOT];
    void* pipe_out[NSLOT];
    size_t pipe_in_bytes[NSLOT];
    size_t pipe_out_bytes[NSLOT];
    // TMA descriptor encode entry point (driver API, resolved at ctx creation)
    void* encode_tiled;
    // rectification tile plans (rectify.cu: RectPlan), most recent NPLAN parameter sets
    static const int NPLAN = 64;
    void* rect_plans[NPLAN];
    int rect_plan_next;
    // tile plans of groups of views rectified in one launch (rectify.cu: MultiPlan), most recent NMULTI groups
    static const int NMULTI = 4;
    void* multi_plans[NMULTI];
    int multi_plan_next;
    // ticket counters of the persistent rectification kernels (rectify.cu: RectSched)
    static const int NSCHED = 16;
    void* sched_pool;
    cudaEvent_t sched_event[NSCHED];
    cudaStream_t sched_stream[NSCHED];
    unsigned char sched_used[NSCHED];
    unsigned sched_next;
    // NCCL communicator of this context (comm.cu); NULL / 1 rank: a world of one
    void* nccl_comm;
    int comm_nranks, comm_rank;
    unsigned long long collectives;       // all-reduces issued so far
    // workspace of the device-resident LM loop (lm.cu: LmWorkspace), grown on demand
    void* lm_ws;
    // nvJPEG handle / state / raster scratch of the image ingest (ingest.cu: Ingest), created on first use
    void* ingest;
};

#define CC_CUDA(call)                                                     \
    do {                                                                  \
        cudaError_t _e = (call);                                          \
        if (_e != cudaSuccess) return cc::cuda_fail(_e, #call);           \
    } while (0)

#define CC_REQUIRE(cond, msg)                                             \
    do {                                                                  \
        if (!(cond)) return cc::set_error(CC_ERR_INVALID_ARG, "%s", msg); \
    } while (0)

// ---- device helpers -------------------------------------------------------
namespace cc {

// streaming 128-bit accesses: inputs are read once, outputs written once
__device__ __forceinline__ double2 ldg_stream(const double2* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];"
                 : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream(double2* p, double2 v) {
    asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" :: "l"(p), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void stg_stream(float4* p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void stg_stream(uint4* p, uint4 v) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// 2^-23-accurate reciprocal seed on the SFU (MUFU.RCP64H)
__device__ __forceinline__ double rcp_approx(double a) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    return r;
}

}  // namespace cc
