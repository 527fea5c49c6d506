// ingest.cu -- image-file ingest on the device: JPEG streams -> u8c3 frames in the layout the
// rectification kernels read ("the step upstream of rectification").
//
// Reference: the plot loop reads every calibration image with FileIO.load and converts it to RGB
// (src/plot_calibration.jl:37, also src/detect_fit.jl:6,64) before warp (:40).  Here the compressed
// bytes go to the GPU once: nvJPEG (CUDA toolkit LIBRARY code, loaded with dlopen on first use so the
// shared library itself has no link-time dependency on it) decodes into a raster (width contiguous),
// and a small hand-written kernel transposes the raster into the memory of the Julia array img[r, c]
// (first RowCol axis = image row contiguous), which is what cc_rectify_u8c3 / _views take.  Nothing
// comes back to the host.  The hardware video path (NVDEC) needs the Video Codec SDK headers, which
// this image does not have; nvJPEG covers still images and MJPEG streams.
#include <dlfcn.h>
#include <nvjpeg.h>

#include <mutex>

#include "common.cuh"

namespace cc {

struct NvJpegApi {
    void* lib;
    nvjpegStatus_t (*CreateSimple)(nvjpegHandle_t*);
    nvjpegStatus_t (*Destroy)(nvjpegHandle_t);
    nvjpegStatus_t (*StateCreate)(nvjpegHandle_t, nvjpegJpegState_t*);
    nvjpegStatus_t (*StateDestroy)(nvjpegJpegState_t);
    nvjpegStatus_t (*GetImageInfo)(nvjpegHandle_t, const unsigned char*, size_t, int*, nvjpegChromaSubsampling_t*, int*, int*);
    nvjpegStatus_t (*Decode)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char*, size_t, nvjpegOutputFormat_t,
                             nvjpegImage_t*, cudaStream_t);
};

static NvJpegApi g_nvjpeg;
static std::once_flag g_nvjpeg_once;

static const NvJpegApi* nvjpeg_api() {
    std::call_once(g_nvjpeg_once, [] {
        const char* names[] = {"libnvjpeg.so.12", "libnvjpeg.so"};
        for (const char* n : names)
            if ((g_nvjpeg.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break;
        if (!g_nvjpeg.lib) return;
#define CC_SYM(field, name) \
    g_nvjpeg.field = reinterpret_cast<decltype(g_nvjpeg.field)>(dlsym(g_nvjpeg.lib, name)); \
    if (!g_nvjpeg.field) { dlclose(g_nvjpeg.lib); g_nvjpeg.lib = nullptr; return; }
        CC_SYM(CreateSimple, "nvjpegCreateSimple")
        CC_SYM(Destroy, "nvjpegDestroy")
        CC_SYM(StateCreate, "nvjpegJpegStateCreate")
        CC_SYM(StateDestroy, "nvjpegJpegStateDestroy")
        CC_SYM(GetImageInfo, "nvjpegGetImageInfo")
        CC_SYM(Decode, "nvjpegDecode")
#undef CC_SYM
    });
    return g_nvjpeg.lib ? &g_nvjpeg : nullptr;
}

// raster[h][w][3] (w contiguous, `rpitch` bytes per raster line) -> dst[(c * pitch + r) * 3 + ch]:
// pixel (row r, column c) of the image lands where the Julia array img[r + 1, c + 1] lives.
// 32 x 32 pixel tiles through shared memory: raster lines are read as 96 contiguous bytes per warp,
// frame lines are written as 96 contiguous bytes per warp.
constexpr int kTr = 32;
__global__ void __launch_bounds__(kTr * 8)
raster_to_frame_u8c3(const uint8_t* __restrict__ raster, size_t rpitch, int h, int w, uint8_t* __restrict__ dst,
                     size_t pitch) {
    __shared__ uint8_t tile[kTr][kTr * 3 + 4];          // [raster line][3 bytes x 32 columns] (+4: bank spread)
    const int r0 = blockIdx.y * kTr, c0 = blockIdx.x * kTr;
    const int nr = min(kTr, h - r0), nc = min(kTr, w - c0);
    // load: thread t reads byte t % 96 of raster line t / 96 (+ 8 lines per pass)
    for (int i = threadIdx.x; i < kTr * kTr * 3; i += blockDim.x) {
        const int line = i / (kTr * 3), b = i - line * (kTr * 3);
        if (line < nr && b < nc * 3) tile[line][b] = raster[(size_t)(r0 + line) * rpitch + (size_t)c0 * 3 + b];
    }
    __syncthreads();
    // store: frame line c (fixed image column) holds the rows r0..r0+nr-1 contiguously, 3 bytes each
    for (int i = threadIdx.x; i < kTr * kTr * 3; i += blockDim.x) {
        const int col = i / (kTr * 3), b = i - col * (kTr * 3);
        const int row = b / 3, ch = b - 3 * row;
        if (col < nc && row < nr) dst[((size_t)(c0 + col) * pitch + (size_t)(r0 + row)) * 3 + ch] = tile[row][col * 3 + ch];
    }
}

struct Ingest {
    nvjpegHandle_t handle;
    nvjpegJpegState_t state;
    uint8_t* raster;
    size_t raster_bytes;
};

void ingest_free(cc_ctx* ctx) {
    Ingest* in = static_cast<Ingest*>(ctx->ingest);
    if (!in) return;
    const NvJpegApi* api = nvjpeg_api();
    if (api) {
        if (in->state) api->StateDestroy(in->state);
        if (in->handle) api->Destroy(in->handle);
    }
    if (in->raster) cudaFree(in->raster);
    delete in;
    ctx->ingest = nullptr;
}

static int ingest_get(cc_ctx* ctx, Ingest** out) {
    const NvJpegApi* api = nvjpeg_api();
    if (!api) return set_error(CC_ERR_UNSUPPORTED, "libnvjpeg.so.12 not found (CUDA toolkit library)");
    if (!ctx->ingest) {
        Ingest* in = new (std::nothrow) Ingest();
        if (!in) return set_error(CC_ERR_CUDA, "out of host memory");
        in->handle = nullptr; in->state = nullptr; in->raster = nullptr; in->raster_bytes = 0;
        if (api->CreateSimple(&in->handle) != NVJPEG_STATUS_SUCCESS ||
            api->StateCreate(in->handle, &in->state) != NVJPEG_STATUS_SUCCESS) {
            ctx->ingest = in;
            ingest_free(ctx);
            return set_error(CC_ERR_CUDA, "nvjpegCreateSimple / nvjpegJpegStateCreate failed");
        }
        ctx->ingest = in;
    }
    *out = static_cast<Ingest*>(ctx->ingest);
    return CC_OK;
}

int jpeg_info(const uint8_t* data, size_t length, int* sz1, int* sz2, int* channels) {
    const NvJpegApi* api = nvjpeg_api();
    if (!api) return set_error(CC_ERR_UNSUPPORTED, "libnvjpeg.so.12 not found (CUDA toolkit library)");
    // GetImageInfo only parses the headers, but it wants a handle: a process-wide one
    static nvjpegHandle_t info_handle = nullptr;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (!info_handle && api->CreateSimple(&info_handle) != NVJPEG_STATUS_SUCCESS) {
        info_handle = nullptr;
        return set_error(CC_ERR_CUDA, "nvjpegCreateSimple failed");
    }
    int ncomp = 0, widths[NVJPEG_MAX_COMPONENT] = {}, heights[NVJPEG_MAX_COMPONENT] = {};
    nvjpegChromaSubsampling_t ss;
    if (api->GetImageInfo(info_handle, data, length, &ncomp, &ss, widths, heights) != NVJPEG_STATUS_SUCCESS)
        return set_error(CC_ERR_INVALID_ARG, "not a JPEG stream nvJPEG can parse");
    if (sz1) *sz1 = heights[0];
    if (sz2) *sz2 = widths[0];
    if (channels) *channels = ncomp;
    return CC_OK;
}

int jpeg_decode_u8c3(cc_ctx* ctx, const uint8_t* const* jpegs, const size_t* lengths, int n, uint8_t* dst, int sz1,
                     int sz2, size_t pitch, size_t frame_stride, cudaStream_t st) {
    Ingest* in = nullptr;
    int rc = ingest_get(ctx, &in);
    if (rc) return rc;
    const NvJpegApi* api = nvjpeg_api();
    const size_t rpitch = ((size_t)sz2 * 3 + 255) / 256 * 256;       // raster line: width x RGB
    const size_t need = rpitch * (size_t)sz1;
    if (in->raster_bytes < need) {
        if (in->raster) { CC_CUDA(cudaStreamSynchronize(st)); CC_CUDA(cudaFree(in->raster)); in->raster = nullptr; in->raster_bytes = 0; }
        CC_CUDA(cudaMalloc(&in->raster, need));
        in->raster_bytes = need;
    }
    for (int i = 0; i < n; ++i) {
        int ncomp = 0, widths[NVJPEG_MAX_COMPONENT] = {}, heights[NVJPEG_MAX_COMPONENT] = {};
        nvjpegChromaSubsampling_t ss;
        if (api->GetImageInfo(in->handle, jpegs[i], lengths[i], &ncomp, &ss, widths, heights) != NVJPEG_STATUS_SUCCESS)
            return set_error(CC_ERR_INVALID_ARG, "image %d: not a JPEG stream nvJPEG can parse", i);
        if (heights[0] != sz1 || widths[0] != sz2)
            return set_error(CC_ERR_INVALID_ARG, "image %d is %d x %d, the frames are %d x %d (rows x columns)", i,
                             heights[0], widths[0], sz1, sz2);
        nvjpegImage_t img;
        memset(&img, 0, sizeof(img));
        img.channel[0] = in->raster;
        img.pitch[0] = rpitch;
        // grey JPEGs come out as R = G = B, like RGB.(load(file)) (src/plot_calibration.jl:37)
        if (api->Decode(in->handle, in->state, jpegs[i], lengths[i], NVJPEG_OUTPUT_RGBI, &img, st) != NVJPEG_STATUS_SUCCESS)
            return set_error(CC_ERR_CUDA, "image %d: nvjpegDecode failed", i);
        const dim3 grid((sz2 + kTr - 1) / kTr, (sz1 + kTr - 1) / kTr);
        raster_to_frame_u8c3<<<grid, kTr * 8, 0, st>>>(in->raster, rpitch, sz1, sz2, dst + (size_t)i * frame_stride * 3, pitch);
        ctx->launches++;
    }
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

}  // namespace cc
