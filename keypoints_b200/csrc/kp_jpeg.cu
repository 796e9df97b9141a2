// Real-data ingestion on the GPU: JPEG decode with nvJPEG + the CelebA transform of the reference
// (datasets.py:297-300: transforms.Resize((128,128)) on the PIL image = Pillow's antialiased BILINEAR, then ToTensor) as
// two small kernels, writing straight into a slot of the fp32 NCHW batch the trainer consumes.  libnvjpeg is opened with
// dlopen at first use: the training library itself carries no dependency on it.
#include "kp_common.cuh"
#include <dlfcn.h>
#include <nvjpeg.h>

namespace {

struct NvjpegApi {
    void* lib = nullptr;
    nvjpegStatus_t (*CreateSimple)(nvjpegHandle_t*) = nullptr;
    nvjpegStatus_t (*Destroy)(nvjpegHandle_t) = nullptr;
    nvjpegStatus_t (*StateCreate)(nvjpegHandle_t, nvjpegJpegState_t*) = nullptr;
    nvjpegStatus_t (*StateDestroy)(nvjpegJpegState_t) = nullptr;
    nvjpegStatus_t (*GetImageInfo)(nvjpegHandle_t, const unsigned char*, size_t, int*, nvjpegChromaSubsampling_t*, int*, int*) = nullptr;
    nvjpegStatus_t (*Decode)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char*, size_t, nvjpegOutputFormat_t, nvjpegImage_t*,
                             cudaStream_t) = nullptr;
};

static NvjpegApi* nvjpeg_api() {
    static NvjpegApi api;
    static bool tried = false;
    if (tried) return api.lib ? &api : nullptr;
    tried = true;
    const char* names[] = {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12", "/usr/local/cuda/lib64/libnvjpeg.so"};
    for (const char* n : names) {
        api.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (api.lib) break;
    }
    if (!api.lib) return nullptr;
#define KP_SYM(field, name)                                                 \
    *(void**)(&api.field) = dlsym(api.lib, name);                           \
    if (!api.field) { dlclose(api.lib); api.lib = nullptr; return nullptr; }
    KP_SYM(CreateSimple, "nvjpegCreateSimple");
    KP_SYM(Destroy, "nvjpegDestroy");
    KP_SYM(StateCreate, "nvjpegJpegStateCreate");
    KP_SYM(StateDestroy, "nvjpegJpegStateDestroy");
    KP_SYM(GetImageInfo, "nvjpegGetImageInfo");
    KP_SYM(Decode, "nvjpegDecode");
#undef KP_SYM
    return &api;
}

struct JpegCtx {
    nvjpegHandle_t handle = nullptr;
    nvjpegJpegState_t state = nullptr;
};

// Pillow's resampling (Resample.c, ImagingResampleHorizontal_8bpc / Vertical_8bpc with the BILINEAR = triangle filter):
// support = max(scale, 1), taps [xmin, xmax) around centre (o + 0.5) * scale, weights triangle((x - centre + 0.5) / support)
// normalised to sum 1, 8-bit result rounded after EACH pass (horizontal first, then vertical).
__device__ __forceinline__ void taps(int o, float scale, int in_size, int& xmin, int& xmax, float& centre, float& inv_fs) {
    const float fs = fmaxf(scale, 1.0f);
    centre = (o + 0.5f) * scale;
    inv_fs = 1.0f / fs;
    xmin = (int)(centre - fs + 0.5f);
    if (xmin < 0) xmin = 0;
    xmax = (int)(centre + fs + 0.5f);
    if (xmax > in_size) xmax = in_size;
}
__device__ __forceinline__ float tri(float x) { x = fabsf(x); return x < 1.0f ? 1.0f - x : 0.0f; }

// in: [H][W][C] uint8 (pitch bytes per row) -> tmp: [H][OW][C] uint8
__global__ void resize_h_k(const unsigned char* __restrict__ in, int in_pitch, int H, int W, int C, unsigned char* __restrict__ tmp,
                           int OW, float scale) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= H * OW) return;
    const int y = idx / OW, ox = idx - y * OW;
    int xmin, xmax;
    float centre, inv_fs;
    taps(ox, scale, W, xmin, xmax, centre, inv_fs);
    float ww = 0.f, acc[4] = {0.f, 0.f, 0.f, 0.f};
    const unsigned char* row = in + (size_t)y * in_pitch;
    for (int x = xmin; x < xmax; ++x) {
        const float w = tri((x - centre + 0.5f) * inv_fs);
        ww += w;
        for (int c = 0; c < C; ++c) acc[c] = fmaf(w, (float)row[x * C + c], acc[c]);
    }
    const float inv = ww > 0.f ? 1.0f / ww : 0.f;
    for (int c = 0; c < C; ++c) {
        const float v = fminf(fmaxf(acc[c] * inv + 0.5f, 0.f), 255.f);
        tmp[((size_t)y * OW + ox) * C + c] = (unsigned char)(int)v;
    }
}

// tmp: [H][OW][C] uint8 -> out: [C][OH][OW] fp32 = round8(vertical filter) / 255  (ToTensor)
__global__ void resize_v_k(const unsigned char* __restrict__ tmp, int H, int OW, int C, float* __restrict__ out, int OH, float scale) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= OH * OW) return;
    const int oy = idx / OW, ox = idx - oy * OW;
    int ymin, ymax;
    float centre, inv_fs;
    taps(oy, scale, H, ymin, ymax, centre, inv_fs);
    float ww = 0.f, acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int y = ymin; y < ymax; ++y) {
        const float w = tri((y - centre + 0.5f) * inv_fs);
        ww += w;
        const unsigned char* px = tmp + ((size_t)y * OW + ox) * C;
        for (int c = 0; c < C; ++c) acc[c] = fmaf(w, (float)px[c], acc[c]);
    }
    const float inv = ww > 0.f ? 1.0f / ww : 0.f;
    for (int c = 0; c < C; ++c) {
        const float v = fminf(fmaxf(acc[c] * inv + 0.5f, 0.f), 255.f);
        out[((size_t)c * OH + oy) * OW + ox] = (float)(int)v * (1.0f / 255.0f);
    }
}

}  // namespace

extern "C" int kp_jpeg_create(void** ctx) {
    KP_CHECK_ARG(ctx, "kp_jpeg_create: null argument");
    NvjpegApi* api = nvjpeg_api();
    if (!api) {
        kp_set_error("kp_jpeg_create: libnvjpeg could not be loaded (%s)", dlerror() ? dlerror() : "symbols missing");
        return KP_ERR_UNSUPPORTED;
    }
    JpegCtx* c = new JpegCtx();
    if (api->CreateSimple(&c->handle) != NVJPEG_STATUS_SUCCESS || api->StateCreate(c->handle, &c->state) != NVJPEG_STATUS_SUCCESS) {
        if (c->handle) api->Destroy(c->handle);
        delete c;
        kp_set_error("kp_jpeg_create: nvjpegCreateSimple / nvjpegJpegStateCreate failed");
        return KP_ERR_CUDA;
    }
    *ctx = c;
    return KP_OK;
}

extern "C" int kp_jpeg_destroy(void* ctx) {
    if (!ctx) return KP_OK;
    NvjpegApi* api = nvjpeg_api();
    JpegCtx* c = (JpegCtx*)ctx;
    if (api) {
        if (c->state) api->StateDestroy(c->state);
        if (c->handle) api->Destroy(c->handle);
    }
    delete c;
    return KP_OK;
}

extern "C" int kp_jpeg_info(void* ctx, const uint8_t* data, int64_t len, int* width, int* height, int* components) {
    KP_CHECK_ARG(ctx && data && len > 0 && width && height, "kp_jpeg_info: bad arguments");
    NvjpegApi* api = nvjpeg_api();
    JpegCtx* c = (JpegCtx*)ctx;
    int ncomp = 0, ws[NVJPEG_MAX_COMPONENT], hs[NVJPEG_MAX_COMPONENT];
    nvjpegChromaSubsampling_t ss;
    const nvjpegStatus_t s = api->GetImageInfo(c->handle, data, (size_t)len, &ncomp, &ss, ws, hs);
    if (s != NVJPEG_STATUS_SUCCESS) {
        kp_set_error("kp_jpeg_info: nvjpegGetImageInfo failed (%d) - not a baseline/progressive JPEG stream?", (int)s);
        return KP_ERR_ARG;
    }
    *width = ws[0];
    *height = hs[0];
    if (components) *components = ncomp;
    return KP_OK;
}

extern "C" int kp_jpeg_decode(void* ctx, kp_stream stream, const uint8_t* data, int64_t len, uint8_t* out_rgb, int width,
                              int height) {
    KP_CHECK_ARG(ctx && data && len > 0 && out_rgb && width > 0 && height > 0, "kp_jpeg_decode: bad arguments");
    NvjpegApi* api = nvjpeg_api();
    JpegCtx* c = (JpegCtx*)ctx;
    nvjpegImage_t img;
    for (int i = 0; i < NVJPEG_MAX_COMPONENT; ++i) { img.channel[i] = nullptr; img.pitch[i] = 0; }
    img.channel[0] = out_rgb;
    img.pitch[0] = (size_t)width * 3;
    const nvjpegStatus_t s = api->Decode(c->handle, c->state, data, (size_t)len, NVJPEG_OUTPUT_RGBI, &img, (cudaStream_t)stream);
    if (s != NVJPEG_STATUS_SUCCESS) {
        kp_set_error("kp_jpeg_decode: nvjpegDecode failed (%d)", (int)s);
        return KP_ERR_CUDA;
    }
    return KP_OK;
}

extern "C" int kp_resize_to_f32(kp_stream stream, const uint8_t* in_hwc, int in_h, int in_w, int channels, uint8_t* tmp,
                                float* out_chw, int out_h, int out_w) {
    KP_CHECK_ARG(in_hwc && tmp && out_chw && in_h > 0 && in_w > 0 && out_h > 0 && out_w > 0 && channels >= 1 && channels <= 4,
                 "kp_resize_to_f32: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const int n1 = in_h * out_w, n2 = out_h * out_w;
    resize_h_k<<<(n1 + 255) / 256, 256, 0, st>>>(in_hwc, in_w * channels, in_h, in_w, channels, tmp, out_w, (float)in_w / (float)out_w);
    resize_v_k<<<(n2 + 255) / 256, 256, 0, st>>>(tmp, in_h, out_w, channels, out_chw, out_h, (float)in_h / (float)out_h);
    KP_LAUNCH_CHECK();
    return KP_OK;
}
