// C-ABI entry points: error reporting, device info and the convolution dispatch.
#include "common.cuh"
#include <string.h>

namespace cg {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
int g_tc_mode = 7;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int validate_geom(const cg_conv_geom& g) {
    CG_REQUIRE(g.G >= 1 && (g.x_groups == 1 || g.x_groups == g.G), "conv: x_groups=%d must be 1 or G=%d", g.x_groups, g.G);
    CG_REQUIRE(g.Cin % 4 == 0 && g.Cin > 0, "conv: Cin=%d must be a positive multiple of 4", g.Cin);
    CG_REQUIRE(g.Cout > 0 && g.B > 0 && g.H > 0 && g.W > 0, "conv: empty tensor");
    CG_REQUIRE(g.stride >= 1 && g.KH >= 1 && g.KW >= 1 && g.pad >= 0, "conv: bad kernel geometry");
    int Hin = g.ups ? 2 * g.H : g.H, Win = g.ups ? 2 * g.W : g.W;
    int Ho = (Hin + 2 * g.pad - g.KH) / g.stride + 1, Wo = (Win + 2 * g.pad - g.KW) / g.stride + 1;
    CG_REQUIRE(Ho == g.Ho && Wo == g.Wo, "conv: output size %dx%d inconsistent with geometry (expected %dx%d)", g.Ho, g.Wo, Ho, Wo);
    return CG_OK;
}

}  // namespace cg

using namespace cg;

extern "C" const char* cg_last_error(void) { return g_err; }

extern "C" int cg_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        set_error("cudaGetDevice: %s", cudaGetErrorString(e));
        return CG_ERR_NO_DEVICE;
    }
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) {
        set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e));
        return CG_ERR_NO_DEVICE;
    }
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return prop.multiProcessorCount;
}

extern "C" int cg_set_tensor_core_mode(int mode) {
    int prev = g_tc_mode;
    g_tc_mode = mode == 1 ? 7 : mode;  // 1 = everything; otherwise a bit mask: 1 forward, 2 data gradient, 4 weight gradient
    return prev;
}

extern "C" uint64_t cg_launch_count(void) { return g_launches.load(); }

extern "C" size_t cg_conv_workspace_bytes(const cg_conv_geom* g, int which) {
    if (!g) return 0;
    ConvDims d = conv_dims(*g);
    size_t need = 0;
    if (which == 1) {
        if (g->ups) need = (size_t)g->G * g->B * d.Hin * d.Win * g->Cin * sizeof(float);
        if ((g_tc_mode & 2) && tc_dgrad_supported(*g)) { size_t t = tc_dgrad_ws(*g); need = t > need ? t : need; }
    }
    if (which == 2) {
        size_t a = ((g_tc_mode & 4) && tc_wgrad_supported(*g)) ? tc_wgrad_ws(*g) : simt_wgrad_ws(*g);
        size_t b = colsum_ws(g->G, d.Mpix, g->Cout);
        need = a > b ? a : b;
    }
    return need;
}

extern "C" int cg_conv_fwd(const cg_conv_geom* g, const float* x, const float* w, const float* bias, float* y, int act,
                           float slope, void* ws, size_t ws_bytes, void* stream) {
    if (int rc = validate_geom(*g)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if ((g_tc_mode & 1) && tc_fwd_supported(*g)) return tc_conv_fwd(*g, x, w, bias, y, act, slope, ws, ws_bytes, st);
    return simt_conv_fwd(*g, x, w, bias, y, act, slope, st);
}

extern "C" int cg_conv_dgrad(const cg_conv_geom* g, const float* dy, const float* w, float* dx, const float* addend,
                             const float* mask_src, float mask_slope, void* ws, size_t ws_bytes, void* stream) {
    if (int rc = validate_geom(*g)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if ((g_tc_mode & 2) && tc_dgrad_supported(*g)) return tc_conv_dgrad(*g, dy, w, dx, addend, mask_src, mask_slope, ws, ws_bytes, st);
    if (!g->ups) return simt_conv_dgrad(*g, dy, w, dx, addend, mask_src, mask_slope, st);
    size_t need = cg_conv_workspace_bytes(g, 1);
    if (need > ws_bytes) {
        set_error("conv_dgrad: workspace %zu < %zu bytes", ws_bytes, need);
        return CG_ERR_WORKSPACE;
    }
    if (int rc = simt_conv_dgrad(*g, dy, w, (float*)ws, nullptr, nullptr, 0.f, st)) return rc;
    return pool2x2_sum((const float*)ws, dx, addend, mask_src, mask_slope, (long)g->G * g->B, g->H, g->W, g->Cin, st);
}

extern "C" int cg_conv_wgrad(const cg_conv_geom* g, const float* x, const float* dy, float* dw, float* db, void* ws,
                             size_t ws_bytes, void* stream) {
    if (int rc = validate_geom(*g)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if ((g_tc_mode & 4) && tc_wgrad_supported(*g)) {
        if (int rc = tc_conv_wgrad(*g, x, dy, dw, ws, ws_bytes, st)) return rc;
    } else {
        if (int rc = simt_conv_wgrad(*g, x, dy, dw, ws, ws_bytes, st)) return rc;
    }
    if (db) {
        ConvDims d = conv_dims(*g);
        return colsum(dy, db, g->G, d.Mpix, g->Cout, ws, ws_bytes, st);
    }
    return CG_OK;
}

extern "C" int cg_upsample2x_bwd(const float* d_up, float* dx, int N, int H, int W, int C, void* stream) {
    CG_REQUIRE(C % 4 == 0, "upsample2x_bwd: C=%d must be a multiple of 4", C);
    return pool2x2_sum(d_up, dx, nullptr, nullptr, 0.f, N, H, W, C, (cudaStream_t)stream);
}
