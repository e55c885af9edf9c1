// Image-space, loss and optimiser kernels (all HBM-bound; coalesced, vectorised where the layout allows).
#include "common.cuh"

namespace cg {

__device__ __forceinline__ float4 f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// block-wide sum of `n` floats per thread; result valid in thread 0
template <int N>
__device__ __forceinline__ void block_reduce(float (&v)[N], float* smem /* >= N*32 */) {
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int i = 0; i < N; i++) smem[i * 32 + warp] = v[i];
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < N; i++) {
            float t = lane < nw ? smem[i * 32 + lane] : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            v[i] = t;
        }
    }
}

// ---- attention-mask head: Decoder_V2_atten.forward networks.py:398-407 -------------------------
// h = tanh output of dec.model.9, 12 lanes: [o0 rgb | o1 rgb | o2 rgb | m0 m1 m2]
// mask_k = (tanh(10*h[9+k])+1)/2;  im <- (1-mask_k)*im + mask_k*o_k, k = 0..2, starting from x_in
__global__ void mask_head_fwd_kernel(const float* __restrict__ h, const float* __restrict__ x_in, float* __restrict__ x_fake,
                                     float* __restrict__ mask, long total, long per_group) {
    pdl_trigger();
    pdl_wait();
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float* hp = h + i * 12;
    float4 a = f4(hp), b = f4(hp + 4), c = f4(hp + 8);
    float o[9] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x};
    float mk[3] = {(tanhf(10.f * c.y) + 1.f) * 0.5f, (tanhf(10.f * c.z) + 1.f) * 0.5f, (tanhf(10.f * c.w) + 1.f) * 0.5f};
    float4 xi = f4(x_in + (i % per_group) * 4);
    float im[3] = {xi.x, xi.y, xi.z};
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
        for (int ch = 0; ch < 3; ch++) im[ch] = (1.f - mk[k]) * im[ch] + mk[k] * o[3 * k + ch];
    reinterpret_cast<float4*>(x_fake)[i] = make_float4(im[0], im[1], im[2], 0.f);
    reinterpret_cast<float4*>(mask)[i] = make_float4(mk[0], mk[1], mk[2], 0.f);
}

__global__ void mask_head_bwd_kernel(const float* __restrict__ h, const float* __restrict__ x_in,
                                     const float* __restrict__ d_xfake, const float* __restrict__ d_mask,
                                     float* __restrict__ dh_pre, long total, long per_group) {
    pdl_trigger();
    pdl_wait();
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float* hp = h + i * 12;
    float4 a = f4(hp), b = f4(hp + 4), c = f4(hp + 8);
    float hv[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
    float tk[3], mk[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        tk[k] = tanhf(10.f * hv[9 + k]);
        mk[k] = (tk[k] + 1.f) * 0.5f;
    }
    float4 xi = f4(x_in + (i % per_group) * 4);
    float im[4][3];
    im[0][0] = xi.x; im[0][1] = xi.y; im[0][2] = xi.z;
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
        for (int ch = 0; ch < 3; ch++) im[k + 1][ch] = (1.f - mk[k]) * im[k][ch] + mk[k] * hv[3 * k + ch];
    float4 dx = f4(d_xfake + i * 4);
    float dim[3] = {dx.x, dx.y, dx.z};
    float dm[3] = {0.f, 0.f, 0.f};
    if (d_mask) {
        float4 d = f4(d_mask + i * 4);
        dm[0] = d.x; dm[1] = d.y; dm[2] = d.z;
    }
    float dh[12];
#pragma unroll
    for (int k = 2; k >= 0; k--) {
        float dmk = dm[k];
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            dh[3 * k + ch] = mk[k] * dim[ch];
            dmk += dim[ch] * (hv[3 * k + ch] - im[k][ch]);
            dim[ch] *= (1.f - mk[k]);
        }
        dh[9 + k] = dmk * 5.f * (1.f - tk[k] * tk[k]);  // d/dh (tanh(10h)+1)/2
    }
    float out[12];
#pragma unroll
    for (int j = 0; j < 12; j++) out[j] = dh[j] * (1.f - hv[j] * hv[j]);  // through the layer's own tanh
    float* op = dh_pre + i * 12;
    *reinterpret_cast<float4*>(op) = make_float4(out[0], out[1], out[2], out[3]);
    *reinterpret_cast<float4*>(op + 4) = make_float4(out[4], out[5], out[6], out[7]);
    *reinterpret_cast<float4*>(op + 8) = make_float4(out[8], out[9], out[10], out[11]);
}

// ---- AvgPool2d(3, stride 2, pad 1, count_include_pad=False): networks.py:32,129 ------------------
__global__ void avgpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long total, int H, int W, int C4) {
    pdl_trigger();
    pdl_wait();
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int Ho = H / 2, Wo = W / 2;
    int c = (int)(i % C4);
    long t = i / C4;
    int ow = (int)(t % Wo);
    t /= Wo;
    int oh = (int)(t % Ho);
    long n = t / Ho;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    int cnt = 0;
    for (int dh = -1; dh <= 1; dh++) {
        int ih = 2 * oh + dh;
        if (ih < 0 || ih >= H) continue;
        for (int dw = -1; dw <= 1; dw++) {
            int iw = 2 * ow + dw;
            if (iw < 0 || iw >= W) continue;
            float4 v = __ldg(reinterpret_cast<const float4*>(x) + ((n * H + ih) * W + iw) * C4 + c);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            cnt++;
        }
    }
    float r = 1.f / (float)cnt;
    reinterpret_cast<float4*>(y)[i] = make_float4(s.x * r, s.y * r, s.z * r, s.w * r);
}

__device__ __forceinline__ int pool_cnt(int o, int L) {  // valid taps of output index o along one axis
    int c = 0;
    for (int d = -1; d <= 1; d++) {
        int i = 2 * o + d;
        c += (i >= 0 && i < L);
    }
    return c;
}
// one thread per input pixel; gathers from the <=2x2 output windows that cover it
__global__ void avgpool_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, long total, int H, int W, int Cy,
                                   int Cx, int nch, int accumulate) {
    pdl_trigger();
    pdl_wait();
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int Ho = H / 2, Wo = W / 2;
    int iw = (int)(i % W);
    long t = i / W;
    int ih = (int)(t % H);
    long n = t / H;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    // outputs oh with |ih - 2*oh| <= 1
    for (int oh = (ih - 1 + 1) / 2; oh <= (ih + 1) / 2; oh++) {
        if (oh < 0 || oh >= Ho || abs(ih - 2 * oh) > 1) continue;
        int ch = pool_cnt(oh, H);
        for (int ow = iw / 2; ow <= (iw + 1) / 2; ow++) {
            if (ow < 0 || ow >= Wo || abs(iw - 2 * ow) > 1) continue;
            float r = 1.f / (float)(ch * pool_cnt(ow, W));
            const float* p = dy + ((n * Ho + oh) * Wo + ow) * Cy;
            for (int c = 0; c < nch; c++) acc[c] += __ldg(p + c) * r;
        }
    }
    float* o = dx + i * Cx;
    for (int c = 0; c < nch; c++) o[c] = accumulate ? o[c] + acc[c] : acc[c];
}

__global__ void acc_slice_kernel(float* __restrict__ dst, const float* __restrict__ src, long npix, int Cd, int Cs, int nch) {
    pdl_trigger();
    pdl_wait();
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    for (int c = 0; c < nch; c++) dst[i * Cd + c] += __ldg(src + i * Cs + c);
}

__global__ void gather_images_kernel(const float* __restrict__ pool0, int n0, const float* __restrict__ pool1,
                                     const int32_t* __restrict__ idx, const float* __restrict__ x_in, float* __restrict__ y,
                                     long total, int Bt, int B, int HW) {
    pdl_trigger();
    pdl_wait();
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int pix = (int)(i % HW);
    long gn = i / HW;  // g*Bt + n
    int n = (int)(gn % Bt);
    int slot = __ldg(idx + gn);
    float4 v = slot < n0 ? f4(pool0 + ((long)slot * HW + pix) * 4) : f4(pool1 + ((long)(slot - n0) * HW + pix) * 4);
    if (x_in) {
        float4 u = f4(x_in + ((long)(n % B) * HW + pix) * 4);
        reinterpret_cast<float4*>(y)[i * 2] = v;
        reinterpret_cast<float4*>(y)[i * 2 + 1] = u;
    } else {
        reinterpret_cast<float4*>(y)[i] = v;
    }
}

__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, long total, int C, int HW, int Cp) {
    pdl_trigger();
    pdl_wait();
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;  // over N*HW
    if (i >= total) return;
    long n = i / HW;
    int pix = (int)(i - n * HW);
    for (int c = 0; c < Cp; c++) y[i * Cp + c] = c < C ? __ldg(x + (n * C + c) * HW + pix) : 0.f;
}
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ x, float* __restrict__ y, long total, int C, int HW, int Cp) {
    pdl_trigger();
    pdl_wait();
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;  // over N*HW
    if (i >= total) return;
    long n = i / HW;
    int pix = (int)(i - n * HW);
    for (int c = 0; c < C; c++) y[(n * C + c) * HW + pix] = __ldg(x + i * Cp + c);
}

// ---- LSGAN: networks.py:64,90,166,194 -----------------------------------------------------------
__global__ void __launch_bounds__(256) lsgan_fwd_kernel(const float* __restrict__ out, const float* __restrict__ targets,
                                                        const float* __restrict__ weights, float* __restrict__ sums,
                                                        float* __restrict__ loss, int nseg, int n_per_seg, int accumulate) {
    pdl_trigger();
    pdl_wait();
    __shared__ float sm[32];
    const int g = blockIdx.x;
    float total = 0.f;
    for (int seg = 0; seg < nseg; seg++) {
        float t = __ldg(targets + seg);
        const float* p = out + ((long)g * nseg + seg) * n_per_seg;
        float v[1] = {0.f};
        for (int i = threadIdx.x; i < n_per_seg; i += blockDim.x) {
            float d = __ldg(p + i) - t;
            v[0] += d * d;
        }
        block_reduce<1>(v, sm);
        if (threadIdx.x == 0) {
            sums[g * nseg + seg] = v[0];
            total += __ldg(weights + g * nseg + seg) * (v[0] / (float)n_per_seg);
        }
    }
    if (threadIdx.x == 0) loss[g] = accumulate ? loss[g] + total : total;
}
__global__ void lsgan_bwd_kernel(const float* __restrict__ out, const float* __restrict__ targets, const float* __restrict__ coef,
                                 float* __restrict__ dout, long total, int nseg, int n_per_seg) {
    pdl_trigger();
    pdl_wait();
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int gs = (int)(i / n_per_seg);
    dout[i] = __ldg(coef + gs) * (__ldg(out + i) - __ldg(targets + gs % nseg));
}

// ---- focus losses: trainer_council.py:230-250 -----------------------------------------------------
constexpr int FC_PIX = 2048;  // pixels per block
__global__ void __launch_bounds__(256) focus_fwd_kernel(const float* __restrict__ mask, float* __restrict__ part, int B, int H,
                                                        int W, float center, float eps) {
    pdl_trigger();
    pdl_wait();
    __shared__ float sm[4 * 32];
    const int g = blockIdx.y;
    const long npix = (long)B * H * W;
    const long p0 = (long)blockIdx.x * FC_PIX, p1 = min(npix, p0 + FC_PIX);
    const float* mb = mask + (long)g * npix * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    for (long px = p0 + threadIdx.x; px < p1; px += blockDim.x) {
        int w = (int)(px % W);
        int h = (int)((px / W) % H);
        float4 m = f4(mb + px * 4);
        v[0] += 1.f / (fabsf(m.x - center) + eps) + 1.f / (fabsf(m.y - center) + eps) + 1.f / (fabsf(m.z - center) + eps);
        v[1] += m.x + m.y + m.z;
        if (h + 1 < H) {
            float4 d = f4(mb + (px + W) * 4);
            v[2] += fabsf(d.x - m.x) + fabsf(d.y - m.y) + fabsf(d.z - m.z);
        }
        if (w + 1 < W) {
            float4 r = f4(mb + (px + 1) * 4);
            v[3] += fabsf(r.x - m.x) + fabsf(r.y - m.y) + fabsf(r.z - m.z);
        }
    }
    block_reduce<4>(v, sm);
    if (threadIdx.x == 0) {
        float* o = part + ((long)blockIdx.x * gridDim.y + g) * 4;
        o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; o[3] = v[3];
    }
}
__global__ void focus_final_kernel(const float* __restrict__ part, float* __restrict__ sums, int G4, int nchunks) {
    pdl_trigger();
    pdl_wait();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G4) return;
    double s = 0.0;
    for (int k = 0; k < nchunks; k++) s += (double)part[(long)k * G4 + i];
    sums[i] = (float)s;
}
__device__ __forceinline__ float sgn(float x) { return (x > 0.f) - (x < 0.f); }
__global__ void focus_bwd_kernel(const float* __restrict__ mask, const float* __restrict__ coef, float* __restrict__ dmask,
                                 long total, long npix, int H, int W, float center, float eps) {
    pdl_trigger();
    pdl_wait();
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;  // over G*npix
    if (i >= total) return;
    int g = (int)(i / npix);
    long px = i - (long)g * npix;
    int w = (int)(px % W);
    int h = (int)((px / W) % H);
    float c01 = __ldg(coef + g * 3), cs = __ldg(coef + g * 3 + 1), ctv = __ldg(coef + g * 3 + 2);
    const float* mp = mask + i * 4;
    float4 m = f4(mp);
    float mm[3] = {m.x, m.y, m.z}, o[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        float d = mm[c] - center;
        float den = fabsf(d) + eps;
        o[c] = cs - c01 * sgn(d) / (den * den);
    }
    if (ctv != 0.f) {
        float tv[3] = {0.f, 0.f, 0.f};
        if (h + 1 < H) { float4 q = f4(mp + (long)W * 4); tv[0] -= sgn(q.x - m.x); tv[1] -= sgn(q.y - m.y); tv[2] -= sgn(q.z - m.z); }
        if (h > 0)     { float4 q = f4(mp - (long)W * 4); tv[0] += sgn(m.x - q.x); tv[1] += sgn(m.y - q.y); tv[2] += sgn(m.z - q.z); }
        if (w + 1 < W) { float4 q = f4(mp + 4);           tv[0] -= sgn(q.x - m.x); tv[1] -= sgn(q.y - m.y); tv[2] -= sgn(q.z - m.z); }
        if (w > 0)     { float4 q = f4(mp - 4);           tv[0] += sgn(m.x - q.x); tv[1] += sgn(m.y - q.y); tv[2] += sgn(m.z - q.z); }
#pragma unroll
        for (int c = 0; c < 3; c++) o[c] += ctv * tv[c];
    }
    reinterpret_cast<float4*>(dmask)[i] = make_float4(o[0], o[1], o[2], 0.f);
}

// ---- Adam: torch.optim.Adam (trainer_council.py:170-179): L2 decay into the gradient, eps 1e-8 ----
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long n, float lr_c1, float b1, float b2, float eps, float wd, float rsq_c2, float gscale) {
    pdl_trigger();
    pdl_wait();
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long stride = (long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        float pv = p[i];
        float gv = __ldg(g + i) * gscale + wd * pv;
        float mv = b1 * m[i] + (1.f - b1) * gv;
        float vv = b2 * v[i] + (1.f - b2) * gv * gv;
        m[i] = mv;
        v[i] = vv;
        float denom = sqrtf(vv) * rsq_c2 + eps;
        p[i] = pv - lr_c1 * (mv / denom);
    }
}

}  // namespace cg

using namespace cg;
#define ST ((cudaStream_t)stream)

extern "C" int cg_mask_head_fwd(const float* h, const float* x_in, float* x_fake, float* mask, int G, int B, int HW, void* stream) {
    long per = (long)B * HW, total = per * G;
    launch_k(mask_head_fwd_kernel, cdiv(total, 256), 256, 0, ST, h, x_in, x_fake, mask, total, per);
    return check_launch("mask_head_fwd");
}
extern "C" int cg_mask_head_bwd(const float* h, const float* x_in, const float* d_xfake, const float* d_mask, float* dh_pre,
                                int G, int B, int HW, void* stream) {
    long per = (long)B * HW, total = per * G;
    launch_k(mask_head_bwd_kernel, cdiv(total, 256), 256, 0, ST, h, x_in, d_xfake, d_mask, dh_pre, total, per);
    return check_launch("mask_head_bwd");
}
extern "C" int cg_avgpool_fwd(const float* x, float* y, int N, int H, int W, int C, void* stream) {
    CG_REQUIRE(C % 4 == 0 && H % 2 == 0 && W % 2 == 0, "avgpool_fwd: C=%d H=%d W=%d unsupported", C, H, W);
    long total = (long)N * (H / 2) * (W / 2) * (C / 4);
    launch_k(avgpool_fwd_kernel, cdiv(total, 256), 256, 0, ST, x, y, total, H, W, C / 4);
    return check_launch("avgpool_fwd");
}
extern "C" int cg_avgpool_bwd(const float* dy, float* dx, int N, int H, int W, int Cy, int Cx, int nch, int accumulate, void* stream) {
    CG_REQUIRE(nch <= 4 && nch <= Cy && nch <= Cx && H % 2 == 0 && W % 2 == 0, "avgpool_bwd: bad lanes/sizes");
    long total = (long)N * H * W;
    launch_k(avgpool_bwd_kernel, cdiv(total, 256), 256, 0, ST, dy, dx, total, H, W, Cy, Cx, nch, accumulate);
    return check_launch("avgpool_bwd");
}
extern "C" int cg_acc_slice(float* dst, const float* src, long npix, int Cd, int Cs, int nch, void* stream) {
    launch_k(acc_slice_kernel, cdiv(npix, 256), 256, 0, ST, dst, src, npix, Cd, Cs, nch);
    return check_launch("acc_slice");
}
extern "C" int cg_gather_images(const float* pool0, int n0, const float* pool1, const int32_t* idx, const float* x_in, float* y,
                                int G, int Bt, int B, int HW, void* stream) {
    long total = (long)G * Bt * HW;
    launch_k(gather_images_kernel, cdiv(total, 256), 256, 0, ST, pool0, n0, pool1, idx, x_in, y, total, Bt, B, HW);
    return check_launch("gather_images");
}
extern "C" int cg_nchw_to_nhwc(const float* x, float* y, int N, int C, int HW, int Cp, void* stream) {
    long total = (long)N * HW;
    launch_k(nchw_to_nhwc_kernel, cdiv(total, 256), 256, 0, ST, x, y, total, C, HW, Cp);
    return check_launch("nchw_to_nhwc");
}
extern "C" int cg_nhwc_to_nchw(const float* x, float* y, int N, int C, int HW, int Cp, void* stream) {
    long total = (long)N * HW;
    launch_k(nhwc_to_nchw_kernel, cdiv(total, 256), 256, 0, ST, x, y, total, C, HW, Cp);
    return check_launch("nhwc_to_nchw");
}
extern "C" int cg_lsgan_fwd(const float* out, const float* targets, const float* weights, float* sums, float* loss, int G,
                            int nseg, int n_per_seg, int accumulate, void* stream) {
    launch_k(lsgan_fwd_kernel, G, 256, 0, ST, out, targets, weights, sums, loss, nseg, n_per_seg, accumulate);
    return check_launch("lsgan_fwd");
}
extern "C" int cg_lsgan_bwd(const float* out, const float* targets, const float* coef, float* dout, int G, int nseg, int n_per_seg,
                            void* stream) {
    long total = (long)G * nseg * n_per_seg;
    launch_k(lsgan_bwd_kernel, cdiv(total, 256), 256, 0, ST, out, targets, coef, dout, total, nseg, n_per_seg);
    return check_launch("lsgan_bwd");
}
extern "C" int cg_focus_fwd(const float* mask, float* sums, int G, int B, int H, int W, float center, float eps, void* ws,
                            size_t ws_bytes, void* stream) {
    long npix = (long)B * H * W;
    int nchunks = cdiv(npix, FC_PIX);
    size_t need = (size_t)nchunks * G * 4 * sizeof(float);
    if (need > ws_bytes) {
        set_error("focus_fwd: workspace %zu < %zu bytes", ws_bytes, need);
        return CG_ERR_WORKSPACE;
    }
    launch_k(focus_fwd_kernel, dim3(nchunks, G), 256, 0, ST, mask, (float*)ws, B, H, W, center, eps);
    if (int rc = check_launch("focus_fwd")) return rc;
    launch_k(focus_final_kernel, cdiv(G * 4, 64), 64, 0, ST, (const float*)ws, sums, G * 4, nchunks);
    return check_launch("focus_final");
}
extern "C" int cg_focus_bwd(const float* mask, const float* coef, float* dmask, int G, int B, int H, int W, float center,
                            float eps, void* stream) {
    long npix = (long)B * H * W, total = npix * G;
    launch_k(focus_bwd_kernel, cdiv(total, 256), 256, 0, ST, mask, coef, dmask, total, npix, H, W, center, eps);
    return check_launch("focus_bwd");
}
extern "C" int cg_adam_step(float* p, const float* g, float* m, float* v, long n, float lr, float beta1, float beta2, float eps,
                            float weight_decay, int step, float grad_scale, void* stream) {
    double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
    float lr_c1 = (float)((double)lr / bc1), rsq_c2 = (float)(1.0 / sqrt(bc2));
    int blocks = cdiv(n, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    launch_k(adam_kernel, blocks, 256, 0, ST, p, g, m, v, n, lr_c1, beta1, beta2, eps, weight_decay, rsq_c2, grad_scale);
    return check_launch("adam");
}
