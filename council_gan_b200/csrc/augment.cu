// Device-side training-time image transforms: what the reference's data loader does per image on the CPU with
// torchvision + Pillow (utils.py:122-181: RandomGrayscale -> ColorJitter -> RandomHorizontalFlip -> Resize -> RandomCrop ->
// ToTensor -> Normalize), restated for a batch of decoded uint8 RGB images already in device memory.  Bit-exact with Pillow
// 12 / torchvision 0.26 for every op (oracle/augment_oracle.py is the CPU restatement both sides are checked against):
//   * L = (R*19595 + G*38470 + B*7471 + 0x8000) >> 16                           (Convert.c)
//   * ImageEnhance = Image.blend(degenerate, image, f): float32 `in1 + f * (in2 - in1)`, truncated / clipped   (Blend.c)
//   * hue through Pillow's float RGB<->HSV conversion, including its float/double promotions                  (Convert.c)
//   * bilinear resize = separable triangle filter in 22-bit fixed point, uint8 between the passes (Resample.c)
// The output is written directly in the layout the training step consumes: channels-last fp32 [B][H][W][4], values in [-1, 1].
#include "common.cuh"

namespace cg {

__device__ __forceinline__ int lum(int r, int g, int b) { return (r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16; }

// Image.blend for one channel: never let the compiler contract the multiply-add (Pillow's C code is not built with FMA)
__device__ __forceinline__ int blend1(int in1, int in2, float a, bool interp) {
    float temp = __fadd_rn((float)in1, __fmul_rn(a, (float)(in2 - in1)));
    if (interp) return (int)temp;
    return temp <= 0.f ? 0 : (temp >= 255.f ? 255 : (int)temp);
}

__device__ __forceinline__ void rgb2hsv(int r, int g, int b, int& uh, int& us, int& uv) {
    int maxc = max(r, max(g, b)), minc = min(r, min(g, b));
    uv = maxc;
    if (minc == maxc) { uh = 0; us = 0; return; }
    float cr = (float)(maxc - minc);
    float s = __fdiv_rn(cr, (float)maxc);
    float rc = __fdiv_rn((float)(maxc - r), cr), gc = __fdiv_rn((float)(maxc - g), cr), bc = __fdiv_rn((float)(maxc - b), cr);
    float h;
    if (r == maxc) h = __fsub_rn(bc, gc);
    else if (g == maxc) h = (float)(2.0 + (double)rc - (double)bc);
    else h = (float)(4.0 + (double)gc - (double)rc);
    h = (float)fmod((double)h / 6.0 + 1.0, 1.0);
    int ih = (int)((double)h * 255.0), is = (int)((double)s * 255.0);
    uh = min(max(ih, 0), 255);
    us = min(max(is, 0), 255);
}
__device__ __forceinline__ int round8(float x) { return min(max((int)floor((double)x + 0.5), 0), 255); }
__device__ __forceinline__ void hsv2rgb(int h, int s, int v, int& r, int& g, int& b) {
    if (s == 0) { r = g = b = v; return; }
    float fh = __fdiv_rn(__fmul_rn((float)h, 6.0f), 255.0f);
    float fs = __fdiv_rn((float)s, 255.0f);
    int i = (int)floorf(fh);
    float f = __fsub_rn(fh, (float)i);
    float vv = (float)v;
    int p = round8(__fmul_rn(vv, __fsub_rn(1.0f, fs)));
    int q = round8(__fmul_rn(vv, __fsub_rn(1.0f, __fmul_rn(fs, f))));
    int t = round8(__fmul_rn(vv, __fsub_rn(1.0f, __fmul_rn(fs, __fsub_rn(1.0f, f)))));
    switch (i % 6) {
        case 0: r = v; g = t; b = p; break;
        case 1: r = q; g = v; b = p; break;
        case 2: r = p; g = v; b = t; break;
        case 3: r = p; g = q; b = v; break;
        case 4: r = t; g = p; b = v; break;
        default: r = v; g = p; b = q; break;
    }
}

// per-image L sums (exact integers: order-independent atomics) for the images whose current op is the contrast adjustment
__global__ void __launch_bounds__(256) aug_lsum_kernel(const uint8_t* __restrict__ imgs, const int32_t* __restrict__ desc,
                                                       const int32_t* __restrict__ opcode, unsigned long long* __restrict__ lsum) {
    const int b = blockIdx.y;
    if (opcode[b] != CG_AUG_CONTRAST) return;
    const int npix = desc[b * 4 + 1] * desc[b * 4 + 2];
    const uint8_t* p = imgs + desc[b * 4];
    unsigned long long s = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) s += (unsigned)lum(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(lsum + b, s);
}

__global__ void __launch_bounds__(256) aug_color_kernel(uint8_t* __restrict__ imgs, const int32_t* __restrict__ desc,
                                                        const int32_t* __restrict__ opcode, const float* __restrict__ param,
                                                        const unsigned long long* __restrict__ lsum) {
    const int b = blockIdx.y;
    const int op = opcode[b];
    if (op == CG_AUG_NONE) return;
    const int npix = desc[b * 4 + 1] * desc[b * 4 + 2];
    uint8_t* p = imgs + desc[b * 4];
    const float f = param[b];
    const bool interp = f >= 0.f && f <= 1.f;
    int mean = 0;
    if (op == CG_AUG_CONTRAST) mean = (int)((double)lsum[b] / (double)npix + 0.5);  // int(ImageStat.Stat(L).mean[0] + 0.5)
    const int hshift = op == CG_AUG_HUE ? (int)(f * 255.f) : 0;                     // np.int32(hue_factor * 255): truncation
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
        int r = p[3 * i], g = p[3 * i + 1], bb = p[3 * i + 2];
        if (op == CG_AUG_GRAY) {
            r = g = bb = lum(r, g, bb);
        } else if (op == CG_AUG_BRIGHTNESS) {
            if (f != 1.f) { r = f == 0.f ? 0 : blend1(0, r, f, interp); g = f == 0.f ? 0 : blend1(0, g, f, interp); bb = f == 0.f ? 0 : blend1(0, bb, f, interp); }
        } else if (op == CG_AUG_CONTRAST) {
            if (f != 1.f) { r = f == 0.f ? mean : blend1(mean, r, f, interp); g = f == 0.f ? mean : blend1(mean, g, f, interp); bb = f == 0.f ? mean : blend1(mean, bb, f, interp); }
        } else if (op == CG_AUG_SATURATION) {
            const int l = lum(r, g, bb);
            if (f != 1.f) { r = f == 0.f ? l : blend1(l, r, f, interp); g = f == 0.f ? l : blend1(l, g, f, interp); bb = f == 0.f ? l : blend1(l, bb, f, interp); }
        } else if (op == CG_AUG_HUE) {
            int h, s, v;
            rgb2hsv(r, g, bb, h, s, v);
            h = (h + hshift) & 255;  // uint8 wrap-around
            hsv2rgb(h, s, v, r, g, bb);
        }
        p[3 * i] = (uint8_t)r; p[3 * i + 1] = (uint8_t)g; p[3 * i + 2] = (uint8_t)bb;
    }
}

// ---- Pillow's two-pass resize for n images of one source size -----------------------------------------------------------------
constexpr int RS_BITS = 32 - 8 - 2;
__device__ __forceinline__ uint8_t rs_clip8(long long v) {
    long long r = v >> RS_BITS;
    return (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
}
// horizontal pass (+ optional mirror of the source columns: RandomHorizontalFlip precedes Resize): tmp[n][H][ow][3]
__global__ void __launch_bounds__(256) aug_resize_h_kernel(const uint8_t* __restrict__ imgs, const int32_t* __restrict__ src_off,
                                                           const int32_t* __restrict__ flip, uint8_t* __restrict__ tmp,
                                                           const int32_t* __restrict__ bounds, const int32_t* __restrict__ kk, int ksize,
                                                           int H, int W, int ow) {
    const int n = blockIdx.y;
    const int total = H * ow;
    const uint8_t* src = imgs + src_off[n];
    const bool fl = flip[n] != 0;
    uint8_t* dst = tmp + (long)n * total * 3;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int y = i / ow, xx = i - y * ow;
        const int x0 = bounds[2 * xx], cnt = bounds[2 * xx + 1];
        long long a0 = 1ll << (RS_BITS - 1), a1 = a0, a2 = a0;
        for (int x = 0; x < cnt; x++) {
            const int sx = fl ? W - 1 - (x0 + x) : x0 + x;
            const uint8_t* px = src + ((long)y * W + sx) * 3;
            const long long k = kk[xx * ksize + x];
            a0 += px[0] * k; a1 += px[1] * k; a2 += px[2] * k;
        }
        dst[3 * i] = rs_clip8(a0); dst[3 * i + 1] = rs_clip8(a1); dst[3 * i + 2] = rs_clip8(a2);
    }
}
// vertical pass fused with RandomCrop / CenterCrop, ToTensor (x / 255) and Normalize((x - 0.5) / 0.5): only the cropped window of the
// resized image is ever produced.  out[slot][ch][cw][4] fp32 channels-last, lane 3 = 0; optionally also NCHW [slot][3][ch][cw].
__global__ void __launch_bounds__(256) aug_resize_v_crop_kernel(const uint8_t* __restrict__ tmp, const int32_t* __restrict__ slot,
                                                                const int32_t* __restrict__ crop, float* __restrict__ out_nhwc,
                                                                float* __restrict__ out_nchw, const int32_t* __restrict__ bounds,
                                                                const int32_t* __restrict__ kk, int ksize, int H, int ow, int ch, int cw) {
    const int n = blockIdx.y;
    const int total = ch * cw;
    const uint8_t* src = tmp + (long)n * H * ow * 3;
    const int ci = crop[2 * n], cj = crop[2 * n + 1];
    const long sl = slot[n];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int y = i / cw, x = i - y * cw;
        const int yy = ci + y, xx = cj + x;
        const int y0 = bounds[2 * yy], cnt = bounds[2 * yy + 1];
        long long a0 = 1ll << (RS_BITS - 1), a1 = a0, a2 = a0;
        for (int t = 0; t < cnt; t++) {
            const uint8_t* px = src + ((long)(y0 + t) * ow + xx) * 3;
            const long long k = kk[yy * ksize + t];
            a0 += px[0] * k; a1 += px[1] * k; a2 += px[2] * k;
        }
        float v[3] = {(float)rs_clip8(a0), (float)rs_clip8(a1), (float)rs_clip8(a2)};
#pragma unroll
        for (int c = 0; c < 3; c++) v[c] = __fdiv_rn(__fsub_rn(__fdiv_rn(v[c], 255.0f), 0.5f), 0.5f);
        reinterpret_cast<float4*>(out_nhwc)[sl * total + i] = make_float4(v[0], v[1], v[2], 0.f);
        if (out_nchw) {
            float* o = out_nchw + sl * 3 * total + i;
            o[0] = v[0]; o[total] = v[1]; o[2 * total] = v[2];
        }
    }
}

}  // namespace cg

using namespace cg;

extern "C" int cg_aug_color(uint8_t* imgs, const int32_t* desc, const int32_t* opcode, const float* param, unsigned long long* lsum, int B,
                            int max_pixels, int any_contrast, void* stream) {
    CG_REQUIRE(B >= 1 && max_pixels >= 1, "aug_color: empty batch");
    cudaStream_t st = (cudaStream_t)stream;
    int blocks = cdiv(max_pixels, 256 * 4);
    if (blocks > 64) blocks = 64;
    if (any_contrast) {
        cudaError_t e = cudaMemsetAsync(lsum, 0, (size_t)B * sizeof(unsigned long long), st);
        if (e != cudaSuccess) {
            set_error("aug_color: %s", cudaGetErrorString(e));
            return CG_ERR_CUDA;
        }
        aug_lsum_kernel<<<dim3(blocks, B), 256, 0, st>>>(imgs, desc, opcode, lsum);
        if (int rc = check_launch("aug_lsum")) return rc;
    }
    aug_color_kernel<<<dim3(blocks, B), 256, 0, st>>>(imgs, desc, opcode, param, lsum);
    return check_launch("aug_color");
}

extern "C" int cg_aug_resize_crop(const uint8_t* imgs, const int32_t* src_off, const int32_t* flip, const int32_t* slot, const int32_t* crop,
                                  int n, int H, int W, int oh, int ow, int ch, int cw, const int32_t* bounds_h, const int32_t* kk_h, int ksize_h,
                                  const int32_t* bounds_v, const int32_t* kk_v, int ksize_v, uint8_t* tmp, float* out_nhwc, float* out_nchw,
                                  void* stream) {
    CG_REQUIRE(n >= 1 && H >= 1 && W >= 1 && ch <= oh && cw <= ow, "aug_resize_crop: bad geometry (%d x %d -> %d x %d, crop %d x %d)", H, W, oh, ow, ch, cw);
    cudaStream_t st = (cudaStream_t)stream;
    int b1 = cdiv((long)H * ow, 256);
    if (b1 > 256) b1 = 256;
    aug_resize_h_kernel<<<dim3(b1, n), 256, 0, st>>>(imgs, src_off, flip, tmp, bounds_h, kk_h, ksize_h, H, W, ow);
    if (int rc = check_launch("aug_resize_h")) return rc;
    int b2 = cdiv((long)ch * cw, 256);
    if (b2 > 256) b2 = 256;
    aug_resize_v_crop_kernel<<<dim3(b2, n), 256, 0, st>>>(tmp, slot, crop, out_nhwc, out_nchw, bounds_v, kk_v, ksize_v, H, ow, ch, cw);
    return check_launch("aug_resize_v_crop");
}
