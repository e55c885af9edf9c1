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
#include "augment_math.cuh"

namespace cg {

// per-image L sums (exact integers: order-independent atomics) for the images whose current op is the contrast adjustment
__global__ void __launch_bounds__(256) aug_lsum_kernel(const uint8_t* __restrict__ imgs, const int32_t* __restrict__ desc,
                                                       const int32_t* __restrict__ opcode, unsigned long long* __restrict__ lsum) {
    pdl_trigger();
    pdl_wait();
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
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.y;
    const int op = opcode[b];
    if (op == CG_AUG_NONE) return;
    const int npix = desc[b * 4 + 1] * desc[b * 4 + 2];
    uint8_t* p = imgs + desc[b * 4];
    const float f = param[b];
    const int mean = op == CG_AUG_CONTRAST ? contrast_mean(lsum[b], npix) : 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
        int r = p[3 * i], g = p[3 * i + 1], bb = p[3 * i + 2];
        color_px(op, f, mean, r, g, bb);
        p[3 * i] = (uint8_t)r; p[3 * i + 1] = (uint8_t)g; p[3 * i + 2] = (uint8_t)bb;
    }
}

// ---- Pillow's two-pass resize for n images of one source size -----------------------------------------------------------------
// horizontal pass (+ optional mirror of the source columns: RandomHorizontalFlip precedes Resize): tmp[n][H][ow][3]
__global__ void __launch_bounds__(256) aug_resize_h_kernel(const uint8_t* __restrict__ imgs, const int32_t* __restrict__ src_off,
                                                           const int32_t* __restrict__ flip, uint8_t* __restrict__ tmp,
                                                           const int32_t* __restrict__ bounds, const int32_t* __restrict__ kk, int ksize,
                                                           int H, int W, int ow) {
    pdl_trigger();
    pdl_wait();
    const int n = blockIdx.y;
    const int total = H * ow;
    const uint8_t* src = imgs + src_off[n];
    const bool fl = flip[n] != 0;
    uint8_t* dst = tmp + (long)n * total * 3;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int y = i / ow, xx = i - y * ow;
        resize_h_px(src + (long)y * W * 3, W, fl, bounds, kk, ksize, xx, dst + 3 * (long)i);
    }
}
// vertical pass fused with RandomCrop / CenterCrop, ToTensor (x / 255) and Normalize((x - 0.5) / 0.5): only the cropped window of the
// resized image is ever produced.  out[slot][ch][cw][4] fp32 channels-last, lane 3 = 0; optionally also NCHW [slot][3][ch][cw].
__global__ void __launch_bounds__(256) aug_resize_v_crop_kernel(const uint8_t* __restrict__ tmp, const int32_t* __restrict__ slot,
                                                                const int32_t* __restrict__ crop, float* __restrict__ out_nhwc,
                                                                float* __restrict__ out_nchw, const int32_t* __restrict__ bounds,
                                                                const int32_t* __restrict__ kk, int ksize, int H, int ow, int ch, int cw) {
    pdl_trigger();
    pdl_wait();
    const int n = blockIdx.y;
    const int total = ch * cw;
    const uint8_t* src = tmp + (long)n * H * ow * 3;
    const int ci = crop[2 * n], cj = crop[2 * n + 1];
    const long sl = slot[n];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int y = i / cw, x = i - y * cw;
        float v[3];
        resize_v_px(src, ow, bounds, kk, ksize, ci + y, cj + x, v);
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
        launch_k(aug_lsum_kernel, dim3(blocks, B), 256, 0, st, imgs, desc, opcode, lsum);
        if (int rc = check_launch("aug_lsum")) return rc;
    }
    launch_k(aug_color_kernel, dim3(blocks, B), 256, 0, st, imgs, desc, opcode, param, lsum);
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
    launch_k(aug_resize_h_kernel, dim3(b1, n), 256, 0, st, imgs, src_off, flip, tmp, bounds_h, kk_h, ksize_h, H, W, ow);
    if (int rc = check_launch("aug_resize_h")) return rc;
    int b2 = cdiv((long)ch * cw, 256);
    if (b2 > 256) b2 = 256;
    launch_k(aug_resize_v_crop_kernel, dim3(b2, n), 256, 0, st, tmp, slot, crop, out_nhwc, out_nchw, bounds_v, kk_v, ksize_v, H, ow, ch, cw);
    return check_launch("aug_resize_v_crop");
}
