// Host build of the device-side transform arithmetic (council_gan_b200/csrc/augment_math.cuh): the SAME per-pixel functions the CUDA
// kernels call, driven by plain loops with the kernels' indexing, behind the C ABI's argument lists.  Test infrastructure
// (tests/test_augment_cpu.py compiles it with g++ -ffp-contract=off); lets the CPU suite check the kernels' arithmetic bit-for-bit.
#include "../council_gan_b200/csrc/augment_math.cuh"
using namespace cg;

extern "C" int h_aug_color(uint8_t* imgs, const int32_t* desc, const int32_t* opcode, const float* param, int B) {
    for (int b = 0; b < B; b++) {
        const int op = opcode[b];
        if (op == CG_AUG_NONE) continue;
        const int npix = desc[b * 4 + 1] * desc[b * 4 + 2];
        uint8_t* p = imgs + desc[b * 4];
        unsigned long long lsum = 0;
        if (op == CG_AUG_CONTRAST)
            for (int i = 0; i < npix; i++) lsum += (unsigned)lum(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
        const int mean = op == CG_AUG_CONTRAST ? contrast_mean(lsum, npix) : 0;
        for (int i = 0; i < npix; i++) {
            int r = p[3 * i], g = p[3 * i + 1], bb = p[3 * i + 2];
            color_px(op, param[b], mean, r, g, bb);
            p[3 * i] = (uint8_t)r; p[3 * i + 1] = (uint8_t)g; p[3 * i + 2] = (uint8_t)bb;
        }
    }
    return 0;
}

extern "C" int h_aug_resize_crop(const uint8_t* imgs, const int32_t* src_off, const int32_t* flip, const int32_t* slot, const int32_t* crop,
                                 int n_img, int H, int W, int oh, int ow, int ch, int cw, const int32_t* bounds_h, const int32_t* kk_h,
                                 int ksize_h, const int32_t* bounds_v, const int32_t* kk_v, int ksize_v, uint8_t* tmp, float* out_nhwc,
                                 float* out_nchw) {
    if (ch > oh || cw > ow) return 1;
    for (int n = 0; n < n_img; n++) {
        const uint8_t* src = imgs + src_off[n];
        uint8_t* dst = tmp + (long)n * H * ow * 3;
        for (int i = 0; i < H * ow; i++) {
            const int y = i / ow, xx = i - y * ow;
            resize_h_px(src + (long)y * W * 3, W, flip[n] != 0, bounds_h, kk_h, ksize_h, xx, dst + 3 * (long)i);
        }
        const int total = ch * cw;
        const long sl = slot[n];
        for (int i = 0; i < total; i++) {
            const int y = i / cw, x = i - y * cw;
            float v[3];
            resize_v_px(dst, ow, bounds_v, kk_v, ksize_v, crop[2 * n] + y, crop[2 * n + 1] + x, v);
            float* o4 = out_nhwc + (sl * total + i) * 4;
            o4[0] = v[0]; o4[1] = v[1]; o4[2] = v[2]; o4[3] = 0.f;
            if (out_nchw) {
                float* o = out_nchw + sl * 3 * total + i;
                o[0] = v[0]; o[total] = v[1]; o[2 * total] = v[2];
            }
        }
    }
    return 0;
}

// every colour through the hue path (rgb -> hsv -> shift -> rgb), for the exhaustive check
extern "C" void h_hue_all(const uint8_t* rgb, uint8_t* out, long n, int shift) {
    for (long i = 0; i < n; i++) {
        int r = rgb[3 * i], g = rgb[3 * i + 1], b = rgb[3 * i + 2];
        color_px(CG_AUG_HUE, (float)shift, 0, r, g, b);
        out[3 * i] = (uint8_t)r; out[3 * i + 1] = (uint8_t)g; out[3 * i + 2] = (uint8_t)b;
    }
}
