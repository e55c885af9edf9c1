// Per-pixel arithmetic of the device-side image transforms (augment.cu), written so that the SAME source also compiles as host C++:
// tests/test_augment_cpu.py builds it with g++ and checks every function bit-for-bit against oracle/augment_oracle.py (itself pinned
// to Pillow / torchvision), so the only thing left to the GPU tests is indexing.  Every float operation is an explicit
// round-to-nearest intrinsic on the device (no FMA contraction: Pillow's C code is built without it).
#pragma once
#include <math.h>
#include <stdint.h>
#include "../../include/council_b200.h"
#ifdef __CUDACC__
#define CG_HD __host__ __device__ __forceinline__
#else
#define CG_HD static inline
#endif
namespace cg {
#ifdef __CUDA_ARCH__
CG_HD float f_add(float a, float b) { return __fadd_rn(a, b); }
CG_HD float f_sub(float a, float b) { return __fsub_rn(a, b); }
CG_HD float f_mul(float a, float b) { return __fmul_rn(a, b); }
CG_HD float f_div(float a, float b) { return __fdiv_rn(a, b); }
#else  // host build: compiled with -ffp-contract=off
CG_HD float f_add(float a, float b) { volatile float r = a + b; return r; }
CG_HD float f_sub(float a, float b) { volatile float r = a - b; return r; }
CG_HD float f_mul(float a, float b) { volatile float r = a * b; return r; }
CG_HD float f_div(float a, float b) { volatile float r = a / b; return r; }
#endif
CG_HD int i_min(int a, int b) { return a < b ? a : b; }
CG_HD int i_max(int a, int b) { return a > b ? a : b; }

CG_HD int lum(int r, int g, int b) { return (r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16; }

// Image.blend for one channel: never let the compiler contract the multiply-add (Pillow's C code is not built with FMA)
CG_HD int blend1(int in1, int in2, float a, bool interp) {
    float temp = f_add((float)in1, f_mul(a, (float)(in2 - in1)));
    if (interp) return (int)temp;
    return temp <= 0.f ? 0 : (temp >= 255.f ? 255 : (int)temp);
}

CG_HD void rgb2hsv(int r, int g, int b, int& uh, int& us, int& uv) {
    int maxc = i_max(r, i_max(g, b)), minc = i_min(r, i_min(g, b));
    uv = maxc;
    if (minc == maxc) { uh = 0; us = 0; return; }
    float cr = (float)(maxc - minc);
    float s = f_div(cr, (float)maxc);
    float rc = f_div((float)(maxc - r), cr), gc = f_div((float)(maxc - g), cr), bc = f_div((float)(maxc - b), cr);
    float h;
    if (r == maxc) h = f_sub(bc, gc);
    else if (g == maxc) h = (float)(2.0 + (double)rc - (double)bc);
    else h = (float)(4.0 + (double)gc - (double)rc);
    h = (float)fmod((double)h / 6.0 + 1.0, 1.0);
    int ih = (int)((double)h * 255.0), is = (int)((double)s * 255.0);
    uh = i_min(i_max(ih, 0), 255);
    us = i_min(i_max(is, 0), 255);
}
CG_HD int round8(float x) { return i_min(i_max((int)floor((double)x + 0.5), 0), 255); }
CG_HD void hsv2rgb(int h, int s, int v, int& r, int& g, int& b) {
    if (s == 0) { r = g = b = v; return; }
    float fh = f_div(f_mul((float)h, 6.0f), 255.0f);
    float fs = f_div((float)s, 255.0f);
    int i = (int)floorf(fh);
    float f = f_sub(fh, (float)i);
    float vv = (float)v;
    int p = round8(f_mul(vv, f_sub(1.0f, fs)));
    int q = round8(f_mul(vv, f_sub(1.0f, f_mul(fs, f))));
    int t = round8(f_mul(vv, f_sub(1.0f, f_mul(fs, f_sub(1.0f, f)))));
    switch (i % 6) {
        case 0: r = v; g = t; b = p; break;
        case 1: r = q; g = v; b = p; break;
        case 2: r = p; g = v; b = t; break;
        case 3: r = p; g = q; b = v; break;
        case 4: r = t; g = p; b = v; break;
        default: r = v; g = p; b = q; break;
    }
}

constexpr int RS_BITS = 32 - 8 - 2;
CG_HD uint8_t rs_clip8(long long v) {
    long long r = v >> RS_BITS;
    return (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
}
// ToTensor (x / 255) + Normalize((x - 0.5) / 0.5)
CG_HD float to_unit(float v) { return f_div(f_sub(f_div(v, 255.0f), 0.5f), 0.5f); }

// one colour op on one pixel (aug_color_kernel's body).  f: blend factor, or np.int32(hue_factor * 255) for the hue shift
CG_HD void color_px(int op, float f, int mean, int& r, int& g, int& b) {
    const bool interp = f >= 0.f && f <= 1.f;
    if (op == CG_AUG_GRAY) {
        r = g = b = lum(r, g, b);
    } else if (op == CG_AUG_BRIGHTNESS || op == CG_AUG_CONTRAST || op == CG_AUG_SATURATION) {
        if (f == 1.f) return;                                   // Image.blend returns im2 itself
        const int d = op == CG_AUG_BRIGHTNESS ? 0 : (op == CG_AUG_CONTRAST ? mean : lum(r, g, b));   // the degenerate image
        if (f == 0.f) { r = g = b = d; return; }                // ... and im1 itself
        r = blend1(d, r, f, interp); g = blend1(d, g, f, interp); b = blend1(d, b, f, interp);
    } else if (op == CG_AUG_HUE) {
        int h, s, v;
        rgb2hsv(r, g, b, h, s, v);
        h = (h + (int)f) & 255;                                 // uint8 wrap-around of np_h += np.uint8(shift)
        hsv2rgb(h, s, v, r, g, b);
    }
}
CG_HD int contrast_mean(unsigned long long lsum, int npix) { return (int)((double)lsum / (double)npix + 0.5); }  // int(Stat(L).mean[0] + 0.5)

// horizontal resize of output column xx in source row `row` (W pixels, RGB), mirrored when the image was flipped
CG_HD void resize_h_px(const uint8_t* row, int W, bool flip, const int32_t* bounds, const int32_t* kk, int ksize, int xx, uint8_t* out) {
    const int x0 = bounds[2 * xx], cnt = bounds[2 * xx + 1];
    long long a0 = 1ll << (RS_BITS - 1), a1 = a0, a2 = a0;
    for (int x = 0; x < cnt; x++) {
        const int sx = flip ? W - 1 - (x0 + x) : x0 + x;
        const uint8_t* px = row + (long)sx * 3;
        const long long k = kk[xx * ksize + x];
        a0 += px[0] * k; a1 += px[1] * k; a2 += px[2] * k;
    }
    out[0] = rs_clip8(a0); out[1] = rs_clip8(a1); out[2] = rs_clip8(a2);
}
// vertical resize of output row yy at column xx of the horizontally resized image (row pitch ow pixels) + ToTensor + Normalize
CG_HD void resize_v_px(const uint8_t* img, int ow, const int32_t* bounds, const int32_t* kk, int ksize, int yy, int xx, float* v) {
    const int y0 = bounds[2 * yy], cnt = bounds[2 * yy + 1];
    long long a0 = 1ll << (RS_BITS - 1), a1 = a0, a2 = a0;
    for (int t = 0; t < cnt; t++) {
        const uint8_t* px = img + ((long)(y0 + t) * ow + xx) * 3;
        const long long k = kk[yy * ksize + t];
        a0 += px[0] * k; a1 += px[1] * k; a2 += px[2] * k;
    }
    v[0] = to_unit((float)rs_clip8(a0)); v[1] = to_unit((float)rs_clip8(a1)); v[2] = to_unit((float)rs_clip8(a2));
}
}  // namespace cg
