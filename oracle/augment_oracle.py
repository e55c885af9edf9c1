"""CPU oracle of the reference's training-time image transforms  --  TEST INFRASTRUCTURE ONLY.

Restates, in plain numpy integer / float32 arithmetic, what the reference's input pipeline does to ONE decoded image
(`utils.py:122-181` `get_data_loader_folder`: a torchvision `Compose` over PIL images) for the transforms that are live in
the three shipped configurations (`configs/*_council_folder.yaml:92-119`):

    RandomGrayscale(p) -> ColorJitter(b, c, s, h) -> RandomHorizontalFlip -> Resize(new_size) -> RandomCrop(h, w)
    -> ToTensor -> Normalize(0.5, 0.5)                                  (train)
    Resize(new_size) -> RandomCrop -> CenterCrop -> ToTensor -> Normalize   (test loaders)

The arithmetic lives in third-party code that is not under /root/reference: torchvision (pinned by the reference's
conda_requirements.yml as torchvision=0.6.0; 0.26.0 in this image) and Pillow (7.1; 12.2.0 here).  Pinning: the
functions below are checked bit-for-bit against torchvision + Pillow as installed here
(tests/test_augment_cpu.py), on random images and on every parameter combination the configs can draw.

Integer formulas restated from Pillow's C sources (libImaging): Convert.c rgb2l / rgb2hsv / hsv2rgb, Blend.c, Resample.c
(precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc / Vertical_8bpc, PRECISION_BITS = 32 - 8 - 2).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


# ---- Convert.c ---------------------------------------------------------------------------------------------------------
def rgb_to_l(img):
    """img uint8 [H,W,3] -> uint8 [H,W]: L = (R*19595 + G*38470 + B*7471 + 0x8000) >> 16   (ITU-R 601-2 luma, L24 macro)"""
    i = img.astype(np.int64)
    return ((i[..., 0] * 19595 + i[..., 1] * 38470 + i[..., 2] * 7471 + 0x8000) >> 16).astype(np.uint8)


def grayscale3(img):
    """RandomGrayscale / F.rgb_to_grayscale(num_output_channels=3) on an RGB PIL image"""
    l = rgb_to_l(img)
    return np.stack([l, l, l], -1)


def rgb_to_hsv(img):
    """Convert.c rgb2hsv_row (float arithmetic as written there: C `float`)"""
    r, g, b = (img[..., k].astype(np.int32) for k in range(3))
    maxc = np.maximum(r, np.maximum(g, b))
    minc = np.minimum(r, np.minimum(g, b))
    cr = (maxc - minc).astype(F32)
    safe = np.where(cr == 0, F32(1), cr)
    s = cr / np.where(maxc == 0, 1, maxc).astype(F32)
    rc = (maxc - r).astype(F32) / safe
    gc = (maxc - g).astype(F32) / safe
    bc = (maxc - b).astype(F32) / safe
    # `h = 2.0 + rc - bc` : the literal is a C double, so the sum is formed in double and rounded to the float variable
    h = np.where(r == maxc, bc - gc, np.where(g == maxc, (2.0 + rc.astype(np.float64) - bc.astype(np.float64)).astype(F32),
                                              (4.0 + gc.astype(np.float64) - rc.astype(np.float64)).astype(F32))).astype(F32)
    # h = fmod((h / 6.0 + 1.0), 1.0): the constants are C doubles, so the right-hand side is evaluated in double precision and
    # ROUNDED BACK to the float variable; uh = (int)(h * 255.0) promotes that float to double again
    h = np.fmod(h.astype(np.float64) / 6.0 + 1.0, 1.0).astype(F32)
    uh = np.clip((h.astype(np.float64) * 255.0).astype(np.int64), 0, 255)
    us = np.clip((s.astype(np.float64) * 255.0).astype(np.int64), 0, 255)
    gray = minc == maxc
    uh = np.where(gray, 0, uh)
    us = np.where(gray, 0, us)
    return np.stack([uh, us, maxc], -1).astype(np.uint8)


def hsv_to_rgb(hsv):
    """Convert.c hsv2rgb_row"""
    h, s, v = (hsv[..., k] for k in range(3))
    fh = (h.astype(F32) * F32(6.0) / F32(255.0)).astype(F32)
    fs = (s.astype(F32) / F32(255.0)).astype(F32)
    i = np.floor(fh).astype(np.int32)
    f = (fh - i.astype(F32)).astype(F32)
    vv = v.astype(F32)

    def r8(x):  # CLIP8(round(x)): C `round` = half away from zero, values are >= 0 here
        return np.clip(np.floor(x.astype(np.float64) + 0.5).astype(np.int64), 0, 255).astype(np.uint8)
    p = r8(vv * (F32(1.0) - fs))
    q = r8(vv * (F32(1.0) - fs * f))
    t = r8(vv * (F32(1.0) - fs * (F32(1.0) - f)))
    up, uq, ut, uv = p, q, t, v
    sel = i % 6
    r = np.choose(sel, [uv, uq, up, up, ut, uv])
    g = np.choose(sel, [ut, uv, uv, uq, up, up])
    b = np.choose(sel, [up, up, ut, uv, uv, uq])
    out = np.stack([r, g, b], -1).astype(np.uint8)
    gray = s == 0
    out[gray] = np.stack([v, v, v], -1)[gray]
    return out


# ---- Blend.c / ImageEnhance ---------------------------------------------------------------------------------------------
def blend(im1, im2, alpha):
    """Image.blend(im1, im2, alpha): out = im1 + alpha * (im2 - im1) in float, truncated; clipped when extrapolating"""
    a = F32(alpha)
    if a == 0:
        return im1.copy()
    if a == 1:
        return im2.copy()
    temp = (im1.astype(np.int32).astype(F32) + a * (im2.astype(np.int32) - im1.astype(np.int32)).astype(F32)).astype(F32)
    if 0 <= a <= 1:
        return temp.astype(np.int64).astype(np.uint8)        # (UINT8)temp: truncation, no clipping needed
    out = np.where(temp <= 0, 0, np.where(temp >= 255, 255, temp.astype(np.int64)))
    return out.astype(np.uint8)


def adjust_brightness(img, f):
    return blend(np.zeros_like(img), img, f)                  # ImageEnhance.Brightness: degenerate = black


def adjust_contrast(img, f):
    l = rgb_to_l(img)
    mean = int(l.astype(np.float64).sum() / l.size + 0.5)     # int(ImageStat.Stat(L).mean[0] + 0.5)
    return blend(np.full_like(img, mean), img, f)


def adjust_saturation(img, f):
    return blend(grayscale3(img), img, f)                     # ImageEnhance.Color: degenerate = L converted back to RGB


def adjust_hue(img, f):
    hsv = rgb_to_hsv(img)
    hsv[..., 0] = (hsv[..., 0].astype(np.int64) + int(f * 255)) % 256   # np_h += np.int32(f * 255).astype(np.uint8): wraps (F_pil.adjust_hue)
    return hsv_to_rgb(hsv)


def color_jitter(img, order, b, c, s, h):
    for fn in order:  # transforms.ColorJitter.forward: 0 brightness, 1 contrast, 2 saturation, 3 hue
        if fn == 0 and b is not None:
            img = adjust_brightness(img, b)
        elif fn == 1 and c is not None:
            img = adjust_contrast(img, c)
        elif fn == 2 and s is not None:
            img = adjust_saturation(img, s)
        elif fn == 3 and h is not None:
            img = adjust_hue(img, h)
    return img


# ---- Resample.c (bilinear) ----------------------------------------------------------------------------------------------
PRECISION_BITS = 32 - 8 - 2


def _triangle(x):
    x = abs(x)
    return 1.0 - x if x < 1.0 else 0.0


def precompute_coeffs(insize, outsize):
    """-> (bounds [outsize][2] = (xmin, count), coefficient table int32 [outsize][ksize]) for the bilinear filter"""
    scale = filterscale = insize / outsize
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((outsize, 2), np.int32)
    kk = np.zeros((outsize, ksize), np.int32)
    for xx in range(outsize):
        center = (xx + 0.5) * scale
        ss = 1.0 / filterscale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > insize:
            xmax = insize
        xmax -= xmin
        k = [_triangle((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = sum(k)
        for x in range(xmax):
            v = k[x] / ww if ww != 0.0 else k[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _clip8(v):
    return np.clip(v >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_bilinear(img, out_h, out_w):
    """Image.resize((out_w, out_h), BILINEAR) on uint8 RGB: horizontal pass then vertical pass, uint8 in between"""
    H, W, _ = img.shape
    cur = img
    if out_w != W:
        bounds, kk = precompute_coeffs(W, out_w)
        tmp = np.zeros((H, out_w, 3), np.uint8)
        for xx in range(out_w):
            x0, n = bounds[xx]
            acc = np.full((H, 3), 1 << (PRECISION_BITS - 1), np.int64)
            for x in range(n):
                acc += cur[:, x0 + x, :].astype(np.int64) * int(kk[xx, x])
            tmp[:, xx, :] = _clip8(acc)
        cur = tmp
    if out_h != H:
        bounds, kk = precompute_coeffs(H, out_h)
        tmp = np.zeros((out_h, cur.shape[1], 3), np.uint8)
        for yy in range(out_h):
            y0, n = bounds[yy]
            acc = np.full((cur.shape[1], 3), 1 << (PRECISION_BITS - 1), np.int64)
            for y in range(n):
                acc += cur[y0 + y, :, :].astype(np.int64) * int(kk[yy, y])
            tmp[yy] = _clip8(acc)
        cur = tmp
    return cur


def resized_size(h, w, size):
    """transforms.Resize(int): shortest side -> size, the other side int(size * long / short)"""
    if w <= h:
        return int(size * h / w), size
    return size, int(size * w / h)


# ---- the whole per-image pipeline -----------------------------------------------------------------------------------------
def train_transform(img, params, new_size, crop_h, crop_w):
    """img uint8 [H,W,3]; params: dict(gray: bool, jitter: None | (order, b, c, s, h), flip: bool, crop: (i, j)).
    -> float32 [3, crop_h, crop_w] in [-1, 1]  (ToTensor + Normalize(0.5, 0.5))"""
    if params.get('gray'):
        img = grayscale3(img)
    if params.get('jitter') is not None:
        img = color_jitter(img, *params['jitter'])
    if params.get('flip'):
        img = img[:, ::-1, :]
    if new_size is not None:
        oh, ow = resized_size(img.shape[0], img.shape[1], new_size)
        img = resize_bilinear(np.ascontiguousarray(img), oh, ow)
    i, j = params['crop']
    img = img[i:i + crop_h, j:j + crop_w, :]
    x = img.astype(F32) / F32(255.0)
    x = (x - F32(0.5)) / F32(0.5)
    return np.ascontiguousarray(x.transpose(2, 0, 1))
