"""Device-side input pipeline for the training step (reference: utils.py:122-181, data.py:108-177, train.py:225-228).

The reference decodes with PIL and augments every image on the CPU inside DataLoader workers (torchvision ``Compose`` over PIL
images), collates to NCHW float tensors and copies them to the GPU.  At ~200 images/s per GPU (two domains, eight GPUs) the
PIL transforms, not the decode, are the cost.  Here the workers only DECODE; the uint8 pixels of a whole minibatch go up in one
pinned asynchronous copy and a handful of kernels (csrc/augment.cu) apply exactly the reference's transforms -- bit-identical to
Pillow / torchvision for every op, see oracle/augment_oracle.py -- and write the result straight in the layout the kernels of
the step read (channels-last fp32 [1, B, H, W, 4] in [-1, 1]); the NCHW tensor of the reference API is produced on request.

Transforms covered = the ones the three shipped configurations enable (configs/*_council_folder.yaml:92-119):
RandomGrayscale, ColorJitter, RandomHorizontalFlip, Resize, RandomCrop (+ CenterCrop of the test loaders), ToTensor, Normalize.
The others (vertical flip, rotation, affine, perspective, RandomResizedCrop; all ``False`` in the shipped configs) raise.
Random parameters are drawn from the torch CPU generator with the same calls, in the same order, as torchvision's transforms,
so a single process seeded like the reference draws the same parameters.
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch

AUG_NONE, AUG_GRAY, AUG_BRIGHTNESS, AUG_CONTRAST, AUG_SATURATION, AUG_HUE = 0, 1, 2, 3, 4, 5
_UNSUPPORTED = ('do_VerticalFlip', 'do_RandomRotation', 'do_RandomAffine', 'do_RandomPerspective', 'do_RandomResizedCrop')
IMG_EXTENSIONS = ('.jpg', '.JPG', '.jpeg', '.JPEG', '.png', '.PNG', '.ppm', '.PPM', '.bmp', '.BMP')  # data.py:82-86


def _jitter_range(v, center=1.0, lo_bound=0.0, clip_first=True):
    """transforms.ColorJitter._check_input for a scalar setting"""
    if v is None or v == 0:
        return None
    if center == 0:  # hue
        return (-float(v), float(v))
    lo, hi = center - float(v), center + float(v)
    return (max(lo, lo_bound) if clip_first else lo, hi)


def precompute_coeffs(insize, outsize):
    """Pillow Resample.c precompute_coeffs + normalize_coeffs_8bpc for the bilinear (triangle) filter.
    -> (bounds int32 [outsize, 2] = (first source index, tap count), coefficients int32 [outsize, ksize] in 22-bit fixed point)"""
    scale = filterscale = insize / outsize
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((outsize, 2), np.int32)
    kk = np.zeros((outsize, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(outsize):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), insize) - xmin
        k = [max(0.0, 1.0 - abs((x + xmin - center + 0.5) * ss)) for x in range(xmax)]
        ww = sum(k)
        for x in range(xmax):
            v = k[x] / ww if ww != 0.0 else k[x]
            kk[xx, x] = int(0.5 + v * (1 << 22)) if v >= 0 else int(-0.5 + v * (1 << 22))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


class DeviceAugment:
    """The transform stack of ``get_data_loader_folder(..., train, new_size, height, width, crop=True, config, is_data_A)`` applied
    to a list of decoded uint8 RGB images on the device."""

    def __init__(self, ops, config, is_data_A, train=True, new_size=None, height=None, width=None):
        for k in _UNSUPPORTED:
            if config.get(k, False) and train:
                raise NotImplementedError('%s is not on the device-side input pipeline (off in the shipped configs)' % k)
        self.ops, self.train = ops, train
        self.new_size = new_size if new_size is not None else config.get('new_size', config.get('new_size_a' if is_data_A else 'new_size_b'))
        self.ch = height if height is not None else (config['crop_image_height'] if train else self.new_size)
        self.cw = width if width is not None else (config['crop_image_width'] if train else self.new_size)
        self.flip = bool(config.get('do_HorizontalFlip', False)) and train
        self.gray_p = float(config.get('RandomGrayscale_P', 0.0)) if (config.get('do_RandomGrayscale', False) and train) else None
        jit = config.get('do_ColorJitter_A', False) if is_data_A else config.get('do_ColorJitter_B', False)
        self.jitter = None
        if jit and train:  # utils.py:152-156
            self.jitter = (_jitter_range(config['ColorJitter_brightness']), _jitter_range(config['ColorJitter_contrast']),
                           _jitter_range(config['ColorJitter_saturation']), _jitter_range(config['ColorJitter_hue'], center=0))
        self._coef = {}

    # ---- random parameters, torchvision's draw order -------------------------------------------------------------------
    def sample(self, h, w):
        """Parameters of one image of size h x w, drawn exactly as torchvision 0.26 draws them inside
        Compose([RandomGrayscale, ColorJitter, RandomHorizontalFlip, Resize, RandomCrop, ...])."""
        p = {'gray': False, 'jitter': None, 'flip': False}
        if self.gray_p is not None:
            p['gray'] = bool(torch.rand(1) < self.gray_p)
        if self.jitter is not None:
            order = torch.randperm(4).tolist()
            vals = [None if r is None else float(torch.empty(1).uniform_(r[0], r[1])) for r in self.jitter]
            p['jitter'] = (order, vals[0], vals[1], vals[2], vals[3])
        if self.flip:
            p['flip'] = bool(torch.rand(1) < 0.5)
        oh, ow = self.resized_size(h, w)
        if oh < self.ch or ow < self.cw:
            raise ValueError('Required crop size %s is larger than input image size %s' % ((self.ch, self.cw), (oh, ow)))
        if ow == self.cw and oh == self.ch:
            p['crop'] = (0, 0)
        else:  # RandomCrop.get_params
            i = int(torch.randint(0, oh - self.ch + 1, size=(1,)).item())
            j = int(torch.randint(0, ow - self.cw + 1, size=(1,)).item())
            p['crop'] = (i, j)
        return p

    def resized_size(self, h, w):
        if self.new_size is None:
            return h, w
        s = self.new_size
        return (int(s * h / w), s) if w <= h else (s, int(s * w / h))

    # ---- the device pipeline -----------------------------------------------------------------------------------------------
    def __call__(self, images, params=None, want_nchw=False):
        """images: list of uint8 [H, W, 3] CPU tensors / arrays (decoded RGB).  -> channels-last [1, B, ch, cw, 4] on the device
        (and the NCHW [B, 3, ch, cw] tensor of the reference API when want_nchw)."""
        ops = self.ops
        imgs = [torch.as_tensor(np.ascontiguousarray(im)) for im in images]
        B = len(imgs)
        if params is None:
            params = [self.sample(im.shape[0], im.shape[1]) for im in imgs]
        sizes = [(int(im.shape[0]), int(im.shape[1])) for im in imgs]
        offs, total = [], 0
        for h, w in sizes:
            offs.append(total)
            total += (h * w * 3 + 15) // 16 * 16
        packed = torch.empty(total, dtype=torch.uint8)
        for im, o, (h, w) in zip(imgs, offs, sizes):
            packed[o:o + h * w * 3] = im.reshape(-1)
        desc = torch.tensor([[o, h, w, 0] for o, (h, w) in zip(offs, sizes)], dtype=torch.int32)
        # colour phases: phase 0 = RandomGrayscale, phases 1..4 = the four ColorJitter ops in each image's own order
        phases = []
        if any(p['gray'] for p in params):
            phases.append(([AUG_GRAY if p['gray'] else AUG_NONE for p in params], [0.0] * B))
        if any(p['jitter'] is not None for p in params):
            for k in range(4):
                codes, vals = [], []
                for p in params:
                    if p['jitter'] is None:
                        codes.append(AUG_NONE), vals.append(0.0)
                        continue
                    order, b, c, s, h = p['jitter']
                    fn = order[k]
                    v = (b, c, s, h)[fn]
                    if v is None:
                        codes.append(AUG_NONE), vals.append(0.0)
                    elif fn == 3:
                        codes.append(AUG_HUE), vals.append(float(int(v * 255)))  # np.int32(hue_factor * 255): host double, truncation
                    else:
                        codes.append((AUG_BRIGHTNESS, AUG_CONTRAST, AUG_SATURATION)[fn]), vals.append(float(v))
                phases.append((codes, vals))
        # groups of images with one source size share the resize coefficient tables
        groups = {}
        for b, hw in enumerate(sizes):
            groups.setdefault(hw, []).append(b)
        host = [packed.view(torch.int32) if total % 4 == 0 else packed, desc]
        for codes, vals in phases:
            host += [torch.tensor(codes, dtype=torch.int32), torch.tensor(vals, dtype=torch.float32)]
        gmeta = []
        for (h, w), members in groups.items():
            oh, ow = self.resized_size(h, w)
            bh, kh = self._coeffs(w, ow)
            bv, kv = self._coeffs(h, oh)
            host += [torch.tensor([offs[b] for b in members], dtype=torch.int32),
                     torch.tensor([int(params[b]['flip']) for b in members], dtype=torch.int32),
                     torch.tensor(members, dtype=torch.int32),
                     torch.tensor([list(params[b]['crop']) for b in members], dtype=torch.int32),
                     torch.from_numpy(bh), torch.from_numpy(kh), torch.from_numpy(bv), torch.from_numpy(kv)]
            gmeta.append((h, w, oh, ow, len(members), kh.shape[1], kv.shape[1]))
        dev = ops.stage(host)  # ONE pinned buffer, ONE async H2D copy for pixels and every table
        d_pix = dev[0].view(torch.uint8)
        d_desc = dev[1]
        k = 2
        for codes, vals in phases:
            ops.aug_color(d_pix, d_desc, dev[k], dev[k + 1], B, max(h * w for h, w in sizes), any(c == AUG_CONTRAST for c in codes))
            k += 2
        out = ops.empty(1, B, self.ch, self.cw, 4)
        nchw = ops.empty(B, 3, self.ch, self.cw) if want_nchw else None
        for (h, w, oh, ow, n, ksh, ksv) in gmeta:
            src_off, flip, slot, crop, bh, kh, bv, kv = dev[k:k + 8]
            k += 8
            ops.aug_resize_crop(d_pix, src_off, flip, slot, crop, n, h, w, oh, ow, self.ch, self.cw, bh, kh, ksh, bv, kv, ksv, out, nchw)
        return (out, nchw) if want_nchw else out

    def _coeffs(self, insize, outsize):
        key = (insize, outsize)
        if key not in self._coef:
            self._coef[key] = precompute_coeffs(insize, outsize)
        return self._coef[key]


def make_dataset(root):
    """data.py:96-107: every image file under root (recursively), sorted."""
    out = []
    for d, _, fnames in sorted(os.walk(root)):
        for f in fnames:
            if f.endswith(IMG_EXTENSIONS):
                out.append(os.path.join(d, f))
    return sorted(out)


class DeviceFolderLoader:
    """``get_data_loader_folder`` (utils.py:122-181) with decode on CPU threads and the transforms on the device.
    Yields what ``DeviceAugment.__call__`` returns; shuffle / drop_last follow DataLoader(shuffle=train, drop_last=True)."""

    def __init__(self, ops, input_folder, batch_size, train, config, is_data_A, num_workers=4, want_nchw=False):
        from concurrent.futures import ThreadPoolExecutor
        self.files = make_dataset(input_folder)
        if not self.files:
            raise RuntimeError('Found 0 images in: ' + input_folder)
        self.bs, self.train, self.want_nchw = batch_size, train, want_nchw
        self.aug = DeviceAugment(ops, config, is_data_A, train)
        self.pool = ThreadPoolExecutor(max(1, num_workers))

    def __len__(self):
        return len(self.files) // self.bs

    @staticmethod
    def _decode(path):
        from PIL import Image
        with open(path, 'rb') as f:
            return np.array(Image.open(f).convert('RGB'))  # data.py default_loader (a writable copy)

    def __iter__(self):
        order = torch.randperm(len(self.files)).tolist() if self.train else list(range(len(self.files)))
        nxt = None
        for b in range(len(self)):
            idx = order[b * self.bs:(b + 1) * self.bs]
            cur = nxt if nxt is not None else [self.pool.submit(self._decode, self.files[i]) for i in idx]
            if b + 1 < len(self):  # decode the next minibatch while this one is augmented and consumed
                nidx = order[(b + 1) * self.bs:(b + 2) * self.bs]
                nxt = [self.pool.submit(self._decode, self.files[i]) for i in nidx]
            else:
                nxt = None
            yield self.aug([f.result() for f in cur], want_nchw=self.want_nchw)


def get_all_data_loaders(ops, conf, want_nchw=False):
    """``utils.get_all_data_loaders(conf)`` (utils.py:45-104) for the folder layout the shipped configs use
    (``data_root``/{trainA,trainB,testA,testB}): -> (train_loader_a, train_loader_b, test_loader_a, test_loader_b), each a one-element
    list like the reference's.  Train loaders crop to crop_image_height x crop_image_width, test loaders to new_size x new_size."""
    if 'data_root' not in conf:
        raise NotImplementedError('list-file datasets (data_folder_train_a / data_list_train_a) are not on the device-side input pipeline')
    if conf.get('inbalenceDataSets', {}).get('imbalance_sub_dataset', False):
        raise NotImplementedError('imbalanced sub-dataset sampling (inbalenceDataSets) is not on the device-side input pipeline')
    if conf.get('input_dim_a', 3) == 1 or conf.get('input_dim_b', 3) == 1:
        raise NotImplementedError('single-channel domains (dim3to1) are not on the device-side input pipeline')
    bs, nw, root = conf['batch_size'], conf.get('num_workers', 4), conf['data_root']

    def mk(sub, train, is_a):
        return [DeviceFolderLoader(ops, os.path.join(root, sub), bs, train, conf, is_a, num_workers=nw, want_nchw=want_nchw)]
    return mk('trainA', True, True), mk('trainB', True, False), mk('testA', False, True), mk('testB', False, False)
