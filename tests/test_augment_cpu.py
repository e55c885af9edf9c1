"""Input pipeline: the CPU oracle of the reference's per-image transforms (oracle/augment_oracle.py) must be
bit-identical to torchvision + Pillow as installed (the libraries the reference's data loader calls, utils.py:122-181), and the
product's parameter sampler / batch packing (council_gan_b200/data.py) must reproduce a seeded torchvision Compose exactly.
CPU only; the CUDA kernels are checked against the same oracle in tests/test_augment_gpu.py."""
import numpy as np
import pytest
import torch

import augment_oracle as ao
from common import config_for
from council_gan_b200.data import DeviceAugment, precompute_coeffs
from ops_torch import TorchOps

PIL = pytest.importorskip('PIL.Image')
T = pytest.importorskip('torchvision.transforms')
TF = pytest.importorskip('torchvision.transforms.functional')


def rnd_img(rng, h, w, smooth=False):
    a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    if smooth:
        a = np.array(PIL.fromarray(a).resize((max(w // 4, 1), max(h // 4, 1))).resize((w, h), PIL.BICUBIC))
    return a


def test_colour_ops_bit_exact_vs_pillow():
    rng = np.random.default_rng(0)
    for smooth in (False, True):
        img = rnd_img(rng, 109, 89, smooth)
        pil = PIL.fromarray(img)
        assert np.array_equal(ao.grayscale3(img), np.array(TF.rgb_to_grayscale(pil, 3)))
        for f in (0.9, 1.1, 0.0, 1.0, 0.5, 1.7, 1.0321, 0.9993):
            assert np.array_equal(ao.adjust_brightness(img, f), np.array(TF.adjust_brightness(pil, f))), ('brightness', f)
            assert np.array_equal(ao.adjust_contrast(img, f), np.array(TF.adjust_contrast(pil, f))), ('contrast', f)
            assert np.array_equal(ao.adjust_saturation(img, f), np.array(TF.adjust_saturation(pil, f))), ('saturation', f)
        for f in (-0.1, 0.1, 0.05, -0.0371, 0.5, -0.5, 0.0):
            assert np.array_equal(ao.adjust_hue(img, f), np.array(TF.adjust_hue(pil, f))), ('hue', f)


def test_hsv_conversion_bit_exact_on_a_colour_lattice():
    """every 3rd value of each channel (614 k colours) through Pillow's float RGB<->HSV code, both directions"""
    v = np.arange(0, 256, 3)
    r, g, b = np.meshgrid(v, v, v, indexing='ij')
    allc = np.stack([r, g, b], -1).astype(np.uint8).reshape(len(v), -1, 3)
    assert np.array_equal(ao.rgb_to_hsv(allc), np.array(PIL.fromarray(allc).convert('HSV')))
    assert np.array_equal(ao.hsv_to_rgb(allc), np.array(PIL.fromarray(allc, 'HSV').convert('RGB')))


@pytest.mark.parametrize('h,w,size', [(218, 178, 256), (218, 178, 128), (300, 400, 256), (64, 48, 128), (500, 375, 128), (256, 256, 256),
                                      (178, 218, 256)])
def test_resize_bit_exact_vs_pillow(h, w, size):
    rng = np.random.default_rng(h * 1000 + w)
    img = rnd_img(rng, h, w, True)
    oh, ow = ao.resized_size(h, w, size)
    want = np.array(TF.resize(PIL.fromarray(img), size))
    assert want.shape == (oh, ow, 3)
    assert np.array_equal(ao.resize_bilinear(img, oh, ow), want)
    b1, k1 = precompute_coeffs(w, ow)   # the product's table builder == the oracle's
    b2, k2 = ao.precompute_coeffs(w, ow)
    assert np.array_equal(b1, b2) and np.array_equal(k1, k2)


def reference_compose(hp, is_data_A, train, new_size, height, width):
    """The torchvision Compose get_data_loader_folder builds (utils.py:122-176) for the flags of the shipped configs."""
    tl = [T.ToTensor(), T.Normalize((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))]
    if not train:
        tl = [T.CenterCrop(new_size)] + tl
    tl = [T.RandomCrop((height, width))] + tl
    tl = [T.Resize(new_size)] + tl
    if hp['do_HorizontalFlip'] and train:
        tl = [T.RandomHorizontalFlip()] + tl
    jit = hp['do_ColorJitter_A'] if is_data_A else hp['do_ColorJitter_B']
    if jit and train:
        tl = [T.ColorJitter(brightness=hp['ColorJitter_brightness'], contrast=hp['ColorJitter_contrast'],
                            saturation=hp['ColorJitter_saturation'], hue=hp['ColorJitter_hue'])] + tl
    if hp['do_RandomGrayscale'] and train:
        tl = [T.RandomGrayscale(p=hp['RandomGrayscale_P'])] + tl
    return T.Compose(tl)


AUG = {'do_HorizontalFlip': True, 'do_VerticalFlip': False, 'do_ColorJitter_A': True, 'do_ColorJitter_B': True, 'ColorJitter_hue': 0.1,
       'ColorJitter_brightness': 0.1, 'ColorJitter_saturation': 0.1, 'ColorJitter_contrast': 0.1, 'do_RandomGrayscale': True,
       'RandomGrayscale_P': 0.3, 'do_RandomRotation': False, 'do_RandomAffine': False, 'do_RandomPerspective': False,
       'do_RandomResizedCrop': False}   # configs/male2female_council_folder.yaml:96-117 (grayscale probability raised to exercise it)


@pytest.mark.parametrize('train', [True, False])
def test_whole_pipeline_equals_seeded_torchvision_compose(train):
    """Same torch seed -> same random parameters (same calls in the same order) -> bit-identical tensors, image after image, through
    the product's DeviceAugment (host packing / grouping / phases) with the oracle standing in for the CUDA kernels."""
    hp = dict(config_for('male2female'), **AUG)
    hp['new_size'], hp['crop_image_height'], hp['crop_image_width'] = 64, 48, 56
    rng = np.random.default_rng(5)
    imgs = [rnd_img(rng, 109, 89, True), rnd_img(rng, 70, 120, True), rnd_img(rng, 109, 89, True), rnd_img(rng, 64, 64, True)]
    new_size = hp['new_size']
    ch, cw = (hp['crop_image_height'], hp['crop_image_width']) if train else (new_size, new_size)
    comp = reference_compose(hp, True, train, new_size, ch, cw)
    torch.manual_seed(1234)
    want = [comp(PIL.fromarray(im)) for im in imgs]
    aug = DeviceAugment(TorchOps('cpu'), hp, is_data_A=True, train=train)
    torch.manual_seed(1234)
    params = [aug.sample(im.shape[0], im.shape[1]) for im in imgs]
    for im, p, w in zip(imgs, params, want):  # oracle, image by image
        got = ao.train_transform(im, p, new_size, ch, cw)
        assert np.array_equal(got, w.numpy()), 'oracle pipeline differs from torchvision'
    out, nchw = aug(imgs, params=params, want_nchw=True)  # product host logic, whole (ragged) batch
    for b, w in enumerate(want):
        assert torch.equal(nchw[b], w)
        assert torch.equal(out[0, b, :, :, :3].permute(2, 0, 1), w) and float(out[0, b, :, :, 3].abs().max()) == 0


def test_unsupported_transforms_raise():
    hp = dict(config_for('male2female'), **AUG)
    for k in ('do_VerticalFlip', 'do_RandomRotation', 'do_RandomAffine', 'do_RandomPerspective', 'do_RandomResizedCrop'):
        with pytest.raises(NotImplementedError):
            DeviceAugment(TorchOps('cpu'), dict(hp, **{k: True}), True, True)


def test_trainer_takes_the_device_batches_as_they_are(tmp_path):
    """folder -> DeviceFolderLoader -> Council_Trainer.dis_update: the channels-last batch goes into the step without a layout pass and
    gives the same losses as the NCHW tensor of the reference API."""
    import council_oracle as co
    from common import load_golden, setup_case
    from council_gan_b200.data import DeviceFolderLoader
    from council_gan_b200.trainer_council import Council_Trainer
    from test_trainer_host_cpu import load_states
    gold = load_golden('glasses64_n2_b2_early')
    hp, states, _, _ = setup_case(gold)
    hp = dict(hp, **AUG)
    hp['new_size'], hp['crop_image_height'], hp['crop_image_width'] = 64, 64, 64
    rng = np.random.default_rng(1)
    for dom in ('trainA', 'trainB'):
        (tmp_path / dom).mkdir()
        for k in range(4):
            PIL.fromarray(rnd_img(rng, 80, 72, True)).save(str(tmp_path / dom / ('%d.png' % k)))
    ops = TorchOps('cpu')
    # utils.get_all_data_loaders(conf): four one-element lists; test loaders crop to new_size and neither shuffle nor augment
    from council_gan_b200.data import get_all_data_loaders
    for dom in ('testA', 'testB'):
        (tmp_path / dom).mkdir()
        PIL.fromarray(rnd_img(rng, 80, 72, True)).save(str(tmp_path / dom / '0.png'))
    conf = dict(hp, data_root=str(tmp_path), batch_size=1, num_workers=1)
    tra, trb, tea, teb = get_all_data_loaders(ops, conf)
    assert [len(v) for v in (tra, trb, tea, teb)] == [1, 1, 1, 1] and len(tra[0]) == 4 and len(tea[0]) == 1
    torch.manual_seed(0)
    t1 = next(iter(tea[0]))
    torch.manual_seed(0)
    t2 = next(iter(tea[0]))   # the reference's test stack still draws a RandomCrop window (71 x 64 -> 64 x 64 here): same seed, same batch
    assert t1.shape == (1, 1, 64, 64, 4) and torch.equal(t1, t2)
    with pytest.raises(NotImplementedError):
        get_all_data_loaders(ops, dict(conf, inbalenceDataSets={'imbalance_sub_dataset': True}))
    losses = []
    for want_nchw in (False, True):
        torch.manual_seed(3)
        la = DeviceFolderLoader(ops, str(tmp_path / 'trainA'), 2, True, hp, True, num_workers=2, want_nchw=want_nchw)
        lb = DeviceFolderLoader(ops, str(tmp_path / 'trainB'), 2, True, hp, False, num_workers=2, want_nchw=want_nchw)
        assert len(la) == 2
        ba, bb = next(iter(la)), next(iter(lb))
        if want_nchw:
            (ca, ba), (cb, bb) = ba, bb
            assert ba.shape == (2, 3, 64, 64) and ca.shape == (1, 2, 64, 64, 4)
        co.seed_all(3)
        tr = Council_Trainer(hp, 'cpu', _ops=ops)
        load_states(tr, states)
        co.seed_all(7)
        tr.dis_update(ba, bb, hp)
        tr.gen_update(ba, bb, hp, hp['iteration'])
        losses.append([float(v) for v in tr.loss_dis_total_s] + [float(v) for v in tr.loss_gen_total_s])
        assert tr.img_cache_misses == (2 if want_nchw else 0)
    assert losses[0] == losses[1]


# ---- the kernels' own arithmetic, compiled for the host ---------------------------------------------------------------------------
@pytest.fixture(scope='module')
def hostlib(tmp_path_factory):
    """g++ build of csrc/augment_math.cuh (the per-pixel functions the CUDA kernels call) behind the C ABI's argument lists"""
    import ctypes
    import os
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    so = str(tmp_path_factory.mktemp('aughost') / 'libaughost.so')
    subprocess.check_call(['g++', '-O2', '-ffp-contract=off', '-shared', '-fPIC', '-o', so, os.path.join(here, 'augment_host.cpp')])
    return ctypes.CDLL(so)


class HostKernelOps(TorchOps):
    """TorchOps with the two input-pipeline ops running the kernels' arithmetic (host build) instead of the numpy oracle"""

    def __init__(self, lib):
        super().__init__('cpu')
        self.lib = lib

    def aug_color(self, pix, desc, opcode, param, B, max_pixels, any_contrast):
        assert self.lib.h_aug_color(pix.data_ptr(), desc.data_ptr(), opcode.data_ptr(), param.data_ptr(), B) == 0

    def aug_resize_crop(self, pix, src_off, flip, slot, crop, n, H, W, oh, ow, ch, cw, bh, kh, ksh, bv, kv, ksv, out, nchw):
        import ctypes
        tmp = torch.empty(n * H * ow * 3, dtype=torch.uint8)
        P = ctypes.c_void_p
        args = [P(t.data_ptr()) for t in (pix, src_off, flip, slot, crop)] + [n, H, W, oh, ow, ch, cw, P(bh.data_ptr()), P(kh.data_ptr()), ksh,
                                                                             P(bv.data_ptr()), P(kv.data_ptr()), ksv, P(tmp.data_ptr()),
                                                                             P(out.data_ptr()), P(nchw.data_ptr() if nchw is not None else None)]
        assert self.lib.h_aug_resize_crop(*args) == 0


def test_kernel_hue_arithmetic_every_colour(hostlib):
    """all 2^24 colours through the kernels' RGB -> HSV -> shift -> RGB code (host build) == the oracle (== Pillow)"""
    import ctypes
    v = np.arange(256, dtype=np.uint8)
    for shift in (0, -14, 25, 127, -128):
        for r0 in range(0, 256, 32):
            r, g, b = np.meshgrid(v[r0:r0 + 32], v, v, indexing='ij')
            rgb = np.ascontiguousarray(np.stack([r, g, b], -1).reshape(-1, 3))
            out = np.empty_like(rgb)
            hostlib.h_hue_all(ctypes.c_void_p(rgb.ctypes.data), ctypes.c_void_p(out.ctypes.data), ctypes.c_long(rgb.shape[0]), shift)
            hsv = ao.rgb_to_hsv(rgb[None])
            hsv[..., 0] = (hsv[..., 0].astype(np.int64) + shift) % 256
            assert np.array_equal(out, ao.hsv_to_rgb(hsv)[0]), (shift, r0)


@pytest.mark.parametrize('train', [True, False])
def test_kernel_arithmetic_whole_pipeline_on_host(hostlib, train):
    """DeviceAugment driving the kernels' arithmetic (host build): bit-identical to the oracle on ragged batches, every jitter order,
    interpolating and extrapolating factors -- what tests/test_augment_gpu.py checks on the GPU, minus the grid indexing"""
    import itertools
    hp = dict(config_for('male2female'), **AUG)
    hp['new_size'], hp['crop_image_height'], hp['crop_image_width'] = 96, 80, 88
    rng = np.random.default_rng(11)
    orders = list(itertools.permutations(range(4)))
    sizes = [(109, 89), (120, 160), (96, 96)] * 8
    imgs = [rnd_img(rng, h, w, k % 2 == 0) for k, (h, w) in enumerate(sizes)]
    aug = DeviceAugment(HostKernelOps(hostlib), hp, is_data_A=True, train=train)
    ch, cw = (80, 88) if train else (96, 96)
    torch.manual_seed(2)
    params = [aug.sample(h, w) for h, w in sizes]
    if train:
        for k, (p, o) in enumerate(zip(params, orders)):
            p['gray'] = k % 5 == 0
            p['jitter'] = (list(o), 0.6 + 0.07 * k, 1.9 - 0.06 * k, 0.0 if k == 3 else (1.0 if k == 4 else 0.4 + 0.1 * k), -0.5 + k / 23.0)
    out, nchw = aug(imgs, params=params, want_nchw=True)
    for b, (im, p) in enumerate(zip(imgs, params)):
        want = torch.from_numpy(ao.train_transform(im, p, 96, ch, cw))
        assert torch.equal(nchw[b], want), (b, p, float((nchw[b] - want).abs().max()))
        assert torch.equal(out[0, b, :, :, :3].permute(2, 0, 1), want)
