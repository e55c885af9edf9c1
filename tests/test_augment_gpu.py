"""Device-side input pipeline (csrc/augment.cu through council_gan_b200/data.py) against the CPU oracle of the reference's transforms
(oracle/augment_oracle.py, itself bit-identical to torchvision + Pillow: tests/test_augment_cpu.py).  Bit-exact: integer pixel
arithmetic and the same float operations in the same order."""
import numpy as np
import pytest
import torch

import augment_oracle as ao
from common import config_for
from council_gan_b200.data import DeviceAugment
from test_augment_cpu import AUG

pytestmark = pytest.mark.gpu


def smooth_img(rng, h, w):
    a = rng.integers(0, 256, (max(h // 4, 1), max(w // 4, 1), 3), dtype=np.uint8)
    return ao.resize_bilinear(a, h, w)  # smooth content without needing Pillow on the GPU box


@pytest.mark.parametrize('seed', [0, 1, 2])
@pytest.mark.parametrize('train', [True, False])
def test_device_pipeline_bit_exact(seed, train):
    from council_gan_b200.ops import CudaOps
    ops = CudaOps('cuda:0')
    hp = dict(config_for('male2female'), **AUG)
    hp['RandomGrayscale_P'] = 0.4
    hp['new_size'], hp['crop_image_height'], hp['crop_image_width'] = 256, 256, 256
    rng = np.random.default_rng(seed)
    sizes = [(218, 178)] * 5 + [(300, 260), (256, 256), (178, 218)]  # CelebA-sized images plus a ragged tail
    imgs = [smooth_img(rng, h, w) if i % 2 else rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for i, (h, w) in enumerate(sizes)]
    aug = DeviceAugment(ops, hp, is_data_A=True, train=train)
    torch.manual_seed(100 + seed)
    params = [aug.sample(h, w) for h, w in sizes]
    out, nchw = aug(imgs, params=params, want_nchw=True)
    torch.cuda.synchronize()
    ch = cw = 256
    for b, (im, p) in enumerate(zip(imgs, params)):
        want = torch.from_numpy(ao.train_transform(im, p, hp['new_size'], ch, cw))
        got = nchw[b].cpu()
        assert torch.equal(got, want), ('image', b, p, float((got - want).abs().max()))
        assert torch.equal(out[0, b, :, :, :3].permute(2, 0, 1).cpu(), want)
        assert float(out[0, b, :, :, 3].abs().max()) == 0


def test_every_jitter_order_and_extreme_factors():
    """all 24 op orders; factors outside [0, 1] (extrapolating blend with clipping), hue at both ends"""
    import itertools
    from council_gan_b200.ops import CudaOps
    ops = CudaOps('cuda:0')
    hp = dict(config_for('male2female'), **AUG)
    hp['new_size'], hp['crop_image_height'], hp['crop_image_width'] = 64, 64, 64
    rng = np.random.default_rng(9)
    orders = list(itertools.permutations(range(4)))
    imgs = [rng.integers(0, 256, (80, 72, 3), dtype=np.uint8) for _ in orders]
    aug = DeviceAugment(ops, hp, is_data_A=True, train=True)
    params = []
    for k, o in enumerate(orders):
        params.append({'gray': k % 5 == 0, 'flip': k % 2 == 1, 'crop': (k % 8, 0),  # 80 x 72 -> 71 x 64: rows 0..7, column 0
                       'jitter': (list(o), 0.6 + 0.07 * k, 1.9 - 0.06 * k, 0.0 if k == 3 else 0.4 + 0.1 * k, -0.5 + k / 23.0)})
    _, nchw = aug(imgs, params=params, want_nchw=True)
    for b, (im, p) in enumerate(zip(imgs, params)):
        want = torch.from_numpy(ao.train_transform(im, p, 64, 64, 64))
        assert torch.equal(nchw[b].cpu(), want), (b, p)


def test_device_batch_feeds_the_training_step():
    """the channels-last tensor the pipeline writes is what Council_Trainer._img would have produced from the NCHW tensor"""
    from council_gan_b200.ops import CudaOps
    ops = CudaOps('cuda:0')
    hp = dict(config_for('male2female'), **AUG)
    hp['new_size'], hp['crop_image_height'], hp['crop_image_width'] = 64, 64, 64
    rng = np.random.default_rng(3)
    imgs = [rng.integers(0, 256, (70, 66, 3), dtype=np.uint8) for _ in range(4)]
    aug = DeviceAugment(ops, hp, is_data_A=False, train=True)
    torch.manual_seed(5)
    out, nchw = aug(imgs, want_nchw=True)
    assert torch.equal(ops.nchw_to_nhwc(nchw, 4)[None], out)
