"""Input pipeline timing: DeviceAugment (one staged H2D copy + colour / resize kernels) vs the reference's torchvision Compose on the
host, CelebA-sized images (218x178) -> 256x256, batch 8 (male2female configuration).  Decode is excluded on both sides."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
sys.path.insert(0, 'oracle')
from common import config_for  # noqa: E402
from council_gan_b200.data import DeviceAugment  # noqa: E402
from council_gan_b200.ops import CudaOps  # noqa: E402
from test_augment_cpu import AUG, reference_compose  # noqa: E402


def main():
    import PIL.Image as PIL
    ops = CudaOps('cuda:0')
    hp = dict(config_for('male2female'), **AUG)
    hp['RandomGrayscale_P'] = 0.0
    hp['new_size'], hp['crop_image_height'], hp['crop_image_width'] = 256, 256, 256
    rng = np.random.default_rng(0)
    B = 8
    imgs = [rng.integers(0, 256, (218, 178, 3), dtype=np.uint8) for _ in range(B)]
    aug = DeviceAugment(ops, hp, is_data_A=True, train=True)
    torch.manual_seed(0)
    for _ in range(5):
        aug(imgs)
    torch.cuda.synchronize()
    n = 50
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.current_stream())
    for _ in range(n):
        aug(imgs)
    e1.record(torch.cuda.current_stream())
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n
    dev_ms = e0.elapsed_time(e1) / n
    comp = reference_compose(hp, True, True, 256, 256, 256)
    pil = [PIL.fromarray(im) for im in imgs]
    torch.set_num_threads(1)
    for im in pil:
        comp(im)
    t0 = time.perf_counter()
    reps = 10
    for _ in range(reps):
        for im in pil:
            comp(im)
    cpu = (time.perf_counter() - t0) / reps
    print(json.dumps({'batch': B, 'src': [218, 178], 'dst': [256, 256],
                      'device_pipeline_ms_per_batch_wall': wall * 1e3, 'device_pipeline_ms_per_batch_events': dev_ms,
                      'device_images_per_s': B / wall,
                      'torchvision_one_worker_ms_per_batch': cpu * 1e3, 'torchvision_one_worker_images_per_s': B / cpu}))


if __name__ == '__main__':
    main()
