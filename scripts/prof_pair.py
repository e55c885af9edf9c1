"""CTA-pair vs single-CTA forward kernel on the dominant 3x3 256->256 layer: CUDA-event timing per launch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from council_gan_b200.ops import CudaOps
ops = CudaOps('cuda:0')
G, B, H, W, C = 4, 8, 64, 64, 256
g = torch.Generator().manual_seed(0)
x = torch.randn(G, B, H, W, C, generator=g).cuda()
w = (torch.randn(G, C, 3, 3, C, generator=g) * 0.02).cuda()
b = torch.randn(G, C, generator=g).cuda()


def timeit(mode, n=20):
    ops.set_tensor_core_mode(mode)
    for _ in range(3):
        ops.conv_fwd(x, w, b, 1, 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        ops.conv_fwd(x, w, b, 1, 1)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


if len(sys.argv) > 1 and sys.argv[1] == 'one':
    ops.set_tensor_core_mode(7)
    ops.conv_fwd(x, w, b, 1, 1)
    torch.cuda.synchronize()
    sys.exit(0)
print('single-CTA  ms', timeit(7 | 8))
print('pair auto   ms', timeit(7))
for cap in (16, 32, 48, 64, 66, 68, 70, 72, 74):
    print('pair cap', cap, 'ms', timeit(7 | (cap << 8)))
