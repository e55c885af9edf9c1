"""Tiny driver for ncu: the dominant 3x3 256->256 layer (male2female B=8, council of 4) forward, dgrad, wgrad."""
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
dy = torch.randn(G, B, H, W, C, generator=g).cuda()
dw = torch.empty_like(w)
which = sys.argv[1] if len(sys.argv) > 1 else 'all'
for it in range(3):
    if which in ('all', 'fwd'):
        y = ops.conv_fwd(x, w, b, 1, 1)
    if which in ('all', 'dgrad'):
        dx = ops.conv_dgrad(dy, w, x.shape, 1, 1)
    if which in ('all', 'wgrad'):
        ops.conv_wgrad(x, dy, dw, None, 1, 1)
torch.cuda.synchronize()
print('done')
