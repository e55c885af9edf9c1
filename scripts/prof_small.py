"""ncu driver for the small-N / small-K layers: 1x1 64->64 and 3x3 128->64 at 256x256 (G=4, B=8)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from council_gan_b200.ops import CudaOps
ops = CudaOps('cuda:0')
which = sys.argv[1]
g = torch.Generator().manual_seed(0)
if which == 'h1':
    G, B, H, W, Ci, Co, K, pad = 4, 8, 256, 256, 64, 64, 1, 0
else:
    G, B, H, W, Ci, Co, K, pad = 4, 8, 256, 256, 128, 64, 3, 1
x = torch.randn(G, B, H, W, Ci, generator=g).cuda()
w = (torch.randn(G, Co, K, K, Ci, generator=g) * 0.05).cuda()
b = torch.randn(G, Co, generator=g).cuda()
for it in range(3):
    y = ops.conv_fwd(x, w, b, 1, pad, act=1)
torch.cuda.synchronize()
print('done')
