"""Driver for ncu / timing of ONE convolution geometry:  python scripts/prof_layer.py KIND G B H W Cin Cout K stride pad [iters]
KIND in fwd | dgrad | wgrad | all.  Prints the CUDA-event time per launch (ms) after warm-up."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from council_gan_b200.ops import CudaOps
kind = sys.argv[1]
G, B, H, W, Cin, Cout, K, stride, pad = [int(v) for v in sys.argv[2:11]]
iters = int(sys.argv[11]) if len(sys.argv) > 11 else 5
ops = CudaOps('cuda:0')
if os.environ.get('COUNCIL_TC_MODE'):
    ops.set_tensor_core_mode(int(os.environ['COUNCIL_TC_MODE'], 0))
g = torch.Generator().manual_seed(0)
x = torch.randn(G, B, H, W, Cin, generator=g).cuda()
w = (torch.randn(G, Cout, K, K, Cin, generator=g) * 0.05).cuda()
bias = torch.randn(G, Cout, generator=g).cuda()
y = ops.conv_fwd(x, w, bias, stride, pad, act=2)
dy = torch.randn(*y.shape, generator=g).cuda() if False else torch.randn_like(y)
dw = torch.empty_like(w)
db = torch.empty_like(bias)
def run(k):
    if k == 'fwd':
        ops.conv_fwd(x, w, bias, stride, pad, act=2)
    elif k == 'dgrad':
        ops.conv_dgrad(dy, w, x.shape, stride, pad)
    else:
        ops.conv_wgrad(x, dy, dw, db, stride, pad)
for k in (['fwd', 'dgrad', 'wgrad'] if kind == 'all' else [kind]):
    for _ in range(2):
        run(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run(k)
    e1.record()
    torch.cuda.synchronize()
    print('mode %s %s G%d B%d %dx%d Cin%d Cout%d k%d s%d: %.3f ms per call' % (os.environ.get('COUNCIL_TC_MODE', 'default'), k, G, B, H, W, Cin, Cout, K, stride, e0.elapsed_time(e1) / iters))
