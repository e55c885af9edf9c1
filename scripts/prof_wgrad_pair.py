"""CTA-pair vs single-CTA weight-gradient kernel on the 3x3 256->256 layer (and two strided layers): CUDA-event timing."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from council_gan_b200.ops import CudaOps
ops = CudaOps('cuda:0')
CASES = [(4, 8, 64, 64, 256, 256, 3, 1, 1), (4, 32, 128, 128, 128, 256, 4, 2, 1), (4, 32, 64, 64, 256, 512, 4, 2, 1), (4, 8, 128, 128, 256, 128, 3, 1, 1)]
one = len(sys.argv) > 1 and sys.argv[1] == 'one'
for (G, B, H, W, Cin, Cout, K, s, pad) in CASES[:1] if one else CASES:
    g = torch.Generator().manual_seed(0)
    x = torch.randn(G, B, H, W, Cin, generator=g).cuda()
    Ho = (H + 2 * pad - K) // s + 1
    dy = torch.randn(G, B, Ho, Ho, Cout, generator=g).cuda()
    dw = torch.empty(G, Cout, K, K, Cin, device='cuda')

    def timeit(mode, n=10):
        ops.set_tensor_core_mode(mode)
        for _ in range(2):
            ops.conv_wgrad(x, dy, dw, None, s, pad)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            ops.conv_wgrad(x, dy, dw, None, s, pad)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    if one:
        ops.set_tensor_core_mode(7)
        ops.conv_wgrad(x, dy, dw, None, s, pad)
        torch.cuda.synchronize()
        break
    t1 = timeit(7)
    ref = dw.clone()
    t2 = timeit(7 | (1 << 16))
    rel = ((dw - ref).abs().max() / ref.abs().max()).item()
    print((G, B, H, W, Cin, Cout, K, s), 'one CTA/SM ms %.4f' % t1, 'two CTAs/SM ms %.4f' % t2, 'max rel diff %.2e' % rel, flush=True)
