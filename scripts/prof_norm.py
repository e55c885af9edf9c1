"""Driver for ncu / timing of the normalisation passes: python scripts/prof_norm.py G B H W C [iters]
times in_stats + norm_act_fwd vs norm_fused_fwd, and norm_act_bwd vs norm_fused_bwd (CUDA events, ms per call)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from council_gan_b200.ops import CudaOps
G, B, H, W, C = [int(v) for v in sys.argv[1:6]]
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 10
ops = CudaOps('cuda:0')
y = torch.randn(G, B, H, W, C, device='cuda')
res = torch.randn_like(y)
dz = torch.randn_like(y)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
def timed(fn):
    for _ in range(2):
        fn()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()  # evict L2 between timed iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters
def two_fwd():
    m, r = ops.in_stats(y)
    return ops.norm_act_fwd(y, m, r, None, 0, res, 1, False)
m, r = ops.in_stats(y)
gb = y.numel() * 4 / 1e9
t = timed(two_fwd); print('G%d B%d %dx%d C%d  in_stats + norm_act_fwd(res): %.3f ms  (%.0f GB/s algorithmic 4 units)' % (G, B, H, W, C, t, 4 * gb / t * 1e3))
t = timed(lambda: ops.norm_fused_fwd(y, None, 0, res, 1, False)); print('  norm_fused_fwd(res): %.3f ms  (%.0f GB/s algorithmic 3 units)' % (t, 3 * gb / t * 1e3))
t = timed(lambda: ops.norm_act_bwd(dz, y, m, r, None, 0, 1, False, None)); print('  norm_act_bwd: %.3f ms  (%.0f GB/s algorithmic 5 units)' % (t, 5 * gb / t * 1e3))
t = timed(lambda: ops.norm_fused_bwd(dz, y, m, r, None, 0, 1, False, None)); print('  norm_fused_bwd: %.3f ms  (%.0f GB/s algorithmic 3 units)' % (t, 3 * gb / t * 1e3))
