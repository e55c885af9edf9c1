"""Host-side cost of one training step on the launch-bound small configuration: cProfile over 20 steps (GPU work is asynchronous, so the
profile shows what the Python / ctypes launch path costs), plus wall-clock per step with and without a synchronize per step."""
import cProfile
import io
import os
import pstats
import random
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from council_gan_b200 import Council_Trainer  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else 'glasses_128_n2_b1'
hp, n, batch, size, it = bench.load_hp(workload)
random.seed(1); np.random.seed(1); torch.manual_seed(1)
tr = Council_Trainer(hp, 'cuda:0')
xa, xb = bench.synth(batch, size, 123)
xa, xb = xa.cuda(), xb.cuda()


def step():
    tr.dis_update(xa, xb, hp)
    tr.dis_council_update(xa, xb, hp)
    tr.gen_update(xa, xb, hp, it)
    tr.update_learning_rate()


for _ in range(5):
    step()
torch.cuda.synchronize()
l0 = tr.ops.launch_count()
t0 = time.perf_counter()
for _ in range(20):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print('launches/step %d; host issue time %.3f ms/step; incl. final drain %.3f ms/step' % ((tr.ops.launch_count() - l0) / 20, (t1 - t0) * 50, (t2 - t0) * 50))
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(45)
print(s.getvalue())
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(45)
print(s.getvalue())
