"""2 ranks x half batch must equal 1 rank x full batch on real GPUs over NCCL (exact-fp32 SIMT kernels so the
comparison is tight): rank 0 also runs the single-process global-batch iteration and compares parameters."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')]
import torch, torch.distributed as dist
import council_oracle as co
from common import load_golden, setup_case
from test_trainer_host_cpu import load_states

rank, world, lr = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dist.init_process_group('nccl', device_id=torch.device('cuda:%d' % lr))
from council_gan_b200 import Council_Trainer
gold = load_golden('m2f64_n4_b2')
hp, states, x_a, x_b = setup_case(gold)

def run(xa, xb):
    co.seed_all(hp['random_seed'])
    tr = Council_Trainer(hp, 'cuda:%d' % lr)
    tr.ops.set_tensor_core_mode(0)
    load_states(tr, states)
    co.seed_all(gold['rng_seed'])
    tr.dis_update(xa, xb, hp); tr.dis_council_update(xa, xb, hp); tr.gen_update(xa, xb, hp, gold['iteration'])
    tr.synchronize()  # under data parallelism the last family's all-reduce + Adam are deferred
    torch.cuda.synchronize()
    return tr

b = x_a.size(0) // world
tr = run(x_a[rank * b:(rank + 1) * b], x_b[rank * b:(rank + 1) * b])
dp = {n: net.bank.data.clone() for n, net in tr._nets.items()}
dp_loss = [float(v) for v in tr.loss_gen_total_s]
if rank == 0:
    # single-process reference on the global batch: temporarily pretend there is no process group
    import council_gan_b200.trainer_council as tc
    saved = tc._dist
    tc._dist = lambda: None
    ref = run(x_a, x_b)
    tc._dist = saved
    print('gen loss dp', dp_loss, 'single', [float(v) for v in ref.loss_gen_total_s])
    for n, net in ref._nets.items():
        diff = (net.bank.data - dp[n]).abs()
        frac = (diff > 0.5 * hp['lr']).float().mean().item()
        print('%-20s max |dp - single| = %.3e   fraction off by > lr/2: %.4f' % (n, diff.max().item(), frac))
dist.barrier()
dist.destroy_process_group()
