"""Is the TF32 gradient deviation a bug or the network's own noise amplification?  Per-layer cosine / relative L2 of
the generator gradients vs the fp32 CPU oracle for: exact fp32 kernels, exact kernels with weights perturbed by
2^-11 relative noise, and the tensor-core path enabled for forward / dgrad / wgrad separately and together."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')]
import torch
import council_oracle as co
from common import load_golden, setup_case
from test_trainer_host_cpu import load_states, run_oracle
from council_gan_b200 import Council_Trainer

case = sys.argv[1] if len(sys.argv) > 1 else 'glasses64_n2_b2_early'
gold = load_golden(case)
orc, hp = run_oracle(gold, torch.float32)
d0 = orc.dirs[0]
KEYS = ['enc_content.model.0.conv.weight', 'enc_content.model.2.conv.weight', 'enc_content.model.3.model.4.model.1.conv.weight',
        'dec.model.0.model.0.model.0.conv.weight', 'dec.model.0.model.4.model.1.conv.weight', 'dec.model.3.conv.weight',
        'dec.model.6.conv.weight', 'dec.model.8.conv.weight', 'dec.model.9.conv.weight', 'mlp.model.2.fc.weight']

def run(mode, perturb=0.0):
    hp2, states, x_a, x_b = setup_case(gold)
    if perturb:
        g = torch.Generator().manual_seed(99)
        states = {k: [{kk: (vv * (1 + perturb * torch.randn(vv.shape, generator=g)) if vv.dim() > 1 else vv) for kk, vv in sd.items()}
                      for sd in lst] for k, lst in states.items()}
    co.seed_all(hp2['random_seed'])
    tr = Council_Trainer(hp2, 'cuda:0')
    tr.ops.set_tensor_core_mode(mode)
    load_states(tr, states)
    co.seed_all(gold['rng_seed'])
    tr.dis_update(x_a, x_b, hp2); tr.dis_council_update(x_a, x_b, hp2); tr.gen_update(x_a, x_b, hp2, gold['iteration'])
    torch.cuda.synchronize()
    net = tr._nets['gen_' + d0]
    out = {}
    for spec in net.live_specs:
        if spec.wname in KEYS:
            g = spec.export_weight(net.bank.g(spec.wname)[0]).cpu()
            og = orc.P['gen_' + d0][0][spec.wname].grad
            out[spec.wname] = (float((g * og).sum() / (g.norm() * og.norm())), float((g - og).norm() / og.norm()))
    losses = [float(v) for v in tr.loss_gen_total_s]
    return out, losses

print('oracle gen losses', [float(v) for v in orc.loss_gen_total_s])
for name, mode, pert in (('fp32 kernels', 0, 0.0), ('fp32 kernels, weights*(1+2^-11 n)', 0, 2.0 ** -11), ('tc fwd only', 3 - 2, 0.0),
                         ('tc dgrad only', 2, 0.0), ('tc wgrad only', 4, 0.0), ('tc all', 7, 0.0)):
    try:
        res, losses = run(mode, pert)
        print('==', name, 'mode', mode, 'losses', losses)
        for k in KEYS:
            if k in res:
                print('   %-55s cos %.4f  relL2 %.3e' % (k, res[k][0], res[k][1]))
    except Exception as e:
        print('==', name, 'FAILED', repr(e)[:300])
