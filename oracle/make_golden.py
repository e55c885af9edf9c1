"""Generate golden fixtures from the UNMODIFIED reference  --  TEST INFRASTRUCTURE ONLY.

Runs in the build container only (needs /root/reference, which does not exist on the GPU box).
Imports the reference read-only with the two out-of-tree shims of SURVEY.md section 8c:
  1. a stub ``torchfile`` module (utils.py:6 imports it; used only by load_vgg16, dead at vgg_w=0);
  2. ``Tensor.cuda(dev)`` / ``Module.cuda(dev)`` rebound to ``.to(dev)`` so ``cuda_device='cpu'`` works;
  3. ``Tensor.cpu()`` returns a COPY, as it does for the CUDA tensors the reference is written for.  Without it the CPU run
     differs from the CUDA run from the second iteration on: trainer_council.py:578 appends
     ``dis_council_loss_ab.detach().cpu().numpy()`` to the loss history and :582/:589 then scale that very tensor in place --
     for a CPU tensor ``.cpu().numpy()`` is a view, so the history would record loss * w_match * council_w instead of the loss.
     (Single-iteration fixtures are bit-identical with or without this shim: the ratio is taken before the in-place scaling.)
Loads deterministic synthetic parameters (``council_oracle.synth_all_states``) into the reference's
``Council_Trainer``, runs one training iteration (dis_update -> dis_council_update -> gen_update,
train.py:241-250) and writes the observed values to ``tests/golden/<case>.json``.

    python oracle/make_golden.py            # regenerates every case in CASES
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import council_oracle as co  # noqa: E402

REF = os.environ.get('COUNCIL_REF_DIR', '/root/reference')

# case name -> (config yaml, overrides, image size, batch, iteration)
CASES = {
    # BASELINE.json configs[0]: glasses 128x128, council_size=2, batch=1, all gates open
    'glasses128_n2_b1': ('glasses', {'council.council_size': 2}, 128, 1, 20001),
    # same network before either gate opens (no council loss, no focus loss, no dis_council step)
    'glasses64_n2_b2_early': ('glasses', {'council.council_size': 2}, 64, 2, 100),
    # male2female hyper-parameters (K=4 with N=4 -> one duplicated peer; mask_tv_w=0) at a CPU-friendly size
    'm2f64_n4_b2': ('male2female', {}, 64, 2, 60001),
    # selfie2anime: b2a-only direction (gan_w asymmetry trainer_council.py:775-777), focus weights 0
    'anime64_n3_b2': ('selfie2anime', {'council.council_size': 3}, 64, 2, 2001),
    # BASELINE.json configs[1] at its real resolution (256x256: every layer geometry of the headline workload), council and batch
    # reduced so that the CPU reference and the oracle finish in seconds
    'm2f256_n2_b1': ('male2female', {'council.council_size': 2}, 256, 1, 60001),
    # ---- round 2: the BASELINE configurations at their real council size, batch and resolution -----------------------------
    # BASELINE.json configs[1]: male2female 256x256, council_size=4, batch=8 (the configuration the metric is quoted on)
    'm2f256_n4_b8': ('male2female', {}, 256, 8, 60001),
    # BASELINE.json configs[2]: selfie2anime 256x256, council_size=4, batch=4 (b2a only, gan_w asymmetry, no focus loss)
    'anime256_n4_b4': ('selfie2anime', {}, 256, 4, 2001),
    # BASELINE.json configs[4] per GPU: male2female 512x512, council_size=6, 2 images per GPU
    'm2f512_n6_b2': ('male2female', {'council.council_size': 6}, 512, 2, 60001),
    # both directions at once (do_a2b and do_b2a): loss accumulation over directions, two Adam banks per family
    'glasses64_n2_b2_both': ('glasses', {'council.council_size': 2, 'do_b2a': True}, 64, 2, 20001),
    # three consecutive iterations with the on/off flip of the council loss live (2 on / 1 off), StepLR decay between them
    # (step_size 2) and the loss histories evolving: pins every piece of state carried from one iteration to the next
    'glasses64_n2_b2_iter3': ('glasses', {'council.council_size': 2, 'council.flipOnOff': True,
                                          'council.flipOnOff_On_iteration': 2, 'council.flipOnOff_Off_iteration': 1,
                                          'step_size': 2}, 64, 2, 20001, 3),
}

PROBE_PARAMS = {
    'gen': ['enc_content.model.0.conv.weight', 'enc_content.model.3.model.2.model.1.conv.weight',
            'dec.model.0.model.0.model.0.conv.weight', 'dec.model.2.conv.weight', 'dec.model.9.conv.weight',
            'dec.model.9.conv.bias', 'mlp.model.2.fc.weight', 'mlp.model.0.fc.bias'],
    'dis': ['cnns.0.0.conv.weight', 'cnns.1.3.conv.weight', 'cnns.0.4.weight', 'cnns.1.4.bias'],
    'dis_council': ['cnns.0.0.conv.weight', 'cnns.0.4.weight', 'cnns.1.5.weight', 'cnns.1.2.conv.bias'],
}


def load_config(name, overrides):
    hp = yaml.safe_load(open(os.path.join(ROOT, 'configs', name + '.yaml')))
    for k, v in overrides.items():
        d = hp
        ks = k.split('.')
        for kk in ks[:-1]:
            d = d[kk]
        d[ks[-1]] = v
    return hp


def probe(t, n=8):
    """A few deterministic samples + moments of a tensor: enough to pin it, small enough to commit."""
    f = t.detach().double().flatten()
    idx = torch.linspace(0, f.numel() - 1, n).long()
    return {'mean': f.mean().item(), 'absmean': f.abs().mean().item(), 'l2': f.norm().item(),
            'samples': [f[i].item() for i in idx]}


def import_reference():
    sys.modules.setdefault('torchfile', types.ModuleType('torchfile'))
    sys.path.insert(0, REF)
    import torch.nn as nn
    torch.Tensor.cuda = lambda self, device=None, *a, **k: self.to(device if device is not None else 'cpu')
    nn.Module.cuda = lambda self, device=None: self.to(device if device is not None else 'cpu')
    _cpu = torch.Tensor.cpu
    torch.Tensor.cpu = lambda self, *a, **k: _cpu(self, *a, **k).clone() if self.device.type == 'cpu' else _cpu(self, *a, **k)
    from trainer_council import Council_Trainer
    return Council_Trainer


def _dir_losses(trainer, d):
    """The loss attributes gen_update publishes for direction d (trainer_council.py:302-322), as floats."""
    ab = 'ab' if d == 'a2b' else 'ba'
    return {'loss_gen_adv': [float(v) for v in getattr(trainer, 'loss_gen_adv_%s_s' % d)],
            'council_loss': [float(v) for v in getattr(trainer, 'council_loss_%s_s' % ab)],
            'loss_gen_mask_zero_one': [float(v) for v in getattr(trainer, 'loss_gen_mask_zero_one_%s_s' % ab)],
            'loss_gen_mask_total': [float(v) for v in getattr(trainer, 'loss_gen_mask_total_%s_s' % ab)],
            'loss_gen_mask_TV': [float(v) for v in getattr(trainer, 'loss_gen_mask_TV_%s_s' % ab)],
            'w_match': float(getattr(trainer, 'w_match_%s_conf' % d))}


def run_iteration(trainer, hp, x_a, x_b, iteration):
    """One pass of the train.py:241-250 sequence; returns what it published."""
    hp['iteration'] = iteration
    rec = {'iteration': iteration}
    trainer.dis_update(x_a, x_b, hp)
    rec['loss_dis_total'] = [float(v) for v in trainer.loss_dis_total_s]
    trainer.loss_dis_council_total_s = None
    trainer.dis_council_update(x_a, x_b, hp)
    ran = trainer.loss_dis_council_total_s is not None
    rec['dis_council_ran'] = ran
    rec['loss_dis_council_total'] = [float(v) for v in trainer.loss_dis_council_total_s] if ran else []
    trainer.gen_update(x_a, x_b, hp, iteration)
    rec['loss_gen_total'] = [float(v) for v in trainer.loss_gen_total_s]
    return rec


def run_case(Council_Trainer, case):
    cfg, overrides, size, batch, iteration = CASES[case][:5]
    n_iters = CASES[case][5] if len(CASES[case]) > 5 else 1
    hp = load_config(cfg, overrides)
    hp['batch_size'] = batch
    hp['iteration'] = iteration
    torch.set_num_threads(8)
    co.seed_all(hp['random_seed'])
    trainer = Council_Trainer(hp, 'cpu')
    states = co.synth_all_states(hp, seed=7)
    for name, lst in states.items():
        fam, d = name.rsplit('_', 1)
        mods = getattr(trainer, '%s_%s_s' % (fam, d))
        for i, sd in enumerate(lst):
            mods[i].load_state_dict(sd)
    x_a, x_b = co.synth_inputs(batch, size, seed=123)
    co.seed_all(hp['random_seed'] + 1)
    dirs = [d for d in ('a2b', 'b2a') if hp['do_' + d]]
    N = hp['council']['council_size']
    out = {'case': case, 'config': cfg, 'overrides': overrides, 'size': size, 'batch': batch,
           'iteration': iteration, 'state_seed': 7, 'input_seed': 123, 'rng_seed': hp['random_seed'] + 1,
           'torch': torch.__version__, 'reference': 'Onr/Council-GAN @ 7fe8f8a (unmodified, CPU, shims only)'}
    d0 = dirs[0]
    if n_iters > 1:
        # train.py:225-252,399: same images every iteration here (synthetic), config['iteration'] advances, StepLR steps
        out['n_iters'] = n_iters
        out['iters'] = []
        for k in range(n_iters):
            rec = run_iteration(trainer, hp, x_a, x_b, iteration + k)
            rec.update(_dir_losses(trainer, d0))
            out['iters'].append(rec)
            trainer.update_learning_rate()
        out['lr_after'] = {fam: [float(o.param_groups[0]['lr']) for o in getattr(trainer, fam + '_opt_s')]
                           for fam in ('gen', 'dis', 'dis_council')}
        out.update(out['iters'][-1])
        out['iteration'] = iteration
    else:
        out.update(run_iteration(trainer, hp, x_a, x_b, iteration))
        out.update(_dir_losses(trainer, d0))
    if len(dirs) > 1:
        out['dirs'] = {d: _dir_losses(trainer, d) for d in dirs}
        out['dirs']['a2b']['loss_dis'] = [float(v) for v in trainer.loss_dis_a2b_s]
        out['dirs']['b2a']['loss_dis'] = [float(v) for v in trainer.loss_dis_b2a_s]
    # post-iteration parameters (each family stepped n_iters times) and the grads of the last step
    post = {}
    for d in dirs:
        for fam in ('gen', 'dis', 'dis_council'):
            mods = getattr(trainer, '%s_%s_s' % (fam, d), None)
            if mods is None or len(mods) == 0:
                continue
            for i in range(N):
                sd = mods[i].state_dict()
                named = dict(mods[i].named_parameters())
                for key in PROBE_PARAMS[fam]:
                    rec = {'post': probe(sd[key])}
                    if fam == 'gen' and named[key].grad is not None:
                        rec['grad'] = probe(named[key].grad)
                    post[('%s.%d.%s' % (fam, i, key)) if d == d0 else ('%s_%s.%d.%s' % (fam, d, i, key))] = rec
    out['params'] = post
    # a fresh forward of member 0 AFTER the iteration (pins the updated generator end to end)
    with torch.no_grad():
        g0 = getattr(trainer, 'gen_%s_s' % d0)[0]
        src = x_a if d0 == 'a2b' else x_b
        c, _ = g0.encode(src)
        s = torch.randn(batch, hp['gen']['style_dim'], 1, 1, generator=torch.Generator().manual_seed(5))
        xf, mask = g0.decode(c, s, src, return_mask=True)
    out['post_x_fake0'] = probe(xf, 16)
    out['post_mask0'] = probe(mask, 16)
    return out


def main():
    Council_Trainer = import_reference()
    os.makedirs(os.path.join(ROOT, 'tests', 'golden'), exist_ok=True)
    cases = sys.argv[1:] or list(CASES)
    for case in cases:
        out = run_case(Council_Trainer, case)
        path = os.path.join(ROOT, 'tests', 'golden', case + '.json')
        with open(path, 'w') as f:
            json.dump(out, f, indent=1)
        print(case, 'dis', out['loss_dis_total'], 'disc', out['loss_dis_council_total'], 'gen', out['loss_gen_total'])


if __name__ == '__main__':
    main()
