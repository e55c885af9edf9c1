"""Shared helpers for the test-suite (oracle drivers, golden loading)."""
import json
import os

import torch
import yaml

import council_oracle as co

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def load_golden(case):
    with open(os.path.join(GOLDEN, case + '.json')) as f:
        return json.load(f)


def config_for(gold_or_name, overrides=None):
    if isinstance(gold_or_name, dict):
        name, overrides = gold_or_name['config'], gold_or_name['overrides']
    else:
        name = gold_or_name
    hp = yaml.safe_load(open(os.path.join(ROOT, 'configs', name + '.yaml')))
    for k, v in (overrides or {}).items():
        d = hp
        ks = k.split('.')
        for kk in ks[:-1]:
            d = d[kk]
        d[ks[-1]] = v
    return hp


def probe(t, n=8):
    f = t.detach().double().flatten().cpu()
    idx = torch.linspace(0, f.numel() - 1, n).long()
    return {'mean': f.mean().item(), 'absmean': f.abs().mean().item(), 'l2': f.norm().item(),
            'samples': [f[i].item() for i in idx]}


def setup_case(gold):
    """(hp, states, x_a, x_b) for a golden case, exactly as oracle/make_golden.py builds them."""
    hp = config_for(gold)
    hp['batch_size'] = gold['batch']
    hp['iteration'] = gold['iteration']
    states = co.synth_all_states(hp, seed=gold['state_seed'])
    x_a, x_b = co.synth_inputs(gold['batch'], gold['size'], seed=gold['input_seed'])
    return hp, states, x_a, x_b


def run_oracle_iteration(gold):
    """One training iteration of the oracle on a golden case; returns the trainer."""
    hp, states, x_a, x_b = setup_case(gold)
    tr = co.OracleTrainer(hp, states)
    co.seed_all(gold['rng_seed'])
    tr.dis_update(x_a, x_b, hp)
    tr.disc_ran = tr.dis_council_update(x_a, x_b, hp)
    tr.gen_update(x_a, x_b, hp, gold['iteration'])
    return tr, hp, x_a, x_b


def close(a, b, rtol, atol=0.0):
    return abs(a - b) <= atol + rtol * max(abs(a), abs(b))
