"""Reference API surface beyond the three updates: checkpoint files, resume, sample(),
update_learning_rate, unsupported-path errors.  CPU, torch test double for the op-set."""
import os

import pytest
import torch

import council_oracle as co
from common import config_for, load_golden, setup_case
from council_gan_b200.trainer_council import Council_Trainer
from ops_torch import TorchOps
from test_trainer_host_cpu import load_states


def make(gold_name='glasses64_n2_b2_early'):
    gold = load_golden(gold_name)
    hp, states, x_a, x_b = setup_case(gold)
    co.seed_all(3)
    tr = Council_Trainer(hp, 'cpu', _ops=TorchOps('cpu'))
    load_states(tr, states)
    return tr, hp, states, x_a, x_b


def test_checkpoint_files_and_resume(tmp_path):
    tr, hp, states, x_a, x_b = make()
    co.seed_all(5)
    tr.dis_update(x_a, x_b, hp)
    tr.gen_update(x_a, x_b, hp, hp['iteration'])
    tr.save(str(tmp_path), 41)
    names = sorted(os.listdir(tmp_path))
    # reference naming: {a2b,b2a}_{gen,dis,dis_council}_{i}_{iter+1:08d}.pt + optimizer_{i}.pt  (trainer_council.py:969-992)
    for i in range(2):
        for fam in ('gen', 'dis', 'dis_council'):
            assert 'a2b_%s_%d_%08d.pt' % (fam, i, 42) in names
        assert 'optimizer_%d.pt' % i in names
    sd = torch.load(os.path.join(tmp_path, 'a2b_gen_0_%08d.pt' % 42))['a2b']
    ref_keys = [k[:-4] if k.endswith('#buf') else k for k, _ in co.gen_param_shapes(hp)]
    assert list(sd.keys()) == ref_keys
    for (k, shape), kk in zip(co.gen_param_shapes(hp), ref_keys):
        assert tuple(sd[kk].shape) == tuple(shape), kk
    # resume into a fresh trainer: parameters, Adam moments and the iteration come back
    co.seed_all(9)
    tr2 = Council_Trainer(hp, 'cpu', _ops=TorchOps('cpu'))
    it = tr2.resume(str(tmp_path), hp)
    assert it == 42
    for name, net in tr._nets.items():
        assert torch.equal(net.bank.data, tr2._nets[name].bank.data), name
        assert torch.equal(net.bank.exp_avg, tr2._nets[name].bank.exp_avg), name
        assert net.bank.step == tr2._nets[name].bank.step
    # the next update is identical
    co.seed_all(6)
    tr.dis_update(x_a, x_b, hp)
    co.seed_all(6)
    tr2.dis_update(x_a, x_b, hp)
    assert torch.equal(tr._nets['dis_a2b'].bank.data, tr2._nets['dis_a2b'].bank.data)


def test_sample_and_lr_schedule():
    tr, hp, states, x_a, x_b = make()
    out = tr.sample(x_a, x_b)
    assert len(out) == 8 and out[4] is None  # a2b only: (x_a, mask, x_ab1, x_ab2, None, None, None, None)
    n = x_a.size(0) * tr.council_size
    assert out[0].shape == (n, 3, 64, 64) and out[1].shape == (n, 3, 64, 64) and out[2].shape == out[3].shape == (n, 3, 64, 64)
    assert float(out[1].min()) >= 0 and float(out[1].max()) <= 1  # masks
    # StepLR(step_size, gamma): lr halves after step_size scheduler steps (utils.py:392-400)
    hp2 = dict(hp, step_size=3, gamma=0.5)
    tr3 = Council_Trainer(hp2, 'cpu', _ops=TorchOps('cpu'))
    lrs = []
    for _ in range(7):
        lrs.append(tr3._lr('gen'))
        tr3.update_learning_rate()
    assert lrs == [hp['lr']] * 3 + [hp['lr'] * 0.5] * 3 + [hp['lr'] * 0.25]


def test_unsupported_paths_raise():
    hp = config_for('glasses')
    for key in ('recon_x_w', 'vgg_w', 'council_abs_w'):
        with pytest.raises(NotImplementedError):
            Council_Trainer(dict(hp, **{key: 1}), 'cpu', _ops=TorchOps('cpu'))
    bad = dict(hp, dis=dict(hp['dis'], gan_type='nsgan'))
    with pytest.raises(AssertionError):
        Council_Trainer(bad, 'cpu', _ops=TorchOps('cpu'))
    tr = Council_Trainer(dict(hp, council=dict(hp['council'], council_size=2)), 'cpu', _ops=TorchOps('cpu'))
    with pytest.raises(NotImplementedError):
        tr.forward(torch.zeros(1, 3, 64, 64))


def test_gating_before_start_iterations():
    """dis_council_update is a no-op before council_start_at_iter; N<=1 prints the reference's message."""
    tr, hp, states, x_a, x_b = make()
    before = tr._nets['dis_council_a2b'].bank.data.clone()
    tr.loss_dis_council_total_s = 'untouched'
    tr.dis_council_update(x_a, x_b, hp)  # iteration 100 < 10000
    assert tr.loss_dis_council_total_s == 'untouched'
    assert torch.equal(before, tr._nets['dis_council_a2b'].bank.data)
