"""Host logic of the product trainer (gating, RNG order, manual backward wiring, loss weighting, history
matching, Adam bookkeeping) checked on CPU against the oracle, with the torch test double standing in for
the CUDA op-set (tests/ops_torch.py).  The CUDA kernels themselves are covered by the -m gpu tests.

Two flavours:
  * fp64: oracle and product both in double precision -> they must agree to ~1e-9 everywhere (gradients
    and every post-step parameter).  This is the exact check of the hand-written backward wiring.
  * fp32: agreement with the oracle AND the reference's golden numbers at the fp32 noise floor.  The noise
    floor of this network's gradients is ~1.5e-3 relative (fp32 vs fp64 oracle): ReLU/LeakyReLU flips and the
    tanh(10*tanh(.)) mask head amplify rounding, and Adam's first step is lr*sign(g), so individual
    parameters may differ by 2*lr; the post-step check is therefore statistical.
"""
import pytest
import torch

import council_oracle as co
from common import close, load_golden, setup_case
from council_gan_b200.trainer_council import Council_Trainer
from ops_torch import TorchOps

_randn = torch.randn


def _randn32(dtype):
    # style noise is always drawn as float32 (like the reference) and then cast
    return lambda *a, **k: _randn(*a, dtype=torch.float32, **k).to(dtype)


def load_states(tr, states):
    for name, lst in states.items():
        fam, d = name.rsplit('_', 1)
        mods = getattr(tr, '%s_%s_s' % (fam, d))
        for i, sd in enumerate(lst):
            mods[i].load_state_dict(sd)


def run_oracle(gold, dtype=torch.float32, n_iters=1, on_iter=None):
    """n_iters consecutive iterations on the same images (as oracle/make_golden.py runs the reference): config['iteration']
    advances and StepLR steps after every iteration (train.py:241-250,399)."""
    hp, states, x_a, x_b = setup_case(gold)
    states = {k: [{kk: vv.to(dtype) for kk, vv in sd.items()} for sd in lst] for k, lst in states.items()}
    torch.randn = _randn32(dtype)
    try:
        tr = co.OracleTrainer(hp, states)
        co.seed_all(gold['rng_seed'])
        for k in range(n_iters):
            hp['iteration'] = gold['iteration'] + k
            tr.dis_update(x_a.to(dtype), x_b.to(dtype), hp)
            tr.disc_ran = tr.dis_council_update(x_a.to(dtype), x_b.to(dtype), hp)
            tr.gen_update(x_a.to(dtype), x_b.to(dtype), hp, hp['iteration'])
            if on_iter is not None:
                on_iter(k, tr)
            if n_iters > 1:
                tr.update_learning_rate()
    finally:
        torch.randn = _randn
    return tr, hp


def run_product_iteration(gold, ops, n_iters=1, on_iter=None, trainer=None):
    hp, states, x_a, x_b = setup_case(gold)
    co.seed_all(hp['random_seed'])
    tr = trainer
    if tr is None:
        tr = Council_Trainer(hp, str(ops.device), _ops=ops)
        load_states(tr, states)
        co.seed_all(gold['rng_seed'])
    torch.randn = _randn32(torch.float32)
    try:
        for k in range(n_iters):
            hp['iteration'] = gold['iteration'] + k
            tr.dis_update(x_a, x_b, hp)
            tr.loss_dis_council_total_s = None
            tr.dis_council_update(x_a, x_b, hp)
            tr.gen_update(x_a, x_b, hp, hp['iteration'])
            if on_iter is not None:
                on_iter(k, tr)
            if n_iters > 1:
                tr.update_learning_rate()
    finally:
        torch.randn = _randn
    return tr, hp


def compare_with_oracle(tr, orc, hp, rtol_loss, grad_rel_l2, flip_frac, min_cos=None, shallow_only=False):
    """Losses, generator gradients (relative L2 per tensor) and every post-step parameter.
    shallow_only (TF32 runs): relative-L2 / flip checks only for tensors a short backward chain away from the loss
    (the discriminators and the generator head); the deep generator gradients, which amplify ANY rounding
    difference by ~1e5 (fp32 vs fp64 oracle already differ by 1.5e-3), are checked by direction (cosine)."""
    N = tr.council_size
    d0 = orc.dirs[0]
    for i in range(N):
        assert close(float(tr.loss_dis_total_s[i]), float(orc.loss_dis_total_s[i]), rtol_loss), ('dis', i)
        if orc.disc_ran:
            assert close(float(tr.loss_dis_council_total_s[i]), float(orc.loss_dis_council_total_s[i]), rtol_loss), ('disc', i)
        else:
            assert tr.loss_dis_council_total_s is None
        assert close(float(tr.loss_gen_total_s[i]), float(orc.loss_gen_total_s[i]), rtol_loss), \
            ('gen', i, float(tr.loss_gen_total_s[i]), float(orc.loss_gen_total_s[i]))
    worst_g, worst_p = 0.0, 0.0
    tr.synchronize()
    for fam, dd in [(f, d) for d in orc.dirs for f in ('gen', 'dis', 'dis_council')]:
        name = '%s_%s' % (fam, dd)
        if name not in orc.P:
            continue
        net = tr._nets[name]
        dead = getattr(net, 'dead_bias', set())
        for i in range(N):
            sd = getattr(tr, name + '_s')[i].state_dict()
            for spec in net._specs():
                for key, is_w in ((spec.wname, True), (spec.bname, False)):
                    if key in dead:
                        continue
                    ref = orc.P[name][i][key].detach()
                    if key.startswith('enc_style'):
                        assert torch.equal(sd[key].cpu().to(ref.dtype), ref), key  # never stepped
                        continue
                    deep = shallow_only and fam == 'gen'
                    diff = (sd[key].cpu().to(ref.dtype) - ref).abs()
                    worst_p = max(worst_p, diff.max().item())
                    frac = (diff > 0.5 * hp['lr']).double().mean().item()
                    if not deep:
                        assert frac <= flip_frac, (fam, i, key, 'fraction of parameters off by > lr/2', frac)
                    og = orc.P[name][i][key].grad
                    if fam == 'gen' and og is not None:
                        bank = net._bank_of(key)
                        g = bank.g(key)[i]
                        g = (spec.export_weight(g) if is_w else g).cpu().to(og.dtype)
                        rel = ((g - og).norm() / (og.norm() + 1e-30)).item()
                        if not deep:
                            worst_g = max(worst_g, rel)
                            assert rel <= grad_rel_l2, (fam, i, key, 'relative L2 gradient error', rel)
                        if min_cos is not None and og.numel() > 64:
                            cos = ((g * og).sum() / (g.norm() * og.norm() + 1e-30)).item()
                            assert cos >= min_cos, (fam, i, key, 'gradient direction (cosine)', cos)
    return worst_g, worst_p


@pytest.mark.parametrize('case', ['glasses64_n2_b2_early', 'anime64_n3_b2', 'm2f64_n4_b2', 'glasses64_n2_b2_both'])
def test_host_logic_exact_in_fp64(case):
    gold = load_golden(case)
    torch.set_num_threads(8)
    orc, hp = run_oracle(gold, torch.float64)
    tr, _ = run_product_iteration(gold, TorchOps('cpu', torch.float64))
    wg, wp = compare_with_oracle(tr, orc, hp, rtol_loss=1e-7, grad_rel_l2=1e-7, flip_frac=0.0)  # losses are published as fp32 tensors
    print(case, 'fp64 worst grad relL2 %.2e, worst post-step parameter diff %.2e' % (wg, wp))


@pytest.mark.parametrize('case', ['glasses64_n2_b2_early', 'm2f64_n4_b2', 'm2f256_n2_b1'])
def test_host_logic_fp32_vs_oracle_and_golden(case):
    gold = load_golden(case)
    torch.set_num_threads(8)
    orc, hp = run_oracle(gold, torch.float32)
    tr, _ = run_product_iteration(gold, TorchOps('cpu'))
    wg, wp = compare_with_oracle(tr, orc, hp, rtol_loss=5e-5, grad_rel_l2=2e-2, flip_frac=0.02)
    print(case, 'fp32 worst grad relL2 %.2e, worst post-step parameter diff %.2e' % (wg, wp))
    for i in range(tr.council_size):  # and the reference's own numbers
        assert close(float(tr.loss_gen_total_s[i]), gold['loss_gen_total'][i], 1e-4)
        assert close(float(tr.loss_dis_total_s[i]), gold['loss_dis_total'][i], 1e-4)
        if gold['dis_council_ran']:
            assert close(float(tr.loss_dis_council_total_s[i]), gold['loss_dis_council_total'][i], 1e-4)


def test_three_iterations_with_flips_and_lr_decay_exact_in_fp64():
    """State carried from one iteration to the next (loss-history rings on the device, Adam moments / step counts, StepLR
    epoch, the council on/off flip, the per-iteration caches) against the oracle, three iterations, exact in fp64; and the
    per-iteration losses of the product in fp32 against the reference's own numbers (golden fixture)."""
    gold = load_golden('glasses64_n2_b2_iter3')
    torch.set_num_threads(8)
    olog, plog = [], []

    def grab(log):
        def f(k, tr):
            log.append(([float(v) for v in tr.loss_dis_total_s], [float(v) for v in tr.loss_gen_total_s],
                        None if not getattr(tr, 'disc_ran', tr.loss_dis_council_total_s is not None)
                        else [float(v) for v in tr.loss_dis_council_total_s]))
        return f
    orc, hp = run_oracle(gold, torch.float64, n_iters=3, on_iter=grab(olog))
    tr, _ = run_product_iteration(gold, TorchOps('cpu', torch.float64), n_iters=3, on_iter=grab(plog))
    for k in range(3):
        for a, b in zip(olog[k][0] + olog[k][1] + (olog[k][2] or []), plog[k][0] + plog[k][1] + (plog[k][2] or [])):
            assert close(a, b, 1e-7), (k, a, b)
        assert (olog[k][2] is None) == (not gold['iters'][k]['dis_council_ran'])
        assert (plog[k][2] is None) == (olog[k][2] is None)
    # third-iteration gradients: fp64 rounding (1e-16) amplified by ~1e5 per iteration through the mask head
    wg, wp = compare_with_oracle(tr, orc, hp, rtol_loss=1e-7, grad_rel_l2=1e-5, flip_frac=1e-3)
    assert abs(tr._lr('gen') - gold['lr_after']['gen'][0]) < 1e-15 and abs(orc.lr_now() - gold['lr_after']['gen'][0]) < 1e-15
    # histories: the device rings hold what the reference's deques hold
    for i in range(tr.council_size):
        assert [float(v) for v in list(tr.los_hist_gan_a2b_s[i])[-3:]] == pytest.approx([float(v) for v in list(orc.hist_gan['a2b'][i])[-3:]], rel=1e-6)
        assert [float(v) for v in list(tr.los_hist_council_a2b_s[i])[-2:]] == pytest.approx([float(v) for v in list(orc.hist_council['a2b'][i])[-2:]], rel=1e-6)
        assert len(tr.los_hist_council_a2b_s[i]) == hp['loss_matching_hist_size']
    # fp32 product vs the unmodified reference's per-iteration numbers
    plog32 = []
    run_product_iteration(gold, TorchOps('cpu'), n_iters=3, on_iter=grab(plog32))
    for k in range(3):
        g = gold['iters'][k]
        # fp32 noise grows from iteration to iteration (tests/test_oracle_golden.py); the product's hand-written backward rounds
        # differently from autograd, so Adam's sign-like first steps differ on the smallest gradients already after iteration 1
        tol = [1e-4, 2e-3, 1e-2][k]
        for a, b in zip(plog32[k][0], g['loss_dis_total']):
            assert close(a, b, tol), ('dis', k, a, b)
        for a, b in zip(plog32[k][1], g['loss_gen_total']):
            assert close(a, b, tol), ('gen', k, a, b)


def test_write_loss_reflection_like_the_reference():
    """utils.write_loss (utils.py:277-305) reflects over every non-callable trainer attribute whose name contains 'loss', 'grad',
    'conf', 'nwd' or 'do' and hands it to add_scalar(s): each must be a scalar, a bool or a plain list of scalars / 0-d tensors."""
    gold = load_golden('glasses64_n2_b2_early')
    gold = dict(gold, iteration=20001)
    tr, hp = run_product_iteration(gold, TorchOps('cpu'))
    seen = {}

    class Writer:
        def add_scalar(self, name, value, it):
            v = float(value)  # numbers, numpy scalars, 0-d tensors
            seen[name] = v

        def add_scalars(self, name, d, it):
            for k, v in d.items():
                seen[name + '/' + k] = float(v)

    def write_loss(iterations, trainer, train_writer):  # same reflection rule as the reference
        members = [attr for attr in dir(trainer)
                   if not callable(getattr(trainer, attr)) and not attr.startswith("__") and
                   ('loss' in attr or 'grad' in attr or 'conf' in attr or 'nwd' in attr or 'do' in attr)]
        for m in members:
            val = getattr(trainer, m)
            if type(val) is bool:
                val = 1 if val else 0
            if type(val) is list:
                train_writer.add_scalars(m, {str(i): (x.data.cpu().numpy() if type(x) is torch.Tensor else x) for i, x in enumerate(val)},
                                         iterations + 1)
            else:
                train_writer.add_scalar(m, val, iterations + 1)
    write_loss(0, tr, Writer())
    assert 'loss_gen_total_s/0' in seen and 'loss_dis_total_s/1' in seen and 'loss_dis_a2b_s/0' in seen
    assert 'w_match_a2b_conf' in seen and 'council_w_conf' in seen
    assert not hasattr(tr, 'loss_dis_b2a_s')  # the reference only defines it for an active direction (:742-746)
    assert all(type(getattr(tr, a)) is list for a in ('loss_gen_total_s', 'loss_dis_total_s', 'loss_gen_adv_a2b_s'))
