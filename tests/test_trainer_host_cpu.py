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


def run_oracle(gold, dtype=torch.float32):
    hp, states, x_a, x_b = setup_case(gold)
    states = {k: [{kk: vv.to(dtype) for kk, vv in sd.items()} for sd in lst] for k, lst in states.items()}
    torch.randn = _randn32(dtype)
    try:
        tr = co.OracleTrainer(hp, states)
        co.seed_all(gold['rng_seed'])
        tr.dis_update(x_a.to(dtype), x_b.to(dtype), hp)
        tr.disc_ran = tr.dis_council_update(x_a.to(dtype), x_b.to(dtype), hp)
        tr.gen_update(x_a.to(dtype), x_b.to(dtype), hp, gold['iteration'])
    finally:
        torch.randn = _randn
    return tr, hp


def run_product_iteration(gold, ops):
    hp, states, x_a, x_b = setup_case(gold)
    co.seed_all(hp['random_seed'])
    tr = Council_Trainer(hp, str(ops.device), _ops=ops)
    load_states(tr, states)
    co.seed_all(gold['rng_seed'])
    torch.randn = _randn32(torch.float32)
    try:
        tr.dis_update(x_a, x_b, hp)
        tr.loss_dis_council_total_s = None
        tr.dis_council_update(x_a, x_b, hp)
        tr.gen_update(x_a, x_b, hp, gold['iteration'])
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
    for fam in ('gen', 'dis', 'dis_council'):
        name = '%s_%s' % (fam, d0)
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


@pytest.mark.parametrize('case', ['glasses64_n2_b2_early', 'anime64_n3_b2', 'm2f64_n4_b2'])
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
