"""Parity of the CUDA training step with the CPU oracle and the reference's golden numbers (through the
C ABI).  Tolerances follow BASELINE.json's north_star: per-step losses within 1e-3 relative; pixel MAE of the
generated image reported and bounded; post-step parameters compared statistically (Adam's first step is
lr*sign(g), see tests/test_trainer_host_cpu.py)."""
import pytest
import torch

import council_oracle as co
from common import close, load_golden, setup_case
from test_trainer_host_cpu import compare_with_oracle, load_states, run_oracle

pytestmark = pytest.mark.gpu


def run_cuda(gold, tc):
    from council_gan_b200 import Council_Trainer
    hp, states, x_a, x_b = setup_case(gold)
    co.seed_all(hp['random_seed'])
    tr = Council_Trainer(hp, 'cuda:0')
    tr.ops.set_tensor_core_mode(tc)
    load_states(tr, states)
    co.seed_all(gold['rng_seed'])
    tr.dis_update(x_a, x_b, hp)
    tr.loss_dis_council_total_s = None
    tr.dis_council_update(x_a, x_b, hp)
    tr.gen_update(x_a, x_b, hp, gold['iteration'])
    torch.cuda.synchronize()
    tr.ops.set_tensor_core_mode(1)
    return tr, hp


@pytest.mark.parametrize('case', ['glasses64_n2_b2_early', 'anime64_n3_b2', 'm2f64_n4_b2', 'glasses128_n2_b1', 'glasses64_n2_b2_both',
                                  'm2f256_n2_b1'])
@pytest.mark.parametrize('tc', [0, 1])
def test_iteration_matches_oracle_and_golden(case, tc):
    check_iteration(case, tc)


def check_iteration(case, tc):
    gold = load_golden(case)
    torch.set_num_threads(max(1, torch.get_num_threads()))
    orc, hp = run_oracle(gold, torch.float32)
    tr, _ = run_cuda(gold, tc)
    d0 = orc.dirs[0]
    N = tr.council_size
    # losses vs the reference's own numbers (golden) -- the north-star gate: 1e-3 relative
    for i in range(N):
        assert close(float(tr.loss_dis_total_s[i]), gold['loss_dis_total'][i], 1e-3), ('dis', i)
        assert close(float(tr.loss_gen_total_s[i]), gold['loss_gen_total'][i], 1e-3), \
            ('gen', i, float(tr.loss_gen_total_s[i]), gold['loss_gen_total'][i])
        if gold['dis_council_ran']:
            assert close(float(tr.loss_dis_council_total_s[i]), gold['loss_dis_council_total'][i], 1e-3), ('disc', i)
    # pixel MAE of the generated images of gen_update vs the oracle
    for i in range(N):
        xf = tr.ops.nhwc_to_nchw(tr._last_fw[d0]['x_fake'][i], 3).cpu()
        mae = (xf - orc.x_fake_gen[d0][i].detach()).abs().mean().item()
        assert mae < (2e-4 if tc == 0 else 3e-3), ('pixel MAE', i, mae)
    if tc == 0:
        wg, wp = compare_with_oracle(tr, orc, hp, rtol_loss=1e-3, grad_rel_l2=3e-2, flip_frac=0.03, min_cos=0.999)
    else:
        # TF32 operands: scripts/grad_noise.py (profiles/r01_grad_noise.log) shows that perturbing the WEIGHTS by
        # 2^-11 relative noise with exact fp32 kernels already moves the deep generator gradients by 15 % (cos 0.989)
        # -- the same as the tensor-core path does (13.5 %, cos 0.991) -- and with the focus loss live
        # (sign(m-.5)/(|m-.5|+eps)^2 on masks that start at ~0.5) the gradient is discontinuous in the mask.  So the
        # gradient direction is only asserted for the case without focus loss; D / DC (short chains) are checked
        # statistically in every case.
        no_focus = case in ('glasses64_n2_b2_early', 'anime64_n3_b2')  # focus gate closed / focus weights 0
        wg, wp = compare_with_oracle(tr, orc, hp, rtol_loss=1e-3, grad_rel_l2=1.0, flip_frac=0.2,
                                     min_cos=0.98 if no_focus else None, shallow_only=True)
    print('%s tc=%d: worst generator grad relL2 %.2e' % (case, tc, wg))


def test_member_api_encode_decode():
    """gen_a2b_s[i].encode / decode (reference API, NCHW in/out) agree with the oracle networks."""
    from council_gan_b200 import Council_Trainer
    gold = load_golden('glasses64_n2_b2_early')
    hp, states, x_a, x_b = setup_case(gold)
    tr = Council_Trainer(hp, 'cuda:0')
    tr.ops.set_tensor_core_mode(0)
    load_states(tr, states)
    s = torch.randn(x_a.size(0), hp['gen']['style_dim'], 1, 1)
    for i in range(2):
        p = states['gen_a2b'][i]
        c, s_fake = tr.gen_a2b_s[i].encode(x_a)
        oc = co.content_encode(p, hp, x_a)
        assert (c.cpu() - oc).abs().max().item() < 2e-3 * oc.abs().max().item()
        assert (s_fake.cpu() - co.style_encode(p, hp, x_a)).abs().max().item() < 1e-3
        out, mask = tr.gen_a2b_s[i].decode(oc, s, x_a, return_mask=True)
        oo, om = co.decode(p, hp, oc, s, x_a)
        assert (out.cpu() - oo).abs().mean().item() < 5e-4 and (mask.cpu() - om).abs().mean().item() < 5e-4
        assert tr.gen_a2b_s[i].dec.mask_s is mask
        sd = tr.gen_a2b_s[i].state_dict()
        assert list(sd.keys()) == list(p.keys())
        for k in p:
            assert torch.equal(sd[k].cpu(), p[k]), k
    tr.ops.set_tensor_core_mode(1)


def _run_cuda_iters(gold, tc, n_iters, on_iter=None, tr=None, first=0):
    from council_gan_b200 import Council_Trainer
    hp, states, x_a, x_b = setup_case(gold)
    if tr is None:
        co.seed_all(hp['random_seed'])
        tr = Council_Trainer(hp, 'cuda:0')
        load_states(tr, states)
        co.seed_all(gold['rng_seed'])
    tr.ops.set_tensor_core_mode(tc)
    for k in range(first, first + n_iters):
        hp['iteration'] = gold['iteration'] + k
        tr.dis_update(x_a, x_b, hp)
        tr.loss_dis_council_total_s = None
        tr.dis_council_update(x_a, x_b, hp)
        tr.gen_update(x_a, x_b, hp, hp['iteration'])
        if on_iter is not None:
            on_iter(k, tr)
        tr.update_learning_rate()
    torch.cuda.synchronize()
    tr.ops.set_tensor_core_mode(1)
    return tr, hp


@pytest.mark.parametrize('tc', [0, 1])
def test_three_iterations_state_carry_and_resume(tc, tmp_path):
    """Three consecutive iterations on the GPU (council flip 2 on / 1 off, StepLR step 2, device-side loss histories, cached
    encodings, deferred Adam) against the reference's per-iteration numbers; then save() after iteration 2 -> resume() in a fresh
    trainer -> iteration 3 must reproduce the uninterrupted run."""
    import random
    import numpy as np
    gold = load_golden('glasses64_n2_b2_iter3')
    log = []

    def grab(k, tr):
        log.append(([float(v) for v in tr.loss_dis_total_s], [float(v) for v in tr.loss_gen_total_s],
                    tr.loss_dis_council_total_s is not None, float(tr.w_match_a2b_conf)))
    tr, hp = _run_cuda_iters(gold, tc, 2, grab)
    tr.save(str(tmp_path), gold['iteration'] + 1)
    rng = (random.getstate(), np.random.get_state(), torch.get_rng_state())
    tr, hp = _run_cuda_iters(gold, tc, 1, grab, tr=tr, first=2)
    # fp32 noise grows from iteration to iteration (tests/test_oracle_golden.py: fp64 vs fp32 oracle differ by 3e-3 at the third);
    # TF32 operands start from a larger per-step difference
    tol = [1e-3, 3e-3, 2e-2] if tc == 0 else [1e-3, 1e-2, 5e-2]
    for k in range(3):
        g = gold['iters'][k]
        assert log[k][2] == g['dis_council_ran'], k
        for a, b in zip(log[k][0], g['loss_dis_total']):
            assert close(a, b, tol[k]), ('dis', k, a, b)
        for a, b in zip(log[k][1], g['loss_gen_total']):
            assert close(a, b, tol[k]), ('gen', k, a, b)
    assert close(log[1][3], gold['iters'][1]['w_match'], 1e-3)
    assert abs(tr._lr('gen') - gold['lr_after']['gen'][0]) < 1e-15
    # resume: fresh trainer, same checkpoint files, same RNG state -> identical third iteration
    from council_gan_b200 import Council_Trainer
    hp2, _, x_a, x_b = setup_case(gold)
    tr2 = Council_Trainer(hp2, 'cuda:0')
    it = tr2.resume(str(tmp_path), hp2)
    assert it == gold['iteration'] + 2
    for d_ in tr._dirs:  # loss histories are not part of the reference's checkpoint either: carry them over for the comparison
        for kk in ('gan', 'council'):
            tr2._rings[d_][kk].copy_(tr._rings[d_][kk])
        tr2._rings[d_]['head_gan'], tr2._rings[d_]['head_council'] = tr._rings[d_]['head_gan'] - 1, tr._rings[d_]['head_council']
    tr2._sched_epoch = {k: 2 for k in tr2._sched_epoch}
    random.setstate(rng[0])
    np.random.set_state(rng[1])
    torch.set_rng_state(rng[2])
    log2 = []
    _run_cuda_iters(gold, tc, 1, lambda k, t: log2.append([float(v) for v in t.loss_dis_total_s]), tr=tr2, first=2)
    for a, b in zip(log2[0], log[2][0]):
        assert close(a, b, 1e-6), ('dis after resume', a, b)
    tr.synchronize()
    tr2.synchronize()
    assert (tr._nets['dis_a2b'].bank.data - tr2._nets['dis_a2b'].bank.data).abs().max().item() < 1e-6
    assert (tr._nets['gen_a2b'].bank.exp_avg - tr2._nets['gen_a2b'].bank.exp_avg).abs().max().item() < 1e-5


BIG_CASES = ['m2f256_n4_b8', 'anime256_n4_b4', 'm2f512_n6_b2']


@pytest.mark.parametrize('case', BIG_CASES)
def test_baseline_configuration_vs_reference_golden(case):
    """BASELINE.json configs[1], [2] and [4] (per GPU) at their REAL council size, batch and resolution, on the default tensor-core
    path, against numbers of the unmodified reference (tests/golden, oracle/make_golden.py): every per-step loss within 1e-3,
    post-step parameters within Adam's first-step envelope, and a fresh forward of the updated generator."""
    from make_golden import PROBE_PARAMS
    from common import probe
    gold = load_golden(case)
    tr, hp = run_cuda(gold, 1)
    N = tr.council_size
    d0 = tr._dirs[0]
    ab = 'ab' if d0 == 'a2b' else 'ba'
    for i in range(N):
        assert close(float(tr.loss_dis_total_s[i]), gold['loss_dis_total'][i], 1e-3), ('dis', i)
        assert close(float(tr.loss_dis_council_total_s[i]), gold['loss_dis_council_total'][i], 1e-3), ('disc', i)
        assert close(float(tr.loss_gen_total_s[i]), gold['loss_gen_total'][i], 1e-3), \
            ('gen', i, float(tr.loss_gen_total_s[i]), gold['loss_gen_total'][i])
        assert close(float(getattr(tr, 'loss_gen_adv_%s_s' % d0)[i]), gold['loss_gen_adv'][i], 1e-3), ('adv', i)
        assert close(float(getattr(tr, 'council_loss_%s_s' % ab)[i]), gold['council_loss'][i], 1e-3), ('council', i)
        if gold['loss_gen_mask_zero_one']:
            assert close(float(getattr(tr, 'loss_gen_mask_zero_one_%s_s' % ab)[i]), gold['loss_gen_mask_zero_one'][i], 1e-3), ('z01', i)
            assert close(float(getattr(tr, 'loss_gen_mask_total_%s_s' % ab)[i]), gold['loss_gen_mask_total'][i], 3e-3), ('mtot', i)
    assert close(float(getattr(tr, 'w_match_%s_conf' % d0)), gold['w_match'], 1e-4)
    lr = hp['lr']
    worst = 0.0
    for fam in ('gen', 'dis', 'dis_council'):
        for i in range(N):
            sd = getattr(tr, '%s_%s_s' % (fam, d0))[i].state_dict()
            for key in PROBE_PARAMS[fam]:
                rec = gold['params']['%s.%d.%s' % (fam, i, key)]
                got = probe(sd[key])
                # Adam's first step moves every parameter by ~lr*sign(g): samples agree within 2*lr, norms within lr-sized slack
                for a, b in zip(got['samples'], rec['post']['samples']):
                    worst = max(worst, abs(a - b))
                    assert abs(a - b) <= 2.1 * lr + 1e-6, (fam, i, key, a, b)
                assert abs(got['absmean'] - rec['post']['absmean']) <= 1.0 * lr + 1e-4 * abs(rec['post']['absmean']), (fam, i, key)
                # gradient norm of the head layer: only without focus loss -- sign(m-.5)/(|m-.5|+eps)^2 on masks near 0.5 makes the
                # gradient discontinuous in the mask (TF32 vs fp32 pixels flip sides; profiles/r01_grad_noise.log)
                if fam == 'gen' and 'grad' in rec and not gold['loss_gen_mask_zero_one'] and key in ('dec.model.9.conv.weight', 'dec.model.9.conv.bias'):
                    g = tr._nets['gen_' + d0]
                    spec = [s for s in g._specs() if key in (s.wname, s.bname)][0]
                    gg = g.bank.g(key)[i]
                    gg = spec.export_weight(gg) if key.endswith('weight') else gg
                    assert close(probe(gg)['l2'], rec['grad']['l2'], 0.1), (key, probe(gg)['l2'], rec['grad']['l2'])  # norm of the head gradient (TF32 chain through D / DC)
    # a fresh forward of member 0 after the iteration, with the fixture's style seed
    g0 = getattr(tr, 'gen_%s_s' % d0)[0]
    _, _, x_a, x_b = setup_case(gold)
    src = x_a if d0 == 'a2b' else x_b
    c, _ = g0.encode(src)
    s = torch.randn(gold['batch'], hp['gen']['style_dim'], 1, 1, generator=torch.Generator().manual_seed(5))
    xf, mask = g0.decode(c, s, src, return_mask=True)
    for got, want in ((probe(xf, 16), gold['post_x_fake0']), (probe(mask, 16), gold['post_mask0'])):
        assert close(got['absmean'], want['absmean'], 5e-3), (got['absmean'], want['absmean'])
        # TF32 forward of a generator whose parameters each moved by +-lr: the existing per-pixel bound is MAE < 3e-3
        assert close(got['mean'], want['mean'], 5e-3, 2e-3), (got['mean'], want['mean'])
    print('%s: worst post-step parameter sample difference %.2e (lr %.1e)' % (case, worst, lr))


def test_batched_sample_on_gpu_matches_oracle_decode():
    """sample(): all members x all images as one stacked pass on the CUDA kernels; rows against the oracle's own
    encode/decode of the same member and image."""
    from council_gan_b200 import Council_Trainer
    gold = load_golden('glasses64_n2_b2_early')
    hp, states, x_a, x_b = setup_case(gold)
    tr = Council_Trainer(hp, 'cuda:0')
    tr.ops.set_tensor_core_mode(0)
    load_states(tr, states)
    torch.manual_seed(3)
    out = tr.sample(x_a, x_b)
    torch.manual_seed(3)
    s2 = torch.randn(x_a.size(0), hp['gen']['style_dim'], 1, 1)
    N, B = tr.council_size, x_a.size(0)
    assert out[4] is None and out[0].shape == (B * N, 3, 64, 64)
    for i in range(B):
        for j in range(N):
            p = states['gen_a2b'][j]
            oc = co.content_encode(p, hp, x_a[i:i + 1])
            o1, m1 = co.decode(p, hp, oc, tr.s_b[i:i + 1].cpu(), x_a[i:i + 1])
            o2, _ = co.decode(p, hp, oc, s2[i:i + 1], x_a[i:i + 1])
            r = i * N + j
            assert (out[1][r].cpu() - m1[0]).abs().mean().item() < 5e-4 and (out[2][r].cpu() - o1[0]).abs().mean().item() < 5e-4
            assert (out[3][r].cpu() - o2[0]).abs().mean().item() < 5e-4
            assert torch.equal(out[0][r].cpu(), x_a[i])
    tr.ops.set_tensor_core_mode(1)
