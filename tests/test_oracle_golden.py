"""The oracle restatement (oracle/council_oracle.py) must reproduce what the unmodified reference
produced (tests/golden/*.json, written by oracle/make_golden.py).  CPU only."""
import pytest
import torch

import council_oracle as co
from common import close, load_golden, probe, run_oracle_iteration
from make_golden import PROBE_PARAMS

CASES = ['glasses64_n2_b2_early', 'm2f64_n4_b2', 'anime64_n3_b2', 'glasses128_n2_b1', 'm2f256_n2_b1']
RTOL = 2e-5  # both sides are torch-CPU fp32; the slack covers thread-count dependent summation order


@pytest.mark.parametrize('case', CASES)
def test_oracle_matches_reference_golden(case):
    gold = load_golden(case)
    torch.set_num_threads(8)
    tr, hp, x_a, x_b = run_oracle_iteration(gold)
    d0 = tr.dirs[0]
    N = tr.N

    def chk(name, got, want, rtol=RTOL):
        assert len(got) == len(want), name
        for i, (g, w) in enumerate(zip(got, want)):
            assert close(float(g), w, rtol, 1e-7), '%s[%d]: oracle %r reference %r' % (name, i, float(g), w)

    chk('loss_dis_total', tr.loss_dis_total_s, gold['loss_dis_total'])
    assert tr.disc_ran == gold['dis_council_ran']
    if tr.disc_ran:
        chk('loss_dis_council_total', tr.loss_dis_council_total_s, gold['loss_dis_council_total'])
    chk('loss_gen_total', tr.loss_gen_total_s, gold['loss_gen_total'])
    chk('loss_gen_adv', tr.loss_gen_adv_s[d0], gold['loss_gen_adv'])
    if gold['council_loss'] and any(v != 0 for v in gold['council_loss']):
        chk('council_loss', tr.council_loss_s[d0], gold['council_loss'])
        assert close(float(tr.w_match[d0]), gold['w_match'], 1e-6)
    chk('mask01', tr.loss_gen_mask_zero_one_s[d0], gold['loss_gen_mask_zero_one'])
    if any(v != 0 for v in gold['loss_gen_mask_total']):
        chk('mask_total', tr.loss_gen_mask_total_s[d0], gold['loss_gen_mask_total'])
    if any(v != 0 for v in gold['loss_gen_mask_TV']):
        chk('mask_tv', tr.loss_gen_mask_TV_s[d0], gold['loss_gen_mask_TV'])

    # post-step parameters and generator grads
    for fam in ('gen', 'dis', 'dis_council'):
        name = '%s_%s' % (fam, d0)
        if name not in tr.P:
            continue
        for i in range(N):
            for key in PROBE_PARAMS[fam]:
                rec = gold['params']['%s.%d.%s' % (fam, i, key)]
                got = probe(tr.P[name][i][key])
                for f in ('mean', 'absmean', 'l2'):
                    assert close(got[f], rec['post'][f], 1e-4, 1e-9), (fam, i, key, f, got[f], rec['post'][f])
                for a, b in zip(got['samples'], rec['post']['samples']):
                    # Adam's first step moves every weight by ~lr*sign(g): compare with an lr-sized floor
                    assert abs(a - b) <= 2.1 * hp['lr'] * 1.0 + 1e-6, (fam, i, key, a, b)
                if 'grad' in rec:
                    gg = probe(tr.P[name][i][key].grad)
                    if key.endswith('.bias') and 'dec.model.9' not in key and 'mlp' not in key:
                        continue
                    assert close(gg['l2'], rec['grad']['l2'], 2e-3, 1e-9), (fam, i, key, gg['l2'], rec['grad']['l2'])

    # fresh forward of the updated member-0 generator
    with torch.no_grad():
        g0 = tr.state('gen_' + d0, 0)
        src = x_a if d0 == 'a2b' else x_b
        c = co.content_encode(g0, hp, src)
        s = torch.randn(gold['batch'], hp['gen']['style_dim'], 1, 1, generator=torch.Generator().manual_seed(5))
        xf, mask = co.decode(g0, hp, c, s, src)
    for got, want in ((probe(xf, 16), gold['post_x_fake0']), (probe(mask, 16), gold['post_mask0'])):
        assert close(got['absmean'], want['absmean'], 2e-3), (got['absmean'], want['absmean'])
        assert close(got['mean'], want['mean'], 2e-3, 2e-4), (got['mean'], want['mean'])
