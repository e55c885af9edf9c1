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


@pytest.mark.parametrize('case', ['glasses64_n2_b2_early', 'anime64_n3_b2', 'm2f64_n4_b2', 'glasses128_n2_b1'])
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
        early = case == 'glasses64_n2_b2_early'
        wg, wp = compare_with_oracle(tr, orc, hp, rtol_loss=1e-3, grad_rel_l2=1.0, flip_frac=0.2,
                                     min_cos=0.95 if early else None, shallow_only=True)
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


# Kept LAST among the GPU tests: the 256x256 fixture (BASELINE configs[1] geometry, council 2, batch 1) was generated after this
# round's GPU budget was spent.  It runs and its outcome is reported (xfail / xpass) without gating the suite; it becomes a hard
# case of test_iteration_matches_oracle_and_golden once it has been seen green on a B200.
@pytest.mark.xfail(strict=False, reason='not yet run on a GPU (fixture added at the end of round 1)')
@pytest.mark.parametrize('tc', [0, 1])
def test_full_resolution_iteration(tc):
    check_iteration('m2f256_n2_b1', tc)
