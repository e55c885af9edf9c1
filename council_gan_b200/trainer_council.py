"""``Council_Trainer`` -- drop-in for the reference's training step, driven through libcouncil_b200.so.

Same constructor, method names, argument meaning, attribute names and error behaviour as
``/root/reference/trainer_council.py`` for the path train.py:241-250 exercises::

    trainer = Council_Trainer(config, cuda_device)
    trainer.dis_update(images_a, images_b, config)
    trainer.dis_council_update(images_a, images_b, config)
    trainer.gen_update(images_a, images_b, config, iterations)
    trainer.update_learning_rate()

What is different underneath (B200-first):
  * the N council members are stacked: one grouped kernel launch per layer serves all of them
    (the reference loops ``for i in range(self.council_size)`` in Python, :328,558,747,826,858);
  * forward and backward are explicit sequences of our own CUDA kernels; no autograd graph, no
    cuDNN/cuBLAS; the dead work the reference performs is skipped (style encoder :754/829/331, D/DC
    weight gradients inside gen_update, the second/third evaluation of the same content encoding);
  * one host<->device synchronisation per gen_update (the loss-history matching :518-524,576-586 needs
    the loss values on the host) instead of 2N;
  * data parallel: when ``torch.distributed`` is initialised every rank holds all members, takes its
    slice of the global minibatch and the flat gradient buffer of each family is all-reduced (NCCL)
    once per optimiser step.
Paths outside the live configuration space of the reference's three configs (recon_*/vgg/abs losses,
nsgan/RaHinge, do_my_style, gray-scale D, random D/G pairing) raise NotImplementedError.
"""
from __future__ import annotations

import math
import os
import random
from collections import deque

import numpy as np
import torch
import torch.nn as nn

from .networks import IMG_C, CouncilDis, CouncilGen
from .utils import get_model_list

_DIRS = ('a2b', 'b2a')


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist
    return None


class _LossList(list):
    """List of 0-d tensors (what the reference publishes for write_loss, utils.py:277-305)."""


class Council_Trainer(nn.Module):
    def __init__(self, hyperparameters, cuda_device='cuda:0', _ops=None):
        super(Council_Trainer, self).__init__()
        hp = hyperparameters
        # ---- the same bookkeeping attributes as trainer_council.py:23-68 --------------------------------
        self.council_size = hp['council']['council_size']
        self.council_size_conf = self.council_size
        self.do_dis_council = hp['council_w'] != 0
        self.do_ads_council_loss = hp['council_abs_w'] != 0
        self.numberOfCouncil_dis_relative_iteration_conf = hp['council']['numberOfCouncil_dis_relative_iteration']
        self.discriminetro_less_style_by_conf = hp['council']['discriminetro_less_style_by']
        self.cuda_device = cuda_device
        self.recon_x_w_conf = hp['recon_x_w']
        self.recon_c_w_conf = hp['recon_c_w']
        self.recon_s_w_conf = hp['recon_s_w']
        self.recon_x_cyc_w_conf = hp['recon_x_cyc_w']
        self.gan_w_conf = hp['gan_w']
        self.vgg_w_conf = hp['vgg_w']
        self.abs_beginning_end_w_conf = hp['abs_beginning_end']
        self.flipOnOff_On_iteration_conf = hp['council']['flipOnOff_On_iteration']
        self.flipOnOff_Off_iteration_conf = hp['council']['flipOnOff_start_with']  # sic, :46-47
        self.council_abs_w_conf = hp['council_abs_w']
        self.council_w_conf = hp['council_w']
        self.council_start_at_iter_conf = hp['council']['council_start_at_iter']
        self.focus_loss_start_at_iter_conf = hp['focus_loss']['focus_loss_start_at_iter']
        self.mask_zero_or_one_w_conf = hp['mask_zero_or_one_w']
        self.mask_zero_or_one_center_conf = hp['focus_loss']['mask_zero_or_one_center']
        self.mask_zero_or_one_epsilon_conf = hp['focus_loss']['mask_zero_or_one_epsilon']
        self.mask_total_w_conf = hp['mask_total_w']
        self.mask_tv_w_conf = hp['mask_tv_w']
        self.batch_size_conf = hp['batch_size']
        self.do_w_loss_matching = hp['do_w_loss_matching']
        self.do_w_loss_matching_focus = hp['focus_loss']['do_w_loss_matching_focus']
        self.los_matching_hist_size_conf = hp['loss_matching_hist_size']
        self.do_a2b_conf = hp['do_a2b']
        self.do_b2a_conf = hp['do_b2a']
        self.w_match_b2a_conf = 1
        self.w_match_a2b_conf = 1
        self.w_match_focus_a2b_conf = 1
        self.w_match_focus_b2a_conf = 1
        self.w_match_focus_zero_one_a2b_conf = 1
        self.w_match_focus_zero_one_b2a_conf = 1
        self._check_supported(hp)
        self._dirs = [d for d in _DIRS if hp['do_' + d]]
        N, hist = self.council_size, self.los_matching_hist_size_conf
        for d in self._dirs:  # :70-92
            setattr(self, 'los_hist_gan_%s_s' % d, [deque(np.ones(hist)) for _ in range(N)])
            setattr(self, 'los_hist_council_%s_s' % d, [deque(np.ones(hist)) for _ in range(N)])
            setattr(self, 'los_hist_focus_%s_s' % d, [deque(np.ones(hist)) for _ in range(N)])
            setattr(self, 'los_hist_focus_zero_one_%s_s' % d, [deque(np.ones(hist)) for _ in range(N)])
        self.do_council_loss = None

        # ---- device op-set: our CUDA library.  No fallback: without it construction fails loudly. -------
        if _ops is None:
            from .ops import CudaOps
            _ops = CudaOps(cuda_device)
        object.__setattr__(self, 'ops', _ops)
        dist = _dist()
        self.world = dist.get_world_size() if dist else 1
        self.rank = dist.get_rank() if dist else 0

        # ---- networks (:101-133), stacked over the council ---------------------------------------------
        nets = {}
        for d in self._dirs:
            cin = hp['input_dim_a'] if d == 'a2b' else hp['input_dim_b']
            nets['gen_' + d] = CouncilGen(_ops, hp, N, cin)
            nets['dis_' + d] = CouncilDis(_ops, hp, N, cin, council=False)
            if self.do_dis_council:
                nets['dis_council_' + d] = CouncilDis(_ops, hp, N, cin, council=True)
        object.__setattr__(self, '_nets', nets)
        self.gen_a2b_s, self.gen_b2a_s, self.dis_a2b_s, self.dis_b2a_s = [], [], [], []
        if self.do_dis_council:
            self.dis_council_a2b_s, self.dis_council_b2a_s = [], []
        for name, net in nets.items():
            object.__setattr__(self, name + '_s', [net.member(i) for i in range(N)])
        self.style_dim = hp['gen']['style_dim']

        display_size = int(hp['display_size'])  # :136-138
        self.s_a = torch.randn(display_size, self.style_dim, 1, 1).to(_ops.device)
        self.s_b = torch.randn(display_size, self.style_dim, 1, 1).to(_ops.device)

        # ---- optimiser state (:140-183): flat fused Adam per family + StepLR bookkeeping -----------------
        self._lr0 = hp['lr']
        self._betas = (hp['beta1'], hp['beta2'])
        self._wd = hp['weight_decay']
        self._lr_policy = hp.get('lr_policy', 'constant')
        if self._lr_policy not in ('constant', 'step'):
            raise NotImplementedError('learning rate policy [%s] is not implemented' % self._lr_policy)
        self._step_size, self._gamma = hp.get('step_size', 1), hp.get('gamma', 1.0)
        self._sched_epoch = {'gen': 0, 'dis': 0, 'dis_council': 0}

        self._init_weights(hp['init'])  # :186-197
        self._img_cache = {}
        self._enc_cache = {}
        self._idx_cache = {}
        self._const_cache = {}
        self.hyperparameters = hp

    # ------------------------------------------------------------------------------------------------
    @staticmethod
    def _check_supported(hp):
        bad = [k for k in ('recon_x_w', 'recon_s_w', 'recon_c_w', 'recon_x_cyc_w', 'vgg_w', 'abs_beginning_end',
                           'council_abs_w') if hp.get(k, 0) != 0]
        if bad:
            raise NotImplementedError('loss terms %s are not on the accelerated training path' % bad)
        if hp['dis']['gan_type'] != 'lsgan':
            assert 0, "Unsupported GAN type: {}".format(hp['dis']['gan_type'])
        if hp['dis'].get('do_Dis_only_gray') or hp['dis'].get('useRandomGen') or hp['gen'].get('useRandomDis'):
            raise NotImplementedError('gray-scale D / random D-G pairing are not on the accelerated path')
        if hp['focus_loss'].get('do_w_loss_matching_focus'):
            raise NotImplementedError('do_w_loss_matching_focus is not on the accelerated path')
        if not (hp['do_a2b'] or hp['do_b2a']):
            raise ValueError('at least one of do_a2b / do_b2a must be set')

    def _init_weights(self, init_type):
        """weights_init (utils.py:402-422): kaiming fan_in normal (or N(0,0.02)) for generators, N(0,0.02) for
        discriminators, zero biases.  Drawn from the torch CPU generator; the stream is not the reference's
        (module construction order differs) -- parity tests load explicit state_dicts instead."""
        for name, net in self._nets.items():
            kind = init_type if name.startswith('gen_') else 'gaussian'
            if kind not in ('gaussian', 'kaiming', 'default'):
                raise NotImplementedError('init [%s] is not implemented' % kind)
            for spec in net._specs():
                b = net._bank_of(spec.wname)
                w = b.p(spec.wname)
                fan_in = spec.cin * spec.k * spec.k
                std = math.sqrt(2.0 / fan_in) if kind == 'kaiming' else 0.02
                ref = torch.randn(net.G, spec.cout, spec.cin, spec.k, spec.k) * std
                for i in range(net.G):
                    spec.import_weight(w[i], ref[i])
                b.p(spec.bname).zero_()

    # nn.Module surface the reference's callers touch
    def cuda(self, device=None):
        return self

    def _gate(self, hp, for_gen):
        """flip on/off + start gating, trainer_council.py:541-555 (gen) / :787-801 (dis_council)."""
        c = hp['council']
        cyc = hp['iteration'] % (c['flipOnOff_On_iteration'] + c['flipOnOff_Off_iteration'])
        start = c['flipOnOff_On_iteration'] if c['flipOnOff_start_with'] else c['flipOnOff_Off_iteration']
        do = c['flipOnOff_start_with'] if cyc < start else (not c['flipOnOff_start_with'])
        if not c['flipOnOff']:
            do = True if for_gen else c['flipOnOff_start_with']
        if for_gen and hp['iteration'] < c['council_start_at_iter']:
            do = False
        return do

    # ---- small host/device helpers ---------------------------------------------------------------------
    def _img(self, x):
        """NCHW image batch (any device) -> shared channels-last [1,B,H,W,4] on the device (cached per tensor)."""
        key = (x.data_ptr(), x._version, tuple(x.shape), str(x.device))
        hit = self._img_cache.get('k')
        if hit is not None and hit[0] == key:
            return hit[1]
        hit2 = self._img_cache.get('k2')
        if hit2 is not None and hit2[0] == key:
            return hit2[1]
        xd = x.detach().to(self.ops.device, self.ops.dtype, non_blocking=True).contiguous()
        img = self.ops.nchw_to_nhwc(xd, IMG_C)[None]
        self._img_cache['k2'] = self._img_cache.get('k')
        self._img_cache['k'] = (key, img, x)
        return img

    def _noise(self, batch):
        """torch.randn(B, style_dim, 1, 1) on the CPU generator (:284-285,741,744,807,809), moved to the device.
        Under data parallelism the GLOBAL batch is drawn on every rank (same seed) and sliced."""
        s = torch.randn(batch * self.world, self.style_dim, 1, 1)
        s = s[self.rank * batch:(self.rank + 1) * batch]
        return s.reshape(1, batch, 1, 1, self.style_dim).to(self.ops.device, self.ops.dtype, non_blocking=True)

    def _const(self, key, values):
        t = self._const_cache.get(key)
        if t is None:
            t = torch.tensor(values, dtype=self.ops.dtype).to(self.ops.device)
            self._const_cache[key] = t
        return t

    def _idx(self, key, build):
        t = self._idx_cache.get(key)
        if t is None:
            t = torch.tensor(build(), dtype=torch.int32).to(self.ops.device)
            self._idx_cache[key] = t
        return t

    def _encode(self, d, src_img, save):
        """Content encoding shared by the three updates of one iteration: same generator parameters and same
        images give the same result (the reference recomputes it 3x, :754-756, :829-832, :331-335)."""
        gen = self._nets['gen_' + d]
        key = (id(src_img), gen.bank.step, getattr(gen, '_load_epoch', 0))
        hit = self._enc_cache.get(d)
        if hit is not None and hit[0] == key and (hit[3] or not save):
            return hit[1], hit[2]
        saved = []
        c = gen.encode(src_img, saved)
        self._enc_cache[d] = (key, c, saved, True, src_img)
        return c, saved

    def _lr(self, fam):
        if self._lr_policy == 'constant':
            return self._lr0
        return self._lr0 * self._gamma ** (self._sched_epoch[fam] // self._step_size)

    def _adam(self, fam):
        """All-reduce the flat gradient of a family (data parallel) and run the fused Adam kernel on it."""
        dist = _dist()
        for d in self._dirs:
            net = self._nets.get('%s_%s' % (fam, d))
            if net is None:
                continue
            bank = net.bank
            if hasattr(self.ops, 'wgrad_join'):
                self.ops.wgrad_join()  # weight gradients may have been queued on a side stream (COUNCIL_WGRAD_STREAM=1)
            if dist is not None and self.world > 1:
                dist.all_reduce(bank.grad)  # SUM; local coefficients already carry 1/world
            bank.step += 1
            self.ops.adam_step(bank.data, bank.grad, bank.exp_avg, bank.exp_avg_sq, self._lr(fam), self._betas[0],
                               self._betas[1], 1e-8, self._wd, bank.step)
            net.params_changed()

    def _src(self, d, a, b):
        return a if d == 'a2b' else b

    def _global_mean(self, t):
        """Reported loss values are means over the GLOBAL minibatch (equal shards): average over ranks."""
        dist = _dist()
        if dist is not None and self.world > 1:
            dist.all_reduce(t)
            t = t / self.world
        return t

    # ==================================================================================================
    # dis_update   (trainer_council.py:735-780)
    # ==================================================================================================
    def dis_update(self, x_a=None, x_b=None, hyperparameters=None):
        hp = hyperparameters
        self._check_supported(hp)
        ops, N = self.ops, self.council_size
        img_a, img_b = self._img(x_a), self._img(x_b)
        s = {}
        if self.do_a2b_conf:  # :740-745
            s['a2b'] = self._noise(x_b.size(0))
        if self.do_b2a_conf:
            s['b2a'] = self._noise(x_a.size(0))
        total = ops.zeros(N)
        self._dis_sums = {}
        first = True
        for d in self._dirs:
            gen, dis = self._nets['gen_' + d], self._nets['dis_' + d]
            src, real = self._src(d, img_a, img_b), self._src(d, img_b, img_a)
            B, H, W = src.shape[1:4]
            c, _ = self._encode(d, src, save=True)
            x_fake, _ = gen.decode(c, s[d], src)
            # D minibatch per member: [own fake ; real]   (calc_dis_loss networks.py:56-64)
            pool = torch.cat((x_fake.view(N * B, H, W, IMG_C), real[0]), 0)
            idx = self._idx(('dis', N, B), lambda: [[g * B + b for b in range(B)] + [N * B + b for b in range(B)]
                                                    for g in range(N)])
            xin = ops.gather_images(pool, idx, None, N, 2 * B)
            saved = []
            outs = dis.forward(xin, saved)
            wdir = float(hp['gan_w']) if d == 'a2b' else 1.0  # :775 vs :777 (no gan_w on the b2a branch)
            targets = self._const('t01', [0.0, 1.0])
            weights = self._const(('w2', wdir, N), [[wdir, wdir]] * N)
            d_outs = []
            self._dis_sums[d] = []
            for out in outs:
                n_seg = out[0].numel() // 2
                sums = ops.lsgan_fwd(out, targets, weights, 2, total, accumulate=not first)
                first = False
                self._dis_sums[d].append((sums, n_seg))
                cf = wdir * 2.0 / (n_seg * self.world)
                coef = self._const(('c2', cf, N), [[cf, cf]] * N)
                d_outs.append(ops.lsgan_bwd(out, targets, coef, 2))
            dis.backward(d_outs, saved, want_wgrad=True, want_dx=False)
        total = self._global_mean(total)
        self._loss_dis_total = total
        self.loss_dis_total_s = _LossList(total[i] for i in range(N))
        self._adam('dis')

    def _per_dir_dis_loss(self, d):
        tot = 0
        for sums, n in self._dis_sums[d]:
            tot = tot + sums.sum(-1) / n
        return _LossList(tot[i] for i in range(self.council_size))

    @property
    def loss_dis_a2b_s(self):
        return self._per_dir_dis_loss('a2b')

    @property
    def loss_dis_b2a_s(self):
        return self._per_dir_dis_loss('b2a')

    # ==================================================================================================
    # dis_council_update   (trainer_council.py:782-883)
    # ==================================================================================================
    def dis_council_update(self, x_a=None, x_b=None, hyperparameters=None):
        hp = hyperparameters
        cc = hp['council']
        if self.council_size <= 1 or cc['numberOfCouncil_dis_relative_iteration'] == 0:
            print('no council discriminetor is needed (council size <= 1 or numberOfCouncil_dis_relative_iteration == 0)')
            return
        self.do_council_loss = self._gate(hp, for_gen=False)
        if not self.do_council_loss or hp['council_w'] == 0 or hp['iteration'] < cc['council_start_at_iter']:
            return
        self._check_supported(hp)
        ops, N = self.ops, self.council_size
        img_a, img_b = self._img(x_a), self._img(x_b)
        s = {}
        if self.do_b2a_conf:  # :806-809: s_a first, then s_b
            s['b2a'] = self._noise(x_a.size(0))
        if self.do_a2b_conf:
            s['a2b'] = self._noise(x_b.size(0))
        less = cc['discriminetro_less_style_by']
        # peers: python `random`, without replacement, pool refilled when exhausted (:861-868)
        Kcfg = cc['numberOfCouncil_dis_relative_iteration']
        peers = []
        for i in range(N):
            pool_i = list(range(0, i)) + list(range(i + 1, N))
            js = []
            for k in range(Kcfg):
                if k == N:
                    break
                if len(pool_i) == 0:
                    pool_i = list(range(0, i)) + list(range(i + 1, N))
                j = random.choice(pool_i)
                pool_i.remove(j)
                js.append(j)
            peers.append(js)
        # a peer drawn twice (K_cfg >= N refills the pool, :865-866) gives an identical term: evaluate each distinct peer
        # once and weight it by its multiplicity (x + x == 2x exactly)
        uniq = [sorted(set(js)) for js in peers]
        mult = [[js.count(j) for j in u] for js, u in zip(peers, uniq)]
        Krun = len(peers[0])
        U = len(uniq[0])
        assert all(len(u) == U for u in uniq)
        total = ops.zeros(N)
        first = True
        for d in self._dirs:
            gen, disc = self._nets['gen_' + d], self._nets['dis_council_' + d]
            src = self._src(d, img_a, img_b)
            B, H, W = src.shape[1:4]
            c, _ = self._encode(d, src, save=True)
            x_fake, _ = gen.decode(c, s[d], src)
            if less != 0:
                x_less, _ = gen.decode(c, s[d] * less, src)
                pool = torch.cat((x_fake.view(N * B, H, W, IMG_C), x_less.view(N * B, H, W, IMG_C)), 0)
                comp0 = N * B
            else:
                pool = x_fake.view(N * B, H, W, IMG_C)
                comp0 = 0
            idx = torch.tensor([[g * B + b for b in range(B)] +
                                [comp0 + j * B + b for j in uniq[g] for b in range(B)] for g in range(N)],
                               dtype=torch.int32).to(ops.device, non_blocking=True)
            xin = ops.gather_images(pool, idx, src, N, (1 + U) * B)
            saved = []
            outs = disc.forward(xin, saved)
            # sum_k [ mean(D(fake_i)^2) + mean((D(less_jk)-1)^2) ] * council_w / Kcfg   (:872, :878)
            wk = float(hp['council_w']) / Kcfg
            targets = self._const(('t0k', U), [0.0] + [1.0] * U)
            wrows = [[wk * Krun] + [wk * m for m in mult[g]] for g in range(N)]
            weights = torch.tensor(wrows, dtype=ops.dtype).to(ops.device, non_blocking=True)
            d_outs = []
            for out in outs:
                n_seg = out[0].numel() // (1 + U)
                ops.lsgan_fwd(out, targets, weights, 1 + U, total, accumulate=not first)
                first = False
                cf = 2.0 / (n_seg * self.world)
                coef = torch.tensor([[v * cf for v in row] for row in wrows], dtype=ops.dtype).to(ops.device, non_blocking=True)
                d_outs.append(ops.lsgan_bwd(out, targets, coef, 1 + U))
            disc.backward(d_outs, saved, want_wgrad=True, want_dx=False)
        total = self._global_mean(total)
        self._loss_dis_council_total = total
        self.loss_dis_council_total_s = _LossList(total[i] for i in range(N))
        self._adam('dis_council')

    # ==================================================================================================
    # gen_update   (trainer_council.py:280-634)
    # ==================================================================================================
    def gen_update(self, x_a, x_b, hyperparameters, iterations=0):
        hp = hyperparameters
        self.hyperparameters = hp
        self._check_supported(hp)
        ops, N = self.ops, self.council_size
        fl = hp['focus_loss']
        img_a, img_b = self._img(x_a), self._img(x_b)
        s_a = self._noise(x_a.size(0))  # :284-285 both are always drawn, a first
        s_b = self._noise(x_b.size(0))
        s = {'a2b': s_b, 'b2a': s_a}
        it = hp['iteration']
        focus_gate = it > fl['focus_loss_start_at_iter']
        self.council_w_conf = hp['council_w'] if it > hp['council']['council_start_at_iter'] else 0  # :323-326
        self.mask_zero_or_one_w_conf = hp['mask_zero_or_one_w'] if focus_gate else 0
        self.mask_total_w_conf = hp['mask_total_w'] if focus_gate else 0
        self.mask_tv_w_conf = hp['mask_tv_w'] if focus_gate else 0
        focus_on = focus_gate and (hp['mask_zero_or_one_w'] != 0 or hp['mask_total_w'] != 0)  # :390
        if focus_on and hp['mask_total_w'] != 0:
            assert fl['mask_small_use_abs'] or fl['mask_small_use_square'], \
                'at leas one small mask loss should be true, mask_small_use_abs or mask_small_use_square'
        self.do_council_loss = self._gate(hp, for_gen=True)
        council_on = (hp['council_w'] != 0) and self.do_council_loss and N > 1 and self.do_dis_council  # :559,567
        gan_on = hp['gan_w'] != 0

        fw = {}
        scal = []  # device scalars to bring to the host in ONE copy: per dir [adv(N), council(N), focus(N,4)]
        for d in self._dirs:
            gen = self._nets['gen_' + d]
            src = self._src(d, img_a, img_b)
            B, H, W = src.shape[1:4]
            c, enc_saved = self._encode(d, src, save=True)
            dec_saved = []
            x_fake, mask = gen.decode(c, s[d], src, dec_saved)
            rec = {'enc': enc_saved, 'dec': dec_saved, 'x_fake': x_fake, 'mask': mask, 'B': B, 'H': H, 'W': W}
            adv = ops.zeros(N)
            cl = ops.zeros(N)
            fs = ops.zeros(N, 4)
            ones = self._const('t1', [1.0])
            onesw = self._const(('w1', N), [[1.0]] * N)
            if gan_on:  # calc_gen_loss networks.py:84-90
                rec['dis_saved'] = []
                rec['dis_outs'] = self._nets['dis_' + d].forward(x_fake, rec['dis_saved'])
                for k, out in enumerate(rec['dis_outs']):
                    ops.lsgan_fwd(out, ones, onesw, 1, adv, accumulate=k > 0)
            if council_on:  # MsImageDisCouncil.calc_gen_loss networks.py:188-194
                idx = self._idx(('id', N, B), lambda: [[g * B + b for b in range(B)] for g in range(N)])
                xin = ops.gather_images(x_fake.view(N * B, H, W, IMG_C), idx, src, N, B)
                rec['disc_saved'] = []
                rec['disc_outs'] = self._nets['dis_council_' + d].forward(xin, rec['disc_saved'])
                for k, out in enumerate(rec['disc_outs']):
                    ops.lsgan_fwd(out, ones, onesw, 1, cl, accumulate=k > 0)
            if focus_on:
                fs = ops.focus_fwd(mask, fl['mask_zero_or_one_center'], fl['mask_zero_or_one_epsilon'])
            fw[d] = rec
            scal.append(torch.cat((adv.view(N, 1), cl.view(N, 1), fs), 1))
        dev_scal = torch.stack(scal)  # [ndirs, N, 6]
        dist = _dist()
        if dist is not None and self.world > 1:
            dist.all_reduce(dev_scal)  # sums over ranks; means are divided by world below
        host = dev_scal.cpu().double().numpy()  # the one host sync of gen_update

        # ---- host: loss values, history matching, backward coefficients --------------------------------
        tot = np.zeros(N, dtype=np.float64)
        names = {}
        coefs = {}
        for di, d in enumerate(self._dirs):
            ab = 'ab' if d == 'a2b' else 'ba'
            rec = fw[d]
            numel = rec['B'] * self.world * 3 * rec['H'] * rec['W']  # mask.numel() of the GLOBAL batch
            adv = host[di, :, 0] / self.world
            cl = host[di, :, 1] / self.world
            f = host[di, :, 2:]
            l01 = f[:, 0] / numel
            msum = f[:, 1] / numel
            ltv = (f[:, 2] + f[:, 3]) / numel
            ltot = np.zeros(N)
            c01 = csum = ctv = np.zeros(N)
            z01, ztot, ztv = [], [0] * N, [0] * N
            if focus_on:
                if hp['mask_zero_or_one_w'] != 0:  # :392-415
                    z01 = list(l01)
                    tot += hp['mask_zero_or_one_w'] * l01
                    c01 = np.full(N, hp['mask_zero_or_one_w'] / numel)
                if hp['mask_tv_w'] != 0:  # :425-431
                    ztv = list(ltv)
                    tot += hp['mask_tv_w'] * ltv
                    ctv = np.full(N, hp['mask_tv_w'] / numel)
                if hp['mask_total_w'] != 0:  # :418-422 then :447-451
                    csum = np.zeros(N)
                    if fl['mask_small_use_abs']:
                        ltot = ltot + np.abs(msum)
                        csum = csum + hp['mask_total_w'] * np.sign(msum) / numel
                    if fl['mask_small_use_square']:
                        ltot = ltot + msum ** 2
                        csum = csum + hp['mask_total_w'] * 2.0 * msum / numel
                    ztot = list(ltot)
                    tot += hp['mask_total_w'] * ltot
            hist_gan = getattr(self, 'los_hist_gan_%s_s' % d)
            hist_c = getattr(self, 'los_hist_council_%s_s' % d)
            if gan_on:
                if self.do_w_loss_matching:  # :518-524
                    for i in range(N):
                        hist_gan[i].append(np.float32(adv[i]))
                        hist_gan[i].popleft()
                tot += hp['gan_w'] * adv
            cdis = np.zeros(N)
            closs = [0] * N
            if council_on:
                w = np.ones(N)
                if self.do_w_loss_matching:  # :576-586
                    for i in range(N):
                        hist_c[i].append(np.float32(cl[i]))
                        hist_c[i].popleft()
                        w[i] = np.mean(hist_gan[i]) / np.mean(hist_c[i])
                        setattr(self, 'w_match_%s_conf' % d, w[i])
                closs_v = cl * w.astype(np.float32) * hp['council_w']
                closs = list(closs_v)
                tot += closs_v
                cdis = w * hp['council_w']
            names[d] = (ab, adv, z01, ztot, ztv, closs)
            coefs[d] = (c01, csum, ctv, cdis)

        # ---- backward --------------------------------------------------------------------------------
        for d in self._dirs:
            rec = fw[d]
            gen = self._nets['gen_' + d]
            c01, csum, ctv, cdis = coefs[d]
            ones = self._const('t1', [1.0])
            d_x = None
            if gan_on:
                d_outs = []
                for out in rec['dis_outs']:
                    cf = hp['gan_w'] * 2.0 / (out[0].numel() * self.world)
                    coef = self._const(('c1', cf, N), [[cf]] * N)
                    d_outs.append(ops.lsgan_bwd(out, ones, coef, 1))
                d_x = self._nets['dis_' + d].backward(d_outs, rec['dis_saved'], want_wgrad=False, want_dx=True)
            if council_on:
                d_outs = []
                for out in rec['disc_outs']:
                    cf = 2.0 / (out[0].numel() * self.world)
                    coef = torch.tensor((cdis * cf).reshape(N, 1), dtype=ops.dtype).to(ops.device, non_blocking=True)
                    d_outs.append(ops.lsgan_bwd(out, ones, coef, 1))
                d_x8 = self._nets['dis_council_' + d].backward(d_outs, rec['disc_saved'], want_wgrad=False, want_dx=True)
                if d_x is None:
                    d_x = ops.zeros(*rec['x_fake'].shape)
                ops.acc_slice(d_x, d_x8, 4)
            if d_x is None:
                d_x = ops.zeros(*rec['x_fake'].shape)
            d_mask = None
            if focus_on:
                coef = torch.tensor(np.stack((c01, csum, ctv), 1), dtype=ops.dtype).to(ops.device, non_blocking=True)
                d_mask = ops.focus_bwd(rec['mask'], coef, fl['mask_zero_or_one_center'], fl['mask_zero_or_one_epsilon'])
            gen.backward(d_x, d_mask, rec['enc'], rec['dec'])
        self._adam('gen')
        self._enc_cache.clear()

        # ---- publish the reference's loss attributes (:302-322, :556-557) --------------------------------
        def lst(v):
            return _LossList(torch.tensor(float(x)) for x in v)
        self.loss_gen_total_s = lst(tot)
        for d in _DIRS:
            ab = 'ab' if d == 'a2b' else 'ba'
            a2b = d
            if d in names:
                _, adv, z01, ztot, ztv, closs = names[d]
                setattr(self, 'loss_gen_adv_%s_s' % a2b, lst(adv) if gan_on else [])
                setattr(self, 'loss_gen_mask_zero_one_%s_s' % ab, lst(z01))
                setattr(self, 'loss_gen_mask_total_%s_s' % ab, lst(ztot))
                setattr(self, 'loss_gen_mask_TV_%s_s' % ab, lst(ztv))
                setattr(self, 'council_loss_%s_s' % ab, lst(closs))
            else:
                setattr(self, 'loss_gen_adv_%s_s' % a2b, [0] * N if gan_on else [])
                setattr(self, 'loss_gen_mask_zero_one_%s_s' % ab, [])
                setattr(self, 'loss_gen_mask_total_%s_s' % ab, [])
                setattr(self, 'loss_gen_mask_TV_%s_s' % ab, [])
                setattr(self, 'council_loss_%s_s' % ab, [])
        self._last_fw = {d: {'x_fake': fw[d]['x_fake'], 'mask': fw[d]['mask']} for d in self._dirs}

    # ==================================================================================================
    # the rest of the reference surface
    # ==================================================================================================
    def update_learning_rate(self):
        """StepLR.step() on every optimiser (:885-896; get_scheduler utils.py:392-400)."""
        for fam in self._sched_epoch:
            if fam == 'dis_council' and not self.do_dis_council:
                continue
            self._sched_epoch[fam] += 1

    def sample(self, x_a=None, x_b=None, s_a=None, s_b=None, council_member_to_sample_vec=None, return_mask=True):
        """Eval-mode translation of every image by every member (:643-733): returns the same 8-tuple."""
        members = range(self.council_size) if council_member_to_sample_vec is None else council_member_to_sample_vec
        res = {}
        for d in _DIRS:
            if not getattr(self, 'do_%s_conf' % d):
                res[d] = (None, None, None, None)
                continue
            x = x_a if d == 'a2b' else x_b
            fixed = (self.s_b if s_b is None else s_b) if d == 'a2b' else (self.s_a if s_a is None else s_a)
            s2 = torch.randn(x.size(0), self.style_dim, 1, 1).to(self.ops.device)
            gens = getattr(self, 'gen_%s_s' % d)
            xs, second, first, third = [], [], [], []
            for i in range(x.size(0)):
                xi = x[i].unsqueeze(0)
                for j in members:
                    xs.append(xi.to(self.ops.device))
                    c, s_fake = gens[j].encode(xi)
                    if not return_mask:
                        second.append(gens[j].decode(c, s_fake, xi))
                        first.append(gens[j].decode(c, fixed[i].unsqueeze(0), xi))
                    else:
                        o, m = gens[j].decode(c, fixed[i].unsqueeze(0), xi, return_mask=True)
                        second.append(m)
                        first.append(o)
                    third.append(gens[j].decode(c, s2[i].unsqueeze(0), xi))
            res[d] = (torch.cat(xs), torch.cat(second), torch.cat(first), torch.cat(third))
        return res['a2b'] + res['b2a']

    def forward(self, *args, **kwargs):
        raise NotImplementedError('Council_Trainer.forward is broken in the reference (trainer_council.py:267 '
                                  'references a nonexistent self.gen_a2b); use sample()')

    def save(self, snapshot_dir, iterations):
        """Per-member checkpoint files with the reference's names and keys (:969-992)."""
        for i in range(self.council_size):
            for fam in ('gen', 'dis', 'dis_council'):
                if fam == 'dis_council' and not self.do_dis_council:
                    continue
                for d in self._dirs:
                    name = os.path.join(snapshot_dir, '%s_%s_%d_%08d.pt' % (d, fam, i, iterations + 1))
                    torch.save({d: getattr(self, '%s_%s_s' % (fam, d))[i].state_dict()}, name)
            opt = {}
            for fam in ('gen', 'dis', 'dis_council'):
                if fam == 'dis_council' and not self.do_dis_council:
                    continue
                opt[fam] = {d: self._opt_state(fam, d, i) for d in self._dirs}
            torch.save(opt, os.path.join(snapshot_dir, 'optimizer_%d.pt' % i))

    def _opt_state(self, fam, d, i):
        bank = self._nets['%s_%s' % (fam, d)].bank
        sl = {}
        for name, (off, shape, n) in bank.table.items():
            per = n // bank.G
            sl[name] = {'exp_avg': bank.exp_avg[off + i * per: off + (i + 1) * per].clone().cpu(),
                        'exp_avg_sq': bank.exp_avg_sq[off + i * per: off + (i + 1) * per].clone().cpu()}
        return {'step': bank.step, 'state': sl, 'sched_epoch': self._sched_epoch[fam]}

    def resume(self, checkpoint_dir, hyperparameters):
        """Load the latest per-member checkpoints (:898-967); returns the iteration parsed from the file name."""
        iterations = 0
        for i in range(self.council_size):
            for fam in ('gen', 'dis', 'dis_council'):
                if fam == 'dis_council' and not self.do_dis_council:
                    continue
                last = get_model_list(checkpoint_dir, '%s_%d' % (fam, i))
                if last is None:
                    import warnings
                    warnings.warn('Failed to find %s checkpoint, did not load model' % fam)
                    continue
                base = os.path.basename(last)
                for d in self._dirs:
                    path = os.path.join(checkpoint_dir, d + base[3:])
                    state = torch.load(path, map_location='cpu')
                    getattr(self, '%s_%s_s' % (fam, d))[i].load_state_dict(state[d])
                if fam == 'gen':
                    iterations = int(last[-11:-3])
            opt_path = os.path.join(checkpoint_dir, 'optimizer_%d.pt' % i)
            if os.path.exists(opt_path):
                opt = torch.load(opt_path, map_location='cpu')
                for fam, per_dir in opt.items():
                    for d, st in per_dir.items():
                        if not isinstance(st, dict) or 'state' not in st:
                            continue  # a reference-format optimizer file: moments restart from zero
                        bank = self._nets['%s_%s' % (fam, d)].bank
                        bank.step = st['step']
                        self._sched_epoch[fam] = st.get('sched_epoch', iterations)
                        for name, (off, shape, n) in bank.table.items():
                            per = n // bank.G
                            bank.exp_avg[off + i * per: off + (i + 1) * per].copy_(st['state'][name]['exp_avg'])
                            bank.exp_avg_sq[off + i * per: off + (i + 1) * per].copy_(st['state'][name]['exp_avg_sq'])
        if iterations > 0:
            print('Resume from iteration %d' % iterations)
            for fam in self._sched_epoch:
                self._sched_epoch[fam] = max(self._sched_epoch[fam], iterations)
        else:
            import warnings
            warnings.warn('FAILED TO RESUME STARTED FROM 0')
        return iterations
