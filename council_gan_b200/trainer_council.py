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
  * NO host<->device synchronisation inside an iteration: the losses of all scales, the focus terms, the
    loss-history matching (:518-524,576-586: float64 ring buffers on the device) and every loss gradient are
    three fused kernels (csrc/losses.cu); the published loss attributes are 0-d device tensors exactly like the
    reference's, so the only sync is the one the caller makes when it reads them (write_loss, utils.py:277-305);
  * data parallel: when ``torch.distributed`` is initialised every rank holds all members, takes its
    slice of the global minibatch and the flat gradient buffer of each family is all-reduced (NCCL)
    once per optimiser step -- asynchronously: the discriminators' all-reduce + Adam are joined only when that
    family's parameters are next needed (dis: during dis_council_update; dis_council: during gen_update's
    generator forward; gen: decoder bucket during the encoder backward).
Paths outside the live configuration space of the reference's three configs (recon_*/vgg/abs losses,
nsgan/RaHinge, do_my_style, gray-scale D, random D/G pairing) raise NotImplementedError.
"""
from __future__ import annotations

import math
import functools
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


_PDL = os.environ.get('COUNCIL_PDL', 'auto')   # auto | 0 | 1
_PDL_MAX_PIXELS = 128 * 128 * 4             # batch x height x width up to which programmatic dependent launch is switched on


def _pinned(fn):
    """Run an update with the launch stream resolved once (ops.pin_stream) instead of once per kernel."""
    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        ops = self.ops
        if not hasattr(ops, 'pin_stream') or ops._stream_cached is not None:
            return fn(self, *args, **kwargs)
        ops.pin_stream()
        if _PDL != 'auto':
            ops.set_pdl(_PDL == '1')
        else:  # launch-bound small maps only (measured: -4 % at 128x128 x 1, +2 % at 256x256 x 8)
            for x in args:
                if torch.is_tensor(x):
                    pix = x.numel() // (IMG_C if x.dim() == 5 else max(int(x.shape[1]), 1)) if x.dim() >= 4 else 0
                    ops.set_pdl(0 < pix <= _PDL_MAX_PIXELS)
                    break
        try:
            return fn(self, *args, **kwargs)
        finally:
            ops.unpin_stream()
    return wrapper


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
        for d in self._dirs:  # :70-92; the gan / council histories live on the device (see _rings), these two are never updated
            setattr(self, 'los_hist_focus_%s_s' % d, [deque(np.ones(hist)) for _ in range(N)])
            setattr(self, 'los_hist_focus_zero_one_%s_s' % d, [deque(np.ones(hist)) for _ in range(N)])
        self.do_council_loss = None

        # ---- device op-set: our CUDA library.  No fallback: without it construction fails loudly. -------
        if _ops is None:
            from .ops import CudaOps
            _ops = CudaOps(cuda_device)
        object.__setattr__(self, 'ops', _ops)
        # loss histories of the matching (:81-92, deque(np.ones(hist))): float64 rings [N][hist+1] on the device, window start
        # kept on the host; exported as deques by the los_hist_{gan,council}_{a2b,b2a}_s attributes
        rings = {d: {'gan': torch.ones(N, hist + 1, dtype=torch.float64).to(_ops.device),
                     'council': torch.ones(N, hist + 1, dtype=torch.float64).to(_ops.device),
                     'head_gan': 0, 'head_council': 0} for d in self._dirs}
        object.__setattr__(self, '_rings', rings)
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
            net._before_access = self._flush
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
        self.img_cache_misses = 0  # image batches uploaded / converted (bench.py checks its e2e leg really copies)
        object.__setattr__(self, '_pending', {})  # family -> [(net, bucket, async work)] awaiting all-reduce completion + Adam
        self.hyperparameters = hp
        self.sync_parameters()

    def __getattr__(self, name):
        # los_hist_gan_a2b_s etc. (trainer_council.py:81-92): lists of deques, read back from the device rings on demand
        if name.startswith('los_hist_gan_') or name.startswith('los_hist_council_'):
            kind, d = ('gan', name[13:16]) if name.startswith('los_hist_gan_') else ('council', name[17:20])
            rings = self.__dict__.get('_rings', {})
            if d in rings and name.endswith('_s'):
                r = rings[d]
                R = r[kind].shape[1]
                host = r[kind].detach().cpu().numpy()
                head = r['head_' + kind]
                return [deque(host[i, [(head + k) % R for k in range(R - 1)]]) for i in range(host.shape[0])]
        return super().__getattr__(name)

    def sync_parameters(self):
        """Data parallel: every rank must hold the same parameters and optimiser state.  Only gradients are all-reduced
        during training, so rank 0's banks are broadcast after construction / resume() (ranks seeded differently, e.g.
        seed + rank, would otherwise train replicas that never re-synchronise).  No-op without a process group."""
        dist = _dist()
        if dist is None or self.world <= 1:
            return
        self._flush()
        for net in self._nets.values():
            for bank in net._banks():
                dist.broadcast(bank.data, 0)
                if bank.trainable:
                    dist.broadcast(bank.exp_avg, 0)
                    dist.broadcast(bank.exp_avg_sq, 0)
            net.params_changed()

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
        """NCHW image batch (any device) -> shared channels-last [1,B,H,W,4] on the device.  Cached per tensor OBJECT (the
        three updates of one iteration receive the same tensors, train.py:241-250); a new tensor is always uploaded."""
        if x.dim() == 5:  # already the step's layout: a DeviceAugment / DeviceFolderLoader batch (data.py)
            if x.shape[0] != 1 or x.shape[-1] != IMG_C or x.device.type != torch.device(self.ops.device).type or x.dtype != self.ops.dtype:
                raise ValueError('channels-last image batches must be [1, B, H, W, %d] %s tensors on %s' % (IMG_C, self.ops.dtype, self.ops.device))
            return x
        key = (x.data_ptr(), x._version, tuple(x.shape), str(x.device))
        for slot in ('k', 'k2'):
            hit = self._img_cache.get(slot)
            if hit is not None and hit[0] == key and hit[2] is x:
                return hit[1]
        xd = x.detach().to(self.ops.device, self.ops.dtype, non_blocking=True).contiguous()
        img = self.ops.nchw_to_nhwc(xd, IMG_C)[None]
        self.img_cache_misses += 1
        self._img_cache['k2'] = self._img_cache.get('k')
        self._img_cache['k'] = (key, img, x)
        return img

    @staticmethod
    def _bsz(x):
        return x.shape[1] if x.dim() == 5 else x.size(0)

    def _noise(self, batch):
        """torch.randn(B, style_dim, 1, 1) on the CPU generator (:284-285,741,744,807,809) as a HOST tensor [1,B,1,1,S];
        ops.stage() moves it.  Under data parallelism the GLOBAL batch is drawn on every rank (all ranks must be seeded
        identically: same torch / python `random` seeds) and sliced."""
        s = torch.randn(batch * self.world, self.style_dim, 1, 1)
        s = s[self.rank * batch:(self.rank + 1) * batch]
        return s.reshape(1, batch, 1, 1, self.style_dim)

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

    # ---- optimiser step, data-parallel gradient exchange ---------------------------------------------------
    def _reduce_async(self, fam, net, lo=0, hi=None):
        """Queue the all-reduce of grad[lo:hi] of one network (NCCL stream, ordered after everything issued so far)."""
        dist = _dist()
        work = None
        if dist is not None and self.world > 1:
            g = net.bank.grad if (lo == 0 and hi is None) else net.bank.grad[lo:hi]
            work = dist.all_reduce(g, async_op=True)  # SUM; local coefficients already carry 1/world
        self._pending.setdefault(fam, []).append((net, work))

    def _finish(self, fam):
        """Join the family's queued all-reduces (stream-side wait, no host block on NCCL) and run its fused Adam."""
        items = self._pending.pop(fam, None)
        if not items:
            return
        if hasattr(self.ops, 'wgrad_join'):
            self.ops.wgrad_join()  # weight gradients may have been queued on a side stream (COUNCIL_WGRAD_STREAM=1)
        for _, work in items:
            if work is not None:
                work.wait()  # every bucket of the family first
        done = []
        for net, _ in items:
            if any(net is n for n in done):
                continue  # a network with several buckets steps once
            done.append(net)
            bank = net.bank
            bank.step += 1
            self.ops.adam_step(bank.data, bank.grad, bank.exp_avg, bank.exp_avg_sq, self._lr(fam), self._betas[0],
                               self._betas[1], 1e-8, self._wd, bank.step)
            net.params_changed()

    def _flush(self):
        for fam in list(self._pending):
            self._finish(fam)

    def synchronize(self):
        """Join every deferred optimiser step (data parallel: the gradient all-reduce of the last update is still in flight
        when gen_update returns).  Called implicitly by the next update, save(), sample() and every state_dict access."""
        self._flush()

    def _adam(self, fam, defer=False):
        """All-reduce the flat gradient of a family (data parallel) and run the fused Adam kernel on it.  defer: leave both
        queued until the family's parameters are next needed (_finish), so that the all-reduce overlaps the next update."""
        for d in self._dirs:
            net = self._nets.get('%s_%s' % (fam, d))
            if net is None:
                continue
            if not any(net is n for n, _ in self._pending.get(fam, [])):
                self._reduce_async(fam, net)
        if not (defer and self.world > 1) or os.environ.get('COUNCIL_DP_SYNC') == '1':  # COUNCIL_DP_SYNC=1: A/B switch, join immediately
            self._finish(fam)

    def _src(self, d, a, b):
        return a if d == 'a2b' else b

    def _global_sum(self, t):
        """Reported loss values are means over the GLOBAL minibatch (equal shards); the local values already carry 1/world."""
        dist = _dist()
        if dist is not None and self.world > 1:
            dist.all_reduce(t)
        return t

    # ==================================================================================================
    # dis_update   (trainer_council.py:735-780)
    # ==================================================================================================
    @_pinned
    def dis_update(self, x_a=None, x_b=None, hyperparameters=None):
        hp = hyperparameters
        self._check_supported(hp)
        self._flush()
        ops, N = self.ops, self.council_size
        img_a, img_b = self._img(x_a), self._img(x_b)
        noise = []
        if self.do_a2b_conf:  # :740-745
            noise.append(('a2b', self._noise(self._bsz(x_b))))
        if self.do_b2a_conf:
            noise.append(('b2a', self._noise(self._bsz(x_a))))
        s = dict(zip((k for k, _ in noise), ops.stage([v for _, v in noise])))
        total = ops.empty(N)
        inv_world = 1.0 / self.world
        for di, d in enumerate(self._dirs):
            gen, dis = self._nets['gen_' + d], self._nets['dis_' + d]
            src, real = self._src(d, img_a, img_b), self._src(d, img_b, img_a)
            B, H, W = src.shape[1:4]
            c, _ = self._encode(d, src, save=True)
            x_fake, _ = gen.decode(c, s[d], src)
            # D minibatch per member: [own fake ; real]   (calc_dis_loss networks.py:56-64); slots >= N*B read the real batch
            idx = self._idx(('dis', N, B), lambda: [[g * B + b for b in range(B)] + [N * B + b for b in range(B)]
                                                    for g in range(N)])
            xin = ops.gather_images((x_fake.view(N * B, H, W, IMG_C), real[0]), idx, None, N, 2 * B)
            saved = []
            outs = dis.forward(xin, saved)
            wdir = float(hp['gan_w']) if d == 'a2b' else 1.0  # :775 vs :777 (no gan_w on the b2a branch)
            plain = ops.empty(N)
            # loss of both scales + d(loss)/d(out) = wdir * 2 / (n * world) * (out - target): one launch
            d_outs = ops.lsgan_fused(outs, 2, [0.0, 1.0], [[wdir, wdir]] * N, inv_world, inv_world, total, di > 0, plain)
            plain = self._global_sum(plain)
            setattr(self, 'loss_dis_%s_s' % d, [plain[i] for i in range(N)])  # :765-768, only for active directions
            dis.backward(d_outs, saved, want_wgrad=True, want_dx=False)
        total = self._global_sum(total)
        self.loss_dis_total_s = [total[i] for i in range(N)]
        self._adam('dis', defer=True)  # joined when the D parameters are next needed (gen_update / the next dis_update)

    # ==================================================================================================
    # dis_council_update   (trainer_council.py:782-883)
    # ==================================================================================================
    @_pinned
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
        self._finish('dis_council')
        self._finish('gen')
        ops, N = self.ops, self.council_size
        img_a, img_b = self._img(x_a), self._img(x_b)
        noise = []
        if self.do_b2a_conf:  # :806-809: s_a first, then s_b
            noise.append(('b2a', self._noise(self._bsz(x_a))))
        if self.do_a2b_conf:
            noise.append(('a2b', self._noise(self._bsz(x_b))))
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
        # every small host table of this update (style noise, less-style noise, slot tables) goes up in ONE pinned copy
        Bs = {d: self._src(d, img_a, img_b).shape[1] for d in self._dirs}
        comp0 = {d: (N * Bs[d] if less != 0 else 0) for d in self._dirs}
        tables = [torch.tensor([[g * Bs[d] + b for b in range(Bs[d])] +
                                [comp0[d] + j * Bs[d] + b for j in uniq[g] for b in range(Bs[d])] for g in range(N)], dtype=torch.int32)
                  for d in self._dirs]
        host = [v for _, v in noise] + ([v * less for _, v in noise] if less != 0 else []) + tables
        dev = ops.stage(host)
        s = dict(zip((k for k, _ in noise), dev[:len(noise)]))
        s_less = dict(zip((k for k, _ in noise), dev[len(noise):2 * len(noise)])) if less != 0 else {}
        idx = dict(zip(self._dirs, dev[-len(self._dirs):]))
        total = ops.empty(N)
        inv_world = 1.0 / self.world
        for di, d in enumerate(self._dirs):
            gen, disc = self._nets['gen_' + d], self._nets['dis_council_' + d]
            src = self._src(d, img_a, img_b)
            B, H, W = src.shape[1:4]
            c, _ = self._encode(d, src, save=True)
            x_fake, _ = gen.decode(c, s[d], src)
            pools = [x_fake.view(N * B, H, W, IMG_C)]
            if less != 0:
                x_less, _ = gen.decode(c, s_less[d], src)
                pools.append(x_less.view(N * B, H, W, IMG_C))
            if di == 0:
                self._finish('dis')  # dis_update's all-reduce had the generator decodes above to complete behind
            xin = ops.gather_images(pools, idx[d], src, N, (1 + U) * B)
            saved = []
            outs = disc.forward(xin, saved)
            # sum_k [ mean(D(fake_i)^2) + mean((D(less_jk)-1)^2) ] * council_w / Kcfg   (:872, :878)
            wk = float(hp['council_w']) / Kcfg
            wrows = [[wk * Krun] + [wk * m for m in mult[g]] for g in range(N)]
            d_outs = ops.lsgan_fused(outs, 1 + U, [0.0] + [1.0] * U, wrows, inv_world, inv_world, total, di > 0)
            disc.backward(d_outs, saved, want_wgrad=True, want_dx=False)
        total = self._global_sum(total)
        self.loss_dis_council_total_s = [total[i] for i in range(N)]
        self._adam('dis_council', defer=True)  # joined before gen_update evaluates the council discriminators

    # ==================================================================================================
    # gen_update   (trainer_council.py:280-634)
    # ==================================================================================================
    @_pinned
    def gen_update(self, x_a, x_b, hyperparameters, iterations=0):
        hp = hyperparameters
        self.hyperparameters = hp
        self._check_supported(hp)
        self._finish('gen')
        ops, N = self.ops, self.council_size
        fl = hp['focus_loss']
        img_a, img_b = self._img(x_a), self._img(x_b)
        s_a = self._noise(self._bsz(x_a))  # :284-285 both are always drawn, a first
        s_b = self._noise(self._bsz(x_b))
        s_a, s_b = ops.stage([s_a, s_b])
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
        center, eps = float(fl['mask_zero_or_one_center']), float(fl['mask_zero_or_one_epsilon'])

        # ---- forward of every direction; pass 1 of the loss (all reductions, one launch per direction) -----------
        fw = {}
        scal = ops.empty(len(self._dirs), N, 6)  # per direction and member: [adv, council, focus sums x4] of THIS rank
        for di, d in enumerate(self._dirs):
            gen = self._nets['gen_' + d]
            src = self._src(d, img_a, img_b)
            B, H, W = src.shape[1:4]
            c, enc_saved = self._encode(d, src, save=True)
            dec_saved = []
            x_fake, mask = gen.decode(c, s[d], src, dec_saved)
            rec = {'enc': enc_saved, 'dec': dec_saved, 'x_fake': x_fake, 'mask': mask, 'B': B, 'H': H, 'W': W,
                   'dis_outs': [], 'disc_outs': []}
            if gan_on:  # calc_gen_loss networks.py:84-90
                if di == 0:
                    self._finish('dis')
                rec['dis_saved'] = []
                rec['dis_outs'] = self._nets['dis_' + d].forward(x_fake, rec['dis_saved'])
            if council_on:  # MsImageDisCouncil.calc_gen_loss networks.py:188-194
                if di == 0:
                    self._finish('dis_council')
                idx = self._idx(('id', N, B), lambda: [[g * B + b for b in range(B)] for g in range(N)])
                xin = ops.gather_images(x_fake.view(N * B, H, W, IMG_C), idx, src, N, B)
                rec['disc_saved'] = []
                rec['disc_outs'] = self._nets['dis_council_' + d].forward(xin, rec['disc_saved'])
            rec['d_adv'] = ops.gen_loss_fwd(rec['dis_outs'], rec['disc_outs'], mask if focus_on else None, center, eps,
                                            float(hp['gan_w']) / self.world, scal[di])
            fw[d] = rec
        self._flush()  # (a family gated off above still steps here)
        dist = _dist()
        if dist is not None and self.world > 1:
            dist.all_reduce(scal)  # sums over ranks; pass 2 divides the means by world and uses the GLOBAL (sum m / numel)^2

        # ---- pass 2 (loss assembly + history matching on the device, remaining loss gradients) and the backward ------
        total = ops.empty(N)
        pub = ops.empty(len(self._dirs), N, 8)
        matching = bool(self.do_w_loss_matching)
        for di, d in enumerate(self._dirs):
            rec = fw[d]
            gen = self._nets['gen_' + d]
            ring = self._rings[d]
            hpd = {'world': self.world, 'hist_size': self.los_matching_hist_size_conf, 'head_gan': ring['head_gan'],
                   'head_council': ring['head_council'], 'gan_on': int(gan_on), 'council_on': int(council_on),
                   'focus_on': int(focus_on), 'matching': int(matching), 'small_abs': int(bool(fl['mask_small_use_abs'])),
                   'small_square': int(bool(fl['mask_small_use_square'])), 'gan_w': float(hp['gan_w']),
                   'council_w': float(hp['council_w']), 'w01': float(hp['mask_zero_or_one_w']), 'wtot': float(hp['mask_total_w']),
                   'wtv': float(hp['mask_tv_w']), 'numel': float(rec['B'] * self.world * 3 * rec['H'] * rec['W'])}
            d_cl, d_mask = ops.gen_loss_bwd(rec['disc_outs'], rec['mask'] if focus_on else None, center, eps, scal[di], hpd,
                                            ring['gan'], ring['council'], total, di > 0, pub[di], focus_on)
            R = self.los_matching_hist_size_conf + 1
            if gan_on and matching:  # :518-524 append + popleft
                ring['head_gan'] = (ring['head_gan'] + 1) % R
            if council_on and matching:  # :576-586
                ring['head_council'] = (ring['head_council'] + 1) % R
            d_x = None
            if gan_on:
                d_x = self._nets['dis_' + d].backward(rec['d_adv'], rec['dis_saved'], want_wgrad=False, want_dx=True)
            if council_on:
                d_x8 = self._nets['dis_council_' + d].backward(d_cl, rec['disc_saved'], want_wgrad=False, want_dx=True)
                if d_x is None:
                    d_x = ops.zeros(*rec['x_fake'].shape)
                ops.acc_slice(d_x, d_x8, 4)
            if d_x is None:
                d_x = ops.zeros(*rec['x_fake'].shape)
            gen.backward(d_x, d_mask, rec['enc'], rec['dec'],
                         on_decoder_done=(lambda g=gen: self._reduce_async('gen', g, g.enc_end, None))
                         if self.world > 1 and os.environ.get('COUNCIL_DP_SYNC') != '1' else None)
            if self.world > 1 and os.environ.get('COUNCIL_DP_SYNC') != '1':
                self._reduce_async('gen', gen, 0, gen.enc_end)  # encoder bucket; the decoder bucket went out during the encoder backward
        self._adam('gen', defer=True)  # joined at the start of the next update (or by save / state_dict / sample)
        self._enc_cache.clear()

        # ---- publish the reference's loss attributes (:302-322, :556-557): 0-d DEVICE tensors, no sync here -------------
        self.loss_gen_total_s = [total[i] for i in range(N)]
        for d in _DIRS:
            ab = 'ab' if d == 'a2b' else 'ba'
            if d in fw:
                di = self._dirs.index(d)

                def col(k, on=True):
                    return [pub[di, i, k] for i in range(N)] if on else []
                setattr(self, 'loss_gen_adv_%s_s' % d, col(1, gan_on))
                setattr(self, 'loss_gen_mask_zero_one_%s_s' % ab, col(2, focus_on and hp['mask_zero_or_one_w'] != 0))
                setattr(self, 'loss_gen_mask_total_%s_s' % ab, col(3) if focus_on and hp['mask_total_w'] != 0 else [0] * N)
                setattr(self, 'loss_gen_mask_TV_%s_s' % ab, col(4) if focus_on and hp['mask_tv_w'] != 0 else [0] * N)
                setattr(self, 'council_loss_%s_s' % ab, col(5) if council_on else [0] * N)
                if council_on and matching:
                    setattr(self, 'w_match_%s_conf' % d, pub[di, N - 1, 6])  # the reference keeps the last member's ratio (:583)
            else:
                setattr(self, 'loss_gen_adv_%s_s' % d, [0] * N if gan_on else [])
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

    @_pinned
    def sample(self, x_a=None, x_b=None, s_a=None, s_b=None, council_member_to_sample_vec=None, return_mask=True):
        """Translation of every image by every member (:643-733): returns the same 8-tuple, rows ordered image-major /
        member-minor like the reference's double loop.  All members and all images run as ONE stacked pass per network
        (the reference runs batch-1 passes in a Python double loop); instance norm / AdaIN are per-sample, so the result
        per image is the same.  (eval() == train() for this network: no dropout, no running statistics.)"""
        self._flush()
        ops, N = self.ops, self.council_size
        members = list(range(N)) if council_member_to_sample_vec is None else list(council_member_to_sample_vec)
        res = {}
        for d in _DIRS:
            if not getattr(self, 'do_%s_conf' % d):
                res[d] = (None, None, None, None)
                continue
            x = x_a if d == 'a2b' else x_b
            if x.dim() == 5:
                raise ValueError('sample() takes NCHW images like the reference (trainer_council.py:643)')
            B = x.size(0)
            fixed = (self.s_b if s_b is None else s_b) if d == 'a2b' else (self.s_a if s_a is None else s_a)
            s2 = torch.randn(B, self.style_dim, 1, 1)  # :649 / :655
            gen = self._nets['gen_' + d]
            img = ops.nchw_to_nhwc(x.detach().to(ops.device, ops.dtype).contiguous(), IMG_C)[None]

            def st(t):  # [B, S, 1, 1] -> shared style [1, B, 1, 1, S]
                return t[:B].to(ops.device, ops.dtype).reshape(1, B, 1, 1, self.style_dim).contiguous()
            c = gen.encode(img)
            first, mask1 = gen.decode(c, st(fixed), img)
            third, _ = gen.decode(c, st(s2), img)
            if return_mask:
                second = mask1
            else:
                second, _ = gen.decode(c, gen.style_encode(img), img)  # recon with each member's own style code

            def rows(t):  # [N, B, H, W, 4] -> [B * len(members), 3, H, W], image-major
                t = ops.nhwc_to_nchw(t, 3)[members]
                return t.permute(1, 0, 2, 3, 4).reshape(B * len(members), 3, t.shape[-2], t.shape[-1])
            xs = x.detach().to(ops.device, ops.dtype).repeat_interleave(len(members), dim=0)
            res[d] = (xs, rows(second), rows(first), rows(third))
        return res['a2b'] + res['b2a']

    def forward(self, *args, **kwargs):
        raise NotImplementedError('Council_Trainer.forward is broken in the reference (trainer_council.py:267 '
                                  'references a nonexistent self.gen_a2b); use sample()')

    def save(self, snapshot_dir, iterations):
        """Per-member checkpoint files with the reference's names, keys and optimiser layout (:969-992): the reference's
        resume() loads them, and ours loads the reference's."""
        self._flush()
        for i in range(self.council_size):
            for fam in ('gen', 'dis', 'dis_council'):
                if fam == 'dis_council' and not self.do_dis_council:
                    continue
                for d in self._dirs:
                    name = os.path.join(snapshot_dir, '%s_%s_%d_%08d.pt' % (d, fam, i, iterations + 1))
                    torch.save({d: getattr(self, '%s_%s_s' % (fam, d))[i].state_dict()}, name)
            opt = {fam: self._opt_state_dict(fam, i) for fam in ('gen', 'dis', 'dis_council')
                   if fam != 'dis_council' or self.do_dis_council}
            torch.save(opt, os.path.join(snapshot_dir, 'optimizer_%d.pt' % i))

    def _opt_params(self, fam):
        """The parameter list of the reference's per-member optimiser (:152-179): a2b network then b2a network, each in
        nn.Module.parameters() order (= state_dict order without buffers).  -> [(net, spec, is_weight)]"""
        out = []
        for d in self._dirs:
            net = self._nets['%s_%s' % (fam, d)]
            by_key = {}
            for spec in net._specs():
                by_key[spec.wname] = (net, spec, True)
                by_key[spec.bname] = (net, spec, False)
            out += [by_key[k] for k in net.reference_key_order() if k in by_key]
        return out

    def _opt_state_dict(self, fam, i):
        """torch.optim.Adam.state_dict() of member i's optimiser for one family, from the flat moment buffers.
        Parameters that never receive a gradient (the style encoder) have no state entry, as in the reference."""
        state = {}
        plist = self._opt_params(fam)
        for idx, (net, spec, is_w) in enumerate(plist):
            name = spec.wname if is_w else spec.bname
            bank = net._bank_of(name)
            if not bank.trainable or bank.step == 0:
                continue
            m, v = bank._view(bank.exp_avg, name)[i], bank._view(bank.exp_avg_sq, name)[i]
            if is_w:
                m, v = spec.export_weight(m), spec.export_weight(v)
            state[idx] = {'step': torch.tensor(float(bank.step)), 'exp_avg': m.detach().clone().cpu(),
                          'exp_avg_sq': v.detach().clone().cpu()}
        group = {'lr': self._lr(fam), 'betas': tuple(self._betas), 'eps': 1e-8, 'weight_decay': self._wd, 'amsgrad': False,
                 'maximize': False, 'foreach': None, 'capturable': False, 'differentiable': False, 'fused': None,
                 'decoupled_weight_decay': False, 'initial_lr': self._lr0, 'params': list(range(len(plist)))}
        return {'state': state, 'param_groups': [group]}

    def _load_opt_state_dict(self, fam, i, sd):
        """Inverse of _opt_state_dict; accepts what torch.optim.Adam.state_dict() of the reference wrote (:988-992)."""
        plist = self._opt_params(fam)
        steps = []
        for idx, ent in sd.get('state', {}).items():
            net, spec, is_w = plist[int(idx)]
            name = spec.wname if is_w else spec.bname
            bank = net._bank_of(name)
            if not bank.trainable:
                continue
            for key, buf in (('exp_avg', bank.exp_avg), ('exp_avg_sq', bank.exp_avg_sq)):
                dst = bank._view(buf, name)[i]
                if is_w:
                    spec.import_weight(dst, ent[key])
                else:
                    dst.copy_(ent[key].detach().to('cpu', dst.dtype).reshape(-1))
            steps.append(int(float(ent['step'])))
        return max(steps) if steps else None

    def resume(self, checkpoint_dir, hyperparameters):
        """Load the latest per-member checkpoints (:898-967); returns the iteration parsed from the file name."""
        self._flush()
        iterations = 0
        steps = {}  # the flat Adam keeps ONE step count per family (the reference: one per parameter, all equal in practice)
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
            try:
                opt = torch.load(opt_path, map_location='cpu')
                for fam in ('dis', 'gen') + (('dis_council',) if self.do_dis_council else ()):
                    st = self._load_opt_state_dict(fam, i, opt[fam])
                    if st is not None:
                        steps[fam] = max(steps.get(fam, 0), st)
            except Exception as e:  # the reference warns and carries on as well (:958-959)
                import warnings
                warnings.warn('some optimizer FAILED to load (%s: %s): Adam moments restart from zero' % (type(e).__name__, e))
        for fam, st in steps.items():
            for d in self._dirs:
                self._nets['%s_%s' % (fam, d)].bank.step = st
        if iterations > 0:
            print('Resume from iteration %d' % iterations)
            for fam in self._sched_epoch:  # get_scheduler(..., last_epoch=iterations) :953-957
                self._sched_epoch[fam] = iterations
        else:
            import warnings
            warnings.warn('FAILED TO RESUME STARTED FROM 0')
        self.sync_parameters()
        return iterations
