"""CPU oracle for the Council-GAN training step  --  TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch (CPU, fp32, autograd) *restatement* of the reference's
algorithm for the one hot path this repo accelerates:

    dis_update -> dis_council_update -> gen_update          (trainer_council.py:735,782,280)
    over AdaINGen / MsImageDis / MsImageDisCouncil          (networks.py:223,17,116)

It is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product package
(``council_gan_b200``) never imports anything under ``oracle/``.

Pinning: the reference ships no tests and no golden vectors for this path, so
the oracle is pinned against outputs of the *unmodified reference itself*, run in the build
container by ``oracle/make_golden.py`` (which imports /root/reference read-only) and committed as
small JSON fixtures under ``tests/golden/``.  ``tests/test_oracle_golden.py`` checks this
restatement against those fixtures.

Every function cites the reference file:line it restates (paths relative to /root/reference).
Tensors are NCHW fp32 exactly as in the reference; parameters live in plain dicts keyed by the
reference's ``state_dict`` key names.
"""
from __future__ import annotations

import math
import random
from collections import deque

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# parameter inventory (reference state_dict keys and shapes)
# ----------------------------------------------------------------------------------------------

def gen_param_shapes(hp):
    """Key -> shape for one AdaINGen (networks.py:223-254), in ``state_dict()`` order.

    AdaIN ``running_mean/var`` dummy buffers (networks.py:637-638) are listed with a ``#buf`` tag.
    """
    g = hp['gen']
    dim, sd, nd, nr, mlp = g['dim'], g['style_dim'], g['n_downsample'], g['n_res'], g['mlp_dim']
    cin = hp['input_dim_a']
    nm = g['num_of_mask_dim_to_add']
    out = []

    def conv(prefix, co, ci, k):
        out.append((prefix + '.weight', (co, ci, k, k)))
        out.append((prefix + '.bias', (co,)))

    # StyleEncoder(4, ...) networks.py:337-350  (dead compute at the BASELINE configs)
    d = dim
    conv('enc_style.model.0.conv', d, cin, 7)
    for i in range(2):
        conv('enc_style.model.%d.conv' % (1 + i), 2 * d, d, 4)
        d *= 2
    for i in range(4 - 2):
        conv('enc_style.model.%d.conv' % (3 + i), d, d, 4)
    conv('enc_style.model.6', sd, d, 1)
    # ContentEncoder networks.py:355-366
    d = dim
    conv('enc_content.model.0.conv', d, cin, 7)
    for i in range(nd):
        conv('enc_content.model.%d.conv' % (1 + i), 2 * d, d, 4)
        d *= 2
    for r in range(nr):
        for j in range(2):
            conv('enc_content.model.%d.model.%d.model.%d.conv' % (1 + nd, r, j), d, d, 3)
    # Decoder_V2_atten networks.py:374-396
    adain = []
    for r in range(nr):
        for j in range(2):
            p = 'dec.model.0.model.%d.model.%d' % (r, j)
            out.append((p + '.norm.running_mean#buf', (d,)))
            out.append((p + '.norm.running_var#buf', (d,)))
            conv(p + '.conv', d, d, 3)
            adain.append(d)
    idx = 1
    for i in range(nd):
        idx += 1  # nn.Upsample occupies an index
        for co, ci in ((d // 2, d), (d // 2, d // 2)):
            p = 'dec.model.%d' % idx
            out.append((p + '.norm.running_mean#buf', (co,)))
            out.append((p + '.norm.running_var#buf', (co,)))
            conv(p + '.conv', co, ci, 3)
            adain.append(co)
            idx += 1
        d //= 2
    conv('dec.model.%d.conv' % idx, d, d, 1)
    conv('dec.model.%d.conv' % (idx + 1), d, d, 1)
    conv('dec.model.%d.conv' % (idx + 2), cin * nm + nm, d, 1)
    n_adain = 2 * sum(adain)
    # MLP networks.py:432-440
    out.append(('mlp.model.0.fc.weight', (mlp, sd)))
    out.append(('mlp.model.0.fc.bias', (mlp,)))
    out.append(('mlp.model.1.fc.weight', (mlp, mlp)))
    out.append(('mlp.model.1.fc.bias', (mlp,)))
    out.append(('mlp.model.2.fc.weight', (n_adain, mlp)))
    out.append(('mlp.model.2.fc.bias', (n_adain,)))
    return out


def dis_param_shapes(hp, council=False):
    """MsImageDis._make_net (networks.py:37-46) / MsImageDisCouncil._make_net (networks.py:134-145)."""
    dp = hp['dis']
    cin = hp['input_dim_a']
    out = []
    for s in range(dp['num_scales']):
        d = dp['dim']
        if council:
            out.append(('cnns.%d.0.conv.weight' % s, (d, 2 * cin, 3, 3)))
        else:
            out.append(('cnns.%d.0.conv.weight' % s, (d, cin, 4, 4)))
        out.append(('cnns.%d.0.conv.bias' % s, (d,)))
        for i in range(dp['n_layer'] - 1):
            out.append(('cnns.%d.%d.conv.weight' % (s, i + 1), (2 * d, d, 4, 4)))
            out.append(('cnns.%d.%d.conv.bias' % (s, i + 1), (2 * d,)))
            d *= 2
        n = dp['n_layer']
        if council:
            out.append(('cnns.%d.%d.weight' % (s, n), (d, d, 1, 1)))
            out.append(('cnns.%d.%d.bias' % (s, n), (d,)))
            n += 1
        out.append(('cnns.%d.%d.weight' % (s, n), (1, d, 1, 1)))
        out.append(('cnns.%d.%d.bias' % (s, n), (1,)))
    return out


def synth_state(shapes, seed, kind):
    """Deterministic synthetic parameters with the statistics of the reference initialisers.

    kind='kaiming'  -> N(0, sqrt(2/fan_in))  (utils.py:412 ``kaiming_normal_(a=0, mode='fan_in')``)
    kind='gaussian' -> N(0, 0.02)            (utils.py:408)
    Biases get small non-zero values (reference initialises them to 0, utils.py:418-419, but a
    non-zero bias exercises more of the data path).  AdaIN dummy buffers keep 0 / 1.
    One ``torch.Generator`` per tensor so that the stream does not depend on iteration order.
    """
    sd = {}
    for n, (key, shape) in enumerate(shapes):
        gen = torch.Generator().manual_seed(seed * 100003 + n)
        if key.endswith('#buf'):
            k = key[:-4]
            sd[k] = torch.zeros(shape) if k.endswith('running_mean') else torch.ones(shape)
        elif key.endswith('.bias'):
            sd[key] = torch.randn(shape, generator=gen) * 0.01
        else:
            fan_in = int(np.prod(shape[1:]))
            std = math.sqrt(2.0 / fan_in) if kind == 'kaiming' else 0.02
            sd[key] = torch.randn(shape, generator=gen) * std
    return sd


def synth_inputs(batch, size, seed=123, channels=3):
    """x_a, x_b = rand*2-1 from a dedicated generator (utils.py:124-126 range)."""
    gen = torch.Generator().manual_seed(seed)
    x_a = torch.rand(batch, channels, size, size, generator=gen) * 2 - 1
    x_b = torch.rand(batch, channels, size, size, generator=gen) * 2 - 1
    return x_a, x_b


# ----------------------------------------------------------------------------------------------
# blocks (networks.py:463-521, 627-656)
# ----------------------------------------------------------------------------------------------

def conv_block(p, prefix, x, stride, pad, norm='none', act='relu', adain=None):
    """Conv2dBlock.forward networks.py:515-521: ZeroPad2d -> Conv2d(bias) -> norm -> activation."""
    x = F.conv2d(F.pad(x, (pad, pad, pad, pad)), p[prefix + '.weight'], p[prefix + '.bias'], stride)
    if norm == 'in':  # nn.InstanceNorm2d(C): affine=False, eps=1e-5, biased var (networks.py:483)
        x = F.instance_norm(x, eps=1e-5)
    elif norm == 'adain':  # AdaptiveInstanceNorm2d.forward networks.py:640-653
        b, c = x.shape[:2]
        weight, bias = adain
        x = F.batch_norm(x.reshape(1, b * c, *x.shape[2:]), None, None, weight, bias, True, 0.1, 1e-5)
        x = x.view(b, c, *x.shape[2:])
    if act == 'relu':
        x = F.relu(x)
    elif act == 'lrelu':
        x = F.leaky_relu(x, 0.2)
    elif act == 'tanh':
        x = torch.tanh(x)
    return x


def content_encode(p, hp, x):
    """ContentEncoder.forward networks.py:355-369."""
    g = hp['gen']
    x = conv_block(p, 'enc_content.model.0.conv', x, 1, 3, 'in', 'relu')
    for i in range(g['n_downsample']):
        x = conv_block(p, 'enc_content.model.%d.conv' % (1 + i), x, 2, 1, 'in', 'relu')
    base = 'enc_content.model.%d' % (1 + g['n_downsample'])
    for r in range(g['n_res']):  # ResBlock.forward networks.py:457-461
        res = x
        x = conv_block(p, '%s.model.%d.model.0.conv' % (base, r), x, 1, 1, 'in', 'relu')
        x = conv_block(p, '%s.model.%d.model.1.conv' % (base, r), x, 1, 1, 'in', 'none')
        x = x + res
    return x


def style_encode(p, hp, x):
    """StyleEncoder.forward networks.py:337-353 (dead at BASELINE configs; kept for encode() API)."""
    x = conv_block(p, 'enc_style.model.0.conv', x, 1, 3, 'none', 'relu')
    for i in range(1, 5):
        x = conv_block(p, 'enc_style.model.%d.conv' % i, x, 2, 1, 'none', 'relu')
    x = F.adaptive_avg_pool2d(x, 1)
    return F.conv2d(x, p['enc_style.model.6.weight'], p['enc_style.model.6.bias'])


def mlp(p, s):
    """MLP.forward networks.py:442-443; LinearBlock networks.py:562-568."""
    h = s.view(s.size(0), -1)
    h = F.relu(F.linear(h, p['mlp.model.0.fc.weight'], p['mlp.model.0.fc.bias']))
    h = F.relu(F.linear(h, p['mlp.model.1.fc.weight'], p['mlp.model.1.fc.bias']))
    return F.linear(h, p['mlp.model.2.fc.weight'], p['mlp.model.2.fc.bias'])


def decode(p, hp, content, style, images):
    """AdaINGen.decode networks.py:285-301 + assign_adain_params :303-312 + Decoder_V2_atten.forward :398-415.

    Returns (new_im, mask_s).
    """
    g = hp['gen']
    params = mlp(p, style)
    off = [0]

    def take(c):  # networks.py:308-312: first C columns -> bias ("mean"), next C -> weight ("std")
        mean = params[:, off[0]:off[0] + c].contiguous().view(-1)
        std = params[:, off[0] + c:off[0] + 2 * c].contiguous().view(-1)
        off[0] += 2 * c
        return std, mean

    x = content
    d = x.shape[1]
    for r in range(g['n_res']):
        res = x
        x = conv_block(p, 'dec.model.0.model.%d.model.0.conv' % r, x, 1, 1, 'adain', 'relu', take(d))
        x = conv_block(p, 'dec.model.0.model.%d.model.1.conv' % r, x, 1, 1, 'adain', 'none', take(d))
        x = x + res
    idx = 1
    for i in range(g['n_downsample']):
        x = F.interpolate(x, scale_factor=2)  # nn.Upsample(scale_factor=2): nearest (networks.py:385)
        idx += 1
        x = conv_block(p, 'dec.model.%d.conv' % idx, x, 1, 1, 'adain', 'relu', take(d // 2))
        x = conv_block(p, 'dec.model.%d.conv' % (idx + 1), x, 1, 1, 'adain', 'relu', take(d // 2))
        idx += 2
        d //= 2
    x = conv_block(p, 'dec.model.%d.conv' % idx, x, 1, 0, 'none', 'relu')
    x = conv_block(p, 'dec.model.%d.conv' % (idx + 1), x, 1, 0, 'none', 'relu')
    new_x = conv_block(p, 'dec.model.%d.conv' % (idx + 2), x, 1, 0, 'none', 'tanh')
    # networks.py:400-407
    nm = g['num_of_mask_dim_to_add']
    od = images.shape[1]
    mask_s = (torch.tanh(10 * new_x[:, -nm:]) + 1) / 2
    new_im = images
    for k in range(nm):
        o = new_x[:, od * k:od * (k + 1)]
        m = mask_s[:, k:k + 1]
        new_im = (1 - m) * new_im + m * o
    return new_im, mask_s


def _avgpool(x):
    """nn.AvgPool2d(3, stride=2, padding=[1,1], count_include_pad=False) networks.py:32,129."""
    return F.avg_pool2d(x, 3, 2, 1, count_include_pad=False)


def ms_dis(p, hp, x):
    """MsImageDis.forward networks.py:48-54."""
    dp = hp['dis']
    outs = []
    for s in range(dp['num_scales']):
        h = x
        for i in range(dp['n_layer']):
            h = conv_block(p, 'cnns.%d.%d.conv' % (s, i), h, 2, 1, 'none', 'lrelu')
        n = dp['n_layer']
        outs.append(F.conv2d(h, p['cnns.%d.%d.weight' % (s, n)], p['cnns.%d.%d.bias' % (s, n)]))
        x = _avgpool(x)
    return outs


def ms_dis_council(p, hp, x, x_input):
    """MsImageDisCouncil.forward networks.py:147-156."""
    dp = hp['dis']
    outs = []
    for s in range(dp['num_scales']):
        h = torch.cat((x, x_input), 1)
        h = conv_block(p, 'cnns.%d.0.conv' % s, h, 1, 1, 'none', 'lrelu')
        for i in range(1, dp['n_layer']):
            h = conv_block(p, 'cnns.%d.%d.conv' % (s, i), h, 2, 1, 'none', 'lrelu')
        n = dp['n_layer']
        h = F.conv2d(h, p['cnns.%d.%d.weight' % (s, n)], p['cnns.%d.%d.bias' % (s, n)])
        outs.append(F.conv2d(h, p['cnns.%d.%d.weight' % (s, n + 1)], p['cnns.%d.%d.bias' % (s, n + 1)]))
        x = _avgpool(x)
        x_input = _avgpool(x_input)
    return outs


def lsgan_dis_loss(outs_fake, outs_real):
    """calc_dis_loss, lsgan branch networks.py:62-64 / :164-166."""
    loss = 0
    for o0, o1 in zip(outs_fake, outs_real):
        loss = loss + torch.mean((o0 - 0) ** 2) + torch.mean((o1 - 1) ** 2)
    return loss


def lsgan_gen_loss(outs_fake):
    """calc_gen_loss, lsgan branch networks.py:88-90 / :192-194."""
    loss = 0
    for o0 in outs_fake:
        loss = loss + torch.mean((o0 - 1) ** 2)
    return loss


# focus-loss criteria, trainer_council.py:230-250
def mask_zero_one(mask, center, eps):
    return torch.sum(1 / (torch.abs(mask - center) + eps)) / mask.numel()


def mask_small(mask, use_abs, use_square):
    loss = 0
    if use_abs:
        loss = loss + torch.abs(torch.sum(mask)) / mask.numel()
    if use_square:
        loss = loss + (torch.sum(mask) / mask.numel()) ** 2
    return loss


def mask_tv(mask):
    return (torch.sum(torch.abs(mask[:, :, 1:, :] - mask[:, :, :-1, :])) +
            torch.sum(torch.abs(mask[:, :, :, 1:] - mask[:, :, :, :-1]))) / mask.numel()


# ----------------------------------------------------------------------------------------------
# Adam exactly as torch.optim.Adam (the reference's optimiser, trainer_council.py:170-179)
# ----------------------------------------------------------------------------------------------

class Adam:
    """torch.optim.Adam defaults: L2 weight decay added to the gradient, eps=1e-8, no amsgrad.
    Parameters whose ``.grad`` is None are skipped entirely (no decay, no step) like torch does."""

    def __init__(self, params, lr, betas, weight_decay):
        self.params = list(params)
        self.lr, self.betas, self.wd = lr, betas, weight_decay
        self.state = {}

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    @torch.no_grad()
    def step(self):
        b1, b2 = self.betas
        for p in self.params:
            if p.grad is None:
                continue
            st = self.state.setdefault(id(p), {'step': 0, 'm': torch.zeros_like(p), 'v': torch.zeros_like(p)})
            st['step'] += 1
            g = p.grad + self.wd * p if self.wd != 0 else p.grad
            st['m'].mul_(b1).add_(g, alpha=1 - b1)
            st['v'].mul_(b2).addcmul_(g, g, value=1 - b2)
            bc1 = 1 - b1 ** st['step']
            bc2 = 1 - b2 ** st['step']
            denom = (st['v'].sqrt() / math.sqrt(bc2)).add_(1e-8)
            p.addcdiv_(st['m'], denom, value=-self.lr / bc1)


# ----------------------------------------------------------------------------------------------
# the trainer step
# ----------------------------------------------------------------------------------------------

class OracleTrainer:
    """Restates Council_Trainer's three updates (trainer_council.py) for the live configuration
    space: lsgan, do_a2b and/or do_b2a, recon_* = vgg_w = abs_beginning_end = council_abs_w = 0.

    ``states`` = {'gen_a2b': [sd_0..sd_{N-1}], 'dis_a2b': [...], 'dis_council_a2b': [...], and/or *_b2a}
    with reference state_dict keys.  Tensors are cloned into leaf parameters.
    """

    DIRS = ('a2b', 'b2a')

    def __init__(self, hp, states):
        self.hp = hp
        self.N = hp['council']['council_size']
        self.style_dim = hp['gen']['style_dim']
        self.dirs = [d for d in self.DIRS if hp['do_' + d]]
        for k in ('recon_x_w', 'recon_s_w', 'recon_c_w', 'recon_x_cyc_w', 'vgg_w', 'abs_beginning_end', 'council_abs_w'):
            assert hp[k] == 0, 'oracle covers the live configuration space only (%s != 0)' % k
        assert hp['dis']['gan_type'] == 'lsgan'
        self.do_dis_council = hp['council_w'] != 0  # trainer_council.py:31
        self.P = {}
        for name, lst in states.items():
            self.P[name] = [{k: v.clone().requires_grad_(not k.endswith(('running_mean', 'running_var')))
                             for k, v in sd.items()} for sd in lst]
        lr, betas, wd = hp['lr'], (hp['beta1'], hp['beta2']), hp['weight_decay']

        def opt(fam):  # trainer_council.py:152-179 (one optimiser per member covering both directions)
            out = []
            for i in range(self.N):
                ps = []
                for d in self.dirs:
                    ps += [v for v in self.P['%s_%s' % (fam, d)][i].values() if v.requires_grad]
                out.append(Adam(ps, lr, betas, wd))
            return out

        self.gen_opt = opt('gen')
        self.dis_opt = opt('dis')
        self.dis_council_opt = opt('dis_council') if self.do_dis_council else None
        hist = hp['loss_matching_hist_size']
        self.hist_gan = {d: [deque(np.ones(hist)) for _ in range(self.N)] for d in self.dirs}  # :81-92
        self.hist_council = {d: [deque(np.ones(hist)) for _ in range(self.N)] for d in self.dirs}
        self.sched_steps = 0

    # -- helpers ------------------------------------------------------------------------------
    def lr_now(self):
        """StepLR(step_size, gamma) utils.py:392-400, advanced by update_learning_rate :885-896."""
        hp = self.hp
        if hp.get('lr_policy', 'constant') == 'constant':
            return hp['lr']
        return hp['lr'] * hp['gamma'] ** (self.sched_steps // hp['step_size'])

    def update_learning_rate(self):
        self.sched_steps += 1
        lr = self.lr_now()
        for opts in (self.gen_opt, self.dis_opt, self.dis_council_opt or []):
            for o in opts:
                o.lr = lr

    def _src(self, d, x_a, x_b):
        return x_a if d == 'a2b' else x_b

    def _real(self, d, x_a, x_b):
        return x_b if d == 'a2b' else x_a

    def _council_active(self, hp, for_gen):
        """Gating: trainer_council.py:541-555 (gen_update) and :787-801 (dis_council_update)."""
        c = hp['council']
        cyc = hp['iteration'] % (c['flipOnOff_On_iteration'] + c['flipOnOff_Off_iteration'])
        start = c['flipOnOff_On_iteration'] if c['flipOnOff_start_with'] else c['flipOnOff_Off_iteration']
        do = c['flipOnOff_start_with'] if cyc < start else (not c['flipOnOff_start_with'])
        if not c['flipOnOff']:
            do = True if for_gen else c['flipOnOff_start_with']
        if hp['iteration'] < c['council_start_at_iter']:
            do = False
        return do

    # -- dis_update  trainer_council.py:735-780 ------------------------------------------------
    def dis_update(self, x_a, x_b, hp):
        assert not hp['dis']['do_Dis_only_gray'] and not hp['dis']['useRandomGen']
        for o in self.dis_opt:
            o.zero_grad()
        s = {}
        if 'a2b' in self.dirs:  # :740-745, draw order a2b (s_b) then b2a (s_a)
            s['a2b'] = torch.randn(x_b.size(0), self.style_dim, 1, 1).to(x_b.device)
        if 'b2a' in self.dirs:
            s['b2a'] = torch.randn(x_a.size(0), self.style_dim, 1, 1).to(x_a.device)
        self.loss_dis_total_s = []
        self.x_fake_dis = {d: [] for d in self.dirs}
        for i in range(self.N):
            total = 0
            for d in self.dirs:
                g = self.P['gen_' + d][i]
                src = self._src(d, x_a, x_b)
                with torch.no_grad():  # the reference builds then discards this graph (.detach() at :769/771)
                    c = content_encode(g, hp, src)
                    x_fake, _ = decode(g, hp, c, s[d], src)
                self.x_fake_dis[d].append(x_fake)
                dp = self.P['dis_' + d][i]
                loss = lsgan_dis_loss(ms_dis(dp, hp, x_fake), ms_dis(dp, hp, self._real(d, x_a, x_b)))
                # :775 applies gan_w on the a2b branch only; :777 does not on b2a
                total = total + (hp['gan_w'] * loss if d == 'a2b' else loss)
            self.loss_dis_total_s.append(total)
            total.backward()
            self.dis_opt[i].step()

    # -- dis_council_update  trainer_council.py:782-883 ------------------------------------------
    def dis_council_update(self, x_a, x_b, hp):
        c = hp['council']
        if self.N <= 1 or c['numberOfCouncil_dis_relative_iteration'] == 0:
            return False
        if (not self._council_active(hp, for_gen=False)) or hp['council_w'] == 0 or \
                hp['iteration'] < c['council_start_at_iter']:
            return False
        for o in self.dis_council_opt:
            o.zero_grad()
        s = {}
        if 'b2a' in self.dirs:  # :806-809, draw order b2a (s_a) then a2b (s_b)
            s['b2a'] = torch.randn(x_a.size(0), self.style_dim, 1, 1).to(x_a.device)
        if 'a2b' in self.dirs:
            s['a2b'] = torch.randn(x_b.size(0), self.style_dim, 1, 1).to(x_b.device)
        less = c['discriminetro_less_style_by']
        fake = {d: [] for d in self.dirs}
        comp = {d: [] for d in self.dirs}
        for i in range(self.N):  # :826-851
            for d in self.dirs:
                g = self.P['gen_' + d][i]
                src = self._src(d, x_a, x_b)
                with torch.no_grad():
                    cc = content_encode(g, hp, src)
                    xf, _ = decode(g, hp, cc, s[d], src)
                    fake[d].append(xf)
                    if less != 0:
                        xl, _ = decode(g, hp, cc, s[d] * less, src)
                        comp[d].append(xl)
                    else:
                        comp[d].append(xf)
        self.x_fake_disc = fake
        self.loss_dis_council_total_s = []
        for i in range(self.N):  # :858-883
            acc = {d: 0 for d in self.dirs}
            pool = list(range(0, i)) + list(range(i + 1, self.N))
            for k in range(c['numberOfCouncil_dis_relative_iteration']):
                if k == self.N:
                    break
                if len(pool) == 0:
                    pool = list(range(0, i)) + list(range(i + 1, self.N))
                j = random.choice(pool)
                pool.remove(j)
                for d in self.dirs:
                    dc = self.P['dis_council_' + d][i]
                    src = self._src(d, x_a, x_b)
                    acc[d] = acc[d] + lsgan_dis_loss(ms_dis_council(dc, hp, fake[d][i], src),
                                                     ms_dis_council(dc, hp, comp[d][j], src))
            total = 0
            for d in self.dirs:  # :877-880 divides by the configured K
                total = total + hp['council_w'] * acc[d] / c['numberOfCouncil_dis_relative_iteration']
            self.loss_dis_council_total_s.append(total)
            total.backward()
            self.dis_council_opt[i].step()
        return True

    # -- gen_update  trainer_council.py:280-634 --------------------------------------------------
    def gen_update(self, x_a, x_b, hp, iterations=0):
        assert not hp['gen']['useRandomDis'] and not hp['dis']['do_Dis_only_gray']
        assert not hp['focus_loss']['do_w_loss_matching_focus']
        fl = hp['focus_loss']
        for o in self.gen_opt:
            o.zero_grad()
        # the reference's D / DC parameters also accumulate (never used) grads here; clear them after
        s_a = torch.randn(x_a.size(0), self.style_dim, 1, 1).to(x_a.device)  # :284-285 both always drawn, a then b
        s_b = torch.randn(x_b.size(0), self.style_dim, 1, 1).to(x_b.device)
        s = {'a2b': s_b, 'b2a': s_a}
        focus_on = hp['iteration'] > fl['focus_loss_start_at_iter'] and \
            (hp['mask_zero_or_one_w'] != 0 or hp['mask_total_w'] != 0)  # :390
        self.loss_gen_total_s = []
        self.loss_gen_adv_s = {d: [] for d in self.dirs}
        self.loss_gen_mask_zero_one_s = {d: [] for d in self.dirs}
        self.loss_gen_mask_total_s = {d: [] for d in self.dirs}
        self.loss_gen_mask_TV_s = {d: [] for d in self.dirs}
        self.council_loss_s = {d: [] for d in self.dirs}
        self.x_fake_gen = {d: [] for d in self.dirs}
        self.mask_gen = {d: [] for d in self.dirs}
        totals = []
        for i in range(self.N):  # loop 1, :328-538
            total = 0
            for d in self.dirs:
                g = self.P['gen_' + d][i]
                src = self._src(d, x_a, x_b)
                cc = content_encode(g, hp, src)
                xf, mask = decode(g, hp, cc, s[d], src)
                self.x_fake_gen[d].append(xf)
                self.mask_gen[d].append(mask)
            if focus_on:
                for d in self.dirs:
                    mask = self.mask_gen[d][i]
                    if hp['mask_zero_or_one_w'] != 0:  # :392-415
                        l01 = mask_zero_one(mask, fl['mask_zero_or_one_center'], fl['mask_zero_or_one_epsilon'])
                        self.loss_gen_mask_zero_one_s[d].append(l01)
                        total = total + hp['mask_zero_or_one_w'] * l01
                    if hp['mask_tv_w'] != 0:  # :425-431 (added to the total before the mask_total term)
                        ltv = mask_tv(mask)
                        self.loss_gen_mask_TV_s[d].append(ltv)
                        total = total + hp['mask_tv_w'] * ltv
                    if hp['mask_total_w'] != 0:  # :418-422, :447-451
                        lt = mask_small(mask, fl['mask_small_use_abs'], fl['mask_small_use_square'])
                        self.loss_gen_mask_total_s[d].append(lt)
                        total = total + hp['mask_total_w'] * lt
            if hp['gan_w'] != 0:  # :497-529
                for d in self.dirs:
                    adv = lsgan_gen_loss(ms_dis(self.P['dis_' + d][i], hp, self.x_fake_gen[d][i]))
                    self.loss_gen_adv_s[d].append(adv)
                    if hp['do_w_loss_matching']:
                        self.hist_gan[d][i].append(adv.detach().cpu().numpy())
                        self.hist_gan[d][i].popleft()
                    total = total + hp['gan_w'] * adv
            totals.append(total)
        do_council = self._council_active(hp, for_gen=True)
        self.w_match = {d: 1 for d in self.dirs}
        for i in range(self.N):  # loop 2, :558-634
            total = totals[i]
            if (hp['council_w'] != 0) and do_council and self.N > 1:
                for d in self.dirs:
                    src = self._src(d, x_a, x_b)
                    cl = lsgan_gen_loss(ms_dis_council(self.P['dis_council_' + d][i], hp, self.x_fake_gen[d][i], src))
                    if hp['do_w_loss_matching']:  # :576-586
                        self.hist_council[d][i].append(cl.detach().cpu().numpy())
                        self.hist_council[d][i].popleft()
                        self.w_match[d] = np.mean(self.hist_gan[d][i]) / np.mean(self.hist_council[d][i])
                        cl = cl * self.w_match[d]
                    cl = cl * hp['council_w']
                    self.council_loss_s[d].append(cl)
                    total = total + cl
            self.loss_gen_total_s.append(total)
            total.backward()
            self.gen_opt[i].step()
        # reference leaves stale grads on D/DC that the next dis_update zeroes (:738-739, :803-804)
        for fam in ('dis', 'dis_council'):
            for d in self.dirs:
                for sd in self.P.get('%s_%s' % (fam, d), []):
                    for v in sd.values():
                        v.grad = None

    def state(self, name, i):
        return {k: v.detach() for k, v in self.P[name][i].items()}


def synth_all_states(hp, seed=7):
    """Synthetic parameters for every network of a trainer, reference-keyed."""
    N = hp['council']['council_size']
    states = {}
    gshapes = gen_param_shapes(hp)
    dshapes = dis_param_shapes(hp, False)
    cshapes = dis_param_shapes(hp, True)
    for di, d in enumerate(('a2b', 'b2a')):
        if not hp['do_' + d]:
            continue
        states['gen_' + d] = [synth_state(gshapes, seed + 1000 * di + 10 * i + 1, 'kaiming') for i in range(N)]
        states['dis_' + d] = [synth_state(dshapes, seed + 1000 * di + 10 * i + 2, 'gaussian') for i in range(N)]
        if hp['council_w'] != 0:
            states['dis_council_' + d] = [synth_state(cshapes, seed + 1000 * di + 10 * i + 3, 'gaussian') for i in range(N)]
    return states


def seed_all(seed):
    """The three RNGs the step consumes (train.py:55-59): python ``random``, numpy, torch CPU."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
