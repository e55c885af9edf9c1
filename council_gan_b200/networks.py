"""Council-stacked networks: AdaIN generator, multi-scale PatchGAN discriminator, council discriminator.

B200-first restructuring of the reference's ``networks.py``: instead of N independent ``nn.Module``
copies executed one after the other (trainer_council.py:101-119, loops at :328,558,747,826,858), the N
council members of one family live in ONE flat fp32 parameter buffer (stacked ``[N, ...]`` per layer) and
every layer is ONE grouped kernel launch over all members.  Forward *and* backward are written out
explicitly (no autograd): each step below is a call into libcouncil_b200.so through ``ops``.

Layer-by-layer correspondence with the reference (paths relative to /root/reference):
  ContentEncoder       networks.py:355-369      -> CouncilGen.encode
  MLP                  networks.py:432-443      -> CouncilGen._mlp
  Decoder_V2_atten     networks.py:374-415      -> CouncilGen.decode
  assign_adain_params  networks.py:303-312      -> column offsets into the MLP output (no copies)
  MsImageDis           networks.py:17-54        -> CouncilDis(council=False)
  MsImageDisCouncil    networks.py:116-156      -> CouncilDis(council=True)
Per-member ``state_dict`` views keep the reference's key names and OIHW shapes (see MemberView).
"""
from __future__ import annotations

import math
import os
from collections import OrderedDict

import torch

from .ops import ACT_LRELU, ACT_NONE, ACT_RELU, ACT_TANH

IMG_C = 4  # image tensors carry 3 live lanes + 1 zero lane (16-byte pixels)


# ------------------------------------------------------------------------------------------------------
# parameter storage
# ------------------------------------------------------------------------------------------------------
class ParamBank:
    """One flat fp32 buffer (+ grad, Adam moments) holding every parameter of a family for all members.

    ``entries``: list of (name, per-member kernel-layout shape).  Each entry is stored ``[G, *shape]``,
    its offset rounded up to 64 floats (256 B) so any slice is a valid TMA / float4 base address.
    The flat ``grad`` buffer is what data parallelism all-reduces; ``data/grad/exp_avg/exp_avg_sq`` are
    what the fused Adam kernel walks.
    """

    def __init__(self, ops, G, entries, trainable=True):
        self.ops, self.G = ops, G
        self._views = {}
        self.table = OrderedDict()
        off = 0
        for name, shape in entries:
            n = G * int(math.prod(shape))
            self.table[name] = (off, tuple(shape), n)
            off += (n + 63) // 64 * 64
        self.total = off
        self.data = ops.zeros(off)
        self.trainable = trainable
        if trainable:
            self.grad = ops.zeros(off)
            self.exp_avg = ops.zeros(off)
            self.exp_avg_sq = ops.zeros(off)
        self.step = 0

    def _view(self, buf, name):
        # views are cached per (buffer object, name): ~430 lookups per step, each a slice + view (8 % of the host time of a step on the
        # launch-bound small configuration)
        hit = self._views.get((id(buf), name))
        if hit is not None and hit[0] is buf:
            return hit[1]
        off, shape, n = self.table[name]
        v = buf[off:off + n].view((self.G,) + shape)
        self._views[(id(buf), name)] = (buf, v)
        return v

    def p(self, name):
        return self._view(self.data, name)

    def g(self, name):
        return self._view(self.grad, name)


class LayerSpec:
    """A convolution (or linear, as 1x1 on a 1x1 map) with its reference key and lane mapping."""

    def __init__(self, key, cout, cin, k, stride, pad, lanes=None, linear=False):
        self.key, self.cout, self.cin, self.k, self.stride, self.pad = key, cout, cin, k, stride, pad
        self.lanes = lanes  # kernel lane index of each reference input channel (None: identity)
        self.cin_p = cin if lanes is None else (max(lanes) // 4 + 1) * 4
        self.linear = linear
        self.wname = key + '.weight'
        self.bname = key + '.bias'

    def entries(self):
        return [(self.wname, (self.cout, self.k, self.k, self.cin_p)), (self.bname, (self.cout,))]

    # reference <-> kernel layout
    def export_weight(self, w):  # w: [Cout,KH,KW,Cin_p] one member
        if self.lanes is not None:
            w = w[..., self.lanes]
        w = w.permute(0, 3, 1, 2)
        return (w.reshape(self.cout, self.cin) if self.linear else w).contiguous().clone()

    def import_weight(self, dst, ref):  # dst: [Cout,KH,KW,Cin_p] view; ref: reference tensor
        """OIHW -> the kernel layout, assembled on the HOST and moved with one plain copy (no permute / fill kernels on the
        device: checkpoint loading stays out of the kernel launch list)."""
        ref = ref.detach().to('cpu', dst.dtype)
        if self.linear:
            ref = ref.reshape(self.cout, self.cin, 1, 1)
        ref = ref.permute(0, 2, 3, 1)
        if self.lanes is not None:
            full = torch.zeros(dst.shape, dtype=dst.dtype)
            full[..., self.lanes] = ref
            ref = full
        dst.copy_(ref.contiguous())


class MemberView:
    """Reference-shaped view of council member ``i`` of a stacked network (``gen_a2b_s[i]`` etc.).

    Provides what the reference's callers use on those objects (train.py / test_on_folder.py /
    trainer_council.py:898-992): ``state_dict`` / ``load_state_dict`` with the reference's keys and OIHW
    shapes, ``cuda_device``, ``eval`` / ``train``, and for generators ``encode`` / ``decode`` /
    ``dec.mask_s``.
    """

    def __init__(self, net, i):
        self.net, self.i = net, i
        self.cuda_device = str(net.ops.device)
        self.training = True

    def eval(self):
        self.training = False
        return self

    def train(self, mode=True):
        self.training = mode
        return self

    def state_dict(self):
        return self.net.member_state_dict(self.i)

    def load_state_dict(self, sd, strict=True):
        self.net.load_member_state_dict(self.i, sd, strict)

    def parameters(self):
        return [v for k, v in self.state_dict().items() if not k.endswith(('running_mean', 'running_var'))]


class _StackedNet:
    """Shared state_dict plumbing for the three stacked network types."""

    def _specs(self):
        raise NotImplementedError

    def _banks(self):
        raise NotImplementedError

    def _bank_of(self, name):
        for b in self._banks():
            if name in b.table:
                return b
        raise KeyError(name)

    def extra_state(self):
        return OrderedDict()

    _before_access = None  # set by the trainer: joins a pending (data-parallel, deferred) optimiser step of this family

    def _sync(self):
        if self._before_access is not None:
            self._before_access()

    def member_state_dict(self, i):
        self._sync()
        sd = OrderedDict()
        for spec in self._specs():
            b = self._bank_of(spec.wname)
            sd[spec.wname] = spec.export_weight(b.p(spec.wname)[i])
            sd[spec.bname] = b.p(spec.bname)[i].clone()
        for k, v in self.extra_state().items():
            sd[k] = v.clone()
        # reference key order (state_dict of the nn.Module tree)
        order = self.reference_key_order()
        return OrderedDict((k, sd[k]) for k in order)

    def load_member_state_dict(self, i, sd, strict=True):
        self._sync()
        keys = set(self.reference_key_order())
        if strict:
            missing, unexpected = keys - set(sd), set(sd) - keys
            if missing or unexpected:
                raise RuntimeError('state_dict mismatch: missing %s unexpected %s' % (sorted(missing), sorted(unexpected)))
        for spec in self._specs():
            b = self._bank_of(spec.wname)
            if spec.wname in sd:
                spec.import_weight(b.p(spec.wname)[i], sd[spec.wname])
            if spec.bname in sd:
                b.p(spec.bname)[i].copy_(sd[spec.bname].detach().to('cpu', b.data.dtype).reshape(-1))
        self.params_changed()

    def params_changed(self):
        self._load_epoch = getattr(self, '_load_epoch', 0) + 1

    def member(self, i):
        return MemberView(self, i)


# ------------------------------------------------------------------------------------------------------
# generator
# ------------------------------------------------------------------------------------------------------
class _DecView:
    def __init__(self):
        self.mask_s = []


class GenMemberView(MemberView):
    """``gen_a2b_s[i]``: encode/decode of a single member through the stacked kernels (G sliced to 1)."""

    def __init__(self, net, i):
        super().__init__(net, i)
        self.dec = _DecView()

    def encode(self, images):
        """AdaINGen.encode networks.py:278-283 -> (content NCHW, style_fake [B,style_dim,1,1])."""
        return self.net.member_encode(self.i, images)

    def decode(self, content, style, images, return_mask=False):
        """AdaINGen.decode networks.py:285-301; sets ``self.dec.mask_s`` like networks.py:400."""
        out, mask = self.net.member_decode(self.i, content, style, images)
        self.dec.mask_s = mask
        if return_mask:
            return out, mask
        return out


class CouncilGen(_StackedNet):
    def __init__(self, ops, hp, G, input_dim=3):
        g = hp['gen']
        assert not g['do_my_style'], 'do_my_style generators are outside the accelerated path'
        assert g['pad_type'] == 'zero' and g['activ'] == 'relu'
        assert input_dim == 3 and g['num_of_mask_dim_to_add'] == 3, 'mask head kernel is specialised for RGB + 3 masks'
        self.ops, self.hp, self.G = ops, hp, G
        # statistics in the convolution epilogue (cg_conv_fwd_stats) vs a separate pass; see profiles/r01_summary.md
        # '1': every normalised layer; 'auto' (default): the wide layers (>= 128 output channels, K >= 1024), whose main loop is an order of
        # magnitude longer than the epilogue and hides the reduction; '0': never.  Round 1 measured +5.2 ms with '1' (narrow 256x256 layers
        # are epilogue-bound); the CTA-pair kernel that serves the wide layers gained the statistics epilogue in round 2.
        self.fuse_stats = os.environ.get('COUNCIL_FUSE_STATS', 'auto')
        # single-launch normalisation with L2-resident second pass (csrc/norm_coop.cu): correct and tested, but measured equal to the
        # two- / three-kernel forms inside the step (profiles/r02_runB_*), so it is opt-in: COUNCIL_COOP_NORM=1
        coop = os.environ.get('COUNCIL_COOP_NORM', '0')  # 0 | 1 (forward and backward) | bwd (backward only: keeps the statistics epilogue)
        self.coop_norm = coop == '1'
        self.coop_norm_bwd = coop in ('1', 'bwd')
        self.fuse_head = os.environ.get('COUNCIL_FUSE_HEAD', '1') == '1'  # decoder tail of no-grad passes as one kernel (csrc/head_fused.cu)
        self.dim, self.style_dim, self.nd, self.nr, self.mlp_dim = g['dim'], g['style_dim'], g['n_downsample'], g['n_res'], g['mlp_dim']
        dim, nd, nr = self.dim, self.nd, self.nr
        img_lanes = [0, 1, 2]
        # --- content encoder (networks.py:355-366)
        self.enc = [LayerSpec('enc_content.model.0.conv', dim, input_dim, 7, 1, 3, lanes=img_lanes)]
        d = dim
        for i in range(nd):
            self.enc.append(LayerSpec('enc_content.model.%d.conv' % (1 + i), 2 * d, d, 4, 2, 1))
            d *= 2
        self.enc_res = [[LayerSpec('enc_content.model.%d.model.%d.model.%d.conv' % (1 + nd, r, j), d, d, 3, 1, 1)
                         for j in range(2)] for r in range(nr)]
        self.cdim = d
        # --- decoder (networks.py:374-396)
        self.dec_res = [[LayerSpec('dec.model.0.model.%d.model.%d.conv' % (r, j), d, d, 3, 1, 1) for j in range(2)]
                        for r in range(nr)]
        self.dec_up = []
        idx = 1
        for i in range(nd):
            idx += 1
            self.dec_up.append((LayerSpec('dec.model.%d.conv' % idx, d // 2, d, 3, 1, 1),
                                LayerSpec('dec.model.%d.conv' % (idx + 1), d // 2, d // 2, 3, 1, 1)))
            idx += 2
            d //= 2
        self.head = [LayerSpec('dec.model.%d.conv' % idx, d, d, 1, 1, 0),
                     LayerSpec('dec.model.%d.conv' % (idx + 1), d, d, 1, 1, 0),
                     LayerSpec('dec.model.%d.conv' % (idx + 2), 12, d, 1, 1, 0)]
        # AdaIN column offsets in modules() order (networks.py:303-312)
        self.adain_off = {}
        off = 0
        for blk in self.dec_res:
            for s in blk:
                self.adain_off[s.key] = off
                off += 2 * s.cout
        for a, b in self.dec_up:
            for s in (a, b):
                self.adain_off[s.key] = off
                off += 2 * s.cout
        self.n_adain = off
        # --- MLP (networks.py:432-440)
        self.mlp = [LayerSpec('mlp.model.0.fc', self.mlp_dim, self.style_dim, 1, 1, 0, linear=True),
                    LayerSpec('mlp.model.1.fc', self.mlp_dim, self.mlp_dim, 1, 1, 0, linear=True),
                    LayerSpec('mlp.model.2.fc', self.n_adain, self.mlp_dim, 1, 1, 0, linear=True)]
        # --- style encoder (networks.py:337-350): parameters kept for the API / checkpoints, never stepped
        d = dim
        self.sty = [LayerSpec('enc_style.model.0.conv', d, input_dim, 7, 1, 3, lanes=img_lanes)]
        for i in range(2):
            self.sty.append(LayerSpec('enc_style.model.%d.conv' % (1 + i), 2 * d, d, 4, 2, 1))
            d *= 2
        for i in range(2):
            self.sty.append(LayerSpec('enc_style.model.%d.conv' % (3 + i), d, d, 4, 2, 1))
        self.sty_out = LayerSpec('enc_style.model.6', self.style_dim, d, 1, 1, 0)

        live = self.enc + [s for b in self.enc_res for s in b] + [s for b in self.dec_res for s in b] + \
            [s for ab in self.dec_up for s in ab] + self.head + self.mlp
        self.live_specs = live
        self.bank = ParamBank(ops, G, [e for s in live for e in s.entries()])
        # flat-buffer offset where the content encoder's parameters end: the decoder / head / MLP gradients [enc_end:] are
        # complete before the encoder backward starts, so data parallelism all-reduces them while it runs
        self.enc_end = self.bank.table[self.dec_res[0][0].wname][0] if self.dec_res else self.bank.table[self.head[0].wname][0]
        self.frozen = ParamBank(ops, G, [e for s in self.sty + [self.sty_out] for e in s.entries()], trainable=False)
        # biases feeding IN / AdaIN are mathematically dead: their gradient is exactly 0 here
        self.dead_bias = set(s.bname for s in live if s not in self.head and s not in self.mlp)

    # -- state_dict plumbing -------------------------------------------------------------------------
    def _specs(self):
        return self.sty + [self.sty_out] + self.live_specs

    def _banks(self):
        return (self.bank, self.frozen)

    def extra_state(self):
        out = OrderedDict()
        for blk in self.dec_res:
            for s in blk:
                p = s.key[:-5]
                out[p + '.norm.running_mean'] = torch.zeros(s.cout)
                out[p + '.norm.running_var'] = torch.ones(s.cout)
        for ab in self.dec_up:
            for s in ab:
                p = s.key[:-5]
                out[p + '.norm.running_mean'] = torch.zeros(s.cout)
                out[p + '.norm.running_var'] = torch.ones(s.cout)
        return out

    def reference_key_order(self):
        keys = []
        for s in self.sty + [self.sty_out] + self.enc + [s for b in self.enc_res for s in b]:
            keys += [s.wname, s.bname]
        for s in [s for b in self.dec_res for s in b] + [s for ab in self.dec_up for s in ab]:
            p = s.key[:-5]
            keys += [p + '.norm.running_mean', p + '.norm.running_var', s.wname, s.bname]
        for s in self.head + self.mlp:
            keys += [s.wname, s.bname]
        return keys

    def member(self, i):
        return GenMemberView(self, i)

    # -- building blocks -------------------------------------------------------------------------------
    def _w(self, s, sl=None):
        w, b = self.bank.p(s.wname), self.bank.p(s.bname)
        if sl is not None:
            w, b = w[sl:sl + 1], b[sl:sl + 1]
        return w, b

    def _conv_norm(self, x, s, sl, adain, act, res, ups_out, saved, ups_in=False):
        """conv(+bias) -> IN/AdaIN statistics -> normalise(+gamma/beta) -> activation (+residual).
        ups_out: the normalise pass writes its result nearest-upsampled x2 (nn.Upsample, networks.py:385), so
        the following convolution is a plain 3x3 on the materialised tensor (TMA im2col cannot halve indices)."""
        ops = self.ops
        w, b = self._w(s, sl)
        # the bias of a convolution that feeds IN / AdaIN is removed again by the mean subtraction: skip the add
        # ups_in (no-grad passes): the x2 nearest upsample is folded into this convolution (four 2x2 parity classes)
        off = self.adain_off.get(s.key, 0)
        wide = s.cout >= 128 and s.k * s.k * s.cin >= 1024
        if self.fuse_stats == '1' or (self.fuse_stats == 'auto' and wide and not self.coop_norm):
            y, mean, rstd = ops.conv_fwd_stats(x, w, s.stride, s.pad, ups=ups_in)
            z = ops.norm_act_fwd(y, mean, rstd, adain, off, res, act, ups_out)
        elif self.coop_norm:  # statistics + normalise in one launch, second pass over y from L2 (csrc/norm_coop.cu)
            y = ops.conv_fwd(x, w, None, s.stride, s.pad, ups=ups_in)
            z, mean, rstd = ops.norm_fused_fwd(y, adain, off, res, act, ups_out)
        else:
            y = ops.conv_fwd(x, w, None, s.stride, s.pad, ups=ups_in)
            mean, rstd = ops.in_stats(y)
            z = ops.norm_act_fwd(y, mean, rstd, adain, off, res, act, ups_out)
        if saved is not None:
            saved.append((x, y, mean, rstd))
        return z

    def _conv_norm_bwd(self, dz, s, rec, adain, d_adain, act, ups_out, addend, need_dx=True):
        """Backward of _conv_norm: fills the weight gradient, returns d(input).  With ups_out, dz has the
        upsampled shape and the 2x2 fan-in is summed while it is read."""
        ops = self.ops
        x, y, mean, rstd = rec
        off = self.adain_off.get(s.key, 0)
        norm_bwd = ops.norm_fused_bwd if self.coop_norm_bwd else ops.norm_act_bwd
        dy = norm_bwd(dz, y, mean, rstd, adain, off, act, ups_out, d_adain)
        ops.conv_wgrad(x, dy, self.bank.g(s.wname), None, s.stride, s.pad)
        if not need_dx:
            return None
        return ops.conv_dgrad(dy, self.bank.p(s.wname), x.shape, s.stride, s.pad, addend=addend)

    # -- forward ---------------------------------------------------------------------------------------
    def encode(self, x_img, saved=None, sl=None):
        """x_img [1,B,H,W,4] (shared by all members) -> content [G,B,H/2^nd,W/2^nd,C]."""
        x = x_img
        for s in self.enc:
            x = self._conv_norm(x, s, sl, None, ACT_RELU, None, False, saved)
        for blk in self.enc_res:
            res = x
            x = self._conv_norm(x, blk[0], sl, None, ACT_RELU, None, False, saved)
            x = self._conv_norm(x, blk[1], sl, None, ACT_NONE, res, False, saved)
        return x

    def _mlp(self, style, saved=None, sl=None):
        """style [1,B,1,1,style_dim] -> AdaIN parameters [G,B,n_adain]."""
        ops = self.ops
        h = style
        acts = []
        for li, s in enumerate(self.mlp):
            w, b = self._w(s, sl)
            h_in = h
            h = ops.conv_fwd(h, w, b, 1, 0, act=ACT_RELU if li < 2 else ACT_NONE)
            acts.append((h_in, h))
        if saved is not None:
            saved.append(acts)
        return h.view(h.shape[0], h.shape[1], self.n_adain)

    def decode(self, content, style, x_img, saved=None, sl=None):
        """content [G,B,h,w,C], style [1,B,1,1,S], x_img [1,B,H,W,4] -> (x_fake, mask) [G,B,H,W,4]."""
        ops = self.ops
        adain = self._mlp(style, saved, sl)
        x = content
        nres, nup = len(self.dec_res), len(self.dec_up)
        # passes that keep activations for backward materialise the upsampled tensor (the weight / data gradient
        # kernels read it); no-grad passes fold the upsample into the consumer convolution instead
        fold = saved is None
        for r, blk in enumerate(self.dec_res):
            res = x
            x = self._conv_norm(x, blk[0], sl, adain, ACT_RELU, None, False, saved)
            x = self._conv_norm(x, blk[1], sl, adain, ACT_NONE, res, r == nres - 1 and nup > 0 and not fold, saved)
        for u, (a, b) in enumerate(self.dec_up):
            x = self._conv_norm(x, a, sl, adain, ACT_RELU, None, False, saved, ups_in=fold)
            if fold and u + 1 == nup and self.fuse_head and b.cout == 64 and ops.head_fused_supported((1, 1, x.shape[2], x.shape[3], 64)):
                # no-grad pass: the rest of the decoder (AdaIN + ReLU of this block, the three 1x1 head layers, mask compositing)
                # is one kernel; the 64-channel full-resolution map is read once instead of making four HBM round trips
                w, _ = self._w(b, sl)
                y = ops.conv_fwd(x, w, None, b.stride, b.pad)
                mean, rstd = ops.in_stats(y)
                (w1, b1), (w2, b2), (w3, b3) = (self._w(s, sl) for s in self.head)
                return ops.head_fused(y, mean, rstd, adain, self.adain_off[b.key], w1, b1, w2, b2, w3, b3, x_img)
            x = self._conv_norm(x, b, sl, adain, ACT_RELU, None, u + 1 < nup and not fold, saved)
        acts = [x]
        for li, s in enumerate(self.head):
            w, b = self._w(s, sl)
            x = ops.conv_fwd(x, w, b, 1, 0, act=ACT_RELU if li < 2 else ACT_TANH)
            acts.append(x)
        x_fake, mask = ops.mask_head_fwd(x, x_img)
        if saved is not None:
            saved.append((adain, acts, x_img))
        return x_fake, mask

    # -- backward (gen_update only) -------------------------------------------------------------------
    def backward(self, d_xfake, d_mask, enc_saved, dec_saved, on_decoder_done=None):
        """Fills ``self.bank.grad`` for every live parameter given d(loss)/d(x_fake), d(loss)/d(mask).
        on_decoder_done: called when grad[enc_end:] (decoder, head, MLP) is final, before the encoder backward."""
        ops, bank = self.ops, self.bank
        dec_saved = list(dec_saved)
        adain, acts, x_img = dec_saved.pop()
        mlp_acts = dec_saved.pop(0)
        d_adain = ops.empty(*adain.shape)  # every column is written by exactly one AdaIN layer's backward
        # head: 1x1 convs with fused activations
        d = ops.mask_head_bwd(acts[3], x_img, d_xfake, d_mask)  # grad w.r.t. pre-tanh output of head[2]
        for li in (2, 1, 0):
            s = self.head[li]
            ops.conv_wgrad(acts[li], d, bank.g(s.wname), bank.g(s.bname), 1, 0)
            d = ops.conv_dgrad(d, bank.p(s.wname), acts[li].shape, 1, 0,
                               mask_src=acts[li] if li > 0 else None, mask_slope=0.0)
        # upsampling blocks, last to first
        recs = dec_saved  # one record per conv in forward order: dec_res (2*nr) then dec_up (2*nd)
        k = len(recs) - 1
        nup = len(self.dec_up)
        for u in range(nup - 1, -1, -1):
            a, b = self.dec_up[u]
            d = self._conv_norm_bwd(d, b, recs[k], adain, d_adain, ACT_RELU, u + 1 < nup, None)
            d = self._conv_norm_bwd(d, a, recs[k - 1], adain, d_adain, ACT_RELU, False, None)
            k -= 2
        if nup > 0:
            d = ops.upsample2x_bwd(d)  # the last residual block's output was written upsampled
        for blk in reversed(self.dec_res):
            d_out = d
            d = self._conv_norm_bwd(d_out, blk[1], recs[k], adain, d_adain, ACT_NONE, False, None)
            d = self._conv_norm_bwd(d, blk[0], recs[k - 1], adain, d_adain, ACT_RELU, False, d_out)
            k -= 2
        assert k == -1
        # MLP (gradients reach it through every AdaIN gamma/beta)
        dm = d_adain.view(d_adain.shape[0], d_adain.shape[1], 1, 1, self.n_adain)
        for li in (2, 1, 0):
            s = self.mlp[li]
            h_in, _ = mlp_acts[li]
            ops.conv_wgrad(h_in, dm, bank.g(s.wname), bank.g(s.bname), 1, 0)
            if li > 0:
                dm = ops.conv_dgrad(dm, bank.p(s.wname), h_in.shape, 1, 0, mask_src=h_in, mask_slope=0.0)
        if on_decoder_done is not None:
            on_decoder_done()
        # content encoder
        recs = list(enc_saved)
        k = len(recs) - 1
        for blk in reversed(self.enc_res):
            d_out = d
            d = self._conv_norm_bwd(d_out, blk[1], recs[k], None, None, ACT_NONE, False, None)
            d = self._conv_norm_bwd(d, blk[0], recs[k - 1], None, None, ACT_RELU, False, d_out)
            k -= 2
        for li in range(len(self.enc) - 1, -1, -1):
            d = self._conv_norm_bwd(d, self.enc[li], recs[k], None, None, ACT_RELU, False, None, need_dx=li > 0)
            k -= 1
        assert k == -1

    # -- single-member API (reference's gen.encode / gen.decode on NCHW tensors) -----------------------
    def member_encode(self, i, images):
        self._sync()
        ops = self.ops
        x = ops.nchw_to_nhwc(images.to(ops.device, ops.dtype).contiguous(), IMG_C)[None]
        c = self.encode(x, None, sl=i)
        content = ops.nhwc_to_nchw(c[0], self.cdim)
        return content, self.member_style_encode(i, x)

    def style_encode(self, x, sl=None):
        """StyleEncoder networks.py:337-353 (norm none, relu), all members (or member sl) at once:
        x [1,B,H,W,4] -> style codes [G,B,1,1,style_dim].  Not on the training path (its result is discarded there)."""
        ops, fz = self.ops, self.frozen

        def wb(s):
            w, b = fz.p(s.wname), fz.p(s.bname)
            return (w, b) if sl is None else (w[sl:sl + 1], b[sl:sl + 1])
        h = x
        for s in self.sty:
            h = ops.conv_fwd(h, *wb(s), s.stride, s.pad, act=ACT_RELU)
        # global average pool (tiny, plumbing) then the 1x1 conv as a linear layer
        pooled = h.mean(dim=(2, 3), keepdim=True).contiguous()
        return ops.conv_fwd(pooled, *wb(self.sty_out), 1, 0)

    def member_style_encode(self, i, x):
        """-> [B, style_dim, 1, 1] of member i (AdaINGen.encode's second result, networks.py:278-283)."""
        return self.style_encode(x, sl=i)[0].permute(0, 3, 1, 2).contiguous()

    def member_decode(self, i, content, style, images):
        self._sync()
        ops = self.ops
        x = ops.nchw_to_nhwc(images.to(ops.device, ops.dtype).contiguous(), IMG_C)[None]
        c = ops.nchw_to_nhwc(content.to(ops.device, ops.dtype).contiguous(), self.cdim)[None]
        st = style.to(ops.device, ops.dtype).reshape(1, style.shape[0], 1, 1, self.style_dim).contiguous()
        x_fake, mask = self.decode(c, st, x, None, sl=i)
        return ops.nhwc_to_nchw(x_fake[0], 3), ops.nhwc_to_nchw(mask[0], 3)


# ------------------------------------------------------------------------------------------------------
# discriminators
# ------------------------------------------------------------------------------------------------------
class CouncilDis(_StackedNet):
    """MsImageDis (council=False) / MsImageDisCouncil (council=True), all members stacked."""

    def __init__(self, ops, hp, G, input_dim=3, council=False):
        dp = hp['dis']
        assert dp['gan_type'] == 'lsgan', 'only the LSGAN objective is on the accelerated path'
        assert dp['norm'] == 'none' and dp['activ'] == 'lrelu' and dp['pad_type'] == 'zero'
        assert input_dim == 3
        self.ops, self.hp, self.G, self.council = ops, hp, G, council
        self.num_scales, self.n_layer = dp['num_scales'], dp['n_layer']
        self.scales = []
        for sc in range(self.num_scales):
            d = dp['dim']
            if council:  # networks.py:138: 3x3 stride 1 on cat(x, x_input)
                layers = [LayerSpec('cnns.%d.0.conv' % sc, d, 2 * input_dim, 3, 1, 1, lanes=[0, 1, 2, 4, 5, 6])]
            else:        # networks.py:40: 4x4 stride 2
                layers = [LayerSpec('cnns.%d.0.conv' % sc, d, input_dim, 4, 2, 1, lanes=[0, 1, 2])]
            for i in range(self.n_layer - 1):
                layers.append(LayerSpec('cnns.%d.%d.conv' % (sc, i + 1), 2 * d, d, 4, 2, 1))
                d *= 2
            n = self.n_layer
            tail = []
            if council:  # networks.py:142-143: two 1x1 convs, no activation between them
                tail.append(LayerSpec('cnns.%d.%d' % (sc, n), d, d, 1, 1, 0))
                n += 1
            tail.append(LayerSpec('cnns.%d.%d' % (sc, n), 1, d, 1, 1, 0))
            self.scales.append((layers, tail))
        self.specs = [s for layers, tail in self.scales for s in layers + tail]
        self.bank = ParamBank(ops, G, [e for s in self.specs for e in s.entries()])

    def _specs(self):
        return self.specs

    def _banks(self):
        return (self.bank,)

    def reference_key_order(self):
        return [k for s in self.specs for k in (s.wname, s.bname)]

    def forward(self, x, saved=None):
        """x [G,Bt,H,W,4|8] -> list over scales of patch outputs [G,Bt,h,w,1]."""
        ops, bank = self.ops, self.bank
        outs = []
        for sc, (layers, tail) in enumerate(self.scales):
            h = x
            acts = [h]
            for s in layers:
                h = ops.conv_fwd(h, bank.p(s.wname), bank.p(s.bname), s.stride, s.pad, act=ACT_LRELU, slope=0.2)
                acts.append(h)
            for s in tail:
                h = ops.conv_fwd(h, bank.p(s.wname), bank.p(s.bname), 1, 0)
                acts.append(h)
            outs.append(h)
            if saved is not None:
                saved.append(acts)
            if sc + 1 < self.num_scales:
                x = ops.avgpool_fwd(x)
        return outs

    def backward(self, d_outs, saved, want_wgrad, want_dx):
        """d_outs: per-scale d(loss)/d(out).  want_wgrad: fill bank.grad (dis/dis_council update).
        want_dx: return d(loss)/d(x) [G,Bt,H,W,lanes of x] accumulated over scales (gen_update)."""
        ops, bank = self.ops, self.bank
        dx_scales = []
        for sc, (layers, tail) in enumerate(self.scales):
            acts = saved[sc]
            d = d_outs[sc]
            allspecs = layers + tail
            nl = len(layers)
            for li in range(len(allspecs) - 1, -1, -1):
                s = allspecs[li]
                a_in = acts[li]
                if want_wgrad:
                    ops.conv_wgrad(a_in, d, bank.g(s.wname), bank.g(s.bname), s.stride, s.pad)
                if li == 0 and not want_dx:
                    break
                # a_in is the lrelu OUTPUT of layer li-1 when that layer is one of `layers`
                masked = 1 <= li <= nl
                d = ops.conv_dgrad(d, bank.p(s.wname), a_in.shape, s.stride, s.pad,
                                   mask_src=a_in if masked else None, mask_slope=0.2)
            if want_dx:
                dx_scales.append(d)
        if not want_dx:
            return None
        # fold the image pyramid back: x_{s+1} = avgpool(x_s)
        dx = dx_scales[-1]
        for sc in range(self.num_scales - 2, -1, -1):
            up = dx_scales[sc]
            ops.avgpool_bwd(dx, up, min(4, up.shape[-1]), True)
            if up.shape[-1] == 8:  # second image of the pair (x_input) is data: its lanes 4..7 are ignored
                pass
            dx = up
        return dx
