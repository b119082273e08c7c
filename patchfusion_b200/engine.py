"""Device-side execution of the PatchFusion hot path on the libpf_b200 kernels.

`Engine` owns the packed weights and a pool of persistent device buffers (static addresses: TMA tensor maps are
cached per buffer and the whole per-micro-batch sequence is CUDA-graph capturable).  It sequences the kernels for

    coarse/fine branch   reference `estimator/models/patchfusion.py:189-225` -> zoedepth_v1.py:125-233 ->
                         depth_anything.py:262-278 -> dpt.py:97-157 -> dinov2 vision_transformer.py:297-321
    G2L maps             `estimator/models/blocks/swin_layers.py:410-432` (once per image; the reference recomputes
                         them per micro-batch with identical input, guided_fusion_model.py:201)
    fusion               `estimator/models/patchfusion.py:259-340`, guided_fusion_model.py:163-207
    tiling / stitch      `estimator/models/baseline_pretrain.py:143-331`, estimator/models/utils.py:21-47

Nothing here touches torch for arithmetic on the path: torch allocates memory, owns the stream, and runs the
one-off load-time transforms (pos-embed bicubic resample, BatchNorm folding constants).
"""
import ctypes as ct
import math

import torch
import torch.nn.functional as F

from . import ops
from .ops import ACT_GELU, ACT_NONE, ACT_RELU, ACT_SOFTPLUS, call, pad_to, stream_ptr
from .params import WINDOW, branch_hparams, guided_fusion_hparams, _get

BF16, F32 = torch.bfloat16, torch.float32


class Map:
    """NHWC bf16 activation: tensor [B,H,W,ld] + logical channel count."""
    __slots__ = ('t', 'C')

    def __init__(self, t, C):
        self.t, self.C = t, C

    @property
    def B(self):
        return self.t.shape[0]

    @property
    def hw(self):
        return self.t.shape[1], self.t.shape[2]

    def rows(self):
        return self.t.view(-1, self.t.shape[-1])


class Engine:
    def __init__(self, config, state_dict, device):
        self.cfg = config
        self.dev = torch.device(device)
        self.P = tuple(_get(config, 'patch_process_shape'))
        self.hp = {'coarse': branch_hparams(_get(config, 'coarse_branch')),
                   'fine': branch_hparams(_get(config, 'fine_branch'))}
        self.bcfg = {'coarse': _get(config, 'coarse_branch'), 'fine': _get(config, 'fine_branch')}
        self.gf = guided_fusion_hparams(_get(config, 'guided_fusion'), self.P)
        self.bufs = {}
        self.sd = {k: v.to(self.dev) for k, v in state_dict.items()}
        H, W = self.P
        assert H % 14 == 0 and W % 14 == 0
        self.gh, self.gw = H // 14, W // 14
        self.W = {}
        self._pack_all()
        self.sd = None   # fp32 originals are no longer needed on the device

    # ------------------------------------------------------------------ buffers
    def buf(self, key, shape, dtype=BF16):
        k = (key, tuple(shape), dtype)
        t = self.bufs.get(k)
        if t is None:
            t = torch.zeros(shape, dtype=dtype, device=self.dev)
            self.bufs[k] = t
        return t

    def map(self, key, B, h, w, C):
        return Map(self.buf(key, (B, h, w, pad_to(C, 8))), C)

    # ------------------------------------------------------------------ weight packing
    def _w(self, k):
        return self.sd[k]

    def _conv(self, name, src_c=None, bias=True):
        b = self._w(name + '.bias') if bias and (name + '.bias') in self.sd else None
        return ops.pack_weight(self._w(name + '.weight'), b, src_c=src_c)

    def _conv_bn(self, conv, bn):
        g, b = self._w(bn + '.weight'), self._w(bn + '.bias')
        m, v = self._w(bn + '.running_mean'), self._w(bn + '.running_var')
        scale = g / torch.sqrt(v + 1e-5)
        return ops.pack_weight(self._w(conv + '.weight'), None, scale=scale, shift=b - m * scale)

    def _f32(self, k):
        return self._w(k).float().contiguous()

    def _pack_head(self, pre, C, hp, drop_rel):
        Wd = {}
        for n in ['seed_bin_regressor', 'seed_projector'] + ['projectors.%d' % i for i in range(4)] + \
                 ['attractors.%d' % i for i in range(4)]:
            Wd[n + '.0'] = self._conv(pre + n + '._net.0')
            Wd[n + '.2'] = self._conv(pre + n + '._net.2')
            w2 = self._w(pre + n + '._net.2.weight')
            if w2.shape[0] <= 16:       # narrow second layer: fused into the first layer's epilogue (fp32 weights)
                Wd[n + '.tail'] = (w2.reshape(w2.shape[0], -1).float().contiguous(),
                                   self._f32(pre + n + '._net.2.bias'))
        w0 = self._w(pre + 'conditional_log_binomial.mlp.0.weight')
        b0 = self._w(pre + 'conditional_log_binomial.mlp.0.bias')
        E = hp['bin_embedding_dim']
        if drop_rel:      # rel_cond is identically zero in the fusion head (patchfusion.py:300,326-328)
            w0 = torch.cat([w0[:, :32], w0[:, 33:]], 1).contiguous()
            Wd['clb.0'] = ops.pack_weight(w0, b0, src_c=[32, E])
        else:
            Wd['clb.0'] = ops.pack_weight(w0, b0, src_c=[32, 1, E])
        Wd['clb.tail'] = (self._w(pre + 'conditional_log_binomial.mlp.2.weight').reshape(4, -1).float().contiguous(),
                          self._f32(pre + 'conditional_log_binomial.mlp.2.bias'))
        return Wd

    def _pack_branch(self, which):
        pre = which + '_branch.'
        hp = self.hp[which]
        D, C, oc = hp['dim'], hp['features'], hp['out_channels']
        vit = pre + 'core.core.pretrained.'
        Wd = {}
        Wd['patch'] = ops.pack_weight(self._w(vit + 'patch_embed.proj.weight').reshape(D, 588),
                                      self._w(vit + 'patch_embed.proj.bias'))
        # pos-embed resample is input-shape-only: done once here with the reference's exact call
        # (vision_transformer.py:189-210: bicubic, scale_factor=((gh+0.1)/37, (gw+0.1)/37), no antialias)
        pe = self._w(vit + 'pos_embed').float()
        n = pe.shape[1] - 1
        s = int(math.sqrt(n))
        if (self.gh, self.gw) != (s, s):
            grid = pe[:, 1:].reshape(1, s, s, D).permute(0, 3, 1, 2)
            grid = F.interpolate(grid, scale_factor=((self.gh + 0.1) / s, (self.gw + 0.1) / s), mode='bicubic',
                                 antialias=False)
            assert grid.shape[-2:] == (self.gh, self.gw)
            pe = torch.cat([pe[:, :1], grid.permute(0, 2, 3, 1).reshape(1, self.gh * self.gw, D)], 1)
        Wd['pos'] = pe[0].contiguous()
        Wd['cls'] = self._f32(vit + 'cls_token').reshape(D)
        for i in range(hp['depth']):
            p = vit + 'blocks.%d.' % i
            Wd['b%d' % i] = dict(
                n1w=self._f32(p + 'norm1.weight'), n1b=self._f32(p + 'norm1.bias'),
                n2w=self._f32(p + 'norm2.weight'), n2b=self._f32(p + 'norm2.bias'),
                qkv=ops.pack_weight(self._w(p + 'attn.qkv.weight'), self._w(p + 'attn.qkv.bias')),
                proj=ops.pack_weight(self._w(p + 'attn.proj.weight'), self._w(p + 'attn.proj.bias')),
                fc1=ops.pack_weight(self._w(p + 'mlp.fc1.weight'), self._w(p + 'mlp.fc1.bias')),
                fc2=ops.pack_weight(self._w(p + 'mlp.fc2.weight'), self._w(p + 'mlp.fc2.bias')),
                ls1=self._f32(p + 'ls1.gamma'), ls2=self._f32(p + 'ls2.gamma'))
        Wd['nw'], Wd['nb'] = self._f32(vit + 'norm.weight'), self._f32(vit + 'norm.bias')
        dh = pre + 'core.core.depth_head.'
        for i in range(4):
            Wd['proj%d' % i] = self._conv(dh + 'projects.%d' % i)
        Wd['rs0'] = ops.pack_weight_convT(self._w(dh + 'resize_layers.0.weight'), self._w(dh + 'resize_layers.0.bias'), 4)
        Wd['rs1'] = ops.pack_weight_convT(self._w(dh + 'resize_layers.1.weight'), self._w(dh + 'resize_layers.1.bias'), 2)
        rs3 = self._conv(dh + 'resize_layers.3')
        assert oc[3] % 64 == 0
        rs3.taps, rs3.src_c = 1, [9 * oc[3]]          # consumed as a plain GEMM over pf_im2col_3x3_s2 rows
        Wd['rs3'] = rs3
        for i in range(4):
            Wd['rn%d' % i] = self._conv(dh + 'scratch.layer%d_rn' % (i + 1), bias=False)
        for i in range(1, 5):
            r = dh + 'scratch.refinenet%d.' % i
            Wd['ff%d.out' % i] = self._conv(r + 'out_conv')
            for u in (1, 2):
                Wd['ff%d.u%d.c1' % (i, u)] = self._conv(r + 'resConfUnit%d.conv1' % u)
                Wd['ff%d.u%d.c2' % (i, u)] = self._conv(r + 'resConfUnit%d.conv2' % u)
        Wd['oc1'] = self._conv(dh + 'scratch.output_conv1')
        Wd['oc2.0'] = self._conv(dh + 'scratch.output_conv2.0')
        Wd['oc2.tail'] = (self._w(dh + 'scratch.output_conv2.2.weight').reshape(1, -1).float().contiguous(),
                          self._f32(dh + 'scratch.output_conv2.2.bias'))
        Wd['conv2'] = self._conv(pre + 'conv2')
        Wd['head'] = self._pack_head(pre, C, hp, drop_rel=False)
        return Wd

    def _pack_fusion(self):
        hp = self.hp['fine']
        C = hp['features']
        Wd = {}
        for i in range(5):      # level 5's fused map is dead in the U-Net (guided_fusion_model.py:198)
            Wd['fc%d' % i] = self._conv('fusion_conv_list.%d' % i, src_c=[C, C])
        g = 'guided_fusion.'
        ic = self.gf['in_channels']
        Wd['inc.0'] = self._conv_bn(g + 'inc.double_conv.0', g + 'inc.double_conv.1')
        Wd['inc.1'] = self._conv_bn(g + 'inc.double_conv.3', g + 'inc.double_conv.4')
        for i in range(5):
            p = g + 'down_conv_list.%d.maxpool_conv.1.double_conv.' % i
            Wd['down%d.0' % i] = self._conv_bn(p + '0', p + '1')
            Wd['down%d.1' % i] = self._conv_bn(p + '3', p + '4')
        inv = ic[::-1]
        for i in range(1, 6):
            p = g + 'up_conv_list.%d.conv.double_conv.' % (i - 1)
            Wd['up%d.0' % i] = self._conv(p + '0', src_c=[inv[i], inv[i - 1], inv[i - 1]])
            Wd['up%d.1' % i] = self._conv(p + '2')
        depth, heads = self.gf['depth'][::-1], self.gf['num_heads'][::-1]
        for i in range(6):
            c, p = inv[i], g + 'g2l_list.%d.' % i
            L = dict(ape=self._f32(p + 'absolute_pos_embed').reshape(-1, c), C=c, heads=heads[i],
                     nw=self._f32(p + 'g2l_layer_norm.weight'), nb=self._f32(p + 'g2l_layer_norm.bias'),
                     ones=torch.ones(c, dtype=F32, device=self.dev), blocks=[])
            for b in range(depth[i]):
                q = p + 'g2l_layer.blocks.%d.' % b
                L['blocks'].append(dict(
                    n1w=self._f32(q + 'norm1.weight'), n1b=self._f32(q + 'norm1.bias'),
                    n2w=self._f32(q + 'norm2.weight'), n2b=self._f32(q + 'norm2.bias'),
                    table=self._f32(q + 'attn.relative_position_bias_table'),
                    qkv=ops.pack_weight(self._w(q + 'attn.qkv.weight'), self._w(q + 'attn.qkv.bias')),
                    proj=ops.pack_weight(self._w(q + 'attn.proj.weight'), self._w(q + 'attn.proj.bias')),
                    fc1=ops.pack_weight(self._w(q + 'mlp.fc1.weight'), self._w(q + 'mlp.fc1.bias')),
                    fc2=ops.pack_weight(self._w(q + 'mlp.fc2.weight'), self._w(q + 'mlp.fc2.bias'))))
            Wd['g2l%d' % i] = L
            p = g + 'convs.%d.double_conv.' % i
            Wd['cv%d.0' % i] = self._conv(p + '0', src_c=[c, c])
            Wd['cv%d.1' % i] = self._conv(p + '2')
        Wd['head'] = self._pack_head('', C, self.hp['coarse'], drop_rel=True)
        return Wd

    def _pack_all(self):
        self.W['coarse'] = self._pack_branch('coarse')
        self.W['fine'] = self._pack_branch('fine')
        self.W['fusion'] = self._pack_fusion()
        torch.cuda.synchronize(self.dev)

    # ------------------------------------------------------------------ small helpers
    def conv(self, key, pw, srcs, N=None, act=ACT_NONE, res1=None, res2=None, relu_copy=False, out=None,
             out_dtype=BF16, tail=None, tail_out=None, skip_main=False):
        """3x3 / 1x1 conv over NHWC maps `srcs` (list of Map) -> Map (and optionally its ReLU copy)."""
        B, (h, w) = srcs[0].B, srcs[0].hw
        N = pw.N if N is None else N
        if out is None:
            out = Map(self.buf(key, (B, h, w, pad_to(N, 8)), out_dtype), N)
        o2 = Map(self.buf(key + '.relu', (B, h, w, pad_to(N, 8))), N) if relu_copy else None
        ops.gemm(pw, [s.t for s in srcs], out.t, image=(B, h, w), act=act,
                 res1=res1.t if res1 is not None else None, res2=res2.t if res2 is not None else None,
                 out2=o2.t if o2 is not None else None, src_c=[s.C for s in srcs], tail=tail, tail_out=tail_out,
                 skip_main=skip_main)
        return (out, o2) if relu_copy else out

    def resize(self, key, x, size, out=None, out_col0=0):
        if out is None and x.hw == tuple(size):
            return x
        if out is None:
            out = self.map(key, x.B, size[0], size[1], x.C)
        ops.resize_bilinear(x.t, x.C, size[0], size[1], out.t, out_col0=out_col0)
        return out

    # ------------------------------------------------------------------ one branch
    def branch(self, which, images, taps=None, slot=''):
        """images: planar fp32 [B,3,H,W] in [0,1] (un-normalised).  Returns (depth fp32 [B,H,W], feats[6] Maps
        low->high: x_d0, r4, r3, r2, r1, out_conv).  `slot` selects an independent buffer set (double buffering of
        the fine branch against the fusion stage of the previous micro-batch)."""
        Wd, hp = self.W[which], self.hp[which]
        B = images.shape[0]
        H, Wd_ = self.P
        gh, gw, D, C, oc = self.gh, self.gw, hp['dim'], hp['features'], hp['out_channels']
        npatch, seq = gh * gw, gh * gw + 1
        seq_pad = pad_to(seq, 8)
        k = which + slot + '.'
        st = stream_ptr()
        # ---- tokens
        a0 = self.buf(k + 'im2col', (B * npatch, 592))
        call('pf_patch_im2col', images, B, H, Wd_, a0, 592, st)
        patch = self.buf(k + 'patch', (B * npatch, D), F32)
        ops.gemm(Wd['patch'], [a0], patch, src_c=[592])
        x = self.buf(k + 'x', (B * seq, D), F32)
        call('pf_assemble_tokens', patch, Wd['cls'], Wd['pos'], B, npatch, D, x, st)
        if taps is not None:
            taps['tokens'] = x.clone()
        hbuf = self.buf(k + 'h', (B * seq, D))
        qk = self.buf(k + 'qk', (B * seq, 2 * D))
        vt = self.buf(k + 'vt', (B * D, seq_pad))
        att = self.buf(k + 'att', (B * seq, D))
        hid = self.buf(k + 'hid', (B * seq, 4 * D))
        feats = []
        for i in range(hp['depth']):
            bw = Wd['b%d' % i]
            ops.layernorm(x, bw['n1w'], bw['n1b'], 1e-6, hbuf)
            ops.gemm(bw['qkv'], [hbuf], qk, vt=vt, vt_col0=2 * D, vt_seq=seq, vt_seq_pad=seq_pad)
            ops.attention(qk, vt, B, seq, seq_pad, hp['heads'], 64 ** -0.5, att)
            ops.gemm(bw['proj'], [att], x, gamma=bw['ls1'])
            ops.layernorm(x, bw['n2w'], bw['n2b'], 1e-6, hbuf)
            ops.gemm(bw['fc1'], [hbuf], hid, act=ACT_GELU)
            ops.gemm(bw['fc2'], [hid], x, gamma=bw['ls2'])
            if taps is not None:
                taps['block%d' % i] = x.clone()
            if i >= hp['depth'] - 4:
                f = self.buf(k + 'vitout%d' % len(feats), (B, gh, gw, D))
                for b in range(B):      # final LayerNorm of the patch tokens only (cls row skipped)
                    call('pf_layernorm', x[b * seq + 1:], D, Wd['nw'], Wd['nb'], ct.c_float(1e-6), npatch, D,
                         f[b], D, st)
                feats.append(Map(f, D))
        return self.dpt_and_head(which, feats, taps, slot)

    def dpt_and_head(self, which, feats, taps=None, slot=''):
        Wd, hp = self.W[which], self.hp[which]
        B = feats[0].B
        gh, gw, C, oc = self.gh, self.gw, hp['features'], hp['out_channels']
        H, Wimg = self.P
        k = which + slot + '.dpt.'
        st = stream_ptr()
        lay = []
        for i in range(4):
            p = self.map(k + 'proj%d' % i, B, gh, gw, oc[i])
            ops.gemm(Wd['proj%d' % i], [feats[i].rows()], p.rows())
            if i == 0:
                o = self.map(k + 'rs0', B, gh * 4, gw * 4, oc[0])
                ops.gemm_convT(Wd['rs0'], p.rows(), (B, gh, gw), o.t)
            elif i == 1:
                o = self.map(k + 'rs1', B, gh * 2, gw * 2, oc[1])
                ops.gemm_convT(Wd['rs1'], p.rows(), (B, gh, gw), o.t)
            elif i == 2:
                o = p
            else:
                oh, ow = (gh - 1) // 2 + 1, (gw - 1) // 2 + 1
                col = self.buf(k + 'rs3col', (B * oh * ow, 9 * oc[3]))
                call('pf_im2col_3x3_s2', p.t, B, gh, gw, oc[3], p.t.shape[-1], col, st)
                o = self.map(k + 'rs3', B, oh, ow, oc[3])
                ops.gemm(Wd['rs3'], [col], o.rows())
            lay.append(o)
        rn, rn_relu = [], []
        for i in range(4):
            a, b = self.conv(k + 'rn%d' % i, Wd['rn%d' % i], [lay[i]], relu_copy=True)
            rn.append(a)
            rn_relu.append(b)

        def rcu(tag, wi, u, x, x_relu, extra=None, relu_copy=False):
            t = self.conv(k + tag + '.t', Wd['ff%d.u%d.c1' % (wi, u)], [x_relu], act=ACT_RELU)
            return self.conv(k + tag + '.y', Wd['ff%d.u%d.c2' % (wi, u)], [t], res1=x, res2=extra, relu_copy=relu_copy)

        def ffb(wi, path, skip, skip_relu, size):
            if path is None:
                s, s_relu = skip, skip_relu
            else:
                s, s_relu = rcu('ff%d.u1' % wi, wi, 1, skip, skip_relu, extra=path, relu_copy=True)
            y = rcu('ff%d.u2' % wi, wi, 2, s, s_relu)
            # out_conv (1x1) commutes with the bilinear upsample: run it at the low resolution
            y = self.conv(k + 'ff%d.out' % wi, Wd['ff%d.out' % wi], [y])
            return self.resize(k + 'ff%d.up' % wi, y, size)

        p4 = ffb(4, None, rn[3], rn_relu[3], rn[2].hw)
        p3 = ffb(3, p4, rn[2], rn_relu[2], rn[1].hw)
        p2 = ffb(2, p3, rn[1], rn_relu[1], rn[0].hw)
        p1 = ffb(1, p2, rn[0], rn_relu[0], (rn[0].hw[0] * 2, rn[0].hw[1] * 2))
        o = self.conv(k + 'oc1', Wd['oc1'], [p1])
        o = self.resize(k + 'oc1up', o, (H, Wimg))
        rel = Map(self.buf(k + 'rel', (B, H, Wimg, 8), F32), 1)
        # output_conv2: 3x3 C/2->32 + ReLU (the hooked `out_conv` tap) with the 1x1 32->1 + ReLU fused in its epilogue
        out_conv = self.conv(k + 'oc2', Wd['oc2.0'], [o], act=ACT_RELU,
                             tail=Wd['oc2.tail'] + (ACT_RELU,), tail_out=rel.t)
        x_d0 = self.conv(k + 'xd0', Wd['conv2'], [rn[3]])
        blocks = [p4, p3, p2, p1]
        if taps is not None:
            taps['rel'] = rel.t[..., 0].clone()
        depth = self.metric_head(which + slot + '.head.', Wd['head'], hp, self.bcfg[which], x_d0, blocks, out_conv, rel, taps)
        return depth, [x_d0] + blocks + [out_conv]

    def metric_head(self, k, Wh, hp, bcfg, x, x_blocks, last, rel, taps=None, depth_out=None):
        """zoedepth_v1.py:173-219 / patchfusion.py:297-339.  rel: Map fp32 [B,H,W,8] (col 0) or None."""
        st = stream_ptr()
        B = x.B
        nb, E = hp['n_bins'], hp['bin_embedding_dim']

        def mlp(tag, name, src, act2=ACT_NONE, f32_out=False, n_out=None):
            pw = Wh[name + '.2']
            if f32_out and (name + '.tail') in Wh:
                h, w = src.hw
                o = Map(self.buf(k + tag + '.o', (B, h, w, pad_to(pad_to(pw.N, 8), 32)), F32), pw.N)
                self.conv(k + tag + '.t', Wh[name + '.0'], [src], act=ACT_RELU, tail=Wh[name + '.tail'] + (act2,),
                          tail_out=o.t, skip_main=True)
                return o
            t = self.conv(k + tag + '.t', Wh[name + '.0'], [src], act=ACT_RELU)
            if f32_out:
                h, w = src.hw
                o = Map(self.buf(k + tag + '.o', (B, h, w, pad_to(pad_to(pw.N, 8), 32)), F32), pw.N)
                ops.gemm(pw, [t.t], o.t, image=(B, h, w), act=act2, src_c=[t.C])
                return o
            return self.conv(k + tag + '.o', pw, [t], act=act2)

        b_prev = mlp('seed', 'seed_bin_regressor', x, ACT_SOFTPLUS, f32_out=True)      # fp32 [B,h,w,64]
        prev_emb = mlp('seedproj', 'seed_projector', x)
        ph, pw_ = x.hw
        b_t = b_prev.t
        for i, xb in enumerate(x_blocks):
            h, w = xb.hw
            emb = mlp('proj%d' % i, 'projectors.%d' % i, xb)
            s = self.map(k + 'sum%d' % i, B, h, w, E)
            call('pf_add_upsampled', emb.t, B, h, w, E, prev_emb.t, prev_emb.hw[0], prev_emb.hw[1], s.t, st)
            A = mlp('att%d' % i, 'attractors.%d' % i, s, ACT_SOFTPLUS, f32_out=True)
            b_new = self.buf(k + 'b%d' % i, (B, h, w, nb), F32)
            call('pf_attractor', A.t, A.t.shape[-1], hp['n_attractors'][i], b_t, ph, pw_, B, h, w, nb,
                 (1 if hp['attractor_kind'] == 'mean' else 0) | (2 if hp['attractor_type'] == 'exp' else 0), b_new, st)
            b_t, ph, pw_, prev_emb = b_new, h, w, emb
            if taps is not None:
                taps['b%d' % i] = b_new.clone()
        H, Wimg = last.hw
        emb_up = self.resize(k + 'embup', prev_emb, (H, Wimg))
        if rel is not None:
            relb = self.map(k + 'relb', B, H, Wimg, 1)
            call('pf_f32_to_bf16', rel.t, ct.c_int64(rel.t.numel()), relb.t, st)
            srcs = [last, relb, emb_up]
        else:
            srcs = [last, emb_up]
        pt = self.buf(k + 'pt', (B, H, Wimg, 8), F32)
        # CLB MLP: 1x1 (161->80) + GELU with the 80->4 + Softplus layer fused in its epilogue (dist_layers.py:91-98)
        self.conv(k + 'clb0', Wh['clb.0'], srcs, act=ACT_GELU, tail=Wh['clb.tail'] + (ACT_SOFTPLUS,), tail_out=pt,
                  skip_main=True)
        depth = self.buf(k + 'depth', (B, H, Wimg), F32) if depth_out is None else depth_out
        assert depth.dtype == F32 and depth.is_contiguous() and tuple(depth.shape) == (B, H, Wimg)
        call('pf_logbinom_depth', pt, 8, b_t, ph, pw_, B, H, Wimg, nb, ct.c_float(_get(bcfg, 'min_temp')),
             ct.c_float(_get(bcfg, 'max_temp')), depth, st)
        return depth

    # ------------------------------------------------------------------ G2L (once per image)
    def g2l(self, coarse_feats):
        Wf = self.W['fusion']
        st = stream_ptr()
        outs = []
        for i, f in enumerate(coarse_feats):
            L = Wf['g2l%d' % i]
            c, heads = L['C'], L['heads']
            h, w = f.hw
            n = h * w
            Hp, Wp = math.ceil(h / WINDOW) * WINDOW, math.ceil(w / WINDOW) * WINDOW
            k = 'g2l%d.' % i
            x = self.buf(k + 'x', (n, c), F32)
            # the reference adds absolute_pos_embed (1, num_patches, C) to the (1, h*w, C) tokens (swin_layers.py:421-422)
            # and would fail on the broadcast if they differ
            assert L['ape'].shape[0] == n, 'guided_fusion.num_patches[%d] = %d does not match the %dx%d coarse map' % (
                i, L['ape'].shape[0], h, w)
            call('pf_g2l_embed', f.t, f.t.shape[-1], L['ape'], n, c, x, st)
            npad = self.buf(k + 'npad', (Hp * Wp, c))
            qkv = self.buf(k + 'qkv', (Hp * Wp, 3 * c))
            att = self.buf(k + 'att', (Hp * Wp, c))
            prj = self.buf(k + 'prj', (Hp * Wp, c), F32)
            hb = self.buf(k + 'h', (n, c))
            hid = self.buf(k + 'hid', (n, 4 * c))
            for bi, bw in enumerate(L['blocks']):
                shift = 0 if bi % 2 == 0 else WINDOW // 2
                call('pf_swin_norm_pad', x, bw['n1w'], bw['n1b'], ct.c_float(1e-5), h, w, Hp, Wp, c, npad, st)
                ops.gemm(bw['qkv'], [npad], qkv)
                call('pf_window_attention', qkv, bw['table'], Hp, Wp, c, heads, shift, att, st)
                ops.gemm(bw['proj'], [att], prj)
                call('pf_swin_residual_crop', x, prj, h, w, Wp, c, st)
                ops.layernorm(x, bw['n2w'], bw['n2b'], 1e-5, hb)
                ops.gemm(bw['fc1'], [hb], hid, act=ACT_GELU)
                ops.gemm(bw['fc2'], [hid], x, gamma=L['ones'])
            o = self.map(k + 'out', 1, h, w, c)
            ops.layernorm(x, L['nw'], L['nb'], 1e-5, o.rows())
            outs.append(o)
        return outs

    # ------------------------------------------------------------------ fusion of T tiles
    def fusion(self, crops, boxes, fine_depth, fine_feats, coarse_depth, coarse_feats, g2l_maps, taps=None,
               depth_out=None):
        """crops planar fp32 [T,3,H,W]; boxes fp32 [T,4] (device, patch_process units); returns fp32 [T,H,W]."""
        Wf = self.W['fusion']
        st = stream_ptr()
        T = crops.shape[0]
        H, Wimg = self.P
        k = 'fus.'
        # ROI crop-zoom of the whole-image coarse maps, fused 3x3 convs with the fine maps (patchfusion.py:263-267)
        guide = []
        for i in range(5):
            cf = coarse_feats[i]
            h, w = cf.hw
            roi = self.map(k + 'croi%d' % i, T, h, w, cf.C)
            ops.roi_crop_zoom(cf.t, cf.C, boxes, h / self.P[0], roi.t)
            guide.append(self.conv(k + 'guide%d' % i, Wf['fc%d' % i], [roi, fine_feats[i]]))
        droi = self.buf(k + 'droi', (T, H, Wimg), F32)
        ops.roi_crop_zoom(coarse_depth, 1, boxes, 1.0, droi)
        u = self.map(k + 'unet_in', T, H, Wimg, 5)
        call('pf_pack_unet_input', droi, fine_depth, crops, T, H, Wimg, u.t, 8, st)
        # encoder
        x = self.conv(k + 'inc0', Wf['inc.0'], [u], act=ACT_RELU)
        x = self.conv(k + 'inc1', Wf['inc.1'], [x], act=ACT_RELU)
        enc = [x]
        for i in range(5):
            h, w = x.hw
            p = self.map(k + 'pool%d' % i, T, h // 2, w // 2, x.C)
            ops.maxpool2(x.t, x.C, p.t)
            x = self.conv(k + 'down%d.0' % i, Wf['down%d.0' % i], [p], act=ACT_RELU)
            x = self.conv(k + 'down%d.1' % i, Wf['down%d.1' % i], [x], act=ACT_RELU)
            enc.append(x)
        enc = enc[::-1]
        outs, prev = [], None
        for i in range(6):
            h, w = g2l_maps[i].hw
            e = self.resize(k + 'encfix%d' % i, enc[i], (h, w))
            if i > 0:
                up_prev = self.resize(k + 'upprev%d' % i, prev, (h, w))
                up_guide = self.resize(k + 'upguide%d' % i, guide[i - 1], (h, w))
                e = self.conv(k + 'up%d.0' % i, Wf['up%d.0' % i], [e, up_prev, up_guide], act=ACT_RELU)
                e = self.conv(k + 'up%d.1' % i, Wf['up%d.1' % i], [e], act=ACT_RELU)
            gm = g2l_maps[i]
            c = self.map(k + 'groi%d' % i, T, h, w, gm.C)
            ops.roi_crop_zoom(gm.t, gm.C, boxes, h / self.P[0], c.t)
            y = self.conv(k + 'cv%d.0' % i, Wf['cv%d.0' % i], [e, c], act=ACT_RELU)
            prev = self.conv(k + 'cv%d.1' % i, Wf['cv%d.1' % i], [y], act=ACT_RELU)
            outs.append(prev)
            if taps is not None:
                taps['fuse%d' % i] = prev.t.clone()
        return self.metric_head('fus.head.', Wf['head'], self.hp['coarse'], self.bcfg['coarse'], outs[0], outs[1:5],
                                outs[5], None, taps, depth_out=depth_out)
