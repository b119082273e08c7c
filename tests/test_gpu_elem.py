"""HBM-bound kernels vs torch fp32 / the oracle's restatements."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from gpu_util import bf, check, from_nhwc, rb, to_nhwc

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def _gen(seed):
    return torch.Generator(device='cuda').manual_seed(seed)


def test_layernorm(cuda):
    from patchfusion_b200 import ops
    g = _gen(0)
    for rows, C in [(1037, 384), (2074, 1024), (266, 64), (5, 32)]:
        x = torch.randn(rows, C, device=cuda, generator=g) * 3 + 1
        w, b = torch.randn(C, device=cuda, generator=g), torch.randn(C, device=cuda, generator=g)
        out = torch.zeros(rows, C, dtype=torch.bfloat16, device=cuda)
        ops.layernorm(x, w, b, 1e-6, out)
        check('layernorm %dx%d' % (rows, C), out, F.layer_norm(x, (C,), w, b, 1e-6), 1e-2)


def test_patch_im2col_and_tokens(cuda):
    from patchfusion_b200 import ops
    g = _gen(1)
    B, H, W, D = 2, 392, 518, 384
    img = torch.rand(B, 3, H, W, device=cuda, generator=g)
    out = torch.zeros(B * 28 * 37, 592, dtype=torch.bfloat16, device=cuda)
    ops.call('pf_patch_im2col', img, B, H, W, out, 592, ops.stream_ptr())
    mean = torch.tensor([0.485, 0.456, 0.406], device=cuda).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225], device=cuda).view(1, 3, 1, 1)
    ref = F.unfold((img - mean) / std, 14, stride=14).transpose(1, 2).reshape(B * 28 * 37, 588)
    check('patch im2col', out[:, :588], ref, 5e-3)
    assert (out[:, 588:] == 0).all()
    patch = torch.randn(B, 1036, D, device=cuda, generator=g)
    cls, pos = torch.randn(D, device=cuda, generator=g), torch.randn(1037, D, device=cuda, generator=g)
    tok = torch.zeros(B, 1037, D, device=cuda)
    ops.call('pf_assemble_tokens', patch, cls, pos, B, 1036, D, tok, ops.stream_ptr())
    ref = torch.cat([cls.view(1, 1, D).expand(B, 1, D), patch], 1) + pos
    check('assemble tokens', tok, ref, 1e-6)


def test_resize_bilinear(cuda):
    from patchfusion_b200 import ops
    g = _gen(2)
    for (B, C, H, W, OH, OW) in [(2, 64, 14, 19, 28, 37), (1, 128, 224, 296, 392, 518), (2, 32, 49, 64, 56, 74),
                                 (1, 8, 196, 259, 224, 296), (1, 64, 56, 74, 56, 74), (2, 256, 112, 148, 224, 296),
                                 (1, 64, 98, 129, 112, 148), (3, 192, 60, 70, 61, 207)]:
        x = torch.randn(B, C, H, W, device=cuda, generator=g)
        out = torch.zeros(B, OH, OW, C + 16, dtype=torch.bfloat16, device=cuda)
        ops.resize_bilinear(to_nhwc(x), C, OH, OW, out, out_col0=8)
        ref = F.interpolate(rb(x), size=(OH, OW), mode='bilinear', align_corners=True)
        check('bilinear %dx%d->%dx%d' % (H, W, OH, OW), out[..., 8:8 + C].float().permute(0, 3, 1, 2), ref, 1e-2)
    x = torch.randn(2, 14, 19, 64, device=cuda, generator=g)
    out = torch.zeros(2, 28, 37, 64, device=cuda)
    ops.call('pf_resize_bilinear_f32', x, 2, 14, 19, 64, 28, 37, out, ops.stream_ptr())
    ref = F.interpolate(x.permute(0, 3, 1, 2), size=(28, 37), mode='bilinear', align_corners=True).permute(0, 2, 3, 1)
    check('bilinear f32', out, ref, 1e-5)


def test_roi_crop_zoom(cuda):
    from patchfusion_b200 import ops
    from torchvision.ops import roi_align
    g = _gen(3)
    P = (392, 518)
    boxes = torch.tensor([[0, 0, 129.5, 98.0], [64.75, 49.0, 194.25, 147.0], [388.5, 294.0, 518.0, 392.0],
                          [101.3, 250.7, 230.8, 348.7]], device=cuda)
    for (C, h, w) in [(64, 14, 19), (64, 224, 296), (32, 392, 518)]:
        f = torch.randn(1, C, h, w, device=cuda, generator=g)
        out = torch.zeros(4, h, w, C, dtype=torch.bfloat16, device=cuda)
        ops.roi_crop_zoom(to_nhwc(f), C, boxes, h / P[0], out)
        bxs = torch.cat([torch.zeros(4, 1, device=cuda), boxes], 1)
        ref = roi_align(rb(f), bxs, (h, w), h / P[0], aligned=True)
        check('roi crop-zoom %dx%d' % (h, w), from_nhwc(out, C), ref, 1e-2)
    d = torch.rand(392, 518, device=cuda, generator=g)
    out = torch.zeros(4, 392, 518, device=cuda)
    ops.roi_crop_zoom(d, 1, boxes, 1.0, out)
    bxs = torch.cat([torch.zeros(4, 1, device=cuda), boxes], 1)
    ref = roi_align(d[None, None], bxs, (392, 518), 1.0, aligned=True)[:, 0]
    check('roi crop-zoom fp32 depth', out, ref, 1e-5)


def test_maxpool_im2col_s2(cuda):
    from patchfusion_b200 import ops
    g = _gen(4)
    x = torch.randn(2, 32, 49, 64, device=cuda, generator=g)
    out = torch.zeros(2, 24, 32, 32, dtype=torch.bfloat16, device=cuda)
    ops.maxpool2(to_nhwc(x), 32, out)
    check('maxpool2', from_nhwc(out, 32), F.max_pool2d(rb(x), 2), 1e-6)
    x = torch.randn(2, 64, 28, 37, device=cuda, generator=g)
    out = torch.zeros(2 * 14 * 19, 9 * 64, dtype=torch.bfloat16, device=cuda)
    ops.call('pf_im2col_3x3_s2', to_nhwc(x), 2, 28, 37, 64, 64, out, ops.stream_ptr())
    ref = F.unfold(rb(x), 3, padding=1, stride=2)                       # [2, 64*9, 266] (c, tap)
    ref = ref.view(2, 64, 9, 266).permute(0, 3, 2, 1).reshape(2 * 266, 9 * 64)
    check('im2col 3x3 s2', out, ref, 1e-6)


def test_crop_resize_and_unet_input(cuda):
    from patchfusion_b200 import ops
    g = _gen(5)
    img = torch.rand(3, 1080, 1920, device=cuda, generator=g)
    origins = torch.tensor([[0, 0], [270, 480], [540, 960], [133, 777]], dtype=torch.int32, device=cuda)
    out = torch.zeros(4, 3, 392, 518, device=cuda)
    ops.call('pf_crop_resize', img, 1080, 1920, origins, 4, 540, 960, 392, 518, out, ops.stream_ptr())
    ref = torch.cat([F.interpolate(img[None, :, y:y + 540, x:x + 960], size=(392, 518), mode='bilinear',
                                   align_corners=True) for y, x in origins.tolist()])
    check('crop + resize', out, ref, 1e-5)
    cd, fd = torch.rand(4, 392, 518, device=cuda, generator=g), torch.rand(4, 392, 518, device=cuda, generator=g)
    u = torch.zeros(4, 392, 518, 8, dtype=torch.bfloat16, device=cuda)
    ops.call('pf_pack_unet_input', cd, fd, out, 4, 392, 518, u, 8, ops.stream_ptr())
    ref = torch.cat([cd[:, None], fd[:, None], out], 1)
    check('unet input', from_nhwc(u, 5), rb(ref), 1e-6)
    assert (u[..., 5:] == 0).all()


def test_swin_helpers(cuda):
    from patchfusion_b200 import ops
    g = _gen(6)
    H, W, C = 14, 19, 64
    Hp, Wp = 24, 24
    feat = torch.randn(1, C, H, W, device=cuda, generator=g)
    ape = torch.randn(H * W, C, device=cuda, generator=g)
    x = torch.zeros(H * W, C, device=cuda)
    ops.call('pf_g2l_embed', to_nhwc(feat), C, ape, H * W, C, x, ops.stream_ptr())
    xr = rb(feat).flatten(2).transpose(1, 2)[0] + ape
    check('g2l embed', x, xr, 1e-6)
    w, b = torch.randn(C, device=cuda, generator=g), torch.randn(C, device=cuda, generator=g)
    out = torch.full((Hp * Wp, C), 3.0, dtype=torch.bfloat16, device=cuda)
    ops.call('pf_swin_norm_pad', x, w, b, ops.C.c_float(1e-5), H, W, Hp, Wp, C, out, ops.stream_ptr())
    ref = F.pad(F.layer_norm(xr, (C,), w, b, 1e-5).view(H, W, C), (0, 0, 0, Wp - W, 0, Hp - H)).reshape(Hp * Wp, C)
    check('swin norm+pad', out, ref, 1e-2)
    y = torch.randn(Hp * Wp, C, device=cuda, generator=g)
    x2 = x.clone()
    ops.call('pf_swin_residual_crop', x2, y, H, W, Wp, C, ops.stream_ptr())
    check('swin residual crop', x2, x + y.view(Hp, Wp, C)[:H, :W].reshape(H * W, C), 1e-6)


def test_metric_tail(cuda):
    from patchfusion_b200 import ops
    from oracle import pf_oracle as po
    g = _gen(7)
    B, H, W, PH, PW, E = 2, 28, 37, 14, 19, 128
    a = torch.randn(B, E, H, W, device=cuda, generator=g)
    prev = torch.randn(B, E, PH, PW, device=cuda, generator=g)
    out = torch.zeros(B, H, W, E, dtype=torch.bfloat16, device=cuda)
    ops.call('pf_add_upsampled', to_nhwc(a), B, H, W, E, to_nhwc(prev), PH, PW, out, ops.stream_ptr())
    check('emb + up(prev)', from_nhwc(out, E), rb(a) + po.up(rb(prev), (H, W)), 1e-2)
    nA, nb = 16, 64
    A = F.softplus(torch.randn(B, H, W, 32, device=cuda, generator=g))
    bprev = F.softplus(torch.randn(B, PH, PW, nb, device=cuda, generator=g))
    bout = torch.zeros(B, H, W, nb, device=cuda)
    ops.call('pf_attractor', A, 32, nA, bprev, PH, PW, B, H, W, nb, 1, bout, ops.stream_ptr())
    bu = po.up(bprev.permute(0, 3, 1, 2), (H, W))
    An = A[..., :nA].permute(0, 3, 1, 2)
    ref = bu + po.inv_attractor(An.unsqueeze(2) - bu.unsqueeze(1)).mean(1)
    check('attractor', bout.permute(0, 3, 1, 2), ref, 1e-5)
    # log-binomial expectation
    Hh, Ww, BH, BW = 56, 74, 28, 37
    pt = F.softplus(torch.randn(B, Hh, Ww, 8, device=cuda, generator=g))
    bc = F.softplus(torch.randn(B, BH, BW, nb, device=cuda, generator=g)) * 3
    depth = torch.zeros(B, Hh, Ww, device=cuda)
    ops.call('pf_logbinom_depth', pt, 8, bc, BH, BW, B, Hh, Ww, nb, ops.C.c_float(0.0212), ops.C.c_float(50.0),
             depth, ops.stream_ptr())
    p4 = pt[..., :4].permute(0, 3, 1, 2)
    p, t = p4[:, :2] + 1e-4, p4[:, 2:] + 1e-4
    p = (p[:, 0] / (p[:, 0] + p[:, 1])).unsqueeze(1)
    t = (t[:, 0] / (t[:, 0] + t[:, 1])).unsqueeze(1)
    t = (50.0 - 0.0212) * t + 0.0212
    k = torch.arange(nb, dtype=torch.float32, device=cuda).view(1, nb, 1, 1)
    n_, k_ = torch.tensor(63.0, device=cuda) + 1e-7, k + 1e-7
    logc = n_ * torch.log(n_) - k_ * torch.log(k_) - (n_ - k_) * torch.log(n_ - k_ + 1e-7)
    y = logc + k * torch.log(p.clamp(1e-4, 1)) + (63 - k) * torch.log((1 - p).clamp(1e-4, 1))
    ref = (torch.softmax(y / t, 1) * po.up(bc.permute(0, 3, 1, 2), (Hh, Ww))).sum(1)
    check('log-binomial depth', depth, ref, 1e-4)


def test_stitch(cuda):
    from patchfusion_b200 import ops
    from oracle import pf_oracle as po
    g = _gen(8)
    CH, CW, th, tw = 784, 1036, 392, 518
    mask = torch.rand(th, tw, device=cuda, generator=g) + 1e-3
    tiles = torch.rand(5, th, tw, device=cuda, generator=g)
    org = [(0, 0), (0, 518), (392, 0), (392, 518), (196, 259)]
    num, den = torch.zeros(CH, CW, device=cuda), torch.zeros(CH, CW, device=cuda)
    o4 = torch.tensor(org[:4], dtype=torch.int32, device=cuda)
    o1 = torch.tensor(org[4:], dtype=torch.int32, device=cuda)
    ops.call('pf_stitch_accumulate', num, den, CH, CW, tiles, 4, th, tw, o4, mask, 0, 0, ops.stream_ptr())
    ops.call('pf_stitch_accumulate', num, den, CH, CW, tiles[4:], 1, th, tw, o1, mask, 0, 0, ops.stream_ptr())
    out = torch.zeros(CH, CW, device=cuda)
    ops.call('pf_stitch_finalize', num, den, CH * CW, out, ops.stream_ptr())
    # literal running average (oracle)
    cnt, acc = torch.zeros(CH, CW, device=cuda), torch.zeros(CH, CW, device=cuda)
    for (y, x), d in zip(org[:4], tiles[:4]):
        cnt[y:y + th, x:x + tw] = mask
        acc[y:y + th, x:x + tw] = d * mask
    ra = po.RunningAverage(acc, cnt)
    c2, a2 = torch.zeros(CH, CW, device=cuda), torch.zeros(CH, CW, device=cuda)
    c2[196:196 + th, 259:259 + tw] = mask
    a2[196:196 + th, 259:259 + tw] = tiles[4] * mask
    ra.update(a2, c2)
    check('stitch (regular)', out, ra.avg, 1e-5)
    # resize + a random (nearest-upsampled) tile
    OH, OW, uh, uw = 1080, 1920, 540, 960
    n2, d2 = torch.zeros(OH, OW, device=cuda), torch.zeros(OH, OW, device=cuda)
    ops.call('pf_stitch_resize', num, den, CH, CW, OH, OW, n2, d2, ops.stream_ptr())
    mask2 = torch.rand(uh, uw, device=cuda, generator=g) + 1e-3
    o = torch.tensor([[100, 333]], dtype=torch.int32, device=cuda)
    ops.call('pf_stitch_accumulate', n2, d2, OH, OW, tiles[:1], 1, th, tw, o, mask2, uh, uw, ops.stream_ptr())
    out2 = torch.zeros(OH, OW, device=cuda)
    ops.call('pf_stitch_finalize', n2, d2, OH * OW, out2, ops.stream_ptr())
    ra.resize((OH, OW))
    upd = F.interpolate(tiles[:1, None], (uh, uw))[0, 0]
    c3, a3 = torch.zeros(OH, OW, device=cuda), torch.zeros(OH, OW, device=cuda)
    c3[100:100 + uh, 333:333 + uw] = mask2
    a3[100:100 + uh, 333:333 + uw] = upd * mask2
    ra.update(a3, c3)
    check('stitch (resize + random tile)', out2, ra.avg, 1e-5)


def test_stitch_gather_deterministic(cuda):
    """pf_stitch_gather (the product path's stitch): equals the literal sequential RunningAverage to fp32 rounding, is
    bit-identical under any permutation of where the predictions sit (slot table = the gathered per-rank blocks), and
    continues from resized base canvases for the random phase."""
    from patchfusion_b200 import ops
    from patchfusion_b200.parallel import slot_table
    from oracle import pf_oracle as po
    g = _gen(18)
    CH, CW, th, tw = 784, 1036, 392, 518
    mask = torch.rand(th, tw, device=cuda, generator=g) + 1e-3
    org = [(0, 0), (0, 518), (392, 0), (392, 518), (0, 259), (392, 259), (196, 0), (196, 518), (196, 259)]
    n = len(org)
    tiles = torch.rand(n, th, tw, device=cuda, generator=g)

    def gather(preds, slots, up=(0, 0), msk=mask, base=(None, None), shape=(CH, CW), want=('num', 'den', 'avg')):
        tab = torch.tensor([(y, x, s_) for (y, x), s_ in zip(org_l, slots)], dtype=torch.int32, device=cuda)
        o = {k: torch.empty(shape, device=cuda) for k in want}
        ops.call('pf_stitch_gather', preds, tab, len(slots), th, tw, msk, up[0], up[1], base[0], base[1], shape[0],
                 shape[1], o.get('num'), o.get('den'), o.get('avg'), ops.stream_ptr())
        return o

    org_l = org
    a = gather(tiles, list(range(n)))
    # sharded layout: world 4, blocks of ceil(9/4)=3 rows, padding rows NaN
    world, per = 4, 3
    full = torch.full((world * per, th, tw), float('nan'), device=cuda)
    slots = slot_table(n, world)
    for i, s_ in enumerate(slots):
        full[s_] = tiles[i]
    b = gather(full, slots)
    assert torch.equal(a['avg'], b['avg']) and torch.equal(a['num'], b['num']) and torch.equal(a['den'], b['den'])
    cnt, acc = torch.zeros(CH, CW, device=cuda), torch.zeros(CH, CW, device=cuda)
    for (y, x), d in zip(org[:4], tiles[:4]):
        cnt[y:y + th, x:x + tw] = mask
        acc[y:y + th, x:x + tw] = d * mask
    ra = po.RunningAverage(acc, cnt)
    for (y, x), d in zip(org[4:], tiles[4:]):
        c2, a2 = torch.zeros(CH, CW, device=cuda), torch.zeros(CH, CW, device=cuda)
        c2[y:y + th, x:x + tw] = mask
        a2[y:y + th, x:x + tw] = d * mask
        ra.update(a2, c2)
    check('stitch gather (regular, 9 tiles)', a['avg'], ra.avg, 1e-5)
    # random phase on top of the resized canvases
    OH, OW, uh, uw = 1080, 1920, 540, 960
    n2, d2 = torch.empty(OH, OW, device=cuda), torch.empty(OH, OW, device=cuda)
    ops.call('pf_stitch_resize', a['num'], a['den'], CH, CW, OH, OW, n2, d2, ops.stream_ptr())
    mask2 = torch.rand(uh, uw, device=cuda, generator=g) + 1e-3
    org_l = [(100, 333), (400, 333), (17, 333)]
    r = gather(tiles[:3].contiguous(), [0, 1, 2], up=(uh, uw), msk=mask2, base=(n2, d2), shape=(OH, OW), want=('avg',))
    ra.resize((OH, OW))
    for (y, x), d in zip(org_l, F.interpolate(tiles[:3, None], (uh, uw))[:, 0]):
        c3, a3 = torch.zeros(OH, OW, device=cuda), torch.zeros(OH, OW, device=cuda)
        c3[y:y + uh, x:x + uw] = mask2
        a3[y:y + uh, x:x + uw] = d * mask2
        ra.update(a3, c3)
    check('stitch gather (resize + 3 random tiles)', r['avg'], ra.avg, 1e-5)


def test_ingest_and_u16_writer(cuda):
    """the callers either side of the path: bicubic align_corners ingest and the uint16 depth writer"""
    from patchfusion_b200 import imageio
    g = torch.Generator().manual_seed(9)
    img = torch.randint(0, 256, (270, 480, 3), generator=g, dtype=torch.uint8)
    got = imageio.ingest(img.numpy(), (1080, 1920), cuda, bgr=True)
    rgb = img.numpy()[:, :, ::-1].copy()
    ref = F.interpolate(torch.tensor(rgb / 255.0).unsqueeze(0).permute(0, 3, 1, 2), (1080, 1920), mode='bicubic',
                        align_corners=True).float()
    assert got.shape == (1, 3, 1080, 1920)
    assert (got.cpu() - ref).abs().max().item() < 1e-4
    d = torch.rand(784, 1036, generator=g) * 80
    u = imageio.depth_to_u16(d.to(cuda)[None, None], (1080, 1920))
    want = (F.interpolate(d[None, None], (1080, 1920))[0, 0].numpy() * 256).astype('uint16')
    assert u.dtype == torch.uint16 and (u.cpu().numpy().astype(np.int64) - want.astype(np.int64)).__abs__().max() <= 1
