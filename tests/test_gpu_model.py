"""End-to-end parity of the CUDA path against the CPU oracle (and the committed reference fixtures) — vits."""
import json
import os
import random

import numpy as np
import pytest
import torch

from gpu_util import check, rel_err

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(1200)]
GOLD = os.path.join(os.path.dirname(__file__), 'golden')
MAX_DEPTH = 80.0


@pytest.fixture(scope='module')
def setup(cuda):
    from oracle import pf_oracle as po
    from patchfusion_b200.configs import depth_anything_patchfusion
    from patchfusion_b200.model import PatchFusion
    from patchfusion_b200.params import synthetic_state_dict
    case = json.load(open(os.path.join(GOLD, 'vits_case0.json')))
    cfg = depth_anything_patchfusion(case['encoder'], image_raw_shape=case['image_raw_shape'],
                                     patch_split_num=case['patch_split_num'])
    sd = synthetic_state_dict(cfg, seed=case['seed'])
    model = PatchFusion(cfg)
    model.load_state_dict(sd, strict=True)
    model = model.to(cuda).eval()
    g = torch.Generator().manual_seed(case['input_seed'])
    img = torch.rand(1, 3, *case['image_raw_shape'], generator=g)
    torch.set_num_threads(os.cpu_count())
    orc = po.Oracle(sd, cfg)
    return dict(case=case, cfg=cfg, sd=sd, model=model, img=img, orc=orc, po=po,
                gold=np.load(os.path.join(GOLD, 'vits_case0.npz')))


def nchw(m):
    return m.t[..., :m.C].float().permute(0, 3, 1, 2).cpu()


def test_coarse_branch_taps(cuda, setup):
    s = setup
    model, orc, img = s['model'], s['orc'], s['img']
    lr = model.resizer(img)
    taps_o, taps_e = {}, {}
    with torch.no_grad():
        d_o, f_o = orc.coarse(lr, taps_o)
    eng = model.engine()
    d_e, f_e = eng.branch('coarse', lr.to(cuda).contiguous(), taps_e)
    torch.cuda.synchronize()
    B, seq, D = 1, 1037, 384
    bad = []

    def soft(name, got, want, tol):
        e = rel_err(got, want)
        print('%-28s rel-Linf %.3e (tol %.0e)' % (name, e, tol))
        if not e < tol:
            bad.append((name, e))

    soft('tokens', taps_e['tokens'].view(B, seq, D).cpu(), taps_o['tokens'], 1e-2)
    for i in range(12):
        soft('block%d' % i, taps_e['block%d' % i].view(B, seq, D).cpu(), taps_o['block%d' % i], 2e-2)
    names = ['x_d0', 'r4', 'r3', 'r2', 'r1', 'out_conv']
    for n, a, b in zip(names, f_e, f_o):
        soft('feat ' + n, nchw(a), b, 3e-2)
    soft('rel depth', taps_e['rel'].cpu(), taps_o['rel'], 3e-2)
    for i in range(4):
        soft('bin centres %d' % i, taps_e['b%d' % i].permute(0, 3, 1, 2).cpu(), taps_o['b%d' % i], 3e-2)
    err = (d_e.cpu() - d_o[:, 0]).abs().max().item()
    print('coarse depth max-abs %.3e  normalised by max_depth %.3e   (range %.3f..%.3f)' %
          (err, err / MAX_DEPTH, d_o.min().item(), d_o.max().item()))
    assert not bad, bad
    assert err / MAX_DEPTH < 1e-3
    assert err / (d_o.max() - d_o.min()).item() < 2e-2, 'error must also be small against the output range'
    # reference fixture (strided sample of the real reference's output)
    st = s['case']['sample_stride']
    g = torch.tensor(s['gold']['coarse_depth'])
    e2 = (d_e.cpu()[:, None, ::st, ::st] - g).abs().max().item()
    print('coarse depth vs reference fixture max-abs %.3e' % e2)
    assert e2 / MAX_DEPTH < 1e-3


def test_fine_and_fusion(cuda, setup):
    s = setup
    model, orc, img, po, cfg, sd = s['model'], s['orc'], s['img'], s['po'], s['cfg'], s['sd']
    eng = model.engine()
    lr = model.resizer(img)
    H, W = s['case']['image_raw_shape']
    h, w = H // 2, W // 2
    raw = [(0, 0), (h // 2, w // 2)]
    P = cfg['patch_process_shape']
    with torch.no_grad():
        d_o, f_o = orc.coarse(lr)
        g2l_o = po.g2l_all(sd, f_o, cfg['guided_fusion'])
        crops = torch.cat([orc.resizer(img[:, :, y:y + h, x:x + w]) for (y, x) in raw])
        fx, fy = 1 / W * P[1], 1 / H * P[0]
        boxes = torch.tensor([[x, y, x + w, y + h] for (y, x) in raw]).int() * torch.tensor([[fx, fy, fx, fy]])
        fd_o, ff_o = po.branch_forward(sd, 'fine_branch.', crops, cfg['fine_branch'])
        rois = [po.roi_crop_zoom(f, boxes, f.shape[-2] / P[0]) for f in f_o]
        droi = po.roi_crop_zoom(d_o, boxes, 1.0)
        taps_o = {}
        fu_o = po.fusion_forward(sd, cfg, fd_o, crops, ff_o, boxes, droi, rois, g2l_o, taps_o)
    cd, cf = eng.branch('coarse', lr.to(cuda).contiguous())
    cd = cd[0].clone()
    cf = [type(f)(f.t.clone(), f.C) for f in cf]
    g2l = eng.g2l(cf)
    bad = []

    def soft(name, got, want, tol):
        e = rel_err(got, want)
        print('%-28s rel-Linf %.3e (tol %.0e)' % (name, e, tol))
        if not e < tol:
            bad.append((name, e))

    for i in range(6):
        soft('g2l level %d' % i, nchw(g2l[i]), g2l_o[i], 3e-2)
    cr = crops.to(cuda).contiguous()
    fd, ff = eng.branch('fine', cr)
    soft('fine depth', fd.cpu(), fd_o[:, 0], 3e-2)
    taps_e = {}
    fu = eng.fusion(cr, boxes.to(cuda).contiguous(), fd, ff, cd, cf, g2l, taps_e)
    torch.cuda.synchronize()
    for i in range(6):
        C = taps_o['fuse%d' % i].shape[1]
        soft('fusion level %d' % i, taps_e['fuse%d' % i][..., :C].float().permute(0, 3, 1, 2).cpu(), taps_o['fuse%d' % i], 3e-2)
    err = (fu.cpu() - fu_o[:, 0]).abs().max().item()
    print('fusion depth max-abs %.3e  /max_depth %.3e  (range %.3f..%.3f)' % (err, err / MAX_DEPTH, fu_o.min().item(), fu_o.max().item()))
    assert not bad, bad
    assert err / MAX_DEPTH < 1e-3
    st = s['case']['sample_stride']
    e2 = (fu.cpu()[:, None, ::st, ::st] - torch.tensor(s['gold']['fusion_depth'])).abs().max().item()
    print('fusion depth vs reference fixture max-abs %.3e' % e2)
    assert e2 / MAX_DEPTH < 1e-3


@pytest.mark.parametrize('mode', ['m1', 'm2', 'r4'])
def test_infer_vs_reference_fixture(cuda, setup, mode):
    s = setup
    model, img, case = s['model'], s['img'], s['case']
    lr = model.resizer(img)
    random.seed(0)
    y, _ = model(mode='infer', image_lr=lr.to(cuda), image_hr=img.to(cuda), cai_mode=mode,
                 process_num=case['process_num'])
    torch.cuda.synchronize()
    st = case['sample_stride']
    g = torch.tensor(s['gold']['infer_' + mode])
    got = y.cpu()[..., ::st, ::st]
    assert got.shape == g.shape
    err = (got - g).abs().max().item()
    print('%s: max-abs vs reference fixture %.3e (/max_depth %.3e; output range %.3f..%.3f)' %
          (mode, err, err / MAX_DEPTH, g.min().item(), g.max().item()))
    assert torch.isfinite(y).all()
    assert err / MAX_DEPTH < 1e-3
    assert err / (g.max() - g.min()).item() < 2e-2, 'error must also be small against the output range'


def test_micro_batch_grouping_invariance(cuda, setup):
    """Tiles are independent and the stitch sums in a fixed order: m2 with process_num 2 / 4 / 9 (different
    micro-batch sizes, graphs and buffer sets) gives the same canvas bit for bit."""
    s = setup
    model, img = s['model'], s['img']
    lr = model.resizer(img)
    outs = []
    for pn in (2, 4, 9):
        y, _ = model(mode='infer', image_lr=lr.to(cuda), image_hr=img.to(cuda), cai_mode='m2', process_num=pn)
        outs.append(y.clone())
    for y in outs[1:]:
        err = (y - outs[0]).abs().max().item()
        print('process_num invariance: max diff %.3e' % err)
        assert err == 0.0


def test_vitl_coarse_and_two_tiles_with_taps(cuda):
    """Depth-Anything-vitl: coarse branch taps, G2L maps and two fused tiles against the GPU-executed oracle (the full
    P16 / P49 canvases are in tests/test_gpu_configs.py)."""
    from oracle import pf_oracle as po
    from patchfusion_b200.configs import depth_anything_patchfusion
    from patchfusion_b200.model import PatchFusion
    from patchfusion_b200.params import synthetic_state_dict
    cfg = depth_anything_patchfusion('vitl')
    sd = synthetic_state_dict(cfg, seed=0)
    model = PatchFusion(cfg)
    model.load_state_dict(sd, strict=True)
    model = model.to(cuda).eval()
    sdc = {k: v.to(cuda) for k, v in sd.items()}
    del sd
    img = torch.rand(1, 3, 2160, 3840, generator=torch.Generator().manual_seed(3)).to(cuda)
    orc = po.Oracle(sdc, cfg)
    lr = orc.resizer(img)
    raw = [(540, 960), (810, 1440)]
    H, W, h, w = 2160, 3840, 540, 960
    P = cfg['patch_process_shape']
    with torch.no_grad():
        cd_o, cf_o = orc.coarse(lr)
        g2l_o = po.g2l_all(sdc, cf_o, cfg['guided_fusion'])
        tc = po.prepare_tile_cfg((H, W), (4, 4), P)
        fu_o = orc.tiles(img, raw, cd_o, cf_o, g2l_o, 2, tc)
    eng = model.engine()
    cd, cf = eng.branch('coarse', lr.contiguous())
    err_c = (cd[0] - cd_o[0, 0]).abs().max().item()
    print('vitl coarse depth max-abs %.3e (range %.3f..%.3f)' % (err_c, cd_o.min().item(), cd_o.max().item()))
    assert err_c / MAX_DEPTH < 1e-3 and err_c / (cd_o.max() - cd_o.min()).item() < 2e-2
    for a, b in zip(cf, cf_o):
        e = rel_err(a.t[..., :a.C].float().permute(0, 3, 1, 2), b)
        assert e < 3e-2, e
    model._coarse = (cd[0], cf, eng.g2l(cf))
    for a, b in zip(model._coarse[2], g2l_o):
        assert rel_err(a.t[..., :a.C].float().permute(0, 3, 1, 2), b) < 4e-2
    io_raw = torch.tensor(raw, dtype=torch.int32, device=cuda)
    fx, fy = np.float32(1 / W * P[1]), np.float32(1 / H * P[0])
    boxes = torch.tensor([[x * fx, y * fy, (x + w) * fx, (y + h) * fy] for (y, x) in raw], dtype=torch.float32, device=cuda)
    pred = torch.empty((2, P[0], P[1]), device=cuda)
    fine = model._fine_stage(eng, img[0].contiguous(), 2, (H, W, h, w, P[0], P[1]), io_raw)
    model._fusion_stage(eng, fine, boxes, pred)
    err = (pred - fu_o[:, 0]).abs().max().item()
    print('vitl fused tiles max-abs %.3e (range %.3f..%.3f)' % (err, fu_o.min().item(), fu_o.max().item()))
    assert err / MAX_DEPTH < 1e-3 and err / (fu_o.max() - fu_o.min()).item() < 2e-2


def _cuda_oracle(cfg, sd, cuda):
    from oracle import pf_oracle as po
    return po, po.Oracle({k: v.to(cuda) for k, v in sd.items()}, cfg)


def test_tile_cfg_override_4x4_and_r_mode(cuda, setup):
    """`tile_cfg=` override (patchfusion.py:402-405): 16 + random tiles on a 4x4 split of the 1080p image, against the
    oracle executed by torch on the GPU (fp32, TF32 off)."""
    s = setup
    model, img = s['model'], s['img'].to(cuda)
    po, orc = _cuda_oracle(s['cfg'], s['sd'], cuda)
    lr = orc.resizer(img)
    tcfg = {'image_raw_shape': [1080, 1920], 'patch_split_num': [4, 4]}
    for mode, pn in (('m1', 4), ('r8', 4)):
        random.seed(1)
        with torch.no_grad():
            want = orc.infer(lr, img, tile_cfg=tcfg, cai_mode=mode, process_num=pn)
        random.seed(1)
        got, _ = model(mode='infer', image_lr=lr, image_hr=img, tile_cfg=tcfg, cai_mode=mode, process_num=pn)
        assert got.shape == want.shape
        err = (got - want).abs().max().item()
        print('%s 4x4: max-abs %.3e (range %.3f..%.3f)' % (mode, err, want.min().item(), want.max().item()))
        assert err / MAX_DEPTH < 1e-3 and err / (want.max() - want.min()).item() < 2e-2
    with pytest.raises(AssertionError):
        model(mode='infer', image_lr=lr, image_hr=img, tile_cfg={'image_raw_shape': [1080, 1920], 'patch_split_num': [7, 4]})


def test_vitb_branch(cuda):
    """Depth-Anything-vitb (BASELINE.json configs[4] encoder: dim 768, 12 heads, features 128): one branch forward of
    two tiles against the GPU-executed oracle."""
    from patchfusion_b200.configs import depth_anything_patchfusion
    from patchfusion_b200.model import PatchFusion
    from patchfusion_b200.params import synthetic_state_dict
    cfg = depth_anything_patchfusion('vitb')
    sd = synthetic_state_dict(cfg, seed=2)
    model = PatchFusion(cfg)
    model.load_state_dict(sd, strict=True)
    model = model.to(cuda).eval()
    po, orc = _cuda_oracle(cfg, sd, cuda)
    x = torch.rand(2, 3, 392, 518, generator=torch.Generator().manual_seed(5)).to(cuda)
    with torch.no_grad():
        d_o, f_o = po.branch_forward(orc.sd, 'fine_branch.', x, cfg['fine_branch'])
    d, f = model.engine().branch('fine', x.contiguous())
    for a, b in zip(f, f_o):
        assert rel_err(a.t[..., :a.C].float().permute(0, 3, 1, 2), b) < 3e-2
    err = (d - d_o[:, 0]).abs().max().item()
    print('vitb fine depth max-abs %.3e (range %.3f..%.3f)' % (err, d_o.min().item(), d_o.max().item()))
    assert err / MAX_DEPTH < 1e-3 and err / (d_o.max() - d_o.min()).item() < 2e-2


def test_stage_level_api(cuda, setup):
    """Reference-named stage entry points (coarse_forward / coarse_postprocess_test / infer_forward, NCHW fp32 in and
    out) reproduce the tile result of the fused path and torchvision's roi_align."""
    from torchvision.ops import roi_align
    s = setup
    model, img = s['model'], s['img'].to(cuda)
    lr = model.resizer(img)
    cd, cf = model.coarse_forward(lr)
    assert cd.shape == (1, 1, 392, 518) and [tuple(f.shape[-2:]) for f in cf][0] == (14, 19)
    H, W = s['case']['image_raw_shape']
    h, w = H // 2, W // 2
    raw = [(0, 0), (h // 2, w // 2)]
    P = s['cfg']['patch_process_shape']
    fx, fy = 1 / W * P[1], 1 / H * P[0]
    bf5 = torch.tensor([[0, x * fx, y * fy, (x + w) * fx, (y + h) * fy] for (y, x) in raw], device=cuda)
    post = model.coarse_postprocess_test(cd, cf, None, bf5)
    ref = roi_align(cf[3], bf5, cf[3].shape[-2:], cf[3].shape[-2] / P[0], aligned=True)
    assert rel_err(post['coarse_feats_roi'][3], ref) < 1e-2
    refd = roi_align(cd, bf5, (392, 518), 1.0, aligned=True)
    assert (post['coarse_depth_roi'] - refd).abs().max().item() < 1e-5
    crops = torch.cat([model.resizer(img[:, :, y:y + h, x:x + w]) for (y, x) in raw])
    pred = model.infer_forward(crops, bf5, {'coarse_prediction': cd, 'coarse_features': cf}, post)
    assert pred.shape == (2, 1, 392, 518)
    want = torch.tensor(s['gold']['fusion_depth'])
    st = s['case']['sample_stride']
    err = (pred.cpu()[..., ::st, ::st] - want).abs().max().item()
    print('infer_forward vs reference fixture max-abs %.3e' % err)
    assert err / MAX_DEPTH < 1e-3
