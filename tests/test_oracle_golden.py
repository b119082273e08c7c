"""The oracle restatement against fixtures produced by the real reference (oracle/make_golden.py)."""
import json
import os
import random

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
pytestmark = pytest.mark.timeout(900)


@pytest.fixture(scope='module')
def ctx():
    from oracle import pf_oracle as po
    from oracle.make_golden import case_inputs
    case = json.load(open(os.path.join(GOLD, 'vits_case0.json')))
    cfg, sd, img = case_inputs(case)
    torch.set_num_threads(os.cpu_count())
    return dict(case=case, cfg=cfg, sd=sd, img=img, po=po, orc=po.Oracle(sd, cfg),
                gold=np.load(os.path.join(GOLD, 'vits_case0.npz')))


def test_coarse_and_fusion_against_reference(ctx):
    po, orc, img, cfg, sd, g, case = ctx['po'], ctx['orc'], ctx['img'], ctx['cfg'], ctx['sd'], ctx['gold'], ctx['case']
    st = case['sample_stride']
    with torch.no_grad():
        lr = orc.resizer(img)
        d, feats = orc.coarse(lr)
        assert np.abs(d[..., ::st, ::st].numpy() - g['coarse_depth']).max() < 1e-5
        for i, f in enumerate(feats):
            s = st if f.shape[-1] > 100 else 1
            want = g['coarse_feat%d' % i]
            assert np.abs(f[..., ::s, ::s].numpy()[:, :8] - want).max() < 1e-4 * max(1.0, np.abs(want).max())
        H, W = case['image_raw_shape']
        h, w = H // 2, W // 2
        raw = [(0, 0), (h // 2, w // 2)]
        P = cfg['patch_process_shape']
        fx, fy = 1 / W * P[1], 1 / H * P[0]
        boxes = torch.tensor([[x, y, x + w, y + h] for (y, x) in raw]).int() * torch.tensor([[fx, fy, fx, fy]])
        crops = torch.cat([orc.resizer(img[:, :, y:y + h, x:x + w]) for (y, x) in raw])
        fd, ff = po.branch_forward(sd, 'fine_branch.', crops, cfg['fine_branch'])
        assert np.abs(fd[..., ::st, ::st].numpy() - g['fine_depth']).max() < 1e-5
        rois = [po.roi_crop_zoom(f, boxes, f.shape[-2] / P[0]) for f in feats]
        assert np.abs(rois[4][..., ::st, ::st].numpy()[:, :8] - g['roi_feat4']).max() < 1e-3
        g2l = po.g2l_all(sd, feats, cfg['guided_fusion'])
        assert np.abs(g2l[4][..., ::st, ::st].numpy()[:, :8] - g['g2l4']).max() < 1e-4
        fu = po.fusion_forward(sd, cfg, fd, crops, ff, boxes, po.roi_crop_zoom(d, boxes, 1.0), rois, g2l)
        assert np.abs(fu[..., ::st, ::st].numpy() - g['fusion_depth']).max() < 1e-3 * np.abs(g['fusion_depth']).max()


def test_infer_m1_against_reference(ctx):
    orc, img, g, case = ctx['orc'], ctx['img'], ctx['gold'], ctx['case']
    st = case['sample_stride']
    with torch.no_grad():
        random.seed(0)
        y = orc.infer(orc.resizer(img), img, cai_mode='m1', process_num=case['process_num'])
    want = g['infer_m1']
    assert y[..., ::st, ::st].shape == want.shape
    assert np.abs(y[..., ::st, ::st].numpy() - want).max() < 1e-3 * np.abs(want).max()


def test_metrics_oracle_matches_reference_fixture():
    """oracle/metrics_oracle.py (numpy restatement of estimator/utils/metric.py) against the values the real reference
    produced in the build container (oracle/make_golden_metrics.py)."""
    import json
    from oracle import metrics_oracle as mo
    from oracle.make_golden_metrics import case_tensors
    for ent in json.load(open(os.path.join(GOLD, 'metrics_case0.json'))):
        c = ent['case']
        gt, pred, edges = case_tensors(c)
        got = mo.compute_metrics(gt, pred, c['lo'], c['hi'], edges)
        for k, v in ent['reference'].items():
            assert abs(float(got[k]) - v) <= 1e-6 * max(1.0, abs(v)), (c['name'], k)


def test_oracle_matches_default_init_reference_fixture():
    """the restatement against the real reference on weights drawn from the reference constructor's own distributions
    (coarse branch only on CPU; the m1 canvas is checked on the GPU)"""
    from oracle import pf_oracle as po
    from patchfusion_b200.configs import depth_anything_patchfusion
    from patchfusion_b200.params import default_init_state_dict
    case = json.load(open(os.path.join(GOLD, 'vits_default0.json')))
    gold = np.load(os.path.join(GOLD, 'vits_default0.npz'))
    cfg = depth_anything_patchfusion(case['encoder'], image_raw_shape=case['image_raw_shape'],
                                     patch_split_num=case['patch_split_num'])
    sd = default_init_state_dict(cfg, seed=case['seed'])
    img = torch.rand(1, 3, *case['image_raw_shape'], generator=torch.Generator().manual_seed(case['input_seed']))
    orc = po.Oracle(sd, cfg)
    with torch.no_grad():
        d, _ = orc.coarse(orc.resizer(img))
    st = case['sample_stride']
    assert (d[..., ::st, ::st] - torch.tensor(gold['coarse_depth'])).abs().max().item() < 1e-5
