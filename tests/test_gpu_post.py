"""SURVEY.md §8f callers on the GPU: fused depth metrics (vs the numpy oracle pinned to the reference's metric.py),
colour mapping, and the tester loop."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def test_depth_metrics_vs_oracle_and_reference_fixture(cuda):
    from oracle import metrics_oracle as mo
    from oracle.make_golden_metrics import case_tensors
    from patchfusion_b200 import metrics
    for ent in json.load(open(os.path.join(GOLD, 'metrics_case0.json'))):
        c = ent['case']
        gt, pred, edges = case_tensors(c)
        got = metrics.compute_metrics(gt.to(cuda), pred.to(cuda), min_depth_eval=c['lo'], max_depth_eval=c['hi'],
                                      disp_gt_edges=edges)
        want = mo.compute_metrics(gt, pred.clone(), c['lo'], c['hi'], edges)
        for k, v in ent['reference'].items():
            assert abs(got[k] - v) <= 2e-5 * max(1.0, abs(v)), (c['name'], k, got[k], v)
            assert abs(got[k] - float(want[k])) <= 2e-5 * max(1.0, abs(v))
    # full size: 4K ground truth, P49 canvas-sized prediction (resampled inside the kernel)
    g = torch.Generator().manual_seed(3)
    gt = torch.rand(1, 1, 2160, 3840, generator=g) * 60 + 0.5
    pred = torch.rand(1, 1, 1568, 2072, generator=g) * 60 + 0.5
    edges = torch.rand(2160, 3840, generator=g) > 0.9
    got = metrics.compute_metrics(gt.to(cuda), pred.to(cuda), min_depth_eval=1e-3, max_depth_eval=80, disp_gt_edges=edges)
    want = mo.compute_metrics(gt, pred, 1e-3, 80, edges)
    for k in want:
        assert abs(got[k] - float(want[k])) <= 1e-4 * max(1.0, abs(float(want[k]))), (k, got[k], want[k])


def test_colorize_and_tester_loop(cuda, tmp_path):
    import cv2
    from patchfusion_b200 import imageio
    from patchfusion_b200.configs import depth_anything_patchfusion
    from patchfusion_b200.model import PatchFusion
    from patchfusion_b200.tester import Tester
    g = torch.Generator().manual_seed(4)
    d = torch.rand(300, 500, generator=g) * 20
    got = imageio.colorize(d.to(cuda)[None, None], cmap='gray_r').cpu().numpy()
    # reference recipe (color.py:112-132) in numpy
    v = d.numpy()
    vmin, vmax = np.percentile(v, 2), np.percentile(v, 95)
    x = (v - vmin) / (vmax - vmin)
    idx = np.clip((x * 256).astype(np.int64), 0, 255)
    idx[x < 0] = 0
    lut = imageio.colormap_lut('gray_r')
    want = lut[idx]
    assert got.shape == (300, 500, 3) and (np.abs(got.astype(int) - want.astype(int)).max() <= 1)
    assert (got != want).mean() < 1e-3          # percentile interpolation in fp32 vs fp64 may move a bin edge
    # tester loop: two synthetic 1080p images, vits, gray-scale PNG + uint16 PNG + metrics
    cfg = depth_anything_patchfusion('vits', image_raw_shape=(1080, 1920), patch_split_num=(2, 2))
    model = PatchFusion(cfg).init_synthetic_weights(0).to(cuda).eval()
    rng = np.random.default_rng(0)
    samples = [dict(img_file_basename='img%d' % i, image_u8=rng.integers(0, 256, (540, 960, 3), dtype=np.uint8),
                    depth_gt=torch.rand(1, 1, 1080, 1920, generator=g) * 2 + 0.2) for i in range(3)]
    t = Tester(model, work_dir=str(tmp_path), save=True, gray_scale=True)
    res = t.run(samples, cai_mode='m1', process_num=2, image_raw_shape=(1080, 1920), patch_split_num=(2, 2))
    assert len(res) == 3 and all(np.isfinite(list(r.values())).all() for r in res)
    ev = Tester.evaluate(res)
    assert set(ev) >= {'a1', 'abs_rel', 'rmse', 'silog'}
    for i in range(3):
        c = cv2.imread(str(tmp_path / ('img%d.png' % i)))
        u = cv2.imread(str(tmp_path / ('img%d_uint16.png' % i)), cv2.IMREAD_UNCHANGED)
        assert c.shape == (784, 1036, 3) and u.dtype == np.uint16 and u.shape == (784, 1036)
    # the uint16 PNG is depth * 256 of the model output
    y, _ = model(mode='infer', cai_mode='m1', process_num=2, tile_cfg={'image_raw_shape': [1080, 1920], 'patch_split_num': [2, 2]},
                 image_hr=imageio.ingest(samples[2]['image_u8'], (1080, 1920), cuda),
                 image_lr=model.make_lr(imageio.ingest(samples[2]['image_u8'], (1080, 1920), cuda)))
    assert np.abs(u.astype(np.int64) - (y[0, 0].cpu().numpy() * 256).astype('uint16').astype(np.int64)).max() <= 1
