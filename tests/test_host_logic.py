"""Host-side logic of the drop-in class (no GPU): tiling geometry, blend mask, error behaviour, HF round trip."""
import os

import numpy as np
import pytest
import torch

from patchfusion_b200.configs import depth_anything_patchfusion
from patchfusion_b200.model import PatchFusion, generatemask


@pytest.fixture(scope='module')
def model():
    return PatchFusion(depth_anything_patchfusion('vits'))


def test_tile_cfg(model):
    tc = model.prepare_tile_cfg((2160, 3840), (4, 4))
    assert tc['patch_raw_shape'] == (540, 960) and tc['patch_reensemble_shape'] == (1568, 2072)
    assert tc['raw_h_split_point'] == [0, 540, 1080, 1620]
    with pytest.raises(AssertionError):
        model.prepare_tile_cfg((2161, 3840), (4, 4))
    with pytest.raises(AssertionError):
        model.prepare_tile_cfg((2160, 3844), (4, 4))


def test_tile_plan_counts():
    from oracle.pf_oracle import prepare_tile_cfg, tile_plan
    P = (392, 518)
    for split, m1, m2 in [((4, 4), 16, 49), ((8, 8), 64, 225), ((2, 2), 4, 9)]:
        tc = prepare_tile_cfg((4320, 7680), split, P)
        assert sum(len(p) for p in tile_plan(tc, P, 'm1')) == m1
        assert sum(len(p) for p in tile_plan(tc, P, 'm2')) == m2
        assert sum(len(p) for p in tile_plan(tc, P, 'r128')) == m2


@pytest.mark.parametrize('size', [(392, 518), (540, 960), (384, 512)])
def test_blend_mask_matches_opencv(size):
    from oracle.pf_oracle import gaussian_mask
    a, b = generatemask(size), gaussian_mask(size)
    assert a.dtype == np.float32 and a.shape == tuple(size) and a.flags['C_CONTIGUOUS']
    assert np.abs(a - b).max() < 5e-6
    assert a.min() == 0.0 and a.max() == 1.0


def test_roi_restatement_matches_torchvision():
    from torchvision.ops import roi_align
    from oracle.pf_oracle import roi_crop_zoom
    g = torch.Generator().manual_seed(0)
    f = torch.randn(1, 8, 28, 37, generator=g)
    boxes = torch.tensor([[0, 0, 129.5, 98.0], [388.5, 294.0, 518.0, 392.0], [101.3, 250.7, 230.8, 348.7]])
    ref = roi_align(f, torch.cat([torch.zeros(3, 1), boxes], 1), (28, 37), 28 / 392, aligned=True)
    assert (roi_crop_zoom(f, boxes, 28 / 392) - ref).abs().max() < 5e-5


def test_error_behaviour(model):
    cfg = depth_anything_patchfusion('vits')
    cfg['fine_branch']['type'] = 'Other'
    with pytest.raises(NotImplementedError):
        PatchFusion(cfg)
    x = torch.rand(2, 3, 392, 518)
    with pytest.raises(AssertionError):          # batch != 1 (patchfusion.py:407)
        model(mode='infer', image_lr=x[:1], image_hr=torch.rand(2, 3, 2160, 3840))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        model(mode='infer', image_lr=x[:1], image_hr=torch.rand(1, 3, 2160, 3840))
    assert model.resizer(torch.rand(1, 3, 540, 960)).shape == (1, 3, 392, 518)
    assert model.tile_cfg['image_raw_shape'] == [2160, 3840]


def test_hf_round_trip(tmp_path):
    cfg = depth_anything_patchfusion('vits')
    m = PatchFusion(cfg).init_synthetic_weights(1)
    if not hasattr(m, 'save_pretrained'):
        pytest.skip('huggingface_hub absent')
    m.save_pretrained(str(tmp_path))
    m.config.to_json_file(str(tmp_path / 'config.json'))          # tools/convert_huggingface.py:79
    m2 = PatchFusion.from_pretrained(str(tmp_path))
    a, b = m.state_dict(), m2.state_dict()
    assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in a)
    assert m2.config.coarse_branch.midas_model_type == 'vits'


def test_reference_import_path_and_registry():
    from estimator.models import build_model
    from estimator.models.patchfusion import PatchFusion as PF
    assert PF is PatchFusion
    m = build_model(dict(type='PatchFusion', config=depth_anything_patchfusion('vits')))
    assert isinstance(m, PatchFusion) and m.patch_process_shape == [392, 518]
