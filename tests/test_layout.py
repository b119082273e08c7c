"""State-dict layout (drop-in checkpoint compatibility) against the layouts dumped from the reference."""
import json
import os

import pytest
import torch

from patchfusion_b200.configs import depth_anything_patchfusion
from patchfusion_b200.params import state_layout, synthetic_state_dict

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


@pytest.mark.parametrize('enc', ['vits', 'vitl'])
def test_layout_matches_reference(enc):
    ref = json.load(open(os.path.join(GOLD, 'state_dict_layout_%s.json' % enc)))
    L = state_layout(depth_anything_patchfusion(enc))
    assert list(L.keys()) == list(ref.keys())
    for k, (shape, dtype, _) in L.items():
        assert list(shape) == ref[k][0], k
        assert str(dtype).replace('torch.', '') == ref[k][1], k


def test_module_state_dict_and_synthetic_weights():
    from patchfusion_b200.model import PatchFusion
    cfg = depth_anything_patchfusion('vits')
    m = PatchFusion(cfg)
    ref = json.load(open(os.path.join(GOLD, 'state_dict_layout_vits.json')))
    sd = m.state_dict()
    assert list(sd.keys()) == list(ref.keys())
    a, b = synthetic_state_dict(cfg, seed=3), synthetic_state_dict(cfg, seed=3)
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert m.load_state_dict(a, strict=True)
    kept = m.get_save_dict()
    assert kept and not any('coarse_branch' in k or 'fine_branch' in k for k in kept)
    missing = m.load_dict(kept)
    assert len(missing.missing_keys) == len(sd) - len(kept) and not missing.unexpected_keys


def test_unknown_backbone_and_bins():
    bad = depth_anything_patchfusion('vits')
    bad['coarse_branch']['midas_model_type'] = 'DPT_BEiT_L_384'
    with pytest.raises(NotImplementedError):
        state_layout(bad)
    bad = depth_anything_patchfusion('vits')
    bad['coarse_branch']['bin_centers_type'] = 'nope'
    with pytest.raises(ValueError):
        state_layout(bad)
