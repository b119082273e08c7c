"""Parity of the CUDA path on BASELINE.json's own configurations (c1, c2, c4) against the oracle
executed by torch on the GPU in true fp32 (TF32 off; the CPU oracle needs ~100 s per vitl tile), plus the
size-independent properties the tile path offers at full size: bit-identical canvases for any micro-batch grouping
and any shard count (deterministic stitch), partition of unity of the blend weights.

Tolerances (north_star: "max-abs < 1e-3 on normalised depth"): |d - d_ref| / max_depth(80) < 1e-3 AND
|d - d_ref| / (max d_ref - min d_ref) < 2e-2 (bf16 operands, fp32 accumulate)."""
import random

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(1800)]
MAX_DEPTH = 80.0


def _build(enc, cuda, seed, **kw):
    from oracle import pf_oracle as po
    from patchfusion_b200.configs import depth_anything_patchfusion
    from patchfusion_b200.model import PatchFusion
    from patchfusion_b200.params import synthetic_state_dict
    cfg = depth_anything_patchfusion(enc, **kw)
    sd = synthetic_state_dict(cfg, seed=seed)
    model = PatchFusion(cfg)
    model.load_state_dict(sd, strict=True)
    model = model.to(cuda).eval()
    sdc = {k: v.to(cuda) for k, v in sd.items()}
    del sd
    return cfg, model, po.Oracle(sdc, cfg)


def _report(tag, got, want):
    err = (got - want).abs().max().item()
    rng = (want.max() - want.min()).item()
    print('%s: max-abs %.3e  /max_depth %.3e  /range %.3e  (range %.3f..%.3f)' %
          (tag, err, err / MAX_DEPTH, err / rng, want.min().item(), want.max().item()))
    assert torch.isfinite(got).all()
    assert err / MAX_DEPTH < 1e-3, tag
    assert err / rng < 2e-2, tag


@pytest.fixture(scope='module')
def vitl(cuda):
    cfg, model, orc = _build('vitl', cuda, 0)
    img = torch.rand(1, 3, 2160, 3840, generator=torch.Generator().manual_seed(3)).to(cuda)
    return dict(cfg=cfg, model=model, orc=orc, img=img, lr=orc.resizer(img).contiguous())


def test_vitl_4k_p16_full_canvas(cuda, vitl):
    """configs[1]: Depth-Anything-vitl, 4K, 4x4 split, cai_mode m1 (16 tiles, canvas 1568x2072)."""
    v = vitl
    with torch.no_grad():
        want = v['orc'].infer(v['lr'], v['img'], cai_mode='m1', process_num=4)
    got, _ = v['model'](mode='infer', image_lr=v['lr'], image_hr=v['img'], cai_mode='m1', process_num=9)
    assert got.shape == want.shape == (1, 1, 1568, 2072)
    _report('vitl 4K P16 (m1)', got, want)


def test_vitl_4k_p49_full_canvas_and_invariances(cuda, vitl):
    """configs[2]: Depth-Anything-vitl, 4K, P49 (m2: 16+12+12+9 tiles in 9,9,9,9,9,4 micro-batches) - the canvas
    bench.py times - against the oracle; then the canvas is BIT-identical for another micro-batch size and for the
    tile-sharded decomposition over 8 ranks, emulated rank by rank on this one GPU: the coarse-owner plan (rank 0
    computes coarse + G2L into the broadcast pack and takes 4 tiles, the others 7,7,7,6,6,6,6) and the replicated
    round-robin plan (7,6,..,6)."""
    v = vitl
    with torch.no_grad():
        want = v['orc'].infer(v['lr'], v['img'], cai_mode='m2', process_num=4)
    got, _ = v['model'](mode='infer', image_lr=v['lr'], image_hr=v['img'], cai_mode='m2', process_num=9)
    got = got.clone()
    _report('vitl 4K P49 (m2)', got, want)
    g4, _ = v['model'](mode='infer', image_lr=v['lr'], image_hr=v['img'], cai_mode='m2', process_num=4)
    d = (g4 - got).abs().max().item()
    print('process_num 9 vs 4: max diff %.3e' % d)
    assert d == 0.0, 'tiles are independent and the stitch order is fixed: grouping must not change a bit'
    for how in ('owner', 'replicate'):
        v['model'].shard_coarse = how
        g8, _ = v['model'](mode='infer', image_lr=v['lr'], image_hr=v['img'], cai_mode='m2', process_num=9,
                           shard=('emulate', 8))
        d = (g8 - got).abs().max().item()
        print('8-way tile sharding (%s) vs single: max diff %.3e' % (how, d))
        assert d == 0.0
    v['model'].shard_coarse = 'owner'


def test_vitb_8k_8x8_r128(cuda):
    """configs[4]: Depth-Anything-vitb, 8K (4320x7680), custom 8x8 tiling + 128 random patches: m2 passes
    (64+56+56+49 = 225 tiles, canvas 3136x4144) -> RunningAverageMap.resize to 4320x7680 -> 16 calls x 8 random tiles
    (`patchfusion.py:441-448`, `baseline_pretrain.py:143-218`) = 353 tiles; also sharded 8 ways (emulated)."""
    cfg, model, orc = _build('vitb', cuda, 2)
    tcfg = {'image_raw_shape': [4320, 7680], 'patch_split_num': [8, 8]}
    img = torch.rand(1, 3, 4320, 7680, generator=torch.Generator().manual_seed(11)).to(cuda)
    lr = orc.resizer(img).contiguous()
    random.seed(5)
    with torch.no_grad():
        want = orc.infer(lr, img, tile_cfg=tcfg, cai_mode='r128', process_num=8)
    random.seed(5)
    got, _ = model(mode='infer', image_lr=lr, image_hr=img, tile_cfg=tcfg, cai_mode='r128', process_num=8)
    assert got.shape == want.shape == (1, 1, 4320, 7680)
    got = got.clone()
    _report('vitb 8K 8x8 r128', got, want)
    del want
    random.seed(5)
    g8, _ = model(mode='infer', image_lr=lr, image_hr=img, tile_cfg=tcfg, cai_mode='r128', process_num=8,
                  shard=('emulate', 8))
    d = (g8 - got).abs().max().item()
    print('vitb 8K r128, 8-way tile sharding vs single: max diff %.3e' % d)
    assert d == 0.0


def test_partition_of_unity_p49(cuda):
    """A canvas stitched from constant tiles is that constant (blend weights / their sum) at the full P49 geometry."""
    from oracle import pf_oracle as po
    from patchfusion_b200 import ops
    from patchfusion_b200.model import generatemask
    P = (392, 518)
    tc = po.prepare_tile_cfg((2160, 3840), (4, 4), P)
    plan = [t[1] for p in po.tile_plan(tc, P, 'm2') for t in p]
    assert len(plan) == 49
    mask = torch.tensor(generatemask(P) + 1e-3, device=cuda)
    const = torch.full((49, P[0], P[1]), 3.25, device=cuda)
    tab = torch.tensor([(y, x, i) for i, (y, x) in enumerate(plan)], dtype=torch.int32, device=cuda)
    out = torch.empty((P[0] * 4, P[1] * 4), device=cuda)
    ops.call('pf_stitch_gather', const, tab, 49, P[0], P[1], mask, 0, 0, None, None, P[0] * 4, P[1] * 4, None, None, out,
             ops.stream_ptr())
    assert (out - 3.25).abs().max().item() < 1e-5


def test_reference_default_init_fixture(cuda):
    """Weights from the reference constructor's own distributions (params.default_init_state_dict; statistics pinned
    against a freshly built reference model by oracle/make_golden.py): the untrained network's output is nearly
    constant (range ~0.008 m), so only the north-star tolerance |d - d_ref| / max_depth < 1e-3 is asserted; the error
    against that tiny range is printed.  Compared with the REAL reference's stored outputs and with the oracle."""
    import json
    import os
    import numpy as np
    from oracle import pf_oracle as po
    from patchfusion_b200.configs import depth_anything_patchfusion
    from patchfusion_b200.model import PatchFusion
    from patchfusion_b200.params import default_init_state_dict
    gold_dir = os.path.join(os.path.dirname(__file__), 'golden')
    case = json.load(open(os.path.join(gold_dir, 'vits_default0.json')))
    gold = np.load(os.path.join(gold_dir, 'vits_default0.npz'))
    cfg = depth_anything_patchfusion(case['encoder'], image_raw_shape=case['image_raw_shape'],
                                     patch_split_num=case['patch_split_num'])
    sd = default_init_state_dict(cfg, seed=case['seed'])
    model = PatchFusion(cfg)
    model.load_state_dict(sd, strict=True)
    model = model.to(cuda).eval()
    img = torch.rand(1, 3, *case['image_raw_shape'], generator=torch.Generator().manual_seed(case['input_seed'])).to(cuda)
    lr = model.resizer(img)
    st = case['sample_stride']
    cd, _ = model.coarse_forward(lr)
    e0 = (cd.cpu()[..., ::st, ::st] - torch.tensor(gold['coarse_depth'])).abs().max().item()
    y, _ = model(mode='infer', image_lr=lr, image_hr=img, cai_mode='m1', process_num=case['process_num'])
    g = torch.tensor(gold['infer_m1'])
    e1 = (y.cpu()[..., ::st, ::st] - g).abs().max().item()
    rng = (g.max() - g.min()).item()
    print('default-init: coarse max-abs %.3e, m1 max-abs %.3e (/max_depth %.3e; output range only %.4f -> %.2f of range)'
          % (e0, e1, e1 / MAX_DEPTH, rng, e1 / rng))
    with torch.no_grad():
        want = po.Oracle({k: v.to(cuda) for k, v in sd.items()}, cfg).infer(lr, img, cai_mode='m1',
                                                                            process_num=case['process_num'])
    e2 = (y - want).abs().max().item()
    assert torch.isfinite(y).all()
    assert e0 / MAX_DEPTH < 1e-3 and e1 / MAX_DEPTH < 1e-3 and e2 / MAX_DEPTH < 1e-3


def test_vitl_tile_reference_fixture(cuda, vitl):
    """vitl-size reference fixture (oracle/make_golden.py vitl): the real reference's coarse depth and one fused 4K
    tile (origin (540, 960)) against the CUDA path."""
    import json
    import os
    import numpy as np
    gold_dir = os.path.join(os.path.dirname(__file__), 'golden')
    case = json.load(open(os.path.join(gold_dir, 'vitl_tile0.json')))
    gold = np.load(os.path.join(gold_dir, 'vitl_tile0.npz'))
    v = vitl
    model, img, lr = v['model'], v['img'], v['lr']
    assert case['input_seed'] == 3 and case['seed'] == 0
    st = case['sample_stride']
    cd, cf = model.coarse_forward(lr)
    g0 = torch.tensor(gold['coarse_depth'])
    e0 = (cd.cpu()[..., ::st, ::st] - g0).abs().max().item()
    H, W = case['image_raw_shape']
    h, w = H // 4, W // 4
    y, x = case['tile']
    P = v['cfg']['patch_process_shape']
    fx, fy = 1 / W * P[1], 1 / H * P[0]
    bf5 = torch.tensor([[0, x * fx, y * fy, (x + w) * fx, (y + h) * fy]], device=cuda)
    crop = model.resizer(img[:, :, y:y + h, x:x + w])
    pred = model.infer_forward(crop, bf5, {'coarse_prediction': cd, 'coarse_features': cf})
    g1 = torch.tensor(gold['fusion_depth'])
    e1 = (pred.cpu()[..., ::st, ::st] - g1).abs().max().item()
    r1 = (g1.max() - g1.min()).item()
    print('vitl reference fixture: coarse max-abs %.3e, fused tile max-abs %.3e (/range %.3e)' % (e0, e1, e1 / r1))
    assert e0 / MAX_DEPTH < 1e-3 and e1 / MAX_DEPTH < 1e-3 and e1 / r1 < 2e-2
