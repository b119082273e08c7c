import sys, os, json, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
from oracle import pf_oracle as po
from patchfusion_b200.configs import depth_anything_patchfusion
from patchfusion_b200.model import PatchFusion
from patchfusion_b200.params import synthetic_state_dict
from patchfusion_b200 import ops
dev = torch.device('cuda:0')
cfg = depth_anything_patchfusion('vits', image_raw_shape=(1080,1920), patch_split_num=(2,2))
sd = synthetic_state_dict(cfg, seed=0)
model = PatchFusion(cfg); model.load_state_dict(sd); model = model.to(dev).eval()
g = torch.Generator().manual_seed(0)
img = torch.rand(1,3,1080,1920, generator=g)
sdc = {k: v.to(dev) for k, v in sd.items()}
orc = po.Oracle(sdc, cfg)
imgc = img.to(dev)
lr = orc.resizer(imgc)
with torch.no_grad():
    cd_o, cf_o = orc.coarse(lr)
    g2l_o = po.g2l_all(sdc, cf_o, cfg['guided_fusion'])
    tc = po.prepare_tile_cfg((1080,1920),(2,2),(392,518))
    plan = po.tile_plan(tc, (392,518), 'm2')
    raw = [t[0] for p in plan for t in p]; proc=[t[1] for p in plan for t in p]
    print('raw', raw); print('proc', proc)
    preds_o = orc.tiles(imgc, raw, cd_o, cf_o, g2l_o, 2, tc)
eng = model.engine()
cd, cf = eng.branch('coarse', lr.contiguous()); cd = cd[0].clone(); cf = [type(f)(f.t.clone(), f.C) for f in cf]
g2l = eng.g2l(cf)
i = 0
for pred, T in model._tile_batch(eng, imgc[0].contiguous(), raw, (cd, cf, g2l), model.prepare_tile_cfg((1080,1920),(2,2)), 2):
    torch.cuda.synchronize()
    for j in range(T):
        e = (pred[j] - preds_o[i+j,0]).abs()
        print('tile', i+j, raw[i+j], 'max err %.3e mean %.3e' % (e.max().item(), e.mean().item()))
    i += T
for mode in ['m1','m2']:
    y,_ = model(mode='infer', image_lr=lr, image_hr=imgc, cai_mode=mode, process_num=2)
    with torch.no_grad(): yo = orc.infer(lr, imgc, cai_mode=mode, process_num=2)
    e = (y-yo).abs()[0,0]
    print(mode, 'max', e.max().item(), 'argmax', np.unravel_index(e.argmax().item(), e.shape), 'mean', e.mean().item())
    for r0 in range(0, e.shape[0], 196):
        print(' '.join('%.1e' % e[r0:r0+196, c0:c0+259].max().item() for c0 in range(0, e.shape[1], 259)))
print('---- stitch isolation')
mask = model._mask((392,518), dev)
mo = torch.tensor(po.gaussian_mask((392,518)) + 1e-3, device=dev)
print('mask stats', mask.min().item(), mask.max().item(), mask.mean().item(), 'vs cv2', (mask-mo).abs().max().item(), mask.shape, mask.dtype, mask.is_contiguous())
preds = []
for pred, T in model._tile_batch(eng, imgc[0].contiguous(), raw, (cd, cf, g2l), model.prepare_tile_cfg((1080,1920),(2,2)), 2):
    preds.append(pred[:T].clone())
preds = torch.cat(preds)
num = torch.zeros(784,1036, device=dev); den = torch.zeros(784,1036, device=dev)
for (py,px), d in zip(proc, preds):
    num[py:py+392, px:px+518] += mask*d; den[py:py+392, px:px+518] += mask
ref = num/den
n2 = torch.zeros(784,1036, device=dev); d2 = torch.zeros(784,1036, device=dev)
org = torch.tensor(proc, dtype=torch.int32, device=dev)
ops.call('pf_stitch_accumulate', n2, d2, 784, 1036, preds, 9, 392, 518, org, mask, 0, 0, ops.stream_ptr())
out = torch.empty_like(n2); ops.call('pf_stitch_finalize', n2, d2, ops.C.c_int64(n2.numel()), out, ops.stream_ptr())
torch.cuda.synchronize()
print('kernel(all 9 at once) vs torch closed form', (out-ref).abs().max().item())
print('torch closed form vs oracle m2', (ref - yo[0,0]).abs().max().item())
n3 = torch.zeros(784,1036, device=dev); d3 = torch.zeros(784,1036, device=dev)
i=0
for T in [2,2,2,2,1]:
    org = torch.tensor(proc[i:i+T], dtype=torch.int32, device=dev)
    ops.call('pf_stitch_accumulate', n3, d3, 784, 1036, preds[i:i+T].contiguous(), T, 392, 518, org, mask, 0, 0, ops.stream_ptr()); i+=T
out3 = torch.empty_like(n3); ops.call('pf_stitch_finalize', n3, d3, ops.C.c_int64(n3.numel()), out3, ops.stream_ptr())
torch.cuda.synchronize()
print('kernel(batched) vs torch closed form', (out3-ref).abs().max().item())
print('model m2 vs torch closed form', (y[0,0]-ref).abs().max().item())
