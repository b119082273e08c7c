"""Micro-batch (process_num) sweep of the vitl 4K P49 step: ms per image through the CUDA-graph path and the max
difference of the canvas against process_num = 9 (the kernels are batch-invariant, so it must be 0.0).
   python tools/pn_sweep.py 9 13 17 25 49"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_inputs
from patchfusion_b200.model import PatchFusion

pns = [int(a) for a in sys.argv[1:]] or [9, 13, 17, 25, 49]
dev = torch.device('cuda:0')
cfg, sd = build_inputs('vitl')
model = PatchFusion(cfg)
model.load_state_dict(sd, strict=True)
model = model.to(dev).eval()
img = torch.rand(1, 3, 2160, 3840, generator=torch.Generator().manual_seed(100)).to(dev)
ref = None
out = {}
for pn in pns:
    def step():
        lr = model.make_lr(img)
        y, _ = model(mode='infer', image_lr=lr, image_hr=img, cai_mode='m2', process_num=pn)
        return y
    try:
        for _ in range(3):
            y = step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            y = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        y = y.clone()
        if ref is None:
            ref = y
        out[pn] = dict(ms_per_image=ms, tiles_per_s=49e3 / ms, max_diff_vs_first=float((y - ref).abs().max()),
                       mem_gb=torch.cuda.max_memory_allocated() / 2**30)
    except Exception as e:   # noqa
        out[pn] = dict(error=str(e)[:300])
    print(pn, out[pn], flush=True)
os.makedirs('gpurun_out', exist_ok=True)
json.dump(out, open('gpurun_out/pn_sweep.json', 'w'), indent=1)
