"""One eager (graph-free) 4K P49 step of the vitl path between cudaProfilerStart/Stop, for
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file ... python tools/profile_step.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['PF_B200_GRAPHS'] = '0'
import torch
from bench import build_inputs
from patchfusion_b200.model import PatchFusion

enc = sys.argv[1] if len(sys.argv) > 1 else 'vitl'
mode = sys.argv[2] if len(sys.argv) > 2 else 'm2'
pn = int(sys.argv[3]) if len(sys.argv) > 3 else 7
dev = torch.device('cuda:0')
cfg, sd = build_inputs(enc)
model = PatchFusion(cfg)
model.load_state_dict(sd, strict=True)
model = model.to(dev).eval()
img = torch.rand(1, 3, 2160, 3840, generator=torch.Generator().manual_seed(100)).to(dev)
lr = model.make_lr(img)
model(mode='infer', image_lr=lr, image_hr=img, cai_mode='m1' if mode == 'm2' else mode, process_num=pn)   # warm-up (allocations, maps)
torch.cuda.synchronize()
torch.cuda.profiler.start()
model(mode='infer', image_lr=lr, image_hr=img, cai_mode=mode, process_num=pn)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
