"""Interleaved A/B of library tuning switches inside ONE process (box-to-box and run-to-run variance is several
percent): the vitl 4K P49 step through the CUDA-graph path, configurations visited round-robin, graphs re-captured
after every switch.   python tools/ab_opts.py "tma=1,hmc=0" "tma=1,hmc=1" "tma=0,hmc=0" """
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_inputs
from patchfusion_b200 import lib
from patchfusion_b200.model import PatchFusion

OPT = dict(tma=lib.OPT_TMA_EPILOGUE, hmc=lib.OPT_HALO_MULTICAST, gmc=lib.OPT_GEMM_MULTICAST, frs=lib.OPT_FUSED_RESAMPLE, pdl=lib.OPT_PDL, sep=lib.OPT_RESIZE_SEPARABLE)
configs = sys.argv[1:] or ['tma=1,hmc=0', 'tma=1,hmc=1']
dev = torch.device('cuda:0')
cfg, sd = build_inputs('vitl')
model = PatchFusion(cfg)
model.load_state_dict(sd, strict=True)
model = model.to(dev).eval()
img = torch.rand(1, 3, 2160, 3840, generator=torch.Generator().manual_seed(100)).to(dev)


PN = [9]


def step():
    lr = model.make_lr(img)
    y, _ = model(mode='infer', image_lr=lr, image_hr=img, cai_mode='m2', process_num=PN[0])
    return y


res = {c: [] for c in configs}
for rnd in range(3):
    for c in configs:
        for kv in c.split(','):
            k, v = kv.split('=')
            if k == 'pn':
                PN[0] = int(v)
            else:
                lib.call('pf_set_option', OPT[k], int(v))
        model._graphs = {}
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(6):
            step()
        e1.record()
        torch.cuda.synchronize()
        res[c].append(e0.elapsed_time(e1) / 6)
for c in configs:
    print('%-24s ms/image %s   min %.2f  mean %.2f' % (c, ' '.join('%.2f' % t for t in res[c]), min(res[c]),
                                                      sum(res[c]) / len(res[c])), flush=True)
