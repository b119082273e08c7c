"""A/B harness: the vitl 4K P49 step timed through the CUDA-graph path, then one eager step with the per-launch
profiler; writes {ms_per_image, by_label} to gpurun_out/ab_<tag>.json.  Box-to-box variance is ~6 %, so variants are
compared inside ONE gpurun call:
   PF_B200_LIBNAME=libpf_b200_base.so python tools/ab_step.py base; python tools/ab_step.py new; python tools/ab_step.py --diff base new"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if sys.argv[1] == '--diff':
    a, b = [json.load(open('gpurun_out/ab_%s.json' % t)) for t in sys.argv[2:4]]
    print('step: %s %.2f ms   %s %.2f ms   (%+.2f %%)' % (sys.argv[2], a['ms_per_image'], sys.argv[3], b['ms_per_image'],
                                                        100 * (b['ms_per_image'] / a['ms_per_image'] - 1)))
    print('sum of kernel times (eager): %.2f -> %.2f ms' % (sum(v[0] for v in a['by_label'].values()),
                                                           sum(v[0] for v in b['by_label'].values())))
    rows = []
    for k in sorted(set(a['by_label']) | set(b['by_label'])):
        ma, mb = a['by_label'].get(k, [0, 0])[0], b['by_label'].get(k, [0, 0])[0]
        rows.append((mb - ma, k, ma, mb))
    rows.sort()
    for dlt, k, ma, mb in rows:
        if abs(dlt) > 0.05:
            print('%-70s %8.3f -> %8.3f ms  (%+.3f)' % (k, ma, mb, dlt))
    sys.exit(0)

import torch
from bench import build_inputs
from patchfusion_b200 import lib
from patchfusion_b200.model import PatchFusion

tag = sys.argv[1]
pn = int(sys.argv[2]) if len(sys.argv) > 2 else 9
dev = torch.device('cuda:0')
cfg, sd = build_inputs('vitl')
model = PatchFusion(cfg)
model.load_state_dict(sd, strict=True)
model = model.to(dev).eval()
img = torch.rand(1, 3, 2160, 3840, generator=torch.Generator().manual_seed(100)).to(dev)


def step():
    lr = model.make_lr(img)
    y, _ = model(mode='infer', image_lr=lr, image_hr=img, cai_mode='m2', process_num=pn)
    return y


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    y = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
lib.PROFILER = lib.Profiler()
step()
torch.cuda.synchronize()
lib.PROFILER.start()
step()
torch.cuda.synchronize()
recs = lib.PROFILER.stop()
lib.PROFILER = None
by = {}
for fam, label, fl, t in recs:
    d = by.setdefault(label, [0.0, 0.0, 0])
    d[0] += t; d[1] += fl; d[2] += 1
os.makedirs('gpurun_out', exist_ok=True)
json.dump(dict(ms_per_image=ms, tiles_per_s=49e3 / ms, by_label=by, checksum=float(y.double().sum())),
          open('gpurun_out/ab_%s.json' % tag, 'w'))
torch.save(y.cpu(), '/tmp/ab_%s.pt' % tag)
msg = ''
if tag != 'base' and os.path.exists('/tmp/ab_base.pt'):
    ref = torch.load('/tmp/ab_base.pt')
    msg = '  max|y - base| %.3e (range %.3f..%.3f)' % (float((y.cpu() - ref).abs().max()), float(ref.min()), float(ref.max()))
print(tag, 'ms/image %.2f  tiles/s %.1f  checksum %.6f%s' % (ms, 49e3 / ms, float(y.double().sum()), msg))
