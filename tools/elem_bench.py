"""Micro-benchmark of the HBM-bound kernels on the shapes of a vitl micro-batch (9 tiles): achieved GB/s against the
algorithmic bytes (bytes in + bytes out)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from patchfusion_b200 import ops

dev = torch.device('cuda:0')
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
T = 9
only = sys.argv[1:]
res = []


def run(name, fn, nbytes):
    if only and not any(o in name for o in only):
        return
    for _ in range(3):
        fn()
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print('%-58s %8.3f ms %8.1f GB/s' % (name, ms, nbytes / ms / 1e6), flush=True)
    res.append(dict(name=name, ms=ms, gbs=nbytes / ms / 1e6, bytes=nbytes))


# fusion resample: up(prev) 256 ch 224x296 -> 392x518 (guided_fusion_model.py:98)
x = torch.randn(T, 224, 296, 256, device=dev).to(torch.bfloat16)
o = torch.empty(T, 392, 518, 256, dtype=torch.bfloat16, device=dev)
run('resize_bilinear up(prev) 256ch 224x296->392x518 x9', lambda: ops.resize_bilinear(x, 256, 392, 518, o), x.numel() * 2 + o.numel() * 2)
x2 = torch.randn(T, 224, 296, 128, device=dev).to(torch.bfloat16)
o2 = torch.empty(T, 392, 518, 128, dtype=torch.bfloat16, device=dev)
run('resize_bilinear oc1/emb 128ch 224x296->392x518 x9', lambda: ops.resize_bilinear(x2, 128, 392, 518, o2), x2.numel() * 2 + o2.numel() * 2)
# ROI crop-zoom of the coarse r1 tap (patchfusion.py:247)
f = torch.randn(1, 224, 296, 256, device=dev).to(torch.bfloat16)
boxes = torch.tensor([[i * 40.0, i * 30.0, i * 40.0 + 129.5, i * 30.0 + 98.0] for i in range(T)], device=dev)
ro = torch.empty(T, 224, 296, 256, dtype=torch.bfloat16, device=dev)
run('roi_crop_zoom 256ch @224x296 x9', lambda: ops.roi_crop_zoom(f, 256, boxes, 224 / 392, ro), ro.numel() * 2 + f.numel() * 2 / 16 * T)
# layernorm of the ViT stream
xs = torch.randn(T * 1037, 1024, device=dev)
w = torch.ones(1024, device=dev); b = torch.zeros(1024, device=dev)
lo = torch.empty(T * 1037, 1024, dtype=torch.bfloat16, device=dev)
run('layernorm 9333x1024 fp32->bf16', lambda: ops.layernorm(xs, w, b, 1e-6, lo), xs.numel() * 4 + lo.numel() * 2)
# log-binomial depth
pt = torch.rand(T, 392, 518, 8, device=dev)
bc = torch.rand(T, 224, 296, 64, device=dev)
dp = torch.empty(T, 392, 518, device=dev)
run('logbinom_depth 64 bins @392x518 x9', lambda: ops.call('pf_logbinom_depth', pt, 8, bc, 224, 296, T, 392, 518, 64, ops.C.c_float(0.0212), ops.C.c_float(50.0), dp, ops.stream_ptr()),
    pt.numel() * 4 + bc.numel() * 4 + dp.numel() * 4)
json.dump(res, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out', 'elem_bench.json'), 'w'))
