"""Micro-benchmark of pf_gemm on the shapes that dominate a vitl tile (CUDA events, L2-flushed between reps)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from patchfusion_b200 import ops

dev = torch.device('cuda:0')
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

SHAPES = [
    # name, kind, params
    ('vit qkv    M7259 K1024 N3072', 'lin', (7259, 1024, 3072)),
    ('vit proj   M7259 K1024 N1024', 'lin', (7259, 1024, 1024)),
    ('vit fc1    M7259 K1024 N4096', 'lin', (7259, 1024, 4096)),
    ('vit fc2    M7259 K4096 N1024', 'lin', (7259, 4096, 1024)),
    ('up4.0 3x3 [32,256,256]->544 @392x518 x7', 'conv', (7, 392, 518, [32, 256, 256], 544)),
    ('up4.0 3x3 [32,256,256]->544 @392x518 x9', 'conv', (9, 392, 518, [32, 256, 256], 544)),
    ('up4.1 3x3 544->32 @392x518 x7', 'conv', (7, 392, 518, [544], 32)),
    ('up3.0 3x3 [256,256,256]->768 @224x296 x7', 'conv', (7, 224, 296, [256, 256, 256], 768)),
    ('up3.1 3x3 768->256 @224x296 x7', 'conv', (7, 224, 296, [768], 256)),
    ('cv4.0 3x3 [256,256]->256 @224x296 x7', 'conv', (7, 224, 296, [256, 256], 256)),
    ('rcu 3x3 256->256 @112x148 x7', 'conv', (7, 112, 148, [256], 256)),
    ('oc1 3x3 256->128 @224x296 x7', 'conv', (7, 224, 296, [256], 128)),
    ('oc2 3x3 128->32 @392x518 x7', 'conv', (7, 392, 518, [128], 32)),
    ('inc 3x3 8->32 @392x518 x7', 'conv', (7, 392, 518, [8], 32)),
    ('1x1 256->256 @112x148 x7', 'conv1', (7, 112, 148, [256], 256)),
    ('clb0 1x1 [32,128]->80 @392x518 x7', 'conv1', (7, 392, 518, [32, 128], 80)),
    ('clbT 1x1 [32,128]->80 gelu + tail 80->4 softplus @392x518 x9', 'conv1t', (9, 392, 518, [32, 128], 80)),
    ('vit proj   M9333 K1024 N1024 gamma', 'ling', (9333, 1024, 1024)),
    ('vit fc1    M9333 K1024 N4096 gelu', 'linact', (9333, 1024, 4096)),
]
only = sys.argv[1:] 
res = []
for name, kind, p in SHAPES:
    if only and not any(o in name for o in only):
        continue
    if kind in ('ling', 'linact'):
        M, K, N = p
        x = torch.randn(M, K, device=dev).to(torch.bfloat16)
        w = torch.randn(N, K, device=dev) / K ** 0.5
        pw = ops.pack_weight(w, torch.randn(N, device=dev))
        if kind == 'ling':
            out = torch.zeros(M, N, dtype=torch.float32, device=dev)
            gam = torch.rand(N, device=dev)
            fn = lambda: ops.gemm(pw, [x], out, gamma=gam)
        else:
            out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
            fn = lambda: ops.gemm(pw, [x], out, act=ops.ACT_GELU)
        flops = 2.0 * M * K * N
    elif kind == 'conv1t':
        NB, H, W, cs, N = p
        srcs = [torch.randn(NB, H, W, c, device=dev).to(torch.bfloat16) for c in cs]
        w = torch.randn(N, sum(cs), 1, 1, device=dev) / sum(cs) ** 0.5
        pw = ops.pack_weight(w, torch.randn(N, device=dev), src_c=cs)
        w2, b2 = torch.randn(4, N, device=dev) / N ** 0.5, torch.randn(4, device=dev)
        out = torch.empty(NB, H, W, ops.pad_to(N, 8), dtype=torch.bfloat16, device=dev)
        pt = torch.empty(NB, H, W, 8, dtype=torch.float32, device=dev)
        fn = lambda: ops.gemm(pw, srcs, out, image=(NB, H, W), act=ops.ACT_GELU, tail=(w2, b2, ops.ACT_SOFTPLUS), tail_out=pt, skip_main=True)
        flops = 2.0 * NB * H * W * sum(cs) * N
    elif kind == 'lin':
        M, K, N = p
        x = torch.randn(M, K, device=dev).to(torch.bfloat16)
        w = torch.randn(N, K, device=dev) / K ** 0.5
        pw = ops.pack_weight(w, torch.randn(N, device=dev))
        out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        fn = lambda: ops.gemm(pw, [x], out)
        flops = 2.0 * M * K * N
    else:
        NB, H, W, cs, N = p
        taps = 9 if kind == 'conv' else 1
        srcs = [torch.randn(NB, H, W, c, device=dev).to(torch.bfloat16) for c in cs]
        k = 3 if taps == 9 else 1
        w = torch.randn(N, sum(cs), k, k, device=dev) / (taps * sum(cs)) ** 0.5
        pw = ops.pack_weight(w, torch.randn(N, device=dev), src_c=cs)
        out = torch.empty(NB, H, W, ops.pad_to(N, 8), dtype=torch.bfloat16, device=dev)
        fn = lambda: ops.gemm(pw, srcs, out, image=(NB, H, W), act=ops.ACT_RELU)
        flops = 2.0 * NB * H * W * sum(cs) * taps * N
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    REP = 10
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REP):
            fn()
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / REP)
    ms = sorted(ts)[len(ts) // 2]
    d = fn()
    print('%-46s %8.3f ms %8.1f TF/s  (block_n %d, m_tiles %d, n_tiles %d)' % (name, ms, flops / ms / 1e9, d.block_n, d.m_tiles, d.n_tiles), flush=True)
    res.append(dict(name=name, ms=ms, tflops=flops / ms / 1e9))

json.dump(res, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out', 'gemm_bench.json'), 'w'))
