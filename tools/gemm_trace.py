"""Timeline of CTA 0 of pf_gemm_kernel on the ViT linear shapes (debug build:
   PF_B200_LIBNAME=libpf_b200_trace.so PF_B200_NVCC_EXTRA=-DPF_GEMM_TRACE python patchfusion_b200/build.py
   PF_B200_LIBNAME=libpf_b200_trace.so python tools/gemm_trace.py)"""
import ctypes as C
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from patchfusion_b200 import lib, ops
dev = torch.device('cuda:0')
h = lib.load()
rd = h.pf_gemm_trace_read
rd.argtypes = [C.c_void_p]
M = 9333
x1 = torch.randn(M, 1024, device=dev).to(torch.bfloat16)
x4 = torch.randn(M, 4096, device=dev).to(torch.bfloat16)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)


def case(name):
    if name == 'clbT':      # metric-head CLB: 1x1 [32, 8, 128] -> 80 + GELU with the 80 -> 4 + Softplus layer fused (9 tiles of 392x518)
        NB, H, W, cs, N = 9, 392, 518, [32, 8, 128], 80
        srcs = [torch.randn(NB, H, W, c, device=dev).to(torch.bfloat16) for c in cs]
        w = torch.randn(N, sum(cs), 1, 1, device=dev) / sum(cs) ** 0.5
        pw = ops.pack_weight(w, torch.randn(N, device=dev), src_c=cs)
        w2, b2 = torch.randn(4, N, device=dev) / N ** 0.5, torch.randn(4, device=dev)
        out = torch.empty(NB, H, W, ops.pad_to(N, 8), dtype=torch.bfloat16, device=dev)
        pt = torch.empty(NB, H, W, 8, dtype=torch.float32, device=dev)
        return lambda: ops.gemm(pw, srcs, out, image=(NB, H, W), act=ops.ACT_GELU, tail=(w2, b2, ops.ACT_SOFTPLUS),
                                tail_out=pt, skip_main=True)
    K, N = dict(proj=(1024, 1024), fc2=(4096, 1024), fc1=(1024, 4096), qkv=(1024, 3072))[name]
    w = torch.randn(N, K, device=dev) / K ** 0.5
    pw = ops.pack_weight(w, torch.randn(N, device=dev))
    x = x1 if K == 1024 else x4
    if name in ('proj', 'fc2'):
        out = torch.zeros(M, N, dtype=torch.float32, device=dev)
        gam = torch.rand(N, device=dev)
        return lambda: ops.gemm(pw, [x], out, gamma=gam)
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    return lambda: ops.gemm(pw, [x], out, act=ops.ACT_GELU if name == 'fc1' else ops.ACT_NONE)


roles = ['tma', 'mma', 'epi-w0', 'epi-w4']
evs = {0: {9: 'producer start', 0: 'tile begin'}, 1: {0: 'tile begin', 1: 'acc free', 2: 'first stage full', 3: 'tile committed'},
       2: {0: 'wait acc', 1: 'acc ready', 2: 'tile drained'}, 3: {0: 'wait acc', 1: 'acc ready', 2: 'tile drained'}}
for name in sys.argv[1:] or ['proj', 'fc2', 'fc1', 'qkv']:
    fn = case(name)
    for cold in (0, 1):
        fn(); fn()
        if cold:
            flush.zero_()
        torch.cuda.synchronize()
        h.pf_gemm_trace_clear()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        buf = (C.c_ulonglong * (4 * 256 * 2))()
        rd(buf)
        recs = []
        for i in range(4 * 256):
            k, clk = buf[2 * i], buf[2 * i + 1]
            if k >> 63:
                recs.append((clk, (k >> 48) & 0x7fff, (k >> 16) & 0xffffffff, k & 0xffff))
        recs.sort()
        print('==== %s  M%d  %s   %.1f us' % (name, M, 'L2 flushed' if cold else 'warm', e0.elapsed_time(e1) * 1e3))
        t0 = recs[0][0]
        for clk, role, t, ev in recs[:120]:
            print('%8d  %-7s tile %4d  %s' % (clk - t0, roles[role], t, evs[role].get(ev, ev)))
