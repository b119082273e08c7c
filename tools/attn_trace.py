"""Timeline of one CTA of pf_attention_kernel (debug build: PF_B200_LIBNAME=libpf_b200_trace.so built with
-DPF_ATTN_TRACE): prints (clock, role, tile, kv block, event) for the first work items."""
import ctypes as C
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from patchfusion_b200 import lib, ops
dev = torch.device('cuda:0')
B, seq, heads = 9, 1037, 16
D = heads * 64
seq_pad = ops.pad_to(seq, 8)
qk = torch.randn(B * seq, 2 * D, device=dev).to(torch.bfloat16)
vt = torch.randn(B * D, seq_pad, device=dev).to(torch.bfloat16)
out = torch.zeros(B * seq, D, dtype=torch.bfloat16, device=dev)
ops.attention(qk, vt, B, seq, seq_pad, heads, 0.125, out)       # the trace keeps the first 1024 events per role
torch.cuda.synchronize()
h = lib.load()
buf = (C.c_ulonglong * (4096 * 2))()
n = C.c_uint()
get = h.pf_attention_trace_read
get.argtypes = [C.c_void_p, C.POINTER(C.c_uint)]
get(buf, C.byref(n))
recs = []
for i in range(min(n.value, 4096)):
    k, clk = buf[2 * i], buf[2 * i + 1]
    if not (k >> 63):
        continue
    k &= (1 << 63) - 1
    recs.append((clk, k >> 48, (k >> 32) & 0xffff, (k >> 16) & 0xffff, k & 0xffff))
recs.sort()
t0 = recs[0][0]
names = {0: ['S ready', 'scores in regs', 'exps done', 'PV(j-1) seen', 'P published'], 1: ['QK issue'], 2: ['PV issue']}
for clk, role, t, j, ev in recs[:260]:
    print('%8d  %-8s tile %d  blk %2d  %s' % (clk - t0, ['softmax', 'qk-warp', 'pv-warp'][role], t, j, names[role][ev]))
