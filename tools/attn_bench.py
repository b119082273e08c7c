import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from patchfusion_b200 import ops
dev = torch.device('cuda:0')
B, seq, heads = 9, 1037, 16
D = heads * 64
seq_pad = ops.pad_to(seq, 8)
qk = torch.randn(B * seq, 2 * D, device=dev).to(torch.bfloat16)
vt = torch.randn(B * D, seq_pad, device=dev).to(torch.bfloat16)
out = torch.zeros(B * seq, D, dtype=torch.bfloat16, device=dev)
for _ in range(3):
    ops.attention(qk, vt, B, seq, seq_pad, heads, 0.125, out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.attention(qk, vt, B, seq, seq_pad, heads, 0.125, out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print('attention B%d: %.3f ms  %.1f TF/s' % (B, ms, 4.0 * B * heads * seq * seq * 64 / ms / 1e9))
