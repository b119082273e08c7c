"""compute-sanitizer target: the CUDA path only (no oracle), small and short.
   compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py model
   compute-sanitizer --tool racecheck python tools/sanitize_smoke.py kernels
'kernels': one launch each of the tcgen05 GEMM (plain, multicast pair, fused tail), the halo conv, attention and the
window attention on tiny shapes; 'model': vits, 1080p, 2x2 split, m1 through the drop-in class (graphs off)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['PF_B200_GRAPHS'] = '0'
import torch
from patchfusion_b200 import ops

dev = torch.device('cuda:0')
what = sys.argv[1] if len(sys.argv) > 1 else 'kernels'
bf = torch.bfloat16
g = torch.Generator(device=dev).manual_seed(0)
if what == 'kernels':
    x = torch.randn(300, 192, device=dev, generator=g).to(bf)
    pw = ops.pack_weight(torch.randn(160, 192, device=dev, generator=g) / 14, torch.randn(160, device=dev, generator=g))
    out = torch.zeros(300, 160, dtype=bf, device=dev)
    ops.gemm(pw, [x], out, act=ops.ACT_GELU)
    xm = torch.randn(640, 128, device=dev, generator=g).to(bf)                      # 5 m-tiles x 37 n-tiles: multicast pairs
    pm = ops.pack_weight(torch.randn(9472, 128, device=dev, generator=g) / 11, None)
    om = torch.zeros(640, 9472, dtype=bf, device=dev)
    ops.gemm(pm, [xm], om)
    xi = torch.randn(1, 40, 36, 64, device=dev, generator=g).to(bf)
    pc = ops.pack_weight(torch.randn(32, 64, 3, 3, device=dev, generator=g) / 24, torch.randn(32, device=dev, generator=g))
    oc = torch.zeros(1, 40, 36, 32, dtype=bf, device=dev)
    w2, b2 = torch.randn(4, 32, device=dev, generator=g), torch.randn(4, device=dev, generator=g)
    pt = torch.zeros(1, 40, 36, 8, dtype=torch.float32, device=dev)
    ops.gemm(pc, [xi], oc, image=(1, 40, 36), act=ops.ACT_RELU, tail=(w2, b2, ops.ACT_SOFTPLUS), tail_out=pt)
    B, seq, heads = 1, 300, 2
    D, sp = heads * 64, ops.pad_to(seq, 8)
    qk = torch.randn(B * seq, 2 * D, device=dev, generator=g).to(bf)
    vt = torch.zeros(B * D, sp, dtype=bf, device=dev)
    vt[:, :seq] = torch.randn(B * D, seq, device=dev, generator=g).to(bf)
    oa = torch.zeros(B * seq, D, dtype=bf, device=dev)
    ops.attention(qk, vt, B, seq, sp, heads, 0.125, oa)
    qkv = torch.randn(24 * 24, 3 * 64, device=dev, generator=g).to(bf)
    ow = torch.zeros(24 * 24, 64, dtype=bf, device=dev)
    ops.call('pf_window_attention', qkv, torch.randn(529, 8, device=dev, generator=g), 24, 24, 64, 8, 6, ow, ops.stream_ptr())
    # round 2b: fp32 residual stream through the bulk reduce-add epilogue, bf16 bulk-store epilogue on an NHWC 1x1 conv,
    # and a >= 148-tile 3x3 conv (weight-multicast CTA pairs of the halo kernel, odd m-tile count -> one phantom tile)
    xs = torch.zeros(300, 256, dtype=torch.float32, device=dev)
    pg = ops.pack_weight(torch.randn(256, 192, device=dev, generator=g) / 14, torch.randn(256, device=dev, generator=g))
    ops.gemm(pg, [x], xs, gamma=torch.rand(256, device=dev, generator=g))
    p1 = ops.pack_weight(torch.randn(128, 64, 1, 1, device=dev, generator=g) / 8, torch.randn(128, device=dev, generator=g))
    o1 = torch.zeros(1, 40, 36, 128, dtype=bf, device=dev)
    ops.gemm(p1, [xi], o1, image=(1, 40, 36), act=ops.ACT_RELU)
    xc = torch.randn(7, 48, 72, 40, device=dev, generator=g).to(bf)
    pcm = ops.pack_weight(torch.randn(64, 40, 3, 3, device=dev, generator=g) / 19, torch.randn(64, device=dev, generator=g))
    ocm = torch.zeros(7, 48, 72, 64, dtype=bf, device=dev)
    ops.gemm(pcm, [xc], ocm, image=(7, 48, 72), act=ops.ACT_RELU)
    torch.cuda.synchronize()
    print('kernels ok', float(out.float().abs().mean()), float(om.float().abs().mean()), float(oc.float().abs().mean()),
          float(oa.float().abs().mean()), float(ow.float().abs().mean()), float(xs.abs().mean()),
          float(o1.float().abs().mean()), float(ocm.float().abs().mean()))
else:
    from patchfusion_b200.configs import depth_anything_patchfusion
    from patchfusion_b200.model import PatchFusion
    cfg = depth_anything_patchfusion('vits', image_raw_shape=(1080, 1920), patch_split_num=(2, 2))
    model = PatchFusion(cfg).init_synthetic_weights(0).to(dev).eval()
    img = torch.rand(1, 3, 1080, 1920, generator=torch.Generator().manual_seed(0)).to(dev)
    y, _ = model(mode='infer', image_lr=model.make_lr(img), image_hr=img, cai_mode='m1', process_num=2)
    torch.cuda.synchronize()
    print('model ok', tuple(y.shape), float(y.mean()))
