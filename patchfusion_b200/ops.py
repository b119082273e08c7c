"""Thin Python wrappers over the C ABI: weight packing and one function per kernel family.

Tensors are torch CUDA tensors used as typed device buffers: activations bf16 NHWC (`[B, H, W, ld]`, logical
channel count tracked by the caller), token matrices `[rows, ld]`, fp32 where include/pf_b200.h says so.
"""
import ctypes as C

import torch

from . import lib
from .lib import ACT_GELU, ACT_NONE, ACT_RELU, ACT_SOFTPLUS, GemmDesc, call, stream_ptr  # noqa: F401


def pad_to(n, m):
    return (n + m - 1) // m * m


class PackedWeight:
    """bf16 K-major weight panel for pf_gemm + fp32 bias."""

    def __init__(self, w, bias, N, src_c, taps, ps=1, ps_cout=0):
        self.w, self.bias, self.N, self.src_c, self.taps = w, bias, N, list(src_c), taps
        self.Ktot = w.shape[1]
        self.ps, self.ps_cout = ps, ps_cout


def n_pad_for(N):
    """rows of the packed panel: what pf_gemm's automatic block_n choice will read (multiple of block_n)."""
    n32 = pad_to(N, 32)
    if n32 <= 256:
        return n32
    best, bn = None, None
    for c in range(256, 127, -32):
        pad = pad_to(N, c) - N
        if best is None or pad < best:
            best, bn = pad, c
    return pad_to(N, bn)


def pack_weight(weight, bias=None, src_c=None, scale=None, shift=None):
    """weight: fp32 [N, C, kh, kw] or [N, C] on the GPU.  `scale/shift` fold an eval-mode BatchNorm:
    y = scale * conv(x) + shift."""
    weight = weight.float().contiguous()
    N = weight.shape[0]
    taps = 1 if weight.dim() == 2 else weight.shape[2] * weight.shape[3]
    assert taps in (1, 9)
    ctot = weight.shape[1]
    src_c = [ctot] if src_c is None else list(src_c)
    assert sum(src_c) == ctot and 1 <= len(src_c) <= 3
    Ktot = sum(taps * pad_to(c, 64) for c in src_c)
    n_pad = n_pad_for(N)
    dst = torch.empty((n_pad, Ktot), dtype=torch.bfloat16, device=weight.device)
    sc = (C.c_int32 * 3)(*(src_c + [0] * (3 - len(src_c))))
    scale_t = scale.float().contiguous() if scale is not None else None
    call('pf_pack_weight', weight, N, n_pad, len(src_c), sc, taps, scale_t, dst, stream_ptr())
    b = None
    if bias is not None or shift is not None:
        b = torch.zeros(N, dtype=torch.float32, device=weight.device)
        if bias is not None:
            b += bias.float() * (scale.float() if scale is not None else 1.0)
        if shift is not None:
            b += shift.float()
    return PackedWeight(dst, b, N, src_c, taps)


def pack_weight_convT(weight, bias, k):
    """ConvTranspose2d(kernel == stride) weight [Cin, Cout, k, k]."""
    weight = weight.float().contiguous()
    cin, cout = weight.shape[0], weight.shape[1]
    cp = pad_to(cout, 32)
    dst = torch.empty((k * k * cp, pad_to(cin, 64)), dtype=torch.bfloat16, device=weight.device)
    call('pf_pack_weight_convT', weight, cin, cout, k, dst, stream_ptr())
    return PackedWeight(dst, bias.float().contiguous() if bias is not None else None, k * k * cp, [cin], 1, ps=k,
                        ps_cout=cout)


def gemm(pw, srcs, out, *, image=None, ps_image=None, M=None, act=ACT_NONE, res1=None, res2=None, gamma=None, out_col0=0, out2=None,
         vt=None, vt_col0=0, vt_seq=0, vt_seq_pad=0, src_c=None, block_n=0, tail=None, tail_out=None,
         skip_main=False, tile=None, resample=None):
    """srcs: list of bf16 tensors.  image=(NB,H,W) selects NHWC/conv addressing (srcs are [NB,H,W,ld]);
    otherwise srcs are [M, ld] matrices.  `out` may be bf16 or fp32; with `gamma` it is the fp32 residual stream
    updated in place (x += gamma * (acc + bias)).  resample[i] = True: source i is a [NB, h, w, ld] map of another size
    that the 3x3 conv reads through a fused bilinear (align_corners=True) resample to (H, W)."""
    d = GemmDesc()
    d.num_src = len(srcs)
    d.taps = pw.taps
    cs = pw.src_c if src_c is None else src_c
    assert len(cs) == len(srcs)
    for i, s in enumerate(srcs):
        assert s.dtype == torch.bfloat16 and s.is_contiguous()
        d.a_ptr[i] = s.data_ptr()
        d.a_c[i] = pad_to(cs[i], 8)
        d.a_ld[i] = s.shape[-1]
        assert d.a_c[i] <= s.shape[-1]
        if resample is not None and resample[i]:
            d.rs_h[i], d.rs_w[i] = s.shape[1], s.shape[2]
    if image is not None:
        d.a_mode = 1
        d.NB, d.H, d.W = image
        rows = d.NB * d.H * d.W
        if tile is not None:        # pin the pixel tile (bh, bw): selects the per-tap TMA kernel for 3x3 convs
            d.bh, d.bw = tile
    else:
        d.a_mode = 0
        d.M = M if M is not None else srcs[0].shape[0]
        rows = d.M
        if pw.ps > 1:
            d.NB, d.H, d.W = ps_image
    d.w_ptr = pw.w.data_ptr()
    d.N = pw.N
    d.Ktot = pw.Ktot
    d.block_n = block_n
    d.bias = pw.bias.data_ptr() if pw.bias is not None else None
    d.act = act
    if res1 is not None:
        d.res1 = res1.data_ptr()
        d.res_ld = res1.shape[-1]
    if res2 is not None:
        d.res2 = res2.data_ptr()
        assert res2.shape[-1] == d.res_ld
    if gamma is not None:
        assert out.dtype == torch.float32
        d.gamma = gamma.data_ptr()
    d.out = out.data_ptr()
    d.out_f32 = 1 if out.dtype == torch.float32 else 0
    d.out_ld = out.shape[-1]
    d.out_col0 = out_col0
    if out2 is not None:
        d.out2 = out2.data_ptr()
        d.out2_ld = out2.shape[-1]
    d.ps, d.ps_cout = pw.ps, pw.ps_cout
    if vt is not None:
        d.vt = vt.data_ptr()
        d.vt_col0, d.vt_seq, d.vt_seq_pad, d.vt_dim = vt_col0, vt_seq, vt_seq_pad, pw.N - vt_col0
    if tail is not None:
        # fused trailing 1x1 layer: tail = (w2 fp32 [n2, N], b2 fp32 [n2] or None, act2)
        w2, b2, act2 = tail
        assert w2.dtype == torch.float32 and w2.is_contiguous() and w2.shape[1] == pw.N and tail_out.dtype == torch.float32
        d.w2, d.b2 = w2.data_ptr(), (b2.data_ptr() if b2 is not None else None)
        d.n2, d.act2, d.skip_main = w2.shape[0], act2, 1 if skip_main else 0
        d.out3, d.out3_ld = tail_out.data_ptr(), tail_out.shape[-1]
    call('pf_gemm', C.byref(d), stream_ptr())
    return d


def gemm_convT(pw, src, image, out):
    """ConvTranspose k==s: src [NB*H*W, ld] rows in (n,y,x) order, out NHWC [NB, H*k, W*k, ld_out]."""
    return gemm(pw, [src], out, ps_image=image, M=image[0] * image[1] * image[2])


def layernorm(x, w, b, eps, out, rows=None, C_=None):
    rows = x.shape[0] if rows is None else rows
    C_ = w.shape[0] if C_ is None else C_
    call('pf_layernorm', x, x.shape[-1], w, b, C.c_float(eps), rows, C_, out, out.shape[-1], stream_ptr())


def attention(qk, vt, B, seq, seq_pad, heads, scale, out):
    call('pf_attention', qk, qk.shape[-1], vt, B, seq, seq_pad, heads, C.c_float(scale), out, out.shape[-1],
         stream_ptr())


def resize_bilinear(x, C_, OH, OW, out, out_col0=0):
    B, H, W, ld = x.shape
    call('pf_resize_bilinear', x, B, H, W, pad_to(C_, 8), ld, OH, OW, out, out.shape[-1], out_col0, stream_ptr())


def roi_crop_zoom(feat, C_, boxes, scale, out, out_col0=0):
    """feat [1,h,w,ld] bf16 or [h,w] fp32 (depth); boxes [T,4] fp32 device."""
    T = boxes.shape[0]
    if feat.dtype == torch.float32:
        h, w = feat.shape[-2:]
        call('pf_roi_crop_zoom', feat, 1, h, w, 1, 1, boxes, T, C.c_float(scale), out, 1, 0, stream_ptr())
    else:
        _, h, w, ld = feat.shape
        call('pf_roi_crop_zoom', feat, 0, h, w, pad_to(C_, 8), ld, boxes, T, C.c_float(scale), out, out.shape[-1],
             out_col0, stream_ptr())


def maxpool2(x, C_, out):
    B, H, W, ld = x.shape
    call('pf_maxpool2', x, B, H, W, pad_to(C_, 8), ld, out, out.shape[-1], stream_ptr())
