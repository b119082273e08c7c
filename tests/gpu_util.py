"""Helpers for the -m gpu parity tests (torch fp32 references on bf16-rounded operands)."""
import torch

from patchfusion_b200 import ops


def bf(x):
    return x.to(torch.bfloat16)


def rb(x):
    """round-trip through bf16: the value the kernels actually consume"""
    return x.to(torch.bfloat16).float()


def to_nhwc(x, ld=None):
    """fp32 NCHW -> bf16 NHWC with the channel dim zero-padded to ld (default: multiple of 8)"""
    B, C, H, W = x.shape
    ld = ops.pad_to(C, 8) if ld is None else ld
    out = torch.zeros(B, H, W, ld, dtype=torch.bfloat16, device=x.device)
    out[..., :C] = x.permute(0, 2, 3, 1).to(torch.bfloat16)
    return out


def from_nhwc(y, C):
    return y[..., :C].float().permute(0, 3, 1, 2).contiguous()


def rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def check(name, got, want, tol):
    e = rel_err(got, want)
    ok = torch.isfinite(got.float()).all().item() and e < tol
    print('%-40s rel-Linf %.3e (tol %.1e) %s' % (name, e, tol, 'OK' if ok else 'FAIL'))
    assert ok, '%s: rel-Linf %.3e >= %.1e' % (name, e, tol)
