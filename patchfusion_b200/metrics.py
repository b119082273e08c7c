"""Depth metrics on the GPU: the reference's `compute_metrics` (`estimator/utils/metric.py:97-148`,
called per image from `estimator/tester/tester.py:78-86` through `u4k_dataset.py:185-186`) as ONE fused reduction
(`pf_depth_metrics`) instead of ~25 full-image numpy passes after a device->host copy of the 8.3 MP prediction."""
import ctypes as C
import math

import torch

from . import lib

KEYS = ('a1', 'a2', 'a3', 'abs_rel', 'rmse', 'log_10', 'rmse_log', 'silog', 'sq_rel')
_NBLOCKS = 148 * 8


def compute_metrics(gt, pred, interpolate=True, garg_crop=False, eigen_crop=False, dataset='', min_depth_eval=0.1,
                    max_depth_eval=10, disp_gt_edges=None, additional_mask=None):
    """Same signature and result dict as the reference (tensors stay on the GPU; one 96-byte read back).
    The eigen/garg crops are static index boxes: pass them as `additional_mask` (U4K evaluation uses neither,
    u4k_dataset.py:186)."""
    if garg_crop or eigen_crop:
        raise NotImplementedError('garg/eigen crops: pass the crop as additional_mask')
    assert gt.is_cuda and pred.is_cuda, 'libpf_b200 takes device pointers only'
    g = gt.squeeze().float().contiguous()
    p = pred.squeeze().float().contiguous()
    H, W = g.shape
    if tuple(p.shape) != (H, W) and not interpolate:
        raise ValueError('pred and gt shapes differ and interpolate=False')
    dev = g.device
    e = disp_gt_edges.squeeze().to(dev).ne(0).to(torch.uint8).contiguous() if disp_gt_edges is not None else None
    m = additional_mask.squeeze().to(dev).ne(0).to(torch.uint8).contiguous() if additional_mask is not None else None
    part = torch.empty((_NBLOCKS, 12), dtype=torch.float64, device=dev)
    out = torch.empty(12, dtype=torch.float64, device=dev)
    lib.call('pf_depth_metrics', p, p.shape[0], p.shape[1], g, H, W, C.c_float(min_depth_eval), C.c_float(max_depth_eval),
             e, m, part, _NBLOCKS, out, lib.stream_ptr())
    return finalize(out.cpu().tolist(), with_see=disp_gt_edges is not None)


def finalize(s, with_see=False):
    """sums -> the reference's metric dict (metric.py:31-50)."""
    n = s[0]
    if n == 0:
        r = {k: float('nan') for k in KEYS}
    else:
        mean_err = s[8] / n
        r = dict(a1=s[1] / n, a2=s[2] / n, a3=s[3] / n, abs_rel=s[4] / n, rmse=math.sqrt(s[6] / n), log_10=s[9] / n,
                 rmse_log=math.sqrt(s[7] / n), silog=math.sqrt(max(s[7] / n - mean_err * mean_err, 0.0)) * 100,
                 sq_rel=s[5] / n)
    if with_see:
        r['see'] = s[11] / s[10] if s[10] > 0 else 0.0
    return r
