"""Pins oracle/metrics_oracle.py against the reference's own `estimator/utils/metric.py` (imported unmodified through
oracle/shims, build container only) and writes tests/golden/metrics_case0.json.   python -m oracle.make_golden_metrics"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import metrics_oracle as mo      # noqa: E402
from oracle import ref_harness as rh         # noqa: E402

CASES = [dict(name='same_shape', gt=(270, 480), pred=(270, 480), seed=0, lo=1e-3, hi=80.0, edges=True),
         dict(name='resampled', gt=(270, 480), pred=(196, 259), seed=1, lo=1e-3, hi=80.0, edges=True),
         dict(name='clamped_no_edges', gt=(135, 240), pred=(135, 240), seed=2, lo=0.5, hi=10.0, edges=False)]


def case_tensors(c):
    g = torch.Generator().manual_seed(c['seed'])
    gt = torch.rand(1, 1, *c['gt'], generator=g) * 12 + 0.01
    gt[0, 0, :7, :9] = 0.0                                   # invalid (<= min) region
    pred = torch.rand(1, 1, *c['pred'], generator=g) * 14
    pred[0, 0, 3, 5] = float('nan')
    pred[0, 0, 4, 6] = float('inf')
    edges = (torch.rand(*c['gt'], generator=g) > 0.8) if c['edges'] else None
    return gt, pred, edges


def main():
    rh._enter()
    from estimator.utils.metric import compute_metrics as ref_metrics
    out = []
    for c in CASES:
        gt, pred, edges = case_tensors(c)
        r = ref_metrics(gt, pred.clone(), garg_crop=False, eigen_crop=False, dataset='', min_depth_eval=c['lo'],
                        max_depth_eval=c['hi'], disp_gt_edges=edges)
        o = mo.compute_metrics(gt, pred.clone(), c['lo'], c['hi'], edges)
        r = {k: float(v) for k, v in r.items()}
        for k in r:
            assert abs(r[k] - float(o[k])) <= 1e-6 * max(1.0, abs(r[k])), (c['name'], k, r[k], o[k])
        out.append(dict(case=c, reference=r))
        print(c['name'], r)
    json.dump(out, open(os.path.join(ROOT, 'tests', 'golden', 'metrics_case0.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
