"""CPU oracle for the depth metrics (TEST INFRASTRUCTURE, not product): numpy restatement of the reference's
`compute_errors` (estimator/utils/metric.py:10-50), `shift_2d_replace` / `soft_edge_error` (:53-72) and
`compute_metrics` (:97-148) without its cv2 / matplotlib / skimage / kornia imports.  Pinned against the real
reference by oracle/make_golden.py (tests/golden/metrics_case0.json)."""
import numpy as np
import torch
import torch.nn.functional as F


def compute_errors(gt, pred):
    thresh = np.maximum((gt / pred), (pred / gt))
    a1 = (thresh < 1.25).mean()
    a2 = (thresh < 1.25 ** 2).mean()
    a3 = (thresh < 1.25 ** 3).mean()
    abs_rel = np.mean(np.abs(gt - pred) / gt)
    sq_rel = np.mean(((gt - pred) ** 2) / gt)
    rmse = np.sqrt(((gt - pred) ** 2).mean())
    rmse_log = np.sqrt(((np.log(gt) - np.log(pred)) ** 2).mean())
    err = np.log(pred) - np.log(gt)
    silog = np.sqrt(np.mean(err ** 2) - np.mean(err) ** 2) * 100
    log_10 = (np.abs(np.log10(gt) - np.log10(pred))).mean()
    return dict(a1=a1, a2=a2, a3=a3, abs_rel=abs_rel, rmse=rmse, log_10=log_10, rmse_log=rmse_log, silog=silog,
                sq_rel=sq_rel)


def _shift(data, dx, dy):
    out = np.roll(data, dx, axis=1)
    if dx < 0:
        out[:, dx:] = 0
    elif dx > 0:
        out[:, 0:dx] = 0
    out = np.roll(out, dy, axis=0)
    if dy < 0:
        out[dy:, :] = 0
    elif dy > 0:
        out[0:dy, :] = 0
    return out


def soft_edge_error(pred, gt, radius=1):
    return np.minimum.reduce([np.abs(_shift(gt, i, j) - pred) for i in range(-radius, radius + 1)
                              for j in range(-radius, radius + 1)])


def compute_metrics(gt, pred, min_depth_eval=0.1, max_depth_eval=10, disp_gt_edges=None, additional_mask=None):
    """metric.py:97-148 with garg_crop = eigen_crop = False (the U4K call, u4k_dataset.py:186)."""
    if gt.shape[-2:] != pred.shape[-2:]:
        pred = F.interpolate(pred, gt.shape[-2:], mode='bilinear', align_corners=False).squeeze()
    pred = pred.squeeze().cpu().numpy().copy()
    pred[pred < min_depth_eval] = min_depth_eval
    pred[pred > max_depth_eval] = max_depth_eval
    pred[np.isinf(pred)] = max_depth_eval
    pred[np.isnan(pred)] = min_depth_eval
    g = gt.squeeze().cpu().numpy()
    valid = np.logical_and(g > min_depth_eval, g < max_depth_eval)
    if additional_mask is not None:
        valid = np.logical_and(valid, additional_mask.squeeze().cpu().numpy())
    m = compute_errors(g[valid], pred[valid])
    if disp_gt_edges is not None:
        mask = np.logical_and(valid, disp_gt_edges.squeeze().numpy())
        m['see'] = soft_edge_error(pred, g)[mask].mean() if mask.sum() > 0 else 0.0
    return m
