"""GPU versions of the two callers either side of the hot path: image ingest
(`estimator/datasets/general_dataset.py:22-47,188-219`) and the uint16 depth writer
(`estimator/tester/tester.py:66-76`, `tools/test_single_forward.py:26`)."""
import ctypes as C

import numpy as np
import torch

from . import ops


def ingest(image_u8, image_raw_shape, device, bgr=True):
    """image_u8: HxWx3 uint8 (numpy array as returned by cv2.imread, or a torch tensor).  Returns image_hr
    (1,3,H',W') fp32 in [0,1] on `device`, bicubic-resized (align_corners=True) to image_raw_shape."""
    if isinstance(image_u8, np.ndarray):
        image_u8 = torch.from_numpy(np.ascontiguousarray(image_u8))
    assert image_u8.dtype == torch.uint8 and image_u8.dim() == 3 and image_u8.shape[2] == 3
    src = image_u8.to(device, non_blocking=True).contiguous()
    H, W = src.shape[:2]
    OH, OW = image_raw_shape
    out = torch.empty((1, 3, OH, OW), dtype=torch.float32, device=device)
    ops.call('pf_ingest_u8', src, H, W, 1 if bgr else 0, OH, OW, out, ops.stream_ptr())
    return out


def depth_to_u16(depth, size=None, scale=256.0):
    """depth (1,1,h,w) or (h,w) fp32 on the GPU -> uint16 tensor (H,W): nearest resize to `size` then *scale."""
    d = depth.reshape(depth.shape[-2:]).float().contiguous()
    H, W = d.shape
    OH, OW = (H, W) if size is None else size
    out = torch.empty((OH, OW), dtype=torch.uint16, device=d.device)
    ops.call('pf_depth_to_u16', d, H, W, OH, OW, C.c_float(scale), out, ops.stream_ptr())
    return out
