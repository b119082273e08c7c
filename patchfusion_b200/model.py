"""Drop-in `PatchFusion` for the reference's `estimator.models.patchfusion.PatchFusion` (inference path).

Same constructor dict / `from_pretrained()` / `state_dict()` key layout / `forward(mode='infer', image_lr,
image_hr, tile_cfg, cai_mode, process_num)` contract and error behaviour as reference
`estimator/models/patchfusion.py:55-453` + `estimator/models/baseline_pretrain.py:91-331`; the arithmetic runs on
libpf_b200 (sm_100a) through `Engine`.  There is no CPU path: calling forward on CPU tensors raises.
"""
import math
import os
import random

import numpy as np
import torch
import torch.nn as nn

from .params import ParamTree, _get, state_layout, synthetic_state_dict, relative_position_index

try:
    from huggingface_hub import PyTorchModelHubMixin
except Exception:                                   # pragma: no cover
    class PyTorchModelHubMixin:                     # minimal stand-in when huggingface_hub is absent
        pass


class AttrDict(dict):
    """Attribute-style view of nested config dicts (the reference reads `config.coarse_branch.type`)."""

    def __init__(self, d=None):
        super().__init__()
        for k, v in (d or {}).items():
            self[k] = AttrDict(v) if isinstance(v, dict) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, AttrDict) else v) for k, v in self.items()}

    def to_json_file(self, path):                   # `tools/convert_huggingface.py:79`
        import json
        with open(path, 'w') as f:
            json.dump(self.to_dict(), f, indent=2)


class Resize:
    """`model.resizer`: bilinear align_corners=True to patch_process_shape (depth_anything/transform.py:127-129 with
    keep_aspect_ratio=False => always exactly (h, w), as used at patchfusion.py:94)."""

    def __init__(self, width, height):
        self.w, self.h = width, height

    def __call__(self, x):
        return nn.functional.interpolate(x, (int(self.h), int(self.w)), mode='bilinear', align_corners=True)


def _gauss_kernel(ksize, sigma):
    i = np.arange(ksize, dtype=np.float64) - (ksize - 1) / 2
    k = np.exp(-(i * i) / (2.0 * sigma * sigma))
    return (k / k.sum()).astype(np.float32)


def generatemask(size):
    """Gaussian blend mask of `estimator/models/utils.py:38-47` without OpenCV: 1 inside the central 80 %, separable
    Gaussian blur (k = 2*ceil(2*sigma)+1, sigma = int(H/16), BORDER_REFLECT_101), min-max normalised."""
    H, W = int(size[0]), int(size[1])
    mask = np.zeros((H, W), dtype=np.float32)
    sigma = int(H / 16)
    k = int(2 * np.ceil(2 * int(H / 16)) + 1)
    mask[int(0.1 * H):H - int(0.1 * H), int(0.1 * W):W - int(0.1 * W)] = 1
    ker = _gauss_kernel(k, float(sigma))
    r = k // 2

    def blur_rows(a):       # convolve along axis 1
        p = np.pad(a, ((0, 0), (r, r)), mode='reflect')
        cs = np.zeros_like(a, dtype=np.float32)
        for j in range(k):
            cs += ker[j] * p[:, j:j + a.shape[1]]
        return cs

    m = blur_rows(mask)
    m = blur_rows(m.T.copy()).T
    m = (m - m.min()) / (m.max() - m.min())
    return np.ascontiguousarray(m, dtype=np.float32)


class PatchFusion(ParamTree, PyTorchModelHubMixin):
    def __init__(self, config):
        nn.Module.__init__(self)
        if not isinstance(config, dict) and hasattr(config, 'to_dict'):
            config = config.to_dict()
        config = AttrDict(dict(config))
        # the HF-dict constructor branch of the reference (patchfusion.py:70-78)
        config.load_branch = bool(config.get('load_branch', False))
        self.config = config
        self.min_depth = config.min_depth
        self.max_depth = config.max_depth
        self.patch_process_shape = config.patch_process_shape
        self.tile_cfg = self.prepare_tile_cfg(config.image_raw_shape, config.patch_split_num)
        self.coarse_branch_cfg = config.coarse_branch
        for br in (config.coarse_branch, config.fine_branch):
            if br.type not in ('ZoeDepth', 'DA-ZoeDepth'):
                raise NotImplementedError
        self.resizer = Resize(config.patch_process_shape[1], config.patch_process_shape[0])
        self._build_tree(state_layout(config))      # raises ValueError / NotImplementedError like the reference
        self.consistency_training = False
        self._engine = None
        self._graphs = {}
        self.graph_launches = 0
        self._coarse = None
        self.use_cuda_graphs = os.environ.get('PF_B200_GRAPHS', '1') != '0'
        self.partition = os.environ.get('PF_B200_PARTITION', 'greedy')
        self.overlap = os.environ.get('PF_B200_OVERLAP', '0') != '0'    # opt-in: +1 % on one GPU, see DESIGN.md
        self._side_stream = None
        self._fine_out = {}
        self._mask_cache = {}
        if config.load_branch:
            for which, path in zip(('coarse_branch', 'fine_branch'), config.pretrain_model):
                sd = torch.load(path, map_location='cpu')['model_state_dict']
                getattr(self, which).load_state_dict(sd, strict=True)
        # constant buffers exactly as the reference constructs them
        with torch.no_grad():
            for name, buf in self.named_buffers():
                if name.endswith('relative_position_index'):
                    buf.copy_(relative_position_index())
                elif name.endswith('k_idx'):
                    buf.copy_(torch.arange(buf.numel()).view(buf.shape))
                elif name.endswith('K_minus_1'):
                    buf.fill_(float(_get(config.coarse_branch, 'n_bins', 64) - 1))
                elif name.endswith('running_var'):
                    buf.fill_(1.0)

    # ------------------------------------------------------------------ reference API surface
    def prepare_tile_cfg(self, image_raw_shape, patch_split_num):
        assert image_raw_shape[0] % (2 * patch_split_num[0]) == 0, \
            'image height should be divisible by 2 * patch_split_num[0]'
        assert image_raw_shape[1] % (2 * patch_split_num[1]) == 0, \
            'image width should be divisible by 2 * patch_split_num[1]'
        pps = self.patch_process_shape
        patch_reensemble_shape = (pps[0] * patch_split_num[0], pps[1] * patch_split_num[1])
        patch_raw_shape = (image_raw_shape[0] // patch_split_num[0], image_raw_shape[1] // patch_split_num[1])
        return {'patch_split_num': patch_split_num, 'patch_reensemble_shape': patch_reensemble_shape,
                'patch_raw_shape': patch_raw_shape, 'image_raw_shape': image_raw_shape,
                'raw_h_split_point': [int(patch_raw_shape[0] * i) for i in range(patch_split_num[0])],
                'raw_w_split_point': [int(patch_raw_shape[1] * i) for i in range(patch_split_num[1])]}

    def load_dict(self, dict):
        return self.load_state_dict(dict, strict=False)

    def get_save_dict(self):
        return {k: v for k, v in self.state_dict().items() if 'coarse_branch' not in k and 'fine_branch' not in k}

    def load_state_dict(self, *a, **kw):
        self._engine, self._graphs = None, {}
        return super().load_state_dict(*a, **kw)

    def _apply(self, fn, *a, **kw):
        self._engine, self._graphs = None, {}
        return super()._apply(fn, *a, **kw)

    def init_synthetic_weights(self, seed=0):
        """Seeded random weights (no checkpoints are reachable offline)."""
        self.load_state_dict(synthetic_state_dict(self.config, seed=seed), strict=True)
        return self

    # ------------------------------------------------------------------ engine plumbing
    def engine(self, device=None):
        from .engine import Engine
        if self._engine is None:
            p = next(self.parameters())
            if not p.is_cuda:
                raise RuntimeError('PatchFusion (B200): the hot path runs on libpf_b200 only; move the model to a '
                                   'CUDA device (there is no CPU fallback)')
            self._engine = Engine(self.config, self.state_dict(), p.device)
        return self._engine

    @torch.no_grad()
    def make_lr(self, image_hr):
        """`image_lr = model.resizer(image)` of `tools/test_single_forward.py:13-14` on the device (same bilinear
        align_corners=True resample, pf_crop_resize with one whole-image tile)."""
        from . import ops
        img = image_hr[0].float().contiguous()
        H, W = img.shape[-2:]
        ph, pw = self.patch_process_shape
        out = torch.empty((1, 3, ph, pw), dtype=torch.float32, device=img.device)
        org = torch.zeros((1, 2), dtype=torch.int32, device=img.device)
        ops.call('pf_crop_resize', img, H, W, org, 1, H, W, ph, pw, out, ops.stream_ptr())
        return out

    def _mask(self, size, device):
        key = (tuple(size), str(device))
        if key not in self._mask_cache:
            self._mask_cache[key] = torch.tensor(generatemask(size) + 1e-3, device=device).contiguous()
        return self._mask_cache[key]

    # ------------------------------------------------------------------ stage-level entry points (NCHW fp32 views)
    @staticmethod
    def _nchw(m):
        return m.t[..., :m.C].float().permute(0, 3, 1, 2).contiguous()

    @torch.no_grad()
    def coarse_forward(self, image_lr):
        d, feats = self.engine().branch('coarse', image_lr.float().contiguous())
        return d[:, None].clone(), [self._nchw(f) for f in feats]

    @torch.no_grad()
    def fine_forward(self, image_hr_crop):
        d, feats = self.engine().branch('fine', image_hr_crop.float().contiguous())
        return d[:, None].clone(), [self._nchw(f) for f in feats]

    def _to_map(self, x):
        """NCHW fp32 tensor -> engine Map (NHWC bf16, channels padded to 8)."""
        from .engine import Map
        from .ops import pad_to
        B, C, H, W = x.shape
        t = torch.zeros((B, H, W, pad_to(C, 8)), dtype=torch.bfloat16, device=x.device)
        t[..., :C] = x.permute(0, 2, 3, 1).to(torch.bfloat16)
        return Map(t, C)

    @torch.no_grad()
    def coarse_postprocess_test(self, coarse_prediction, coarse_features, bboxs, bboxs_feat):
        """`patchfusion.py:240-257`: ROI crop-zoom of the coarse depth and taps for the boxes `bboxs_feat` (T,5:
        batch index, x1, y1, x2, y2 in patch_process units).  NCHW fp32 in / out like the reference."""
        from . import ops
        boxes = bboxs_feat[:, 1:].float().contiguous()
        T = boxes.shape[0]
        P = self.patch_process_shape
        feats = []
        for f in coarse_features:
            m = self._to_map(f.float())
            h, w = m.hw
            out = torch.zeros((T, h, w, m.t.shape[-1]), dtype=torch.bfloat16, device=f.device)
            ops.roi_crop_zoom(m.t, m.C, boxes, h / P[0], out)
            feats.append(out[..., :m.C].float().permute(0, 3, 1, 2).contiguous())
        d = coarse_prediction.float().contiguous()
        droi = torch.zeros((T,) + tuple(d.shape[-2:]), dtype=torch.float32, device=d.device)
        ops.roi_crop_zoom(d[0, 0].contiguous(), 1, boxes, d.shape[-2] / P[0], droi)
        return {'coarse_depth_roi': droi[:, None], 'coarse_feats_roi': feats}

    @torch.no_grad()
    def infer_forward(self, imgs_crop, bbox_feat_forward, tile_temp, coarse_temp_dict=None):
        """`patchfusion.py:343-356`: fine branch + guided fusion of the given crops (T,3,h,w in [0,1]) against the
        whole-image coarse outputs in `tile_temp` ({'coarse_prediction', 'coarse_features'}, NCHW fp32).  As in the
        reference the G2L maps are recomputed from `tile_temp` on every call; `coarse_temp_dict` (the precomputed ROI
        crops) is accepted for signature compatibility - the kernels re-derive the crops from the boxes."""
        eng = self.engine()
        crops = imgs_crop.float().contiguous()
        boxes = bbox_feat_forward[:, 1:].float().contiguous()
        cf = [self._to_map(f.float()) for f in tile_temp['coarse_features']]
        cd = tile_temp['coarse_prediction'].float()[0, 0].contiguous()
        g2l = eng.g2l(cf)
        fd, ff = eng.branch('fine', crops)
        return eng.fusion(crops, boxes, fd, ff, cd, cf, g2l)[:, None].clone()

    # ------------------------------------------------------------------ tiling
    def _graphed(self, key, fn):
        """Run `fn` (a fixed kernel sequence over static buffers) through a CUDA graph: first call runs eagerly
        (allocating the engine buffers and caching the TMA maps) and captures, later calls replay.  One graph per
        (stage, micro-batch size, geometry): ~900 launches per micro-batch collapse into one host call."""
        from . import lib
        if not self.use_cuda_graphs or lib.PROFILER is not None:
            return fn()
        ent = self._graphs.get(key)
        if ent is None:
            fn()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = lib.launch_count()
            with torch.cuda.graph(g):
                fn()
            self._graphs[key] = (g, lib.launch_count() - n0)
            return None
        ent[0].replay()
        self.graph_launches += ent[1]          # kernels executed by the replay (for bench.py's gpu_launches)
        return None

    def _coarse_stage(self, eng, lr):
        cd, cf = eng.branch('coarse', lr)
        self._coarse = (cd[0], cf, eng.g2l(cf))

    def _fine_stage(self, eng, img, T, geom, par=0):
        """crop+resize -> fine branch for the T tiles whose origins sit in the static buffers of `_tile_io(T, par)`."""
        from . import ops
        H, W, h, w, ph, pw = geom
        io = self._tile_io(eng, T, par)
        crops = eng.buf('tile.crops%d' % par, (T, 3, ph, pw), torch.float32)
        ops.call('pf_crop_resize', img, H, W, io['raw'], T, h, w, ph, pw, crops, ops.stream_ptr())
        fd, ff = eng.branch('fine', crops, slot=str(par) if par else '')
        self._fine_out[par] = (crops, fd, ff)

    def _fusion_stage(self, eng, T, geom, canvas, mask, up, par=0):
        """guided fusion + scatter-stitch of the micro-batch whose fine-branch outputs sit in slot `par`."""
        from . import ops
        H, W, h, w, ph, pw = geom
        num, den, CH, CW = canvas
        io = self._tile_io(eng, T, par)
        cd, cf, g2l = self._coarse
        crops, fd, ff = self._fine_out[par]
        pred = eng.fusion(crops, io['boxes'], fd, ff, cd, cf, g2l)
        ops.call('pf_stitch_accumulate', num, den, CH, CW, pred, T, ph, pw, io['dst'], mask, up[0], up[1],
                 ops.stream_ptr())

    def _tiles_stage(self, eng, img, T, geom, canvas, mask, up):
        self._fine_stage(eng, img, T, geom, 0)
        self._fusion_stage(eng, T, geom, canvas, mask, up, 0)

    def _pair_stage(self, eng, img, geom, canvas, mask, up, fus, fine):
        """fusion of micro-batch k (slot fus[1]) on the current stream while the fine branch of micro-batch k+1
        (slot fine[1]) runs on a side stream: the two have no data dependency, so their kernels fill each other's
        partial waves and launch gaps."""
        cur = torch.cuda.current_stream()
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream()
        side = self._side_stream
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            self._fine_stage(eng, img, fine[0], geom, fine[1])
        self._fusion_stage(eng, fus[0], geom, canvas, mask, up, fus[1])
        cur.wait_stream(side)

    def _combine_canvases(self, num, den, base=None):
        """ONE all-gather of the stacked (num, den) canvases + fixed-order sum (pf_stitch_reduce)."""
        from . import ops
        from .parallel import gather_canvases
        stack = gather_canvases(num, den)
        ops.call('pf_stitch_reduce', stack, stack.shape[0], ops.C.c_int64(num.numel()), ops.stream_ptr())
        n, d = stack[0, 0], stack[0, 1]
        if base is not None:
            n, d = n + base[0], d + base[1]
        return n.contiguous(), d.contiguous()

    def _tile_io(self, eng, T, par=0):
        return dict(raw=eng.buf('tile.raw%d' % par, (T, 2), torch.int32),
                    dst=eng.buf('tile.dst%d' % par, (T, 2), torch.int32),
                    boxes=eng.buf('tile.boxes%d' % par, (T, 4), torch.float32))

    def _run_tiles(self, eng, img, raw, dst, geom, canvas, mask, up, process_num):
        H, W, h, w, ph, pw = geom
        fx = np.float32(1 / W * pw)
        fy = np.float32(1 / H * ph)
        # balanced micro-batches (e.g. 49 tiles, process_num 9 -> 9,8,8,8,8,8): at most two captured graph sizes
        # and no ragged tail; grouping does not change any tile's result
        nchunk = -(-len(raw) // process_num)
        if self.partition == 'balanced':
            sizes = [len(raw) // nchunk + (1 if i < len(raw) % nchunk else 0) for i in range(nchunk)]
        else:                                   # 'greedy': full micro-batches + one remainder (the reference's split)
            sizes = [min(process_num, len(raw) - i * process_num) for i in range(nchunk)]
        def load_io(i, s0, par):
            T = sizes[i]
            chunk = raw[s0:s0 + T]
            io = self._tile_io(eng, T, par)
            io['raw'].copy_(torch.tensor(chunk, dtype=torch.int32))
            io['dst'].copy_(torch.tensor(dst[s0:s0 + T], dtype=torch.int32))
            # boxes exactly as baseline_pretrain.py:268-282: int pixel box * fp32 factor
            bx = np.array([[np.float32(x) * fx, np.float32(y) * fy, np.float32(x + w) * fx, np.float32(y + h) * fy]
                           for (y, x) in chunk], dtype=np.float32)
            io['boxes'].copy_(torch.from_numpy(bx))

        gkey = tuple(geom) + (canvas[2], canvas[3]) + tuple(up)
        starts = [sum(sizes[:i]) for i in range(len(sizes))]
        from . import lib
        if not (self.overlap and self.use_cuda_graphs and lib.PROFILER is None and len(sizes) > 1):
            for i, T in enumerate(sizes):
                load_io(i, starts[i], 0)
                self._graphed(('tiles', T) + gkey, lambda: self._tiles_stage(eng, img, T, geom, canvas, mask, up))
            return
        # software pipeline over micro-batches: fine(0) | fusion(k) || fine(k+1) | fusion(last)
        load_io(0, starts[0], 0)
        self._graphed(('fine', sizes[0], 0) + gkey, lambda: self._fine_stage(eng, img, sizes[0], geom, 0))
        for i in range(len(sizes) - 1):
            pf_, pn_ = i & 1, (i + 1) & 1
            load_io(i + 1, starts[i + 1], pn_)
            self._graphed(('pair', sizes[i], pf_, sizes[i + 1], pn_) + gkey,
                          lambda: self._pair_stage(eng, img, geom, canvas, mask, up, (sizes[i], pf_), (sizes[i + 1], pn_)))
        last = len(sizes) - 1
        self._graphed(('fusion', sizes[last], last & 1) + gkey,
                      lambda: self._fusion_stage(eng, sizes[last], geom, canvas, mask, up, last & 1))

    @torch.no_grad()
    def forward(self, mode, image_lr, image_hr, depth_gt=None, crops_image_hr=None, crop_depths=None, bboxs=None,
                tile_cfg=None, cai_mode='m1', process_num=4, shard=None):
        """`shard=(rank, world)` (extension): process only this rank's tiles and combine the canvases with one
        all-gather (patchfusion_b200/parallel.py); None reproduces the reference's single-device behaviour."""
        if mode == 'train':
            raise NotImplementedError('training is out of scope of the B200 hot-path build (SURVEY.md §2 rows 10,12)')
        from . import ops
        if tile_cfg is None:
            tile_cfg = self.tile_cfg
        else:
            tile_cfg = self.prepare_tile_cfg(tile_cfg['image_raw_shape'], tile_cfg['patch_split_num'])
        assert image_hr.shape[0] == 1
        eng = self.engine()
        dev = image_hr.device
        st = ops.stream_ptr
        H, W = tile_cfg['image_raw_shape']
        assert tuple(image_hr.shape[-2:]) == (H, W), 'image_hr must already be at image_raw_shape'
        h, w = tile_cfg['patch_raw_shape']
        ph, pw = self.patch_process_shape
        RH, RW = tile_cfg['patch_reensemble_shape']
        geom = (H, W, h, w, ph, pw)
        # inputs into static buffers (stable addresses for the captured graphs)
        img = eng.buf('in.image_hr', (3, H, W), torch.float32)
        img.copy_(image_hr[0])
        lr = eng.buf('in.image_lr', (1, 3, ph, pw), torch.float32)
        lr.copy_(image_lr)

        self._graphed(('coarse', ph, pw), lambda: self._coarse_stage(eng, lr))

        num = eng.buf('canvas.num', (RH, RW), torch.float32)
        den = eng.buf('canvas.den', (RH, RW), torch.float32)
        num.zero_()
        den.zero_()
        mask = self._mask((ph, pw), dev)
        offsets = [((0, 0), (0, 0))]
        if cai_mode == 'm2' or cai_mode[0] == 'r':
            offsets += [((0, w // 2), (0, pw // 2)), ((h // 2, 0), (ph // 2, 0)), ((h // 2, w // 2), (ph // 2, pw // 2))]
        # The regular passes (baseline_pretrain.py:221-331, patchfusion.py:417-439) are independent tiles whose
        # stitch is a commutative weighted sum, so all passes are flattened into one tile list and micro-batched.
        raw, proc = [], []
        for (oy, ox), (py, px) in offsets:
            assert ox >= 0 and oy >= 0
            ny, nx = (H - oy) // h, (W - ox) // w
            raw += [(h * a + oy, w * b + ox) for a in range(ny) for b in range(nx)]
            proc += [(ph * a + py, pw * b + px) for a in range(ny) for b in range(nx)]
        if shard is not None:
            from .parallel import shard_indices
            own = shard_indices(len(raw), shard[0], shard[1])
            raw, proc = [raw[i] for i in own], [proc[i] for i in own]
        if raw:
            self._run_tiles(eng, img, raw, proc, geom, (num, den, RH, RW), mask, (0, 0), process_num)
        if shard is not None:
            num, den = self._combine_canvases(num, den)
        if cai_mode[0] == 'r':
            mask = self._mask((h, w), dev)
            n2 = eng.buf('canvas.num_raw', (H, W), torch.float32)
            d2 = eng.buf('canvas.den_raw', (H, W), torch.float32)
            ops.call('pf_stitch_resize', num, den, RH, RW, H, W, n2, d2, st())
            num, den = n2, d2
            if shard is not None:
                # the resized regular-phase canvas is identical on every rank: keep it aside so the all-gather of the
                # random phase only sums the per-rank increments
                n2_base, d2_base = n2.clone(), d2.clone()
                n2.zero_()
                d2.zero_()
            for _ in range(int(cai_mode[1:]) // process_num):
                ys = [random.randint(0, H - h - 1) for _ in range(process_num)]     # baseline_pretrain.py:155-156
                x0 = random.randint(0, W - w - 1)
                raw = [(y, x0) for y in ys]
                if shard is not None:        # every rank draws the same boxes (same `random` state), owns a slice
                    raw = [raw[i] for i in shard_indices(len(raw), shard[0], shard[1])]
                if raw:
                    self._run_tiles(eng, img, raw, raw, geom, (num, den, H, W), mask, (h, w), process_num)
            if shard is not None:
                num, den = self._combine_canvases(num, den, base=(n2_base, d2_base))
        out = torch.empty_like(num)
        ops.call('pf_stitch_finalize', num, den, ops.C.c_int64(num.numel()), out, st())
        depth = out[None, None]
        return depth, {'rgb': image_lr, 'depth_pred': depth, 'depth_gt': depth_gt}
