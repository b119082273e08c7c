"""Model configs as plain dicts, schema-compatible with the reference's
`configs/patchfusion_depthanything/depthanything_{vits,vitb,vitl}_patchfusion_u4k.py:9-90` (`model.config`) and with
the `config.json` the reference's `tools/convert_huggingface.py:78-79` writes next to HF checkpoints."""
import copy

_BRANCH = dict(
    type='DA-ZoeDepth', min_depth=1e-3, max_depth=80, depth_anything=True, midas_model_type='vitl',
    img_size=[392, 518], pretrained_resource=None, use_pretrained_midas=True, train_midas=True,
    freeze_midas_bn=True, do_resize=False,
    attractor_alpha=1000, attractor_gamma=2, attractor_kind='mean', attractor_type='inv',
    bin_centers_type='softplus', bin_embedding_dim=128, force_keep_ar=True, inverse_midas=False,
    max_temp=50.0, memory_efficient=True, min_temp=0.0212, n_attractors=[16, 8, 4, 1], n_bins=64,
    output_distribution='logbinomial')

_FEATURES = {'vits': 64, 'vitb': 128, 'vitl': 256}


def depth_anything_patchfusion(encoder='vitl', image_raw_shape=(2160, 3840), patch_split_num=(4, 4),
                               patch_process_shape=(392, 518)):
    if encoder not in _FEATURES:
        raise NotImplementedError(encoder)
    br = copy.deepcopy(_BRANCH)
    br['midas_model_type'] = encoder
    br['img_size'] = list(patch_process_shape)
    c = _FEATURES[encoder]
    h, w = patch_process_shape
    sizes = [(h, w)]
    cur = (h // 14 * 8, w // 14 * 8)          # DPT pyramid: 8x, 4x, 2x, 1x, ~0.5x of the 14-px patch grid
    for _ in range(4):
        sizes.append(cur)
        cur = (cur[0] // 2, cur[1] // 2)
    sizes.append(((h // 14 - 1) // 2 + 1, (w // 14 - 1) // 2 + 1))
    return dict(
        image_raw_shape=list(image_raw_shape), patch_split_num=list(patch_split_num),
        patch_process_shape=list(patch_process_shape), min_depth=1e-3, max_depth=80,
        load_branch=False, pretrain_model=['', ''],
        coarse_branch=copy.deepcopy(br), fine_branch=copy.deepcopy(br),
        guided_fusion=dict(type='GuidedFusionPatchFusion', patch_process_shape=list(patch_process_shape),
                           in_channels=[32, c, c, c, c, c], num_patches=[a * b for a, b in sizes],
                           n_channels=5, g2l=True),
        sigloss=dict(type='SILogLoss'))
