"""pf_gemm (tcgen05 implicit GEMM) vs torch fp32 on the same bf16-rounded operands."""
import pytest
import torch
import torch.nn.functional as F

from gpu_util import bf, check, from_nhwc, rb, to_nhwc

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def _ops():
    from patchfusion_b200 import ops
    return ops


@pytest.mark.parametrize('M,K,N,act', [(300, 192, 96, 'gelu'), (128, 64, 32, 'none'), (2074, 1024, 3072, 'none'),
                                       (1036, 592, 384, 'relu'), (777, 384, 1536, 'softplus'), (4144, 256, 544, 'none')])
def test_linear(cuda, M, K, N, act):
    ops = _ops()
    g = torch.Generator(device='cuda').manual_seed(M + K + N)
    x = torch.randn(M, K, device=cuda, generator=g)
    w = torch.randn(N, K, device=cuda, generator=g) / K ** 0.5
    b = torch.randn(N, device=cuda, generator=g)
    pw = ops.pack_weight(w, b)
    ld = ops.pad_to(K, 8)
    xa = torch.zeros(M, ld, dtype=torch.bfloat16, device=cuda)
    xa[:, :K] = bf(x)
    out = torch.full((M, ops.pad_to(N, 8)), 7.0, dtype=torch.bfloat16, device=cuda)
    acts = dict(none=ops.ACT_NONE, relu=ops.ACT_RELU, gelu=ops.ACT_GELU, softplus=ops.ACT_SOFTPLUS)
    ops.gemm(pw, [xa], out, act=acts[act])
    torch.cuda.synchronize()
    ref = F.linear(rb(x), rb(w), b)
    ref = dict(none=lambda t: t, relu=F.relu, gelu=F.gelu, softplus=F.softplus)[act](ref)
    check('linear %dx%dx%d %s' % (M, K, N, act), out[:, :N], ref, 1e-2)
    if out.shape[1] > N:
        assert (out[:, N:] == 7.0).all(), 'columns beyond N were written'


def test_linear_f32_and_layerscale(cuda):
    ops = _ops()
    g = torch.Generator(device='cuda').manual_seed(1)
    M, K, N = 1037, 1536, 384
    x = torch.randn(M, K, device=cuda, generator=g)
    w = torch.randn(N, K, device=cuda, generator=g) / K ** 0.5
    b = torch.randn(N, device=cuda, generator=g)
    gamma = torch.rand(N, device=cuda, generator=g)
    res = torch.randn(M, N, device=cuda, generator=g)
    pw = ops.pack_weight(w, b)
    out = torch.zeros(M, N, dtype=torch.float32, device=cuda)
    ops.gemm(pw, [bf(x).contiguous()], out)
    ref = F.linear(rb(x), rb(w), b)
    check('linear fp32 out', out, ref, 2e-5)
    xres = res.clone()
    ops.gemm(pw, [bf(x).contiguous()], xres, gamma=gamma)
    check('x += gamma*(xW+b)', xres, res + gamma * ref, 2e-5)


@pytest.mark.parametrize('M,K,N,mode', [(9333, 1024, 1024, 'gamma'), (4148, 1024, 4096, 'gelu'), (9333, 1024, 3072, 'bf16'),
                                        (20000, 32, 96, 'f32'), (9324, 592, 1024, 'f32'), (2400, 256, 2048, 'bf16')])
def test_linear_weight_multicast(cuda, M, K, N, mode):
    """shapes that take the cluster-of-2 weight-multicast variant (>= 4 m-tiles, >= 148 tiles): odd m-tile counts
    (73, 33: the last pair has an out-of-range m-tile), every epilogue flavour, block_n 256 / 96."""
    ops = _ops()
    g = torch.Generator(device='cuda').manual_seed(M + K + N)
    x = torch.randn(M, K, device=cuda, generator=g)
    w = torch.randn(N, K, device=cuda, generator=g) / K ** 0.5
    b = torch.randn(N, device=cuda, generator=g)
    pw = ops.pack_weight(w, b)
    ld = ops.pad_to(K, 8)
    xa = torch.zeros(M, ld, dtype=torch.bfloat16, device=cuda)
    xa[:, :K] = bf(x)
    ref = F.linear(rb(x), rb(w), b)
    if mode == 'gamma':
        res = torch.randn(M, N, device=cuda, generator=g)
        gamma = torch.rand(N, device=cuda, generator=g)
        out = res.clone()
        d = ops.gemm(pw, [xa], out, gamma=gamma, src_c=[K])
        check('mc x += gamma*(xW+b)', out, res + gamma * ref, 2e-5)
    elif mode == 'gelu':
        out = torch.zeros(M, N, dtype=torch.bfloat16, device=cuda)
        d = ops.gemm(pw, [xa], out, act=ops.ACT_GELU, src_c=[K])
        check('mc gelu', out, F.gelu(ref), 1e-2)
    elif mode == 'f32':
        out = torch.zeros(M, N, dtype=torch.float32, device=cuda)
        d = ops.gemm(pw, [xa], out, src_c=[K])
        check('mc fp32 out', out, ref, 2e-5)
    else:
        out = torch.zeros(M, N, dtype=torch.bfloat16, device=cuda)
        d = ops.gemm(pw, [xa], out, src_c=[K])
        check('mc bf16 out', out, ref, 1e-2)
    assert d.m_tiles >= 4 and d.m_tiles * d.n_tiles >= 148


def test_qkv_split_transposed_v(cuda):
    ops = _ops()
    g = torch.Generator(device='cuda').manual_seed(2)
    B, seq, D, heads = 2, 1037, 384, 6
    seq_pad = ops.pad_to(seq, 8)
    x = torch.randn(B * seq, D, device=cuda, generator=g)
    w = torch.randn(3 * D, D, device=cuda, generator=g) / D ** 0.5
    b = torch.randn(3 * D, device=cuda, generator=g)
    pw = ops.pack_weight(w, b)
    qk = torch.zeros(B * seq, 2 * D, dtype=torch.bfloat16, device=cuda)
    vt = torch.zeros(B * D, seq_pad, dtype=torch.bfloat16, device=cuda)
    ops.gemm(pw, [bf(x).contiguous()], qk, vt=vt, vt_col0=2 * D, vt_seq=seq, vt_seq_pad=seq_pad)
    ref = F.linear(rb(x), rb(w), b)
    check('qk part', qk, ref[:, :2 * D], 1e-2)
    v_ref = ref[:, 2 * D:].reshape(B, seq, D).permute(0, 2, 1)              # [B, D, seq]
    check('v transposed', vt.reshape(B, D, seq_pad)[:, :, :seq], v_ref, 1e-2)
    assert (vt.reshape(B, D, seq_pad)[:, :, seq:] == 0).all()


@pytest.mark.parametrize('NB,H,W,cs,N,act', [(2, 14, 19, [64], 64, 'none'), (1, 37, 50, [128], 256, 'relu'),
                                             (2, 28, 37, [32, 64, 64], 32, 'relu'), (1, 56, 74, [192], 192, 'none'),
                                             (1, 30, 41, [8], 32, 'relu'), (3, 16, 16, [544], 544, 'relu'),
                                             # >= 148 tiles: the weight-multicast pairs of the halo kernel (even and
                                             # odd m-tile counts, three n-tiles with a narrower last one, N = 32)
                                             (5, 50, 70, [96], 64, 'relu'), (7, 48, 72, [40, 64], 544, 'none'),
                                             (7, 48, 72, [136], 32, 'relu')])
@pytest.mark.parametrize('tile', [None, (8, 16)])
def test_conv3x3(cuda, NB, H, W, cs, N, act, tile):
    """tile=None: halo-tile kernel (one A fetch per chunk); tile=(8,16): generic per-tap TMA kernel."""
    ops = _ops()
    g = torch.Generator(device='cuda').manual_seed(H * W + N)
    ctot = sum(cs)
    xs = [torch.randn(NB, c, H, W, device=cuda, generator=g) for c in cs]
    w = torch.randn(N, ctot, 3, 3, device=cuda, generator=g) / (9 * ctot) ** 0.5
    b = torch.randn(N, device=cuda, generator=g)
    pw = ops.pack_weight(w, b, src_c=cs)
    srcs = [to_nhwc(x) for x in xs]
    out = torch.zeros(NB, H, W, ops.pad_to(N, 8), dtype=torch.bfloat16, device=cuda)
    ops.gemm(pw, srcs, out, image=(NB, H, W), act=ops.ACT_RELU if act == 'relu' else ops.ACT_NONE, tile=tile)
    ref = F.conv2d(torch.cat([rb(x) for x in xs], 1), rb(w), b, padding=1)
    if act == 'relu':
        ref = F.relu(ref)
    check('conv3x3 %s -> %d @%dx%d' % (cs, N, H, W), from_nhwc(out, N), ref, 1e-2)


@pytest.mark.parametrize('NB,H,W,cs,N,act', [(5, 50, 70, [96], 64, 'relu'), (7, 48, 72, [40, 64], 544, 'none'),
                                             (7, 48, 72, [136], 32, 'relu')])
@pytest.mark.parametrize('cluster', [0, 2])
def test_conv3x3_halo_multicast_variants(cuda, NB, H, W, cs, N, act, cluster):
    """PF_OPT_HALO_MULTICAST = 0 (one CTA per tile) and 2 (clusters of 4 CTAs): the >= 148-tile shapes that take the
    cluster-of-2 weight-multicast variant by default (see test_conv3x3)"""
    from patchfusion_b200 import lib
    lib.call('pf_set_option', lib.OPT_HALO_MULTICAST, cluster)
    try:
        test_conv3x3(cuda, NB, H, W, cs, N, act, None)
    finally:
        lib.call('pf_set_option', lib.OPT_HALO_MULTICAST, 1)


@pytest.mark.parametrize('NB,H,W,srcs,N', [
    (2, 28, 37, [(14, 19, 64)], 64),                          # x2 up-sample of one source
    (3, 56, 74, [(56, 74, 32), (28, 37, 128), (28, 37, 72)], 96),   # plain + two resampled sources (U-Net `up` conv)
    (2, 40, 52, [(35, 46, 64), (40, 52, 24)], 32),            # non-integer ratio; second source read directly
    (7, 48, 72, [(24, 36, 136)], 544),                        # >= 148 tiles: with the weight-multicast pairs
])
def test_conv3x3_fused_bilinear_resample(cuda, NB, H, W, srcs, N):
    """pf_gemm_desc.rs_h / rs_w: the conv reads its sources through F.interpolate(bilinear, align_corners=True); the
    resampled map is produced in the operand stage of the halo kernel and rounded to bf16 like a materialised one"""
    ops = _ops()
    g = torch.Generator(device='cuda').manual_seed(H * W + N)
    cs = [c for _, _, c in srcs]
    xs = [torch.randn(NB, c, h, w, device=cuda, generator=g) for h, w, c in srcs]
    w_ = torch.randn(N, sum(cs), 3, 3, device=cuda, generator=g) / (9 * sum(cs)) ** 0.5
    b = torch.randn(N, device=cuda, generator=g)
    pw = ops.pack_weight(w_, b, src_c=cs)
    out = torch.zeros(NB, H, W, ops.pad_to(N, 8), dtype=torch.bfloat16, device=cuda)
    ops.gemm(pw, [to_nhwc(x) for x in xs], out, image=(NB, H, W), act=ops.ACT_RELU,
             resample=[(h, w_in) != (H, W) for h, w_in, _ in srcs])
    up = [rb(x) if tuple(x.shape[-2:]) == (H, W) else rb(F.interpolate(rb(x), size=(H, W), mode='bilinear', align_corners=True))
          for x in xs]
    ref = F.relu(F.conv2d(torch.cat(up, 1), rb(w_), b, padding=1))
    check('conv3x3 fused resample %s -> %d @%dx%d' % (srcs, N, H, W), from_nhwc(out, N), ref, 1e-2)


def test_conv3x3_residuals_and_relu_copy(cuda):
    ops = _ops()
    g = torch.Generator(device='cuda').manual_seed(5)
    NB, H, W, Cc = 2, 28, 37, 64
    x = torch.randn(NB, Cc, H, W, device=cuda, generator=g)
    r1 = torch.randn(NB, Cc, H, W, device=cuda, generator=g)
    r2 = torch.randn(NB, Cc, H, W, device=cuda, generator=g)
    w = torch.randn(Cc, Cc, 3, 3, device=cuda, generator=g) / (9 * Cc) ** 0.5
    b = torch.randn(Cc, device=cuda, generator=g)
    pw = ops.pack_weight(w, b)
    out = torch.zeros(NB, H, W, Cc, dtype=torch.bfloat16, device=cuda)
    out2 = torch.zeros_like(out)
    ops.gemm(pw, [to_nhwc(x)], out, image=(NB, H, W), res1=to_nhwc(r1), res2=to_nhwc(r2), out2=out2)
    ref = F.conv2d(rb(x), rb(w), b, padding=1) + rb(r1) + rb(r2)
    check('conv + res1 + res2', from_nhwc(out, Cc), ref, 1e-2)
    check('relu copy', from_nhwc(out2, Cc), F.relu(ref), 1e-2)


def test_conv1x1_bn_fold_and_channel_offset(cuda):
    ops = _ops()
    g = torch.Generator(device='cuda').manual_seed(6)
    NB, H, W, Cin, N = 1, 20, 23, 96, 48
    x = torch.randn(NB, Cin, H, W, device=cuda, generator=g)
    w = torch.randn(N, Cin, 1, 1, device=cuda, generator=g) / Cin ** 0.5
    scale = torch.rand(N, device=cuda, generator=g) + 0.5
    shift = torch.randn(N, device=cuda, generator=g)
    pw = ops.pack_weight(w, None, scale=scale, shift=shift)
    out = torch.zeros(NB, H, W, 128, dtype=torch.bfloat16, device=cuda)
    ops.gemm(pw, [to_nhwc(x)], out, image=(NB, H, W), out_col0=64)
    ref = F.conv2d(rb(x), rb(w * scale.view(-1, 1, 1, 1))) + shift.view(1, -1, 1, 1)
    check('conv1x1 + folded BN at col 64', out[..., 64:64 + N].float().permute(0, 3, 1, 2), ref, 1e-2)
    assert (out[..., :64] == 0).all() and (out[..., 64 + N:] == 0).all()


@pytest.mark.parametrize('k,Cin,Cout,H,W', [(4, 48, 48, 5, 7), (2, 96, 96, 9, 6), (4, 256, 256, 28, 37)])
def test_conv_transpose(cuda, k, Cin, Cout, H, W):
    ops = _ops()
    g = torch.Generator(device='cuda').manual_seed(k + Cin)
    NB = 2
    x = torch.randn(NB, Cin, H, W, device=cuda, generator=g)
    w = torch.randn(Cin, Cout, k, k, device=cuda, generator=g) / Cin ** 0.5
    b = torch.randn(Cout, device=cuda, generator=g)
    pw = ops.pack_weight_convT(w, b, k)
    src = to_nhwc(x).reshape(NB * H * W, -1)
    out = torch.zeros(NB, H * k, W * k, ops.pad_to(Cout, 8), dtype=torch.bfloat16, device=cuda)
    ops.gemm_convT(pw, src, (NB, H, W), out)
    ref = F.conv_transpose2d(rb(x), rb(w), b, stride=k)
    check('convT k%d %d->%d' % (k, Cin, Cout), from_nhwc(out, Cout), ref, 1e-2)


def test_fused_trailing_layer(cuda):
    """conv/linear + activation with a narrow second 1x1 layer folded into the epilogue (heads 80->4, 128->nA, 32->1)."""
    ops = _ops()
    g = torch.Generator(device='cuda').manual_seed(11)
    NB, H, W = 2, 30, 41
    for cs, N, n2, act, act2, skip in [([32, 128], 80, 4, 'gelu', 'softplus', True), ([128], 32, 1, 'relu', 'relu', False),
                                       ([128], 128, 16, 'relu', 'softplus', True)]:
        xs = [torch.randn(NB, c, H, W, device=cuda, generator=g) for c in cs]
        taps = 9 if N == 32 else 1
        k = 3 if taps == 9 else 1
        w = torch.randn(N, sum(cs), k, k, device=cuda, generator=g) / (taps * sum(cs)) ** 0.5
        b = torch.randn(N, device=cuda, generator=g)
        w2 = torch.randn(n2, N, device=cuda, generator=g) / N ** 0.5
        b2 = torch.randn(n2, device=cuda, generator=g)
        pw = ops.pack_weight(w, b, src_c=cs)
        out = torch.full((NB, H, W, ops.pad_to(N, 8)), 5.0, dtype=torch.bfloat16, device=cuda)
        out3 = torch.zeros(NB, H, W, 32, dtype=torch.float32, device=cuda)
        A = dict(relu=ops.ACT_RELU, gelu=ops.ACT_GELU, softplus=ops.ACT_SOFTPLUS)
        ops.gemm(pw, [to_nhwc(x) for x in xs], out, image=(NB, H, W), act=A[act], tail=(w2, b2, A[act2]),
                 tail_out=out3, skip_main=skip)
        fa = dict(relu=F.relu, gelu=F.gelu, softplus=F.softplus)
        mid = fa[act](F.conv2d(torch.cat([rb(x) for x in xs], 1), rb(w), b, padding=k // 2))
        ref = fa[act2](F.conv2d(mid, w2.view(n2, N, 1, 1), b2))
        check('fused tail %s->%d->%d' % (cs, N, n2), out3[..., :n2].permute(0, 3, 1, 2), ref, 2e-3)
        if skip:
            assert (out == 5.0).all(), 'main output must not be written'
        else:
            check('main output', from_nhwc(out, N), mid, 1e-2)
