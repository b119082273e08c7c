// Stage-level sequencing of the PatchFusion hot path: the kernel order behind the reference's
// coarse_forward / fine_forward / fusion_forward / G2L, issued from C++ over caller-owned weights and ONE workspace.
// Every intermediate is bump-allocated from the workspace in program order, so for given (weights, batch) the
// addresses never change: tensor maps are cached, the call is CUDA-graph capturable, and the workspace size is found by
// running the same code in "dry" mode (no launches).  Reference lines are cited at each step.
#include <cuda_bf16.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "pf_kernels.h"

namespace pf {

typedef __nv_bfloat16 bf16;

static inline int pad_to(int n, int m) { return (n + m - 1) / m * m; }

struct Map {            // NHWC bf16 activation
  bf16* p; int B, H, W, C, ld;
  long long rows() const { return static_cast<long long>(B) * H * W; }
};
struct MapF {           // NHWC fp32 activation (metric-bins tail)
  float* p; int B, H, W, C, ld;
};

struct Ctx {
  uint8_t* base; size_t cap, off;
  bool dry;
  void* stream;
  int err;
  pf_tap_fn tap; void* tap_user;

  void* alloc(size_t bytes) {
    off = (off + 255) & ~static_cast<size_t>(255);
    void* p = dry ? nullptr : base + off;
    off += bytes;
    if (!dry && off > cap && !err) err = set_error("workspace too small: need > %zu bytes, have %zu", off, cap);
    return p;
  }
  bool live() const { return !dry && !err; }
  Map map(int B, int H, int W, int C) {
    Map m; m.B = B; m.H = H; m.W = W; m.C = C; m.ld = pad_to(C, 8);
    m.p = static_cast<bf16*>(alloc(static_cast<size_t>(B) * H * W * m.ld * 2));
    return m;
  }
  MapF mapf(int B, int H, int W, int C, int ld) {
    MapF m; m.B = B; m.H = H; m.W = W; m.C = C; m.ld = ld;
    m.p = static_cast<float*>(alloc(static_cast<size_t>(B) * H * W * ld * 4));
    return m;
  }
  void chk(int rc) { if (rc && !err) err = rc; }
  void tap_out(const char* name, const void* ptr, int is_f32, long long rows, int cols, int ld) {
    if (tap && live()) tap(tap_user, name, ptr, is_f32, rows, cols, ld);
  }
};

static Map from_pf(const pf_map& m) {
  Map r; r.p = static_cast<bf16*>(m.ptr); r.B = m.B; r.H = m.H; r.W = m.W; r.C = m.C; r.ld = m.ld;
  return r;
}
static pf_map to_pf(const Map& m) {
  pf_map r; r.ptr = m.p; r.B = m.B; r.H = m.H; r.W = m.W; r.C = m.C; r.ld = m.ld;
  return r;
}

// ---------------------------------------------------------------------------------------------------- gemm wrappers
struct GemmOpt {
  int act = PF_ACT_NONE;
  const Map* res1 = nullptr; const Map* res2 = nullptr;
  const float* gamma = nullptr;
  Map* relu_copy = nullptr;            // second output = relu(v)
  bf16* vt = nullptr; int vt_col0 = 0, vt_seq = 0, vt_seq_pad = 0;
  bool tail = false; int act2 = PF_ACT_NONE; float* out3 = nullptr; int out3_ld = 0; bool skip_main = false;
  int ps_B = 0, ps_H = 0, ps_W = 0;    // ConvTranspose input grid
};

// rows x K matrix sources (a_mode 0) or NHWC images (a_mode 1)
static void gemm_desc_common(pf_gemm_desc& d, const pf_layer& L, const GemmOpt& o, void* out, int out_f32, int out_ld) {
  d.taps = L.taps;
  d.w_ptr = L.w; d.N = L.N; d.Ktot = L.Ktot; d.block_n = 0;
  d.bias = L.bias; d.act = o.act;
  if (o.res1) { d.res1 = o.res1->p; d.res_ld = o.res1->ld; }
  if (o.res2) d.res2 = o.res2->p;
  d.gamma = o.gamma;
  d.out = out; d.out_f32 = out_f32; d.out_ld = out_ld; d.out_col0 = 0;
  if (o.relu_copy) { d.out2 = o.relu_copy->p; d.out2_ld = o.relu_copy->ld; }
  d.ps = L.ps > 1 ? L.ps : 1; d.ps_cout = L.ps_cout;
  if (o.vt) { d.vt = o.vt; d.vt_col0 = o.vt_col0; d.vt_seq = o.vt_seq; d.vt_seq_pad = o.vt_seq_pad; d.vt_dim = L.N - o.vt_col0; }
  if (o.tail) {
    d.w2 = L.w2; d.b2 = L.b2; d.n2 = L.n2; d.act2 = o.act2; d.skip_main = o.skip_main ? 1 : 0;
    d.out3 = o.out3; d.out3_ld = o.out3_ld;
  }
}

// D[M, N] = A[M, K] W^T : A is a bf16 matrix with `cols` logical columns and row stride ld
static void linear(Ctx& c, const pf_layer& L, const bf16* A, long long M, int cols, int ld, void* out, int out_f32,
                   int out_ld, const GemmOpt& o = GemmOpt()) {
  if (!c.live()) return;
  pf_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.num_src = 1; d.a_mode = 0;
  d.a_ptr[0] = A; d.a_c[0] = pad_to(cols, 8); d.a_ld[0] = ld;
  d.M = static_cast<int32_t>(M);
  if (L.ps > 1) { d.NB = o.ps_B; d.H = o.ps_H; d.W = o.ps_W; }
  gemm_desc_common(d, L, o, out, out_f32, out_ld);
  c.chk(pf_gemm(&d, c.stream));
}

// 3x3 / 1x1 conv over up to three channel-concatenated NHWC sources
static void conv_into(Ctx& c, const pf_layer& L, const Map* const* srcs, int ns, void* out, int out_f32, int out_ld,
                      const GemmOpt& o = GemmOpt()) {
  if (!c.live()) return;
  pf_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.num_src = ns; d.a_mode = 1;
  for (int i = 0; i < ns; ++i) { d.a_ptr[i] = srcs[i]->p; d.a_c[i] = pad_to(srcs[i]->C, 8); d.a_ld[i] = srcs[i]->ld; }
  d.NB = srcs[0]->B; d.H = srcs[0]->H; d.W = srcs[0]->W;
  gemm_desc_common(d, L, o, out, out_f32, out_ld);
  c.chk(pf_gemm(&d, c.stream));
}

// 3x3 conv whose sources are read THROUGH F.interpolate(mode='bilinear', align_corners=True) to (H, W): sources of
// another size are resampled inside the conv kernel's operand stage (pf_gemm_desc.rs_h / rs_w) instead of being
// materialised; sources already at (H, W) are read directly.  Opt-in (PF_OPT_FUSED_RESAMPLE = 1): two producer warps per
// CTA cannot keep up with the tensor pipe (measured 215 -> 317 ms per 4K image), so the default materialises them.
static Map resize(Ctx& c, const Map& x, int OH, int OW);
static Map conv_resampled(Ctx& c, const pf_layer& L, const Map* const* srcs, int ns, int H, int W, const GemmOpt& o = GemmOpt()) {
  Map out = c.map(srcs[0]->B, H, W, L.N);
  if (L.taps != 9 || !option(PF_OPT_FUSED_RESAMPLE)) {
    Map tmp[3];
    const Map* a[3];
    for (int i = 0; i < ns; ++i) { tmp[i] = resize(c, *srcs[i], H, W); a[i] = &tmp[i]; }
    conv_into(c, L, a, ns, out.p, 0, out.ld, o);
    return out;
  }
  if (!c.live()) return out;
  pf_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.num_src = ns; d.a_mode = 1;
  for (int i = 0; i < ns; ++i) {
    d.a_ptr[i] = srcs[i]->p; d.a_c[i] = pad_to(srcs[i]->C, 8); d.a_ld[i] = srcs[i]->ld;
    if (srcs[i]->H != H || srcs[i]->W != W) { d.rs_h[i] = srcs[i]->H; d.rs_w[i] = srcs[i]->W; }
  }
  d.NB = srcs[0]->B; d.H = H; d.W = W;
  gemm_desc_common(d, L, o, out.p, 0, out.ld);
  c.chk(pf_gemm(&d, c.stream));
  return out;
}

static Map conv(Ctx& c, const pf_layer& L, const Map* const* srcs, int ns, const GemmOpt& o = GemmOpt()) {
  Map out = c.map(srcs[0]->B, srcs[0]->H, srcs[0]->W, L.N);
  conv_into(c, L, srcs, ns, out.p, 0, out.ld, o);
  return out;
}
static Map conv1(Ctx& c, const pf_layer& L, const Map& s, const GemmOpt& o = GemmOpt()) {
  const Map* a[1] = {&s};
  return conv(c, L, a, 1, o);
}

// F.interpolate(mode='bilinear', align_corners=True)
static Map resize(Ctx& c, const Map& x, int OH, int OW) {
  if (x.H == OH && x.W == OW) return x;
  Map out = c.map(x.B, OH, OW, x.C);
  if (c.live()) c.chk(pf_resize_bilinear(x.p, x.B, x.H, x.W, pad_to(x.C, 8), x.ld, OH, OW, out.p, out.ld, 0, c.stream));
  return out;
}

// ---------------------------------------------------------------------------------------------------- metric-bins head
// zoedepth_v1.py:173-219 (branch heads, with the relative-depth condition) / patchfusion.py:297-339 (fusion head)
static void metric_head(Ctx& c, const pf_head& Hd, const Map& x, const Map* x_blocks, const Map& last, const MapF* rel,
                        float* depth_out) {
  const int B = x.B, nb = Hd.n_bins, E = Hd.bin_embedding_dim;
  // two-layer 1x1 MLP `_net` (localbins_layers.py:84-89,110-114, attractor.py:157-162)
  auto mlp_bf16 = [&](const pf_layer& L0, const pf_layer& L2, const Map& src) -> Map {
    GemmOpt o0; o0.act = PF_ACT_RELU;
    Map t = conv1(c, L0, src, o0);
    return conv1(c, L2, t);
  };
  auto mlp_f32 = [&](const pf_layer& L0, const pf_layer& L2, const Map& src, int act2) -> MapF {
    const int n_out = L0.n2 > 0 ? L0.n2 : L2.N;
    MapF o = c.mapf(B, src.H, src.W, n_out, pad_to(pad_to(n_out, 8), 32));
    if (L0.n2 > 0) {       // narrow second layer fused into the first layer's epilogue
      GemmOpt g; g.act = PF_ACT_RELU; g.tail = true; g.act2 = act2; g.out3 = o.p; g.out3_ld = o.ld; g.skip_main = true;
      conv1(c, L0, src, g);
    } else {
      GemmOpt o0; o0.act = PF_ACT_RELU;
      Map t = conv1(c, L0, src, o0);
      const Map* a[1] = {&t};
      GemmOpt g2; g2.act = act2;
      conv_into(c, L2, a, 1, o.p, 1, o.ld, g2);
    }
    return o;
  };
  MapF b_prev = mlp_f32(Hd.seed0, Hd.seed2, x, PF_ACT_SOFTPLUS);        // seed bin centres, fp32 [B,h,w,64]
  Map prev_emb = mlp_bf16(Hd.seedproj0, Hd.seedproj2, x);
  const float* b_t = b_prev.p;
  int ph = x.H, pw = x.W;
  for (int i = 0; i < 4; ++i) {
    const Map& xb = x_blocks[i];
    Map emb = mlp_bf16(Hd.proj0[i], Hd.proj2[i], xb);
    Map s = c.map(B, xb.H, xb.W, E);
    if (c.live()) c.chk(pf_add_upsampled(emb.p, B, xb.H, xb.W, E, prev_emb.p, prev_emb.H, prev_emb.W, s.p, c.stream));
    MapF A = mlp_f32(Hd.att0[i], Hd.att2[i], s, PF_ACT_SOFTPLUS);
    float* b_new = static_cast<float*>(c.alloc(static_cast<size_t>(B) * xb.H * xb.W * nb * 4));
    if (c.live())
      c.chk(pf_attractor(A.p, A.ld, Hd.n_attractors[i], b_t, ph, pw, B, xb.H, xb.W, nb, Hd.attractor_flags, b_new, c.stream));
    b_t = b_new; ph = xb.H; pw = xb.W; prev_emb = emb;
    char nm[8];
    snprintf(nm, sizeof(nm), "b%d", i);
    c.tap_out(nm, b_new, 1, static_cast<long long>(B) * xb.H * xb.W, nb, nb);
  }
  const int H = last.H, W = last.W;
  Map emb_up = resize(c, prev_emb, H, W);
  float* pt = nullptr;
  GemmOpt g; g.act = PF_ACT_GELU; g.tail = true; g.act2 = PF_ACT_SOFTPLUS; g.out3_ld = 8; g.skip_main = true;
  if (rel != nullptr) {
    Map relb = c.map(B, H, W, 1);
    if (c.live()) c.chk(pf_f32_to_bf16(rel->p, static_cast<int64_t>(B) * H * W * rel->ld, relb.p, c.stream));
    pt = static_cast<float*>(c.alloc(static_cast<size_t>(B) * H * W * 8 * 4));
    g.out3 = pt;
    const Map* a[3] = {&last, &relb, &emb_up};
    conv(c, Hd.clb0, a, 3, g);         // CLB MLP: 1x1 (161->80) + GELU with 80->4 + Softplus fused (dist_layers.py:91-98)
  } else {
    pt = static_cast<float*>(c.alloc(static_cast<size_t>(B) * H * W * 8 * 4));
    g.out3 = pt;
    const Map* a[2] = {&last, &emb_up};
    conv(c, Hd.clb0, a, 2, g);
  }
  if (c.live())
    c.chk(pf_logbinom_depth(pt, 8, b_t, ph, pw, B, H, W, nb, Hd.min_temp, Hd.max_temp, depth_out, c.stream));
}

// ---------------------------------------------------------------------------------------------------- one branch
static void branch_run(Ctx& c, const pf_branch& Wb, const float* images, int B, pf_branch_out* out) {
  const int H = Wb.H, W = Wb.W, gh = H / 14, gw = W / 14;
  const int D = Wb.dim, C = Wb.features;
  const int npatch = gh * gw, seq = npatch + 1, seq_pad = pad_to(seq, 8);
  const long long rows = static_cast<long long>(B) * seq;
  // ---- tokens: normalise + 14x14 patch gather, patch-embed GEMM, cls + pos (vision_transformer.py:212-219)
  bf16* a0 = static_cast<bf16*>(c.alloc(static_cast<size_t>(B) * npatch * 592 * 2));
  if (c.live()) c.chk(pf_patch_im2col(images, B, H, W, a0, 592, c.stream));
  float* patch = static_cast<float*>(c.alloc(static_cast<size_t>(B) * npatch * D * 4));
  linear(c, Wb.patch, a0, static_cast<long long>(B) * npatch, 592, 592, patch, 1, D);
  float* x = static_cast<float*>(c.alloc(static_cast<size_t>(rows) * D * 4));
  if (c.live()) c.chk(pf_assemble_tokens(patch, Wb.cls, Wb.pos, B, npatch, D, x, c.stream));
  c.tap_out("tokens", x, 1, rows, D, D);
  bf16* hbuf = static_cast<bf16*>(c.alloc(static_cast<size_t>(rows) * D * 2));
  bf16* qk = static_cast<bf16*>(c.alloc(static_cast<size_t>(rows) * 2 * D * 2));
  bf16* vt = static_cast<bf16*>(c.alloc(static_cast<size_t>(B) * D * seq_pad * 2));
  bf16* att = static_cast<bf16*>(c.alloc(static_cast<size_t>(rows) * D * 2));
  bf16* hid = static_cast<bf16*>(c.alloc(static_cast<size_t>(rows) * 4 * D * 2));
  if (c.live() && seq_pad > seq) {
    // V^T pad columns [seq, seq_pad) are read by the attention kernel's last KV tile (times P = 0): keep them finite
    cudaError_t e = cudaMemset2DAsync(vt + seq, static_cast<size_t>(seq_pad) * 2, 0, static_cast<size_t>(seq_pad - seq) * 2,
                                      static_cast<size_t>(B) * D, static_cast<cudaStream_t>(c.stream));
    if (e != cudaSuccess) c.chk(set_error("cudaMemset2DAsync(vt pad): %s", cudaGetErrorString(e)));
  }
  Map feats[4];
  int nf = 0;
  for (int i = 0; i < Wb.depth; ++i) {
    const pf_vit_block& bw = Wb.blocks[i];
    // x += ls1 * proj(attn(LN(x)));  x += ls2 * fc2(gelu(fc1(LN(x))))   (dinov2/layers/block.py:82-107)
    if (c.live()) c.chk(pf_layernorm(x, D, bw.n1w, bw.n1b, 1e-6f, static_cast<int32_t>(rows), D, hbuf, D, c.stream));
    GemmOpt oq; oq.vt = vt; oq.vt_col0 = 2 * D; oq.vt_seq = seq; oq.vt_seq_pad = seq_pad;
    linear(c, bw.qkv, hbuf, rows, D, D, qk, 0, 2 * D, oq);
    if (c.live()) c.chk(pf_attention(qk, 2 * D, vt, B, seq, seq_pad, Wb.heads, 0.125f, att, D, c.stream));
    GemmOpt o1; o1.gamma = bw.ls1;
    linear(c, bw.proj, att, rows, D, D, x, 1, D, o1);
    if (c.live()) c.chk(pf_layernorm(x, D, bw.n2w, bw.n2b, 1e-6f, static_cast<int32_t>(rows), D, hbuf, D, c.stream));
    GemmOpt of; of.act = PF_ACT_GELU;
    linear(c, bw.fc1, hbuf, rows, D, D, hid, 0, 4 * D, of);
    GemmOpt o2; o2.gamma = bw.ls2;
    linear(c, bw.fc2, hid, rows, 4 * D, 4 * D, x, 1, D, o2);
    char nm[16];
    snprintf(nm, sizeof(nm), "block%d", i);
    c.tap_out(nm, x, 1, rows, D, D);
    if (i >= Wb.depth - 4) {
      // get_intermediate_layers(x, 4): final LayerNorm of the LAST four blocks' patch tokens, cls dropped
      // (dpt.py:149, vision_transformer.py:297-321) - one batched launch over all images
      Map f = c.map(B, gh, gw, D);
      if (c.live()) c.chk(pf_layernorm_grouped(x, D, Wb.nw, Wb.nb, 1e-6f, B, seq, 1, npatch, D, f.p, D, c.stream));
      feats[nf++] = f;
    }
  }
  // ---- DPT head (dpt.py:97-130, blocks.py:69-153)
  const int* oc = Wb.out_channels;
  Map lay[4];
  for (int i = 0; i < 4; ++i) {
    Map p = c.map(B, gh, gw, oc[i]);
    linear(c, Wb.proj[i], feats[i].p, feats[i].rows(), D, feats[i].ld, p.p, 0, p.ld);
    if (i == 0 || i == 1) {
      const int k = i == 0 ? 4 : 2;
      Map o = c.map(B, gh * k, gw * k, oc[i]);
      GemmOpt g; g.ps_B = B; g.ps_H = gh; g.ps_W = gw;
      linear(c, i == 0 ? Wb.rs0 : Wb.rs1, p.p, p.rows(), oc[i], p.ld, o.p, 0, o.ld, g);
      lay[i] = o;
    } else if (i == 2) {
      lay[i] = p;
    } else {
      const int oh = (gh - 1) / 2 + 1, ow = (gw - 1) / 2 + 1;
      bf16* col = static_cast<bf16*>(c.alloc(static_cast<size_t>(B) * oh * ow * 9 * oc[3] * 2));
      if (c.live()) c.chk(pf_im2col_3x3_s2(p.p, B, gh, gw, oc[3], p.ld, col, c.stream));
      Map o = c.map(B, oh, ow, oc[3]);
      linear(c, Wb.rs3, col, static_cast<long long>(B) * oh * ow, 9 * oc[3], 9 * oc[3], o.p, 0, o.ld);
      lay[i] = o;
    }
  }
  Map rn[4], rn_relu[4];
  for (int i = 0; i < 4; ++i) {
    rn_relu[i] = c.map(lay[i].B, lay[i].H, lay[i].W, C);
    GemmOpt g; g.relu_copy = &rn_relu[i];
    rn[i] = conv1(c, Wb.rn[i], lay[i], g);
  }
  // ResidualConvUnit: y = conv2(relu(conv1(relu(x)))) + x (+ extra); optionally also relu(y)
  auto rcu = [&](int wi, int u, const Map& xin, const Map& xrelu, const Map* extra, Map* relu_copy) -> Map {
    GemmOpt g1; g1.act = PF_ACT_RELU;
    Map t = conv1(c, Wb.ff_c1[wi - 1][u - 1], xrelu, g1);
    GemmOpt g2; g2.res1 = &xin; g2.res2 = extra;
    if (relu_copy) { *relu_copy = c.map(xin.B, xin.H, xin.W, C); g2.relu_copy = relu_copy; }
    return conv1(c, Wb.ff_c2[wi - 1][u - 1], t, g2);
  };
  // FeatureFusionBlock; the 1x1 out_conv commutes with the bilinear upsample and runs at the low resolution
  auto ffb = [&](int wi, const Map* path, const Map& skip, const Map& skip_relu, int OH, int OW) -> Map {
    Map s = skip, s_relu = skip_relu;
    if (path != nullptr) s = rcu(wi, 1, skip, skip_relu, path, &s_relu);
    Map y = rcu(wi, 2, s, s_relu, nullptr, nullptr);
    y = conv1(c, Wb.ff_out[wi - 1], y);
    return resize(c, y, OH, OW);
  };
  Map p4 = ffb(4, nullptr, rn[3], rn_relu[3], rn[2].H, rn[2].W);
  Map p3 = ffb(3, &p4, rn[2], rn_relu[2], rn[1].H, rn[1].W);
  Map p2 = ffb(2, &p3, rn[1], rn_relu[1], rn[0].H, rn[0].W);
  Map p1 = ffb(1, &p2, rn[0], rn_relu[0], rn[0].H * 2, rn[0].W * 2);
  Map o = conv1(c, Wb.oc1, p1);
  o = resize(c, o, H, W);
  MapF rel = c.mapf(B, H, W, 1, 8);
  if (c.live()) {     // only column 0 is produced (fused 32 -> 1 layer); the 7 pad columns feed zero weights and must be finite
    cudaError_t e = cudaMemsetAsync(rel.p, 0, static_cast<size_t>(B) * H * W * 8 * 4, static_cast<cudaStream_t>(c.stream));
    if (e != cudaSuccess) c.chk(set_error("cudaMemsetAsync(rel): %s", cudaGetErrorString(e)));
  }
  // output_conv2: 3x3 C/2 -> 32 + ReLU (the hooked `out_conv` tap) with the 1x1 32 -> 1 + ReLU fused in its epilogue
  GemmOpt go; go.act = PF_ACT_RELU; go.tail = true; go.act2 = PF_ACT_RELU; go.out3 = rel.p; go.out3_ld = rel.ld;
  Map out_conv = conv1(c, Wb.oc2, o, go);
  Map x_d0 = conv1(c, Wb.conv2, rn[3]);
  c.tap_out("rel", rel.p, 1, static_cast<long long>(B) * H * W, 1, 8);
  float* depth = static_cast<float*>(c.alloc(static_cast<size_t>(B) * H * W * 4));
  Map blocks[4] = {p4, p3, p2, p1};
  metric_head(c, Wb.head, x_d0, blocks, out_conv, &rel, depth);
  if (out != nullptr && !c.dry) {
    out->depth = depth;
    out->feats[0] = to_pf(x_d0); out->feats[1] = to_pf(p4); out->feats[2] = to_pf(p3); out->feats[3] = to_pf(p2);
    out->feats[4] = to_pf(p1); out->feats[5] = to_pf(out_conv);
  }
}

// ---------------------------------------------------------------------------------------------------- G2L
static void g2l_run(Ctx& c, const pf_fusion& Wf, const pf_map* coarse, pf_map* outs) {
  const int WS = 12;
  for (int i = 0; i < 6; ++i) {
    const pf_g2l_level& L = Wf.g2l[i];
    const Map f = from_pf(coarse[i]);
    const int cc = L.C, h = f.H, w = f.W, n = h * w;
    const int Hp = (h + WS - 1) / WS * WS, Wp = (w + WS - 1) / WS * WS;
    if (!c.dry && L.ape_rows != n && !c.err)
      c.err = set_error("guided_fusion.num_patches[%d] = %d does not match the %dx%d coarse map", i, L.ape_rows, h, w);
    float* x = static_cast<float*>(c.alloc(static_cast<size_t>(n) * cc * 4));
    if (c.live()) c.chk(pf_g2l_embed(f.p, f.ld, L.ape, n, cc, x, c.stream));
    bf16* npad = static_cast<bf16*>(c.alloc(static_cast<size_t>(Hp) * Wp * cc * 2));
    bf16* qkv = static_cast<bf16*>(c.alloc(static_cast<size_t>(Hp) * Wp * 3 * cc * 2));
    bf16* att = static_cast<bf16*>(c.alloc(static_cast<size_t>(Hp) * Wp * cc * 2));
    float* prj = static_cast<float*>(c.alloc(static_cast<size_t>(Hp) * Wp * cc * 4));
    bf16* hb = static_cast<bf16*>(c.alloc(static_cast<size_t>(n) * cc * 2));
    bf16* hid = static_cast<bf16*>(c.alloc(static_cast<size_t>(n) * 4 * cc * 2));
    for (int b = 0; b < L.depth; ++b) {
      const pf_g2l_block& bw = L.blocks[b];
      const int shift = (b % 2 == 0) ? 0 : WS / 2;
      if (c.live()) c.chk(pf_swin_norm_pad(x, bw.n1w, bw.n1b, 1e-5f, h, w, Hp, Wp, cc, npad, c.stream));
      linear(c, bw.qkv, npad, static_cast<long long>(Hp) * Wp, cc, cc, qkv, 0, 3 * cc);
      if (c.live()) c.chk(pf_window_attention(qkv, bw.table, Hp, Wp, cc, L.heads, shift, att, c.stream));
      linear(c, bw.proj, att, static_cast<long long>(Hp) * Wp, cc, cc, prj, 1, cc);
      if (c.live()) c.chk(pf_swin_residual_crop(x, prj, h, w, Wp, cc, c.stream));
      if (c.live()) c.chk(pf_layernorm(x, cc, bw.n2w, bw.n2b, 1e-5f, n, cc, hb, cc, c.stream));
      GemmOpt g1; g1.act = PF_ACT_GELU;
      linear(c, bw.fc1, hb, n, cc, cc, hid, 0, 4 * cc, g1);
      GemmOpt g2; g2.gamma = L.ones;
      linear(c, bw.fc2, hid, n, 4 * cc, 4 * cc, x, 1, cc, g2);
    }
    Map o = c.map(1, h, w, cc);
    if (c.live()) c.chk(pf_layernorm(x, cc, L.nw, L.nb, 1e-5f, n, cc, o.p, o.ld, c.stream));
    if (!c.dry && outs) outs[i] = to_pf(o);
  }
}

// ---------------------------------------------------------------------------------------------------- fusion
static void fusion_run(Ctx& c, const pf_fusion& Wf, const float* crops, const float* boxes, int T,
                       const float* fine_depth, const pf_map* fine_feats, const float* coarse_depth,
                       const pf_map* coarse_feats, const pf_map* g2l_maps, float* depth_out) {
  const int H = Wf.H, W = Wf.W;
  // ROI crop-zoom of the whole-image coarse maps + fused 3x3 convs with the fine maps (patchfusion.py:240-267)
  Map guide[5];
  for (int i = 0; i < 5; ++i) {
    const Map cf = from_pf(coarse_feats[i]);
    const Map ff = from_pf(fine_feats[i]);
    Map roi = c.map(T, cf.H, cf.W, cf.C);
    if (c.live())
      c.chk(pf_roi_crop_zoom(cf.p, 0, cf.H, cf.W, pad_to(cf.C, 8), cf.ld, boxes, T, static_cast<float>(cf.H) / H, roi.p,
                             roi.ld, 0, c.stream));
    const Map* a[2] = {&roi, &ff};
    guide[i] = conv(c, Wf.fc[i], a, 2);
  }
  float* droi = static_cast<float*>(c.alloc(static_cast<size_t>(T) * H * W * 4));
  if (c.live()) c.chk(pf_roi_crop_zoom(coarse_depth, 1, H, W, 1, 1, boxes, T, 1.0f, droi, 1, 0, c.stream));
  Map u = c.map(T, H, W, 5);
  if (c.live()) c.chk(pf_pack_unet_input(droi, fine_depth, crops, T, H, W, u.p, u.ld, c.stream));
  // encoder (guided_fusion_model.py:179-184)
  GemmOpt gr; gr.act = PF_ACT_RELU;
  Map x = conv1(c, Wf.inc[0], u, gr);
  x = conv1(c, Wf.inc[1], x, gr);
  Map enc[6];
  enc[5] = x;
  for (int i = 0; i < 5; ++i) {
    Map p = c.map(T, x.H / 2, x.W / 2, x.C);
    if (c.live()) c.chk(pf_maxpool2(x.p, T, x.H, x.W, pad_to(x.C, 8), x.ld, p.p, p.ld, c.stream));
    x = conv1(c, Wf.down[i][0], p, gr);
    x = conv1(c, Wf.down[i][1], x, gr);
    enc[4 - i] = x;
  }
  // decoder, low -> high resolution (guided_fusion_model.py:188-205)
  Map outs[6], prev;
  for (int i = 0; i < 6; ++i) {
    const Map gm = from_pf(g2l_maps[i]);
    const int h = gm.H, w = gm.W;
    // F.interpolate(enc / previous level / guide, size=(h, w), bilinear, align_corners=True) feeding the 3x3 convs
    // (guided_fusion_model.py:98-99,191-203)
    Map e = enc[i];
    if (i > 0) {
      const Map* a[3] = {&enc[i], &prev, &guide[i - 1]};
      e = conv_resampled(c, Wf.up[i - 1][0], a, 3, h, w, gr);
      e = conv1(c, Wf.up[i - 1][1], e, gr);
    }
    Map cr = c.map(T, h, w, gm.C);
    if (c.live())
      c.chk(pf_roi_crop_zoom(gm.p, 0, h, w, pad_to(gm.C, 8), gm.ld, boxes, T, static_cast<float>(h) / H, cr.p, cr.ld, 0, c.stream));
    const Map* a2[2] = {&e, &cr};
    Map y = conv_resampled(c, Wf.cv[i][0], a2, 2, h, w, gr);
    prev = conv1(c, Wf.cv[i][1], y, gr);
    outs[i] = prev;
    char nm[16];
    snprintf(nm, sizeof(nm), "fuse%d", i);
    c.tap_out(nm, prev.p, 0, prev.rows(), prev.C, prev.ld);
  }
  metric_head(c, Wf.head, outs[0], &outs[1], outs[5], nullptr, depth_out);
}

static Ctx make_ctx(void* ws, size_t bytes, bool dry, void* stream, pf_tap_fn tap, void* user) {
  Ctx c;
  c.base = static_cast<uint8_t*>(ws); c.cap = bytes; c.off = 0; c.dry = dry; c.stream = stream; c.err = 0;
  c.tap = tap; c.tap_user = user;
  return c;
}

}  // namespace pf

using namespace pf;

extern "C" {

size_t pf_branch_workspace_bytes(const pf_branch* w, int32_t B) {
  Ctx c = make_ctx(nullptr, 0, true, nullptr, nullptr, nullptr);
  branch_run(c, *w, nullptr, B, nullptr);
  return c.off + 256;
}

int pf_branch_forward(const pf_branch* w, const float* images, int32_t B, void* ws, size_t ws_bytes, pf_branch_out* out,
                      pf_tap_fn tap, void* tap_user, void* stream) {
  if (!w || !images || !ws || !out || B < 1) return set_error("pf_branch_forward: null argument");
  if (w->H % 14 || w->W % 14) return set_error("pf_branch_forward: patch_process_shape must be a multiple of 14");
  if (reinterpret_cast<uintptr_t>(ws) & 255) return set_error("pf_branch_forward: workspace must be 256-byte aligned");
  Ctx c = make_ctx(ws, ws_bytes, false, stream, tap, tap_user);
  branch_run(c, *w, images, B, out);
  return c.err;
}

size_t pf_g2l_workspace_bytes(const pf_fusion* w, const pf_map* coarse_feats) {
  Ctx c = make_ctx(nullptr, 0, true, nullptr, nullptr, nullptr);
  g2l_run(c, *w, coarse_feats, nullptr);
  return c.off + 256;
}

int pf_g2l_forward(const pf_fusion* w, const pf_map* coarse_feats, void* ws, size_t ws_bytes, pf_map* out, void* stream) {
  if (!w || !coarse_feats || !ws || !out) return set_error("pf_g2l_forward: null argument");
  if (reinterpret_cast<uintptr_t>(ws) & 255) return set_error("pf_g2l_forward: workspace must be 256-byte aligned");
  Ctx c = make_ctx(ws, ws_bytes, false, stream, nullptr, nullptr);
  g2l_run(c, *w, coarse_feats, out);
  return c.err;
}

size_t pf_fusion_workspace_bytes(const pf_fusion* w, int32_t T, const pf_map* g2l_maps) {
  Ctx c = make_ctx(nullptr, 0, true, nullptr, nullptr, nullptr);
  // shapes only: the fine / coarse maps share the G2L maps' geometry and channel counts
  pf_map fine[6];
  for (int i = 0; i < 6; ++i) { fine[i] = g2l_maps[i]; fine[i].B = T; fine[i].ptr = nullptr; }
  fusion_run(c, *w, nullptr, nullptr, T, nullptr, fine, nullptr, g2l_maps, g2l_maps, nullptr);
  return c.off + 256;
}

int pf_fusion_forward(const pf_fusion* w, const float* crops, const float* boxes, int32_t T, const float* fine_depth,
                      const pf_map* fine_feats, const float* coarse_depth, const pf_map* coarse_feats,
                      const pf_map* g2l_maps, void* ws, size_t ws_bytes, float* depth_out, pf_tap_fn tap,
                      void* tap_user, void* stream) {
  if (!w || !crops || !boxes || !fine_depth || !fine_feats || !coarse_depth || !coarse_feats || !g2l_maps || !ws ||
      !depth_out || T < 1)
    return set_error("pf_fusion_forward: null argument");
  if (reinterpret_cast<uintptr_t>(ws) & 255) return set_error("pf_fusion_forward: workspace must be 256-byte aligned");
  Ctx c = make_ctx(ws, ws_bytes, false, stream, tap, tap_user);
  fusion_run(c, *w, crops, boxes, T, fine_depth, fine_feats, coarse_depth, coarse_feats, g2l_maps, depth_out);
  return c.err;
}

}  // extern "C"
