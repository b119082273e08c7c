// Fused attention for the DINOv2 blocks: softmax(Q K^T * scale) V without materialising the (B,heads,N,N) score
// tensor the reference builds (`dinov2/layers/attention.py:53-59`).  head_dim = 64, no mask, N = 1037 tokens.
//
// Persistent kernel, one CTA per SM, work item = (image b, head h, PAIR of 128-query tiles).  Both contractions run on
// tcgen05 with every accumulator in tensor memory (512 columns):
//   S_t[128x128] = Q_t[128x64] . K_j[128x64]^T     t = 0,1   (TMA loads straight from the qkv GEMM output, K-major)
//   O_t[128x64] += P_t[128x128] . V_j[128x64]      P_t is the A operand IN TMEM (bf16, written by the softmax warps
//                                                  with tcgen05.st - it never touches shared memory); V^T tiles come
//                                                  from the transposed copy the qkv GEMM epilogue writes (K-major B).
//   TMEM columns: S_0 0..127 | S_1 128..255 | O_0 256..319 | O_1 320..383 | P_0 384..447 | P_1 448..511
// At head_dim 64 the softmax, not the tensor pipe, bounds attention (128 exps per row per 2x256-clk MMA), and what
// bounded the softmax was LATENCY (TMEM load / store round trips, barrier hand-offs) with only two warps per
// scheduler (profiles/r02_attention_*).  So the 16 softmax warps (4 per scheduler) STREAM the scores: two threads per
// query row (TMEM lane == row; the warps w and w+4 of a quadrant own columns 0-63 / 64-127), 32 scores at a time
// TMEM -> registers -> exp2 -> bf16 P -> TMEM, using the running maximum of the PREVIOUS blocks (known before the
// scores arrive, so there is no max pass).  The block's own maximum is reduced on the side, exchanged between the two
// half-row warps through shared memory, and only if some row's maximum moved by more than 2^8 is O rescaled (by the
// rows' own threads) and the block redone with the new maximum - exact, and P never exceeds 2^8.  Block 0 (no
// running maximum yet) takes the two passes.  exp2 runs on MUFU for most elements and on the FMA pipe (Cody-Waite +
// cubic, packed FFMA2) for the rest; row sums are packed FADD2.  Warp 16 = TMA producer, warps 17 / 18 =
// single-thread QK^T / PV issuers (separate, so a QK^T is never queued behind a wait for P).  K_j / V_j are shared
// by the two Q tiles.  The last KV block (1037 = 8*128 + 13 keys) runs with N = 16 / K = 16, not a padded 128.
#include <stdlib.h>

#include "pf_common.cuh"
#include "pf_kernels.h"

namespace pf {

constexpr int kQTile = 128, kKTile = 128, kHd = 64;
constexpr int kSoftmaxWarps = 16;
constexpr int kAttnThreads = (kSoftmaxWarps + 3) * 32;   // + TMA producer + QK^T issuer + PV issuer
constexpr int kKS = 3, kVS = 3;                      // K / V smem ring depths
constexpr int kTileBytes = 16384;
constexpr int kOffQ = 0;                             // 2 x [128 q][64]
constexpr int kOffK = 2 * kTileBytes;                // kKS x [128 keys][64]
constexpr int kOffV = kOffK + kKS * kTileBytes;      // kVS x 2 x [64 d][64 keys]
constexpr int kOffBar = kOffV + kVS * kTileBytes;
constexpr int kOffXch = kOffBar + 512;                // fp32 [2 parities][2 tiles][2 halves][128 rows] block maxima + [2][2][128] row sums
constexpr int kAttnSmem = 1024 + kOffXch + 6 * 1024;
constexpr uint32_t kColS = 0, kColO = 256, kColP = 384;
constexpr float kRescaleThreshold = 8.0f;            // log2 units: P stays below 2^8, far inside bf16 / fp32 range
constexpr int kPolyPairs = 5;                        // of the 16 pairs per 32-score chunk, evaluated on the FMA pipe

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x for two values on the FMA/ALU pipes: n = round(x), f = x - n in [-0.5, 0.5], 2^f by a cubic minimax
// (max rel err 7.5e-5, far below the bf16 rounding of P), exponent patched in with an integer add.  x >= -126.
__device__ __forceinline__ uint64_t ex2_poly2(uint64_t x) {
  float x0, x1;
  unpack2f(x, x0, x1);
  const uint64_t xc = pack2f(fmaxf(x0, -126.0f), fmaxf(x1, -126.0f));
  const uint64_t magic = pack2f(12582912.0f, 12582912.0f);          // 1.5 * 2^23: low mantissa bits = round(x)
  const uint64_t t = add2(xc, magic);
  const uint64_t f = sub2(xc, sub2(t, magic));
  uint64_t p = fma2(f, pack2f(0.05517164617776871f, 0.05517164617776871f), pack2f(0.2426111251115799f, 0.2426111251115799f));
  p = fma2(p, f, pack2f(0.6932609677314758f, 0.6932609677314758f));
  p = fma2(p, f, pack2f(0.9999280571937561f, 0.9999280571937561f));
  float t0, t1, p0, p1;
  unpack2f(t, t0, t1);
  unpack2f(p, p0, p1);
  const uint32_t r0 = __float_as_uint(p0) + (__float_as_uint(t0) << 23);
  const uint32_t r1 = __float_as_uint(p1) + (__float_as_uint(t1) << 23);
  return pack2(r0, r1);
}

// 32 scores of one row -> 16 packed bf16x2 probabilities; row-sum partials in two packed accumulators.
__device__ __forceinline__ void softmax_chunk(const uint32_t (&s)[32], uint64_t sc2, uint64_t nm2, uint64_t& sum_a,
                                              uint64_t& sum_b, uint32_t (&pk)[16]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const uint64_t x = fma2(pack2(s[2 * i], s[2 * i + 1]), sc2, nm2);
    uint64_t e;
    // Bresenham spread of the polynomial pairs over the chunk so MUFU and FMA work interleave
    if (((i + 1) * kPolyPairs) / 16 != (i * kPolyPairs) / 16) {
      e = ex2_poly2(x);
    } else {
      float x0, x1;
      unpack2f(x, x0, x1);
      e = pack2f(ex2_approx(x0), ex2_approx(x1));
    }
    if (i & 1) sum_b = add2(sum_b, e); else sum_a = add2(sum_a, e);
    float e0, e1;
    unpack2f(e, e0, e1);
    pk[i] = pack_bf16(e0, e1);
  }
}

// Debug timeline (build with -DPF_ATTN_TRACE, tools/attn_trace.py): CTA 0 records (role, tile, block, event, clock) tuples.
#ifdef PF_ATTN_TRACE
__device__ unsigned long long g_attn_trace[4096 * 2];
__device__ unsigned int g_attn_trace_n;
// no atomics on the traced path: every recording thread owns a 1024-entry region (slot = role/tile) and a private
// counter, so a trace point costs one clock read and two stores
__device__ __forceinline__ void attn_trace(int role, int t, int j, int ev, unsigned int& cnt) {
  if (blockIdx.x != 0) return;
  const unsigned int region = role == 0 ? t : (role + 1);
  if (cnt < 1024) {
    const unsigned int i = region * 1024 + cnt++;
    g_attn_trace[2 * i] = (static_cast<unsigned long long>(role) << 48) | (static_cast<unsigned long long>(t) << 32) |
                          (static_cast<unsigned long long>(j) << 16) | static_cast<unsigned long long>(ev) | (1ull << 63);
    g_attn_trace[2 * i + 1] = clock64();
  }
}
#define ATTN_TRACE(role, t, j, ev) attn_trace(role, t, j, ev, trace_cnt)
#define ATTN_TRACE_DECL unsigned int trace_cnt = 0;
#else
#define ATTN_TRACE(role, t, j, ev)
#define ATTN_TRACE_DECL
#endif

struct AttnParams {
  CUtensorMap tmQK;   // 3-D {2*D, seq, B}, box {64, 128, 1}
  CUtensorMap tmVt;   // 2-D {seq_pad, B*heads*64}, box {64, 64}
  int B, seq, heads, D;
  int n_pairs, n_items;
  float scale_log2;   // scale * log2(e)
  __nv_bfloat16* out;
  int out_ld;
};

__global__ void __launch_bounds__(kAttnThreads, 1) pf_attention_kernel(const __grid_constant__ AttnParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + kOffQ;
  uint8_t* sK = smem + kOffK;
  uint8_t* sV = smem + kOffV;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* k_full = bars + 2;                  // [kKS]
  uint64_t* k_empty = k_full + kKS;
  uint64_t* v_full = k_empty + kKS;             // [kVS]
  uint64_t* v_empty = v_full + kVS;
  uint64_t* s_full = v_empty + kVS;             // [2]  QK^T of the tile's current block is complete
  uint64_t* s_free = s_full + 2;                // [2]  the softmax warps hold the scores in registers
  uint64_t* p_full = s_free + 2;                // [2]  P of the current block is in TMEM
  uint64_t* o_done = p_full + 2;                // [2]  PV of the current block is complete
  uint64_t* o_free = o_done + 2;                // [2]  the item's O has been read out
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seq = P.seq;
  const int nkv = (seq + kKTile - 1) / kKTile;
  ATTN_TRACE_DECL
  pdl_launch_dependents();

  if (warp == kSoftmaxWarps && lane == 0) {
    prefetch_tmap(&P.tmQK);
    prefetch_tmap(&P.tmVt);
    mbar_init(q_full, 1); mbar_init(q_empty, 1);
    for (int s = 0; s < kKS; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
    for (int s = 0; s < kVS; ++s) { mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1); mbar_init(&s_free[t], 8); mbar_init(&p_full[t], 8);
      mbar_init(&o_done[t], 1); mbar_init(&o_free[t], 8);
    }
    fence_barrier_init();
  }
  if (warp == kSoftmaxWarps + 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == kSoftmaxWarps) {
    // ===================== TMA producer =====================
    // The issuing warps walk their loops as WHOLE warps (uniform control flow keeps barrier addresses, descriptors and
    // counters in uniform registers) and one elected lane issues: inside `if (lane == 0)` every tcgen05.mma / TMA
    // paid an ELECT / R2UR.BROADCAST round trip (~13 instructions) on the serial S -> P -> PV critical path.
    {
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0, qe_ph = 0;
      for (int item = blockIdx.x; item < P.n_items; item += gridDim.x) {
        const int pair = item % P.n_pairs, bh = item / P.n_pairs;
        const int h = bh % P.heads, b = bh / P.heads;
        const int q0 = pair * 2 * kQTile;
        mbar_wait(q_empty, qe_ph ^ 1); qe_ph ^= 1;
        if (elect_one()) {
          mbar_expect_tx(q_full, 2 * kTileBytes);
          tma_load_3d(sQ, &P.tmQK, q_full, h * kHd, q0, b);
          tma_load_3d(sQ + kTileBytes, &P.tmQK, q_full, h * kHd, q0 + kQTile, b);   // rows >= seq are zero-filled
        }
        for (int j = 0; j < nkv; ++j) {
          mbar_wait(&k_empty[ks], kph ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&k_full[ks], kTileBytes);
            tma_load_3d(sK + ks * kTileBytes, &P.tmQK, &k_full[ks], P.D + h * kHd, j * kKTile, b);
          }
          if (++ks == kKS) { ks = 0; kph ^= 1; }
          mbar_wait(&v_empty[vs], vph ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&v_full[vs], kTileBytes);
            tma_load_2d(sV + vs * kTileBytes, &P.tmVt, &v_full[vs], j * kKTile, (b * P.heads + h) * kHd);
            tma_load_2d(sV + vs * kTileBytes + 8192, &P.tmVt, &v_full[vs], j * kKTile + 64, (b * P.heads + h) * kHd);
          }
          if (++vs == kVS) { vs = 0; vph ^= 1; }
        }
      }
    }
  } else if (warp == kSoftmaxWarps + 1) {
    // ===================== QK^T issuer: S_t = Q_t K_j^T as soon as the softmax warps have taken S_t =====================
    // (a separate thread from the PV issuer, so a QK^T is never queued behind a wait for P)
    {
      int ks = 0;
      uint32_t kph = 0, qf_ph = 0;
      uint32_t sfree_ph[2] = {0, 0};
      for (int item = blockIdx.x; item < P.n_items; item += gridDim.x) {
        const int q0 = (item % P.n_pairs) * 2 * kQTile;
        const int nt = (q0 + kQTile < seq) ? 2 : 1;            // Q tiles of this item that hold queries
        mbar_wait(q_full, qf_ph); qf_ph ^= 1;
        tc_fence_after();
        const uint64_t dq0 = umma_desc_k128(smem_u32(sQ));
        for (int j = 0; j < nkv; ++j) {
          const int kv_len = min(kKTile, seq - j * kKTile);
          const uint32_t idesc_s = umma_idesc_bf16(128, (kv_len + 15) & ~15);
          mbar_wait(&k_full[ks], kph);
          tc_fence_after();
          const uint64_t dk = umma_desc_k128(smem_u32(sK + ks * kTileBytes));
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            if (t < nt) {
              mbar_wait(&s_free[t], sfree_ph[t] ^ 1); sfree_ph[t] ^= 1;
              tc_fence_after();
              if (lane == 0) ATTN_TRACE(1, t, j, 0);
              const uint64_t dq = dq0 + t * (kTileBytes >> 4);
              if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + kColS + t * 128, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
                umma_commit(&s_full[t]);
              }
            }
          }
          if (elect_one()) umma_commit(&k_empty[ks]);
          if (++ks == kKS) { ks = 0; kph ^= 1; }
        }
        if (elect_one()) umma_commit(q_empty);                 // every QK^T of the item is issued: Q may be replaced
      }
    }
  } else if (warp == kSoftmaxWarps + 2) {
    // ===================== PV issuer: O_t += P_t V_j (A operand = P in TMEM) =====================
    {
      const uint32_t idesc_o = umma_idesc_bf16(128, kHd);
      int vs = 0;
      uint32_t vph = 0;
      uint32_t pfull_ph[2] = {0, 0}, ofree_ph[2] = {0, 0};
      for (int item = blockIdx.x; item < P.n_items; item += gridDim.x) {
        const int q0 = (item % P.n_pairs) * 2 * kQTile;
        const int nt = (q0 + kQTile < seq) ? 2 : 1;
        for (int j = 0; j < nkv; ++j) {
          const int nk = (min(kKTile, seq - j * kKTile) + 15) >> 4;     // 16-key MMA steps of this block
          mbar_wait(&v_full[vs], vph);
          tc_fence_after();
          const uint64_t dv0 = umma_desc_k128(smem_u32(sV + vs * kTileBytes));
          const uint64_t dv1 = umma_desc_k128(smem_u32(sV + vs * kTileBytes + 8192));
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            if (t < nt) {
              mbar_wait(&p_full[t], pfull_ph[t]); pfull_ph[t] ^= 1;
              if (j == 0) { mbar_wait(&o_free[t], ofree_ph[t] ^ 1); ofree_ph[t] ^= 1; }
              tc_fence_after();
              if (lane == 0) ATTN_TRACE(2, t, j, 0);
              const uint32_t tO = tmem_base + kColO + t * 64, tP = tmem_base + kColP + t * 64;
              if (elect_one()) {
                if (nk == 8) {
#pragma unroll
                  for (int k = 0; k < 8; ++k)
                    umma_bf16_ts(tO, tP + k * 8, (k < 4 ? dv0 : dv1) + 2 * (k & 3), idesc_o, (j | k) != 0);
                } else {
                  for (int k = 0; k < nk; ++k)
                    umma_bf16_ts(tO, tP + k * 8, (k < 4 ? dv0 : dv1) + 2 * (k & 3), idesc_o, (j | k) != 0);
                }
                umma_commit(&o_done[t]);
              }
            }
          }
          if (elect_one()) umma_commit(&v_empty[vs]);
          if (++vs == kVS) { vs = 0; vph ^= 1; }
        }
      }
    }
  } else {
    // ===================== softmax warps: two threads per query row, scores streamed 32 at a time =====================
    const int t = warp >> 3, hh = (warp >> 2) & 1, q = warp & 3;
    const int r = q * 32 + lane;                               // query row of the tile (== TMEM lane)
    const uint32_t lane_sel = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t tS = tmem_base + lane_sel + kColS + t * 128 + hh * 64;    // my 64 score columns
    const uint32_t tO = tmem_base + lane_sel + kColO + t * 64 + hh * 32;     // my 32 output columns
    const uint32_t tP = tmem_base + lane_sel + kColP + t * 64 + hh * 32;     // my 32 packed-P columns (64 keys)
    float* xch = reinterpret_cast<float*>(smem + kOffXch);                   // [parity][tile][half][128]
    float* xsum = xch + 2 * 2 * 2 * 128;                                     // [tile][half][128]
    const int pair_bar = 1 + t * 4 + q;                        // named barrier of the two warps sharing these rows
    const float scale = P.scale_log2;
    uint32_t sfull_ph = 0, odone_ph = 0;
    for (int item = blockIdx.x; item < P.n_items; item += gridDim.x) {
      const int pair = item % P.n_pairs, bh = item / P.n_pairs;
      const int h = bh % P.heads, b = bh / P.heads;
      const int row0 = pair * 2 * kQTile + t * kQTile;
      if (row0 >= seq) continue;                               // this tile holds no queries: no barrier traffic at all
      const bool warp_active = row0 + q * 32 < seq;            // warp-uniform, same for both half-row warps
      const int tok = row0 + r;
      if (!warp_active) {
        // rows beyond the sequence: keep the tile's barrier protocol, skip the math
        for (int j = 0; j < nkv; ++j) {
          mbar_wait(&s_full[t], sfull_ph); sfull_ph ^= 1;
          if (j > 0) { mbar_wait(&o_done[t], odone_ph); odone_ph ^= 1; }
          __syncwarp();
          if (lane == 0) { mbar_arrive(&s_free[t]); mbar_arrive(&p_full[t]); }
        }
        mbar_wait(&o_done[t], odone_ph); odone_ph ^= 1;
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_free[t]);
        continue;
      }
      float m_run = -INFINITY, l_run = 0.0f;
      for (int j = 0; j < nkv; ++j) {
        const int kv_len = min(kKTile, seq - j * kKTile);
        const int my_valid = max(0, min(64, kv_len - hh * 64));      // keys of this block among my 64 columns
        float* xm = xch + (((j & 1) * 2 + t) * 2) * 128;             // [half][128] of this parity / tile
        mbar_wait(&s_full[t], sfull_ph); sfull_ph ^= 1;
        tc_fence_after();
        if (q == 0 && lane == 0 && hh == 0) ATTN_TRACE(0, t, j, 0);
        bool pv_waited = j == 0;                               // PV of block j-1 must finish before P / O are touched
        auto wait_pv = [&]() {
          if (!pv_waited) { mbar_wait(&o_done[t], odone_ph); odone_ph ^= 1; tc_fence_after(); pv_waited = true; }
        };
        // exchange a per-row value with the warp that owns the other 64 columns of the same rows
        auto pair_max = [&](float v) -> float {
          xm[hh * 128 + r] = v;
          asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
          return fmaxf(v, xm[(hh ^ 1) * 128 + r]);
        };
        // raw (unscaled) maximum over my valid columns: block 0 only (no running maximum yet)
        auto max_pass = [&]() -> float {
          float mx = -INFINITY;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            if (c * 32 < my_valid) {
              uint32_t v[32];
              tmem_ld32(tS + c * 32, v);
              tmem_ld_wait();
              if (my_valid - c * 32 >= 32) {
                float a0 = -INFINITY, a1 = -INFINITY;
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                  a0 = max3(a0, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
                  a1 = max3(a1, __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
                }
                mx = max3(mx, a0, a1);
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (c * 32 + i < my_valid) mx = fmaxf(mx, __uint_as_float(v[i]));
              }
            }
          }
          return mx;
        };
        // P = exp2(S * scale - m) for my 64 columns, 32 at a time; returns the block's row sum and raw maximum
        auto stream_pass = [&](float m, float& blk_sum, float& raw_max) {
          const uint64_t sc2 = pack2f(scale, scale), nm2 = pack2f(-m, -m);
          uint64_t sum_a = 0, sum_b = 0;                       // bit pattern of (+0.0f, +0.0f)
          float ls = 0.0f, mx = -INFINITY;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t pk[16];
            const int vc = my_valid - c * 32;                  // warp-uniform
            if (vc >= 32) {
              uint32_t v[32];
              tmem_ld32(tS + c * 32, v);
              tmem_ld_wait();
              float a0 = -INFINITY, a1 = -INFINITY;
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                a0 = max3(a0, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
                a1 = max3(a1, __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
              }
              mx = max3(mx, a0, a1);
              softmax_chunk(v, sc2, nm2, sum_a, sum_b, pk);
            } else if (vc > 0) {
              uint32_t v[32];
              tmem_ld32(tS + c * 32, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const bool k0 = 2 * i < vc, k1 = 2 * i + 1 < vc;
                const float e0 = k0 ? ex2_approx(fmaf(__uint_as_float(v[2 * i]), scale, -m)) : 0.0f;
                const float e1 = k1 ? ex2_approx(fmaf(__uint_as_float(v[2 * i + 1]), scale, -m)) : 0.0f;
                if (k0) mx = fmaxf(mx, __uint_as_float(v[2 * i]));
                if (k1) mx = fmaxf(mx, __uint_as_float(v[2 * i + 1]));
                ls += e0 + e1;
                pk[i] = pack_bf16(e0, e1);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) pk[i] = 0u;
            }
            if (c == 0) wait_pv();
            tmem_st16(tP + c * 16, pk);
          }
          float x0, x1, y0, y1;
          unpack2f(sum_a, x0, x1);
          unpack2f(sum_b, y0, y1);
          blk_sum = ls + ((x0 + x1) + (y0 + y1));
          raw_max = mx;
        };
        float blk_sum, raw_max;
        if (j == 0) {
          m_run = pair_max(max_pass()) * scale;                // the row has at least one key: finite
          stream_pass(m_run, blk_sum, raw_max);
        } else {
          stream_pass(m_run, blk_sum, raw_max);                // with the STALE maximum: no max pass on the critical path
          const float m_blk = pair_max(raw_max) * scale;       // -inf * scale = -inf when I own no valid column
          if (__any_sync(0xffffffffu, m_blk - m_run > kRescaleThreshold)) {
            // some row's maximum moved by more than 2^8 (rare): rescale O (PV of block j-1 is complete, wait_pv ran),
            // then redo the block with the new maximum.  Both half-row warps take this branch together.
            const float m_new = fmaxf(m_run, m_blk);
            const float alpha = ex2_approx(m_run - m_new);
            m_run = m_new;
            l_run *= alpha;
            uint32_t o[32];
            tmem_ld32(tO, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(tO, o);
            stream_pass(m_run, blk_sum, raw_max);
          }
        }
        l_run += blk_sum;
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { mbar_arrive(&s_free[t]); mbar_arrive(&p_full[t]); }
        if (q == 0 && lane == 0 && hh == 0) ATTN_TRACE(0, t, j, 4);
      }
      // ---- item epilogue: combine the two half-row sums, O / l -> bf16 -> global
      float* xs = xsum + (t * 2) * 128;
      xs[hh * 128 + r] = l_run;
      mbar_wait(&o_done[t], odone_ph); odone_ph ^= 1;
      tc_fence_after();
      uint32_t o[32];
      tmem_ld32(tO, o);
      tmem_ld_wait();
      tc_fence_before();
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
      const float l_tot = l_run + xs[(hh ^ 1) * 128 + r];
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[t]);                  // PV of the next item may overwrite O
      if (tok < seq) {
        const float inv = 1.0f / l_tot;
        __nv_bfloat16* op = P.out + (static_cast<long long>(b) * seq + tok) * P.out_ld + h * kHd + hh * 32;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          *reinterpret_cast<uint4*>(op + i) = make_uint4(
              pack_bf16(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv),
              pack_bf16(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv),
              pack_bf16(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv),
              pack_bf16(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv));
        }
      }
      // the pair barrier at the top of the next item's first exchange orders the xsum reads before its next write
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kSoftmaxWarps + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace pf

using namespace pf;

#ifdef PF_ATTN_TRACE
extern "C" int pf_attention_trace_read(unsigned long long* out, unsigned int* n) {
  cudaDeviceSynchronize();
  *n = 4096;
  cudaMemcpyFromSymbol(out, g_attn_trace, sizeof(unsigned long long) * 4096 * 2);
  return 0;
}
#endif

extern "C" int pf_attention(const void* qk, int32_t qk_ld, const void* vt, int32_t B, int32_t seq, int32_t seq_pad,
                            int32_t heads, float scale, void* out, int32_t out_ld, void* stream) {
  static bool attr_done[kMaxDevices] = {false};
  static int sm_count[kMaxDevices] = {0};
  const int dev = current_device();
  if (!attr_done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(pf_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
    if (e != cudaSuccess) return set_error("cudaFuncSetAttribute(pf_attention_kernel): %s", cudaGetErrorString(e));
    cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
    attr_done[dev] = true;
  }
  const int D = heads * kHd;
  if (B < 1 || seq < 1 || heads < 1) return set_error("pf_attention: empty problem");
  if (qk_ld % 8 || seq_pad % 8 || out_ld % 8) return set_error("pf_attention: strides must be multiples of 8");
  if (seq_pad < seq) return set_error("pf_attention: seq_pad %d < seq %d", seq_pad, seq);
  AttnParams P;
  if (tmap_3d_bf16(&P.tmQK, qk, 2 * D, seq, B, qk_ld, static_cast<uint64_t>(seq) * qk_ld, 64, 128, 1)) return 1;
  if (tmap_2d_bf16(&P.tmVt, vt, seq_pad, static_cast<uint64_t>(B) * heads * kHd, seq_pad, 64, 64)) return 1;
  P.B = B; P.seq = seq; P.heads = heads; P.D = D;
  P.n_pairs = (seq + 2 * kQTile - 1) / (2 * kQTile);
  P.n_items = B * heads * P.n_pairs;
  P.scale_log2 = scale * 1.4426950408889634f;
  P.out = static_cast<__nv_bfloat16*>(out);
  P.out_ld = out_ld;
  note_work(4.0 * B * heads * static_cast<double>(seq) * seq * kHd, "attention B%d heads%d seq%d", B, heads, seq);
  const int grid = P.n_items < sm_count[dev] ? P.n_items : sm_count[dev];
  cudaError_t le = launch_pdl(pf_attention_kernel, dim3(grid), dim3(kAttnThreads), kAttnSmem, static_cast<cudaStream_t>(stream), P);
  if (le != cudaSuccess) return set_error("pf_attention_kernel launch: %s", cudaGetErrorString(le));
  return check_launch("pf_attention_kernel");
}
