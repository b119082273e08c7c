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
// Warps 0-3 / 4-7: softmax of Q tile 0 / 1, ONE THREAD PER QUERY ROW (TMEM lane == row): the 128 scores of a KV block
// are read from TMEM once into registers, the S buffer is handed back to the tensor core immediately (QK^T of block
// j+1 overlaps softmax of block j), row max by 3-input FMNMX, exp2 on MUFU for most elements and on the FMA pipe
// (Cody-Waite + cubic, packed FFMA2) for the rest - at head_dim 64 the MUFU pipe, not the tensor pipe, bounds
// attention - packed FADD2 row sums, bf16 P back into TMEM.  O stays in TMEM across KV blocks; it is rescaled (by
// the row's own thread) only when the running max moves by more than 2^8 (lazy rescale with a stale max, exact after
// the final 1/l normalisation).  Warp 8 = TMA producer, warps 9 / 10 = single-thread QK^T / PV issuers.  K_j / V_j are shared by
// the two Q tiles.  The last KV block (1037 = 8*128 + 13 keys) runs with N = 16 / K = 16 instead of a padded 128.
#include <stdlib.h>

#include "pf_common.cuh"
#include "pf_kernels.h"

namespace pf {

constexpr int kQTile = 128, kKTile = 128, kHd = 64;
constexpr int kAttnThreads = 352;                   // 8 softmax warps + TMA + QK^T issuer + PV issuer
constexpr int kKS = 3, kVS = 3;                      // K / V smem ring depths
constexpr int kTileBytes = 16384;
constexpr int kOffQ = 0;                             // 2 x [128 q][64]
constexpr int kOffK = 2 * kTileBytes;                // kKS x [128 keys][64]
constexpr int kOffV = kOffK + kKS * kTileBytes;      // kVS x 2 x [64 d][64 keys]
constexpr int kOffBar = kOffV + kVS * kTileBytes;
constexpr int kAttnSmem = 1024 + kOffBar + 512;
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
__device__ __forceinline__ void attn_trace(int role, int t, int j, int ev) {
  if (blockIdx.x != 0) return;
  unsigned int i = atomicAdd(&g_attn_trace_n, 1u);
  if (i < 4096) {
    g_attn_trace[2 * i] = (static_cast<unsigned long long>(role) << 48) | (static_cast<unsigned long long>(t) << 32) |
                          (static_cast<unsigned long long>(j) << 16) | static_cast<unsigned long long>(ev);
    g_attn_trace[2 * i + 1] = clock64();
  }
}
#define ATTN_TRACE(role, t, j, ev) attn_trace(role, t, j, ev)
#else
#define ATTN_TRACE(role, t, j, ev)
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
  pdl_launch_dependents();

  if (warp == 8 && lane == 0) {
    prefetch_tmap(&P.tmQK);
    prefetch_tmap(&P.tmVt);
    mbar_init(q_full, 1); mbar_init(q_empty, 1);
    for (int s = 0; s < kKS; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
    for (int s = 0; s < kVS; ++s) { mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1); mbar_init(&s_free[t], 4); mbar_init(&p_full[t], 4);
      mbar_init(&o_done[t], 1); mbar_init(&o_free[t], 4);
    }
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 8) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0, qe_ph = 0;
      for (int item = blockIdx.x; item < P.n_items; item += gridDim.x) {
        const int pair = item % P.n_pairs, bh = item / P.n_pairs;
        const int h = bh % P.heads, b = bh / P.heads;
        const int q0 = pair * 2 * kQTile;
        mbar_wait(q_empty, qe_ph ^ 1); qe_ph ^= 1;
        mbar_expect_tx(q_full, 2 * kTileBytes);
        tma_load_3d(sQ, &P.tmQK, q_full, h * kHd, q0, b);
        tma_load_3d(sQ + kTileBytes, &P.tmQK, q_full, h * kHd, q0 + kQTile, b);   // rows >= seq are zero-filled
        for (int j = 0; j < nkv; ++j) {
          mbar_wait(&k_empty[ks], kph ^ 1);
          mbar_expect_tx(&k_full[ks], kTileBytes);
          tma_load_3d(sK + ks * kTileBytes, &P.tmQK, &k_full[ks], P.D + h * kHd, j * kKTile, b);
          if (++ks == kKS) { ks = 0; kph ^= 1; }
          mbar_wait(&v_empty[vs], vph ^ 1);
          mbar_expect_tx(&v_full[vs], kTileBytes);
          tma_load_2d(sV + vs * kTileBytes, &P.tmVt, &v_full[vs], j * kKTile, (b * P.heads + h) * kHd);
          tma_load_2d(sV + vs * kTileBytes + 8192, &P.tmVt, &v_full[vs], j * kKTile + 64, (b * P.heads + h) * kHd);
          if (++vs == kVS) { vs = 0; vph ^= 1; }
        }
      }
    }
  } else if (warp == 9) {
    // ===================== QK^T issuer: S_t = Q_t K_j^T as soon as the softmax warps have taken S_t =====================
    // (a separate thread from the PV issuer, so a QK^T is never queued behind a wait for P)
    if (lane == 0) {
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
              ATTN_TRACE(1, t, j, 0);
              const uint64_t dq = dq0 + t * (kTileBytes >> 4);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + kColS + t * 128, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
              umma_commit(&s_full[t]);
            }
          }
          umma_commit(&k_empty[ks]);
          if (++ks == kKS) { ks = 0; kph ^= 1; }
        }
        umma_commit(q_empty);                                  // every QK^T of the item is issued: Q may be replaced
      }
    }
  } else if (warp == 10) {
    // ===================== PV issuer: O_t += P_t V_j (A operand = P in TMEM) =====================
    if (lane == 0) {
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
              ATTN_TRACE(2, t, j, 0);
              const uint32_t tO = tmem_base + kColO + t * 64, tP = tmem_base + kColP + t * 64;
              for (int k = 0; k < nk; ++k)
                umma_bf16_ts(tO, tP + k * 8, (k < 4 ? dv0 : dv1) + 2 * (k & 3), idesc_o, (j | k) != 0);
              umma_commit(&o_done[t]);
            }
          }
          umma_commit(&v_empty[vs]);
          if (++vs == kVS) { vs = 0; vph ^= 1; }
        }
      }
    }
  } else {
    // ===================== softmax warps: one thread per query row =====================
    const int t = warp >> 2, q = warp & 3;
    const uint32_t lane_sel = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t tS = tmem_base + lane_sel + kColS + t * 128;
    const uint32_t tO = tmem_base + lane_sel + kColO + t * 64;
    const uint32_t tP = tmem_base + lane_sel + kColP + t * 64;
    const float scale = P.scale_log2;
    uint32_t sfull_ph = 0, odone_ph = 0;
    for (int item = blockIdx.x; item < P.n_items; item += gridDim.x) {
      const int pair = item % P.n_pairs, bh = item / P.n_pairs;
      const int h = bh % P.heads, b = bh / P.heads;
      const int row0 = pair * 2 * kQTile + t * kQTile;
      if (row0 >= seq) continue;                               // this tile holds no queries: no barrier traffic at all
      const bool warp_active = row0 + q * 32 < seq;            // warp-uniform
      const int tok = row0 + q * 32 + lane;
      float m_run = -INFINITY, l_run = 0.0f;
      for (int j = 0; j < nkv; ++j) {
        const int kv_len = min(kKTile, seq - j * kKTile);
        mbar_wait(&s_full[t], sfull_ph); sfull_ph ^= 1;
        if (!warp_active) {                                    // keep the barrier protocol, skip the math
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_free[t]);
          if (j > 0) { mbar_wait(&o_done[t], odone_ph); odone_ph ^= 1; }
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[t]);
          continue;
        }
        tc_fence_after();
        if (q == 0 && lane == 0) ATTN_TRACE(0, t, j, 0);            // S ready
        bool waited = false;
        // O_t *= alpha, issued by the rows' own threads (rare: only when some row's max moved by > 2^8)
        auto rescale = [&](float m_new) {
          const float alpha = ex2_approx(m_run - m_new);
          m_run = m_new;
          l_run *= alpha;
          mbar_wait(&o_done[t], odone_ph); odone_ph ^= 1;      // PV of block j-1 must have landed in O
          waited = true;
          tc_fence_after();
#pragma unroll 1
          for (int part = 0; part < 4; ++part) {               // 16 columns at a time: the 128 scores stay in registers
            uint32_t o[16];
            tmem_ld16(tO + part * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st16(tO + part * 16, o);
          }
          tmem_st_wait();
        };
        if (kv_len == kKTile) {
          uint32_t s0[32], s1[32], s2[32], s3[32];
          tmem_ld32(tS, s0); tmem_ld32(tS + 32, s1); tmem_ld32(tS + 64, s2); tmem_ld32(tS + 96, s3);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_free[t]);              // QK^T of block j+1 may overwrite S now
          if (q == 0 && lane == 0) ATTN_TRACE(0, t, j, 1);            // scores in registers
          float a0 = -INFINITY, a1 = -INFINITY, a2 = -INFINITY, a3 = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            a0 = max3(a0, __uint_as_float(s0[i]), __uint_as_float(s0[i + 1]));
            a1 = max3(a1, __uint_as_float(s1[i]), __uint_as_float(s1[i + 1]));
            a2 = max3(a2, __uint_as_float(s2[i]), __uint_as_float(s2[i + 1]));
            a3 = max3(a3, __uint_as_float(s3[i]), __uint_as_float(s3[i + 1]));
          }
          const float m_new = fmaxf(m_run, fmaxf(fmaxf(a0, a1), fmaxf(a2, a3)) * scale);
          if (j == 0) m_run = m_new;
          else if (__any_sync(0xffffffffu, m_new - m_run > kRescaleThreshold)) rescale(m_new);
          const uint64_t sc2 = pack2f(scale, scale), nm2 = pack2f(-m_run, -m_run);
          uint64_t sum_a = 0, sum_b = 0;                       // bit pattern of (+0.0f, +0.0f)
          // all 128 probabilities first (the scores die chunk by chunk), THEN the wait for PV of block j-1 (which
          // has had the whole softmax to complete) and the four TMEM stores back to back
          uint32_t pk0[16], pk1[16], pk2[16], pk3[16];
          softmax_chunk(s0, sc2, nm2, sum_a, sum_b, pk0);
          softmax_chunk(s1, sc2, nm2, sum_a, sum_b, pk1);
          softmax_chunk(s2, sc2, nm2, sum_a, sum_b, pk2);
          softmax_chunk(s3, sc2, nm2, sum_a, sum_b, pk3);
          if (q == 0 && lane == 0) ATTN_TRACE(0, t, j, 2);            // exps done
          if (j > 0 && !waited) {                              // PV of block j-1 has consumed the previous P
            mbar_wait(&o_done[t], odone_ph); odone_ph ^= 1;
            tc_fence_after();
          }
          if (q == 0 && lane == 0) ATTN_TRACE(0, t, j, 3);            // PV(j-1) seen complete
          tmem_st16(tP, pk0);
          tmem_st16(tP + 16, pk1);
          tmem_st16(tP + 32, pk2);
          tmem_st16(tP + 48, pk3);
          float x0, x1, y0, y1;
          unpack2f(sum_a, x0, x1);
          unpack2f(sum_b, y0, y1);
          l_run += (x0 + x1) + (y0 + y1);
        } else {
          // ragged last block: only the chunks that hold keys, masked; two passes over TMEM (cheap: <= 1/8 of the row)
          const int nch = (kv_len + 31) >> 5;
          uint32_t v[32];
          float mx = -INFINITY;
          for (int c = 0; c < nch; ++c) {
            tmem_ld32(tS + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i < kv_len) mx = fmaxf(mx, __uint_as_float(v[i]));
          }
          const float m_new = fmaxf(m_run, mx * scale);
          if (j == 0) m_run = m_new;
          else if (__any_sync(0xffffffffu, m_new - m_run > kRescaleThreshold)) rescale(m_new);
          float ls = 0.0f;
          for (int c = 0; c < nch; ++c) {
            tmem_ld32(tS + c * 32, v);
            tmem_ld_wait();
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float e0 = (c * 32 + 2 * i < kv_len) ? ex2_approx(fmaf(__uint_as_float(v[2 * i]), scale, -m_run)) : 0.0f;
              const float e1 = (c * 32 + 2 * i + 1 < kv_len) ? ex2_approx(fmaf(__uint_as_float(v[2 * i + 1]), scale, -m_run)) : 0.0f;
              ls += e0 + e1;
              pk[i] = pack_bf16(e0, e1);
            }
            if (c == 0 && j > 0 && !waited) {
              mbar_wait(&o_done[t], odone_ph); odone_ph ^= 1;
              tc_fence_after();
            }
            tmem_st16(tP + c * 16, pk);
          }
          l_run += ls;
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_free[t]);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
        if (q == 0 && lane == 0) ATTN_TRACE(0, t, j, 4);              // P published
      }
      // ---- item epilogue: O / l -> bf16 -> global
      mbar_wait(&o_done[t], odone_ph); odone_ph ^= 1;
      if (warp_active) {
        tc_fence_after();
        uint32_t o0[32], o1[32];
        tmem_ld32(tO, o0);
        tmem_ld32(tO + 32, o1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_free[t]);                // PV of the next item may overwrite O
        if (tok < seq) {
          const float inv = 1.0f / l_run;
          __nv_bfloat16* op = P.out + (static_cast<long long>(b) * seq + tok) * P.out_ld + h * kHd;
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            *reinterpret_cast<uint4*>(op + i) = make_uint4(
                pack_bf16(__uint_as_float(o0[i]) * inv, __uint_as_float(o0[i + 1]) * inv),
                pack_bf16(__uint_as_float(o0[i + 2]) * inv, __uint_as_float(o0[i + 3]) * inv),
                pack_bf16(__uint_as_float(o0[i + 4]) * inv, __uint_as_float(o0[i + 5]) * inv),
                pack_bf16(__uint_as_float(o0[i + 6]) * inv, __uint_as_float(o0[i + 7]) * inv));
            *reinterpret_cast<uint4*>(op + 32 + i) = make_uint4(
                pack_bf16(__uint_as_float(o1[i]) * inv, __uint_as_float(o1[i + 1]) * inv),
                pack_bf16(__uint_as_float(o1[i + 2]) * inv, __uint_as_float(o1[i + 3]) * inv),
                pack_bf16(__uint_as_float(o1[i + 4]) * inv, __uint_as_float(o1[i + 5]) * inv),
                pack_bf16(__uint_as_float(o1[i + 6]) * inv, __uint_as_float(o1[i + 7]) * inv));
          }
        }
      } else {
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_free[t]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace pf

using namespace pf;

#ifdef PF_ATTN_TRACE
extern "C" int pf_attention_trace_read(unsigned long long* out, unsigned int* n) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(n, g_attn_trace_n, sizeof(unsigned int));
  cudaMemcpyFromSymbol(out, g_attn_trace, sizeof(unsigned long long) * 4096 * 2);
  unsigned int zero = 0;
  cudaMemcpyToSymbol(g_attn_trace_n, &zero, sizeof(zero));
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
