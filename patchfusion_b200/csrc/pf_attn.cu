// Fused attention for the DINOv2 blocks: softmax(Q K^T * scale) V without materialising the (B,heads,N,N) score
// tensor the reference builds (`dinov2/layers/attention.py:53-59`).  head_dim = 64, no mask, N = 1037 tokens.
//
// One CTA per (128-query block, head, image).  Both contractions run on tcgen05 with fp32 accumulators in TMEM:
//   S[128x128]   = Q[128x64] . K_j[128x64]^T          (TMA loads straight from the qkv GEMM output, K-major)
//   Otmp[128x64] = P[128x128] . V_j[128x64]           (P: bf16 written by the softmax warps into a SWIZZLE_128B
//                                                       tile; V^T tiles come from the transposed copy the qkv GEMM
//                                                       epilogue writes, so both operands stay K-major)
// Warps 0-3 own one query row per thread (TMEM lane == row): online softmax in fp32, P -> smem, running rescale of
// the register-resident output.  Warp 4 = TMA producer, warp 5 = MMA issuer.  Two CTAs fit per SM (80 KB smem,
// 256 TMEM columns each) so one CTA's softmax overlaps the other's MMAs.
#include <stdlib.h>

#include "pf_common.cuh"
#include "pf_kernels.h"

namespace pf {

constexpr int kAttnThreads = 192;
constexpr int kQTile = 128, kKTile = 128, kHd = 64;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnParams {
  CUtensorMap tmQK;   // 3-D {2*D, seq, B}, box {64, 128, 1}
  CUtensorMap tmVt;   // 2-D {seq_pad, B*heads*64}, box {64, 64}
  int B, seq, heads, D;
  float scale_log2;   // scale * log2(e)
  __nv_bfloat16* out;
  int out_ld;
};

__global__ void __launch_bounds__(kAttnThreads, 2) pf_attention_kernel(const __grid_constant__ AttnParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                    // 16 KB  [128 q][64]
  uint8_t* sK = smem + 16384;            // 16 KB  [128 keys][64]
  uint8_t* sV = smem + 32768;            // 16 KB  2 x [64 d][64 keys]
  uint8_t* sP = smem + 49152;            // 32 KB  2 x [128 q][64 keys]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 81920);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = bars + 2;
  uint64_t* v_full = bars + 3;
  uint64_t* v_empty = bars + 4;
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* o_full = bars + 7;
  uint64_t* o_empty = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qb * kQTile;
  const int nkv = (P.seq + kKTile - 1) / kKTile;

  if (warp == 4 && lane == 0) {
    prefetch_tmap(&P.tmQK);
    prefetch_tmap(&P.tmVt);
    mbar_init(q_full, 1); mbar_init(k_full, 1); mbar_init(k_empty, 1); mbar_init(v_full, 1); mbar_init(v_empty, 1);
    mbar_init(s_full, 1); mbar_init(p_full, 128); mbar_init(o_full, 1); mbar_init(o_empty, 128);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 128;
  pdl_wait();

  if (warp == 4) {
    if (lane == 0) {
      mbar_expect_tx(q_full, 16384);
      tma_load_3d(sQ, &P.tmQK, q_full, h * kHd, q0, b);
      for (int j = 0; j < nkv; ++j) {
        const uint32_t ph = j & 1;
        mbar_wait(k_empty, ph ^ 1);
        mbar_expect_tx(k_full, 16384);
        tma_load_3d(sK, &P.tmQK, k_full, P.D + h * kHd, j * kKTile, b);
        mbar_wait(v_empty, ph ^ 1);
        mbar_expect_tx(v_full, 16384);
        tma_load_2d(sV, &P.tmVt, v_full, j * kKTile, (b * P.heads + h) * kHd);
        tma_load_2d(sV + 8192, &P.tmVt, v_full, j * kKTile + 64, (b * P.heads + h) * kHd);
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16(128, 128);
      const uint32_t idesc_o = umma_idesc_bf16(128, 64);
      const uint64_t dq = umma_desc_k128(smem_u32(sQ));
      const uint64_t dk = umma_desc_k128(smem_u32(sK));
      const uint64_t dv0 = umma_desc_k128(smem_u32(sV)), dv1 = umma_desc_k128(smem_u32(sV + 8192));
      const uint64_t dp0 = umma_desc_k128(smem_u32(sP)), dp1 = umma_desc_k128(smem_u32(sP + 16384));
      mbar_wait(q_full, 0);
      mbar_wait(k_full, 0);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem_S, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
      umma_commit(k_empty);
      umma_commit(s_full);
      for (int j = 0; j < nkv; ++j) {
        const uint32_t ph = j & 1;
        mbar_wait(p_full, ph);
        mbar_wait(v_full, ph);
        mbar_wait(o_empty, ph ^ 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t pa = (k < 4 ? dp0 : dp1) + 2 * (k & 3);
          const uint64_t vb = (k < 4 ? dv0 : dv1) + 2 * (k & 3);
          umma_bf16(tmem_O, pa, vb, idesc_o, k != 0);
        }
        umma_commit(v_empty);
        umma_commit(o_full);
        if (j + 1 < nkv) {
          mbar_wait(k_full, ph ^ 1);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_S, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
          umma_commit(k_empty);
          umma_commit(s_full);
        }
      }
    }
  } else {
    // ===================== softmax / output warps: thread == query row =====================
    const int r = warp * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
    float o_acc[kHd];
#pragma unroll
    for (int i = 0; i < kHd; ++i) o_acc[i] = 0.0f;
    float m_run = -INFINITY, l_run = 0.0f;
    for (int j = 0; j < nkv; ++j) {
      const uint32_t ph = j & 1;
      const int kvalid = min(kKTile, P.seq - j * kKTile);   // keys of this block that exist
      mbar_wait(s_full, ph);
      tc_fence_after();
      // pass 1: row max of the raw scores (scale > 0, applied once afterwards).  TMEM loads are software
      // pipelined: chunk i+1 is in flight while chunk i is consumed.
      // four independent max / sum chains: the serial 128-long FMNMX / FADD dependency chains were the critical
      // path of a softmax warp (2 warps per scheduler cannot hide 4-cycle-latency chains)
      float mr[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      const bool full = kvalid == kKTile;
      uint32_t va[32], vb[32];
      tmem_ld32(tmem_S + lane_sel, va);
      tmem_ld_wait();
#pragma unroll
      for (int c2 = 0; c2 < kKTile / 32; ++c2) {
        uint32_t (&cur)[32] = (c2 & 1) ? vb : va;
        uint32_t (&nxt)[32] = (c2 & 1) ? va : vb;
        // after the last score chunk, prefetch chunk 0 again for pass 2
        tmem_ld32(tmem_S + lane_sel + ((c2 + 1) & 3) * 32, nxt);
        const int cb = c2 * 32;
        if (full) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mr[i & 3] = fmaxf(mr[i & 3], __uint_as_float(cur[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (cb + i < kvalid) mr[i & 3] = fmaxf(mr[i & 3], __uint_as_float(cur[i]));
        }
        tmem_ld_wait();
      }
      const float m_raw = fmaxf(fmaxf(mr[0], mr[1]), fmaxf(mr[2], mr[3]));
      const float m_new = fmaxf(m_run, m_raw * P.scale_log2);
      const float alpha = ex2_approx(m_run - m_new);
      float ls[4] = {0.f, 0.f, 0.f, 0.f};
      // pass 2: p = 2^(s*scale - m) (one FFMA + one MUFU per score), bf16 pairs -> swizzled smem tile.
      // (va holds chunk 0 again at this point)
#pragma unroll
      for (int c2 = 0; c2 < kKTile / 32; ++c2) {
        uint32_t (&cur)[32] = (c2 & 1) ? vb : va;
        uint32_t (&nxt)[32] = (c2 & 1) ? va : vb;
        if (c2 + 1 < kKTile / 32) tmem_ld32(tmem_S + lane_sel + (c2 + 1) * 32, nxt);
        const int cb = c2 * 32;
        float p[32];
        if (full) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            p[i] = ex2_approx(fmaf(__uint_as_float(cur[i]), P.scale_log2, -m_new));
            ls[i & 3] += p[i];
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float e = ex2_approx(fmaf(__uint_as_float(cur[i]), P.scale_log2, -m_new));
            p[i] = (cb + i < kvalid) ? e : 0.0f;
            ls[i & 3] += p[i];
          }
        }
        uint8_t* sub = sP + (cb >> 6) * 16384 + r * 128;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int chunk = ((cb & 63) >> 3) + g;                 // 16-byte chunk index inside the 128-B row
          uint4 pk = make_uint4(pack_bf16(p[g * 8 + 0], p[g * 8 + 1]), pack_bf16(p[g * 8 + 2], p[g * 8 + 3]),
                                pack_bf16(p[g * 8 + 4], p[g * 8 + 5]), pack_bf16(p[g * 8 + 6], p[g * 8 + 7]));
          *reinterpret_cast<uint4*>(sub + ((chunk ^ (r & 7)) << 4)) = pk;
        }
        if (c2 + 1 < kKTile / 32) tmem_ld_wait();
      }
      l_run = l_run * alpha + ((ls[0] + ls[1]) + (ls[2] + ls[3]));
      m_run = m_new;
      fence_proxy_async_smem();      // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      mbar_arrive(p_full);
      // accumulate this block's P.V
      mbar_wait(o_full, ph);
      tc_fence_after();
      tmem_ld32(tmem_O + lane_sel, va);
      tmem_ld32(tmem_O + lane_sel + 32, vb);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        o_acc[i] = fmaf(o_acc[i], alpha, __uint_as_float(va[i]));
        o_acc[32 + i] = fmaf(o_acc[32 + i], alpha, __uint_as_float(vb[i]));
      }
      tc_fence_before();
      mbar_arrive(o_empty);
    }
    const int tok = q0 + r;
    if (tok < P.seq) {
      const float inv = 1.0f / l_run;
      __nv_bfloat16* op = P.out + (static_cast<long long>(b) * P.seq + tok) * P.out_ld + h * kHd;
#pragma unroll
      for (int i = 0; i < kHd; i += 8) {
        uint4 pk = make_uint4(pack_bf16(o_acc[i] * inv, o_acc[i + 1] * inv), pack_bf16(o_acc[i + 2] * inv, o_acc[i + 3] * inv),
                              pack_bf16(o_acc[i + 4] * inv, o_acc[i + 5] * inv), pack_bf16(o_acc[i + 6] * inv, o_acc[i + 7] * inv));
        *reinterpret_cast<uint4*>(op + i) = pk;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ------------------------------------------------------------------------------------------------------------
// v2: same pipeline, but EIGHT softmax warps per CTA (16 per SM): the two warps of a TMEM lane quadrant split every
// query row - warp half hh owns score columns [64 hh, 64 hh + 64) and output columns [32 hh, 32 hh + 32).  The row
// maximum is exchanged through shared memory once per KV block (one 64-thread named barrier), the row sums are
// combined once at the end.  Twice the warps hide the tcgen05.ld / MUFU / FMA latencies that bounded v1 (ncu: 30 %
// long-scoreboard + 22 % wait stalls at 2 softmax warps per scheduler), and each thread keeps only 32 + 32 registers
// of row state.
constexpr int kAttn2Threads = 320;

__global__ void __launch_bounds__(kAttn2Threads, 2) pf_attention_kernel_v2(const __grid_constant__ AttnParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + 16384;
  uint8_t* sV = smem + 32768;
  uint8_t* sP = smem + 49152;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 81920);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = bars + 2;
  uint64_t* v_full = bars + 3;
  uint64_t* v_empty = bars + 4;
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* o_full = bars + 7;
  uint64_t* o_empty = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  float* xch = reinterpret_cast<float*>(smem + 81920 + 128);      // [2 parities][2 halves][128 rows]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qb * kQTile;
  const int nkv = (P.seq + kKTile - 1) / kKTile;

  if (warp == 8 && lane == 0) {
    prefetch_tmap(&P.tmQK);
    prefetch_tmap(&P.tmVt);
    mbar_init(q_full, 1); mbar_init(k_full, 1); mbar_init(k_empty, 1); mbar_init(v_full, 1); mbar_init(v_empty, 1);
    mbar_init(s_full, 1); mbar_init(p_full, 256); mbar_init(o_full, 1); mbar_init(o_empty, 256);
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 128;
  pdl_wait();

  if (warp == 8) {
    if (lane == 0) {
      mbar_expect_tx(q_full, 16384);
      tma_load_3d(sQ, &P.tmQK, q_full, h * kHd, q0, b);
      for (int j = 0; j < nkv; ++j) {
        const uint32_t ph = j & 1;
        mbar_wait(k_empty, ph ^ 1);
        mbar_expect_tx(k_full, 16384);
        tma_load_3d(sK, &P.tmQK, k_full, P.D + h * kHd, j * kKTile, b);
        mbar_wait(v_empty, ph ^ 1);
        mbar_expect_tx(v_full, 16384);
        tma_load_2d(sV, &P.tmVt, v_full, j * kKTile, (b * P.heads + h) * kHd);
        tma_load_2d(sV + 8192, &P.tmVt, v_full, j * kKTile + 64, (b * P.heads + h) * kHd);
      }
    }
  } else if (warp == 9) {
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16(128, 128);
      const uint32_t idesc_o = umma_idesc_bf16(128, 64);
      const uint64_t dq = umma_desc_k128(smem_u32(sQ));
      const uint64_t dk = umma_desc_k128(smem_u32(sK));
      const uint64_t dv0 = umma_desc_k128(smem_u32(sV)), dv1 = umma_desc_k128(smem_u32(sV + 8192));
      const uint64_t dp0 = umma_desc_k128(smem_u32(sP)), dp1 = umma_desc_k128(smem_u32(sP + 16384));
      mbar_wait(q_full, 0);
      mbar_wait(k_full, 0);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem_S, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
      umma_commit(k_empty);
      umma_commit(s_full);
      for (int j = 0; j < nkv; ++j) {
        const uint32_t ph = j & 1;
        mbar_wait(p_full, ph);
        mbar_wait(v_full, ph);
        mbar_wait(o_empty, ph ^ 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t pa = (k < 4 ? dp0 : dp1) + 2 * (k & 3);
          const uint64_t vb = (k < 4 ? dv0 : dv1) + 2 * (k & 3);
          umma_bf16(tmem_O, pa, vb, idesc_o, k != 0);
        }
        umma_commit(v_empty);
        umma_commit(o_full);
        if (j + 1 < nkv) {
          mbar_wait(k_full, ph ^ 1);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_S, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
          umma_commit(k_empty);
          umma_commit(s_full);
        }
      }
    }
  } else {
    // ===================== softmax / output warps: two threads (one per half) per query row =====================
    const int q = warp & 3, hh = warp >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t s_col = hh * 64;                 // first score column owned by this thread
    float o_acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) o_acc[i] = 0.0f;
    float m_run = -INFINITY, l_run = 0.0f;
    uint8_t* prow = sP + hh * 16384 + r * 128;      // P sub-tile hh holds key columns [64 hh, 64 hh + 64)
    for (int j = 0; j < nkv; ++j) {
      const uint32_t ph = j & 1;
      const int kvalid = min(kKTile, P.seq - j * kKTile) - static_cast<int>(s_col);   // valid keys among my 64 columns
      const bool full = kvalid >= 64;
      mbar_wait(s_full, ph);
      tc_fence_after();
      uint32_t v[32];
      float mr[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      // pass 1: partial row max over my 64 columns (chunk A, then chunk B which stays in registers)
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        tmem_ld32(tmem_S + lane_sel + s_col + c2 * 32, v);
        tmem_ld_wait();
        if (full) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mr[i & 3] = fmaxf(mr[i & 3], __uint_as_float(v[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c2 * 32 + i < kvalid) mr[i & 3] = fmaxf(mr[i & 3], __uint_as_float(v[i]));
        }
      }
      const float m_part = fmaxf(fmaxf(mr[0], mr[1]), fmaxf(mr[2], mr[3]));
      float* x = xch + ph * 256;
      x[hh * 128 + r] = m_part;
      asm volatile("bar.sync %0, 64;" ::"r"(q + 1) : "memory");
      const float m_raw = fmaxf(m_part, x[(hh ^ 1) * 128 + r]);
      const float m_new = fmaxf(m_run, m_raw * P.scale_log2);
      const float alpha = ex2_approx(m_run - m_new);
      float ls[4] = {0.f, 0.f, 0.f, 0.f};
      // pass 2: chunk B is still in v; then chunk A is re-read
#pragma unroll
      for (int c2 = 1; c2 >= 0; --c2) {
        if (c2 == 0) {
          tmem_ld32(tmem_S + lane_sel + s_col, v);
          tmem_ld_wait();
        }
        float p[32];
        if (full) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            p[i] = ex2_approx(fmaf(__uint_as_float(v[i]), P.scale_log2, -m_new));
            ls[i & 3] += p[i];
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float e = ex2_approx(fmaf(__uint_as_float(v[i]), P.scale_log2, -m_new));
            p[i] = (c2 * 32 + i < kvalid) ? e : 0.0f;
            ls[i & 3] += p[i];
          }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int chunk = c2 * 4 + g;                            // 16-byte chunk index inside the 128-B row
          uint4 pk = make_uint4(pack_bf16(p[g * 8 + 0], p[g * 8 + 1]), pack_bf16(p[g * 8 + 2], p[g * 8 + 3]),
                                pack_bf16(p[g * 8 + 4], p[g * 8 + 5]), pack_bf16(p[g * 8 + 6], p[g * 8 + 7]));
          *reinterpret_cast<uint4*>(prow + ((chunk ^ (r & 7)) << 4)) = pk;
        }
      }
      l_run = l_run * alpha + ((ls[0] + ls[1]) + (ls[2] + ls[3]));
      m_run = m_new;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
      mbar_wait(o_full, ph);
      tc_fence_after();
      tmem_ld32(tmem_O + lane_sel + hh * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o_acc[i] = fmaf(o_acc[i], alpha, __uint_as_float(v[i]));
      tc_fence_before();
      mbar_arrive(o_empty);
    }
    // combine the two halves' row sums (both used the same running maximum)
    float* x = xch + 512;
    x[hh * 128 + r] = l_run;
    asm volatile("bar.sync %0, 64;" ::"r"(q + 1) : "memory");
    const float l_tot = l_run + x[(hh ^ 1) * 128 + r];
    const int tok = q0 + r;
    if (tok < P.seq) {
      const float inv = 1.0f / l_tot;
      __nv_bfloat16* op = P.out + (static_cast<long long>(b) * P.seq + tok) * P.out_ld + h * kHd + hh * 32;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 pk = make_uint4(pack_bf16(o_acc[i] * inv, o_acc[i + 1] * inv), pack_bf16(o_acc[i + 2] * inv, o_acc[i + 3] * inv),
                              pack_bf16(o_acc[i + 4] * inv, o_acc[i + 5] * inv), pack_bf16(o_acc[i + 6] * inv, o_acc[i + 7] * inv));
        *reinterpret_cast<uint4*>(op + i) = pk;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace pf

using namespace pf;

extern "C" int pf_attention(const void* qk, int32_t qk_ld, const void* vt, int32_t B, int32_t seq, int32_t seq_pad,
                            int32_t heads, float scale, void* out, int32_t out_ld, void* stream) {
  static bool attr_done_dev[kMaxDevices] = {false};
  bool& attr_done = attr_done_dev[current_device()];
  static const bool use_v1 = getenv("PF_B200_ATTN_V1") != nullptr;
  const int smem_bytes = 1024 + 81920 + 128 + 3072;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(pf_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(pf_attention_kernel_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return set_error("cudaFuncSetAttribute(pf_attention_kernel): %s", cudaGetErrorString(e));
    attr_done = true;
  }
  const int D = heads * kHd;
  if (qk_ld % 8 || seq_pad % 8 || out_ld % 8) return set_error("pf_attention: strides must be multiples of 8");
  AttnParams P;
  if (tmap_3d_bf16(&P.tmQK, qk, 2 * D, seq, B, qk_ld, static_cast<uint64_t>(seq) * qk_ld, 64, 128, 1)) return 1;
  if (tmap_2d_bf16(&P.tmVt, vt, seq_pad, static_cast<uint64_t>(B) * heads * kHd, seq_pad, 64, 64)) return 1;
  P.B = B; P.seq = seq; P.heads = heads; P.D = D;
  P.scale_log2 = scale * 1.4426950408889634f;
  P.out = static_cast<__nv_bfloat16*>(out);
  P.out_ld = out_ld;
  dim3 grid((seq + kQTile - 1) / kQTile, heads, B);
  cudaError_t le = use_v1 ? launch_pdl(pf_attention_kernel, grid, dim3(kAttnThreads), smem_bytes, static_cast<cudaStream_t>(stream), P)
                          : launch_pdl(pf_attention_kernel_v2, grid, dim3(kAttn2Threads), smem_bytes, static_cast<cudaStream_t>(stream), P);
  if (le != cudaSuccess) return set_error("pf_attention_kernel launch: %s", cudaGetErrorString(le));
  return check_launch("pf_attention_kernel");
}
