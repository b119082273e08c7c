// Implicit-GEMM engine of the PatchFusion hot path (tcgen05 / TMEM / TMA, sm_100a).
//
// One persistent, warp-specialised kernel serves every dense contraction of the path:
//   * ViT / Swin linear layers            (`dinov2/layers/attention.py:51,60`, `mlp.py:36-39`, `swin_layers.py:140,162`)
//   * 1x1 convs (NHWC == plain GEMM)      (`depth_anything/dpt.py:30-38`, metric-head MLPs `localbins_layers.py:84-117`)
//   * 3x3 stride-1 pad-1 convs over a channel-concat of up to three NHWC sources
//                                          (`depth_anything/blocks.py:53-58`, `patchfusion.py:122-127,263-267`,
//                                           `guided_fusion_model.py:34-69,98-99,203`)
//   * ConvTranspose k==stride as GEMM + pixel-shuffle store (`depth_anything/dpt.py:40-53`)
//
// D[M,N] = sum_{src,tap,c} A_src[pixel(m)+tap, c] * Wp[n, k(src,tap,c)]   (bf16 x bf16 -> fp32 in TMEM)
//
// A tiles are fetched by TMA straight from the activation tensors: a 3x3 tap is just a (dx,dy) offset of the 4-D
// box {64ch, bw, bh, 1} and the hardware zero-fills out-of-image pixels (conv padding), out-of-range channels and
// rows, so there is no im2col buffer and no halo code.  Weights are pre-packed K-major [N, Ktot] with every
// (source, tap) segment padded to a multiple of 64 channels, matching the producer's enumeration order.
//
// Warp roles (384 threads): warps 0-7 = epilogue (TMEM -> registers -> bias / activation -> a swizzled staging tile in
// shared memory -> bulk tensor store, or a bulk reduce-add for the fp32 residual stream; the fused-tail / residual /
// pixel-shuffle / V^T variants store directly, one full 32-B sector per access), warps 8-9 = operand producers of the
// halo kernel's optional fused bilinear resample (idle otherwise), warp 10 = TMA producer, warp 11 = TMEM owner +
// tcgen05.mma issuer.  Producer and issuer walk their loops as whole warps and one elect.sync lane issues.
// Two TMEM accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.
#include <stdlib.h>

#include "pf_common.cuh"
#include "pf_kernels.h"

namespace pf {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kATileBytes = kBlockM * kBlockK * 2;  // 16 KiB
constexpr int kEpiWarps = 8;                       // two warps per TMEM lane quadrant split the column chunks
// Warp roles: 0..7 epilogue (TMEM lane quadrant = warp & 3), kTmaWarp = TMA producer, kMmaWarp = TMEM owner + tcgen05.mma
// issuer.  The issuer sits in the HIGHEST warp id on purpose: the scheduler arbitrates highest-warp-id-first
// (B300_MICROARCH.md), and the small-N convolutions are bound by the issue rate of this warp - as warp 1 it was queued
// behind the two epilogue warps of its scheduler.
constexpr int kTmaWarp = kEpiWarps + 2, kMmaWarp = kEpiWarps + 3;
constexpr int kGemmThreads = (kEpiWarps + 4) * 32;

struct GemmKernelParams {
  CUtensorMap tmA[3];
  CUtensorMap tmB;
  CUtensorMap tmBh;     // multicast variant: box of block_n / 2 weight rows (each CTA of the pair fetches one half)
  CUtensorMap tmOut;    // d.tma_out: output tensor (bf16 {64 cols, 32 rows} boxes, or fp32 {32 cols, 32 rows})
  GemmDesc d;
  int stages;
  int total_tiles;
  int k_steps;  // per tile
  int kc;       // halo kernel: taps per weight stage (1, 3 or 9)
};

struct TileCoord {
  int n0;               // first output column
  int m0;               // linear: first row
  int img, y0, x0;      // conv: tile origin
};

__device__ __forceinline__ TileCoord decode_tile(const GemmDesc& d, int t) {
  TileCoord c;
  int nt = t % d.n_tiles;
  int mt = t / d.n_tiles;
  c.n0 = nt * d.block_n;
  c.m0 = mt * kBlockM;
  c.img = 0; c.y0 = 0; c.x0 = 0;
  if (d.a_mode == 1) {
    int per_img = d.tiles_y * d.tiles_x;
    c.img = mt / per_img;
    int r = mt - c.img * per_img;
    c.y0 = (r / d.tiles_x) * d.bh;
    c.x0 = (r % d.tiles_x) * d.bw;
  }
  return c;
}


// Columns the MMA of tile t must produce: block_n, or for the last n-tile the remaining columns rounded up to 16.
__device__ __forceinline__ int tile_n_eff(const GemmDesc& d, int t) {
  const int n0 = (t % d.n_tiles) * d.block_n;
  const int rem = ((d.N - n0 + 15) >> 4) << 4;
  return rem < d.block_n ? rem : d.block_n;
}
// K = 16 steps that carry data in the last 64-channel chunk of source s (the rest of the chunk is TMA zero fill).
__device__ __forceinline__ int last_chunk_k16(const GemmDesc& d, int s) {
  return (d.k_true[s] - (d.chunks[s] - 1) * kBlockK + 15) >> 4;
}

// Debug timeline (build with -DPF_GEMM_TRACE, tools/gemm_trace.py): CTA 0 records (role, tile, event, clock).  Every
// recording thread owns a 256-entry region and a private counter: a trace point is one clock read and two stores.
#ifdef PF_GEMM_TRACE
__device__ unsigned long long g_gemm_trace[4 * 256 * 2];
__device__ __forceinline__ void gemm_trace(int role, int t, int ev, unsigned int& cnt) {
  if (blockIdx.x != 0 || cnt >= 256) return;
  const unsigned int i = role * 256 + cnt++;
  g_gemm_trace[2 * i] = (1ull << 63) | (static_cast<unsigned long long>(role) << 48) |
                        (static_cast<unsigned long long>(t) << 16) | static_cast<unsigned long long>(ev);
  g_gemm_trace[2 * i + 1] = clock64();
}
#define GEMM_TRACE(role, t, ev) gemm_trace(role, t, ev, trace_cnt)
#define GEMM_TRACE_DECL unsigned int trace_cnt = 0;
#else
#define GEMM_TRACE(role, t, ev)
#define GEMM_TRACE_DECL
#endif

// One 32-column chunk of one accumulator row: bias -> activation -> residuals -> store.  FULL == all 32 columns exist
// (vector loads/stores, no predication); the tail variant predicates every column but keeps all indices static so
// f[] stays in registers.
constexpr int kMaxTail = 16;   // widest fused trailing 1x1 layer
constexpr int kTailBytes = 512 + 128 * kMaxTail * 4;   // barriers + TMEM slot + [128][kMaxTail] fp32 scratch
// pf_gemm_kernel: one 4 KB staging tile (32 rows x 128 B, SWIZZLE_128B) per epilogue warp for the TMA-store epilogue; the
// fused-tail scratch (never used together with it) aliases the same bytes
constexpr int kStageTile = 32 * 128;
constexpr int kEpiStageBytes = kEpiWarps * kStageTile;  // 32 KB
constexpr int kEpiSmemBytes = 512 + kEpiStageBytes;

// W = chunk width in accumulator columns (32, or 16 when the two warps of a quadrant split an unpaired chunk);
// TAILN = compile-time bound on the fused trailing layer's outputs (0 = no trailing layer): keeps the executed code
// path short - the fully unrolled 16-output variant alone is ~3k instructions and thrashed the instruction cache.
template <bool FULL, int W, int TAILN>
__device__ __forceinline__ void epilogue_chunk(const GemmDesc& d, float (&f)[W], long long orow, int ncol, int lcol,
                                               int nvalid, float (&y2)[kMaxTail], uint32_t (&xv)[TAILN == 0 ? W / 8 : 1][8]) {
  if (d.bias != nullptr) {
    if (FULL) {
      const float4* bp = reinterpret_cast<const float4*>(d.bias + lcol);
#pragma unroll
      for (int j = 0; j < W / 4; ++j) {
        float4 b = __ldg(bp + j);
        f[4 * j] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j) if (j < nvalid) f[j] += __ldg(d.bias + lcol + j);
    }
  }
  if (d.act == PF_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < W; ++j) f[j] = fmaxf(f[j], 0.0f);
  } else if (d.act == PF_ACT_GELU) {
    if (FULL) {
#pragma unroll
      for (int j = 0; j < W; j += 2) gelu_erf2(f[j], f[j + 1]);
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j) if (j < nvalid) f[j] = gelu_erf(f[j]);
    }
  } else if (d.act == PF_ACT_SOFTPLUS) {
#pragma unroll
    for (int j = 0; j < W; ++j) if (FULL || j < nvalid) f[j] = softplus(f[j]);
  }
  // fused trailing 1x1 layer: y[i] += sum_j W2[i][lcol + j] * f[j]  (row-local: the thread owns the whole row)
  if (TAILN > 0) {
#pragma unroll
    for (int i = 0; i < TAILN; ++i) {
      if (i < d.n2) {
        const float* wp = d.w2 + static_cast<long long>(i) * d.n_logical + lcol;
        // four independent partial sums: a single accumulator is a 32-deep FFMA dependency chain, and with two warps
        // per scheduler the fused-tail layers ran at IPC 0.36 (ncu r02b: latency-bound, no pipe above 16 %)
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        if (FULL) {
#pragma unroll
          for (int j = 0; j < W / 4; ++j) {
            float4 w4 = __ldg(reinterpret_cast<const float4*>(wp) + j);
            a0 = fmaf(w4.x, f[4 * j], a0); a1 = fmaf(w4.y, f[4 * j + 1], a1);
            a2 = fmaf(w4.z, f[4 * j + 2], a2); a3 = fmaf(w4.w, f[4 * j + 3], a3);
          }
        } else {
#pragma unroll
          for (int j = 0; j < W; ++j) if (j < nvalid) a0 = fmaf(__ldg(wp + j), f[j], a0);
        }
        y2[i] += (a0 + a1) + (a2 + a3);
      }
    }
    if (d.skip_main) return;
  }
#pragma unroll
  for (int rsel = 0; rsel < 2; ++rsel) {
    const __nv_bfloat16* rbase = rsel == 0 ? d.res1 : d.res2;
    if (rbase == nullptr) continue;
    const __nv_bfloat16* rp = rbase + orow * d.res_ld + lcol;
    if (FULL && d.wide) {
#pragma unroll
      for (int j = 0; j < W / 16; ++j) {
        uint32_t u[8];
        ld_global_256(rp + 16 * j, u);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u[e]));
          f[16 * j + 2 * e] += t.x; f[16 * j + 2 * e + 1] += t.y;
        }
      }
    } else if (FULL) {
#pragma unroll
      for (int j = 0; j < W / 8; ++j) {
        uint4 u = __ldg(reinterpret_cast<const uint4*>(rp) + j);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 t = __bfloat1622float2(h[e]);
          f[8 * j + 2 * e] += t.x; f[8 * j + 2 * e + 1] += t.y;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j) if (j < nvalid) f[j] += __bfloat162float(rp[j]);
    }
  }
  const int oc = lcol + d.out_col0;      // physical output column
#ifdef PF_GEMM_EXP_NOSTORE
  { float acc = 0.f;
#pragma unroll
    for (int j = 0; j < W; ++j) acc += f[j];
    if (acc == 1.2345e-30f) reinterpret_cast<float*>(d.out)[0] = acc;
    return; }
#endif
  if (d.vt != nullptr && ncol >= d.vt_col0) {
    // attention V written transposed: vt[(b*heads + h)*64 + dd][token]
    const int m = static_cast<int>(orow);
    const int b = m / d.vt_seq, tok = m - b * d.vt_seq;
    __nv_bfloat16* vp = d.vt + (static_cast<long long>(b) * d.vt_dim + (ncol - d.vt_col0)) * d.vt_seq_pad + tok;
#pragma unroll
    for (int j = 0; j < W; ++j)
      if (FULL || j < nvalid) vp[static_cast<long long>(j) * d.vt_seq_pad] = __float2bfloat16(f[j]);
  } else if (d.gamma != nullptr) {
    // x <- x + gamma * (acc + bias): fp32 residual stream updated in place
    float* xp = reinterpret_cast<float*>(d.out) + orow * d.out_ld + oc;
    if (FULL && TAILN == 0) {
      // one full 32-B sector per access; the residual-stream segment was fetched by epilogue_cols before the
      // accumulator wait (xv), so only the update + store remain here
#pragma unroll
      for (int j = 0; j < (TAILN == 0 ? W / 8 : 1); ++j) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(d.gamma + lcol + 8 * j));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(d.gamma + lcol + 8 * j + 4));
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) xv[j][e] = __float_as_uint(fmaf(gg[e], f[8 * j + e], __uint_as_float(xv[j][e])));
        st_global_256(xp + 8 * j, xv[j]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j) if (j < nvalid) xp[j] += __ldg(d.gamma + lcol + j) * f[j];
    }
  } else if (d.out_f32) {
    float* op = reinterpret_cast<float*>(d.out) + orow * d.out_ld + oc;
    if (FULL) {
#pragma unroll
      for (int j = 0; j < W; j += 8) {
        const uint32_t u[8] = {__float_as_uint(f[j]), __float_as_uint(f[j + 1]), __float_as_uint(f[j + 2]), __float_as_uint(f[j + 3]),
                               __float_as_uint(f[j + 4]), __float_as_uint(f[j + 5]), __float_as_uint(f[j + 6]), __float_as_uint(f[j + 7])};
        st_global_256(op + j, u);
      }
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j) if (j < nvalid) op[j] = f[j];
    }
  } else {
    __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(d.out) + orow * d.out_ld + oc;
    if (FULL && d.wide) {
#pragma unroll
      for (int j = 0; j < W; j += 16) {
        uint32_t u[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) u[e] = pack_bf16(f[j + 2 * e], f[j + 2 * e + 1]);
        st_global_256(op + j, u);
      }
    } else if (FULL) {
#pragma unroll
      for (int j = 0; j < W; j += 8)
        *reinterpret_cast<uint4*>(op + j) = make_uint4(pack_bf16(f[j], f[j + 1]), pack_bf16(f[j + 2], f[j + 3]),
                                                        pack_bf16(f[j + 4], f[j + 5]), pack_bf16(f[j + 6], f[j + 7]));
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j) if (j < nvalid) op[j] = __float2bfloat16(f[j]);
    }
    if (d.out2 != nullptr) {
      __nv_bfloat16* o2 = d.out2 + orow * d.out2_ld + lcol;
      if (FULL && d.wide) {
#pragma unroll
        for (int j = 0; j < W; j += 16) {
          uint32_t u[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) u[e] = pack_bf16(fmaxf(f[j + 2 * e], 0.f), fmaxf(f[j + 2 * e + 1], 0.f));
          st_global_256(o2 + j, u);
        }
      } else if (FULL) {
#pragma unroll
        for (int j = 0; j < W; j += 8)
          *reinterpret_cast<uint4*>(o2 + j) =
              make_uint4(pack_bf16(fmaxf(f[j], 0.f), fmaxf(f[j + 1], 0.f)), pack_bf16(fmaxf(f[j + 2], 0.f), fmaxf(f[j + 3], 0.f)),
                         pack_bf16(fmaxf(f[j + 4], 0.f), fmaxf(f[j + 5], 0.f)), pack_bf16(fmaxf(f[j + 6], 0.f), fmaxf(f[j + 7], 0.f)));
      } else {
#pragma unroll
        for (int j = 0; j < W; ++j) if (j < nvalid) o2[j] = __float2bfloat16(fmaxf(f[j], 0.0f));
      }
    }
  }
}

// one W-column chunk starting at accumulator column cb of this thread's row
template <int W, int TAILN>
__device__ __forceinline__ void epilogue_cols(const GemmDesc& d, uint32_t taddr, int cb, const TileCoord& c, int ocol0,
                                              long long orow, bool row_ok, float (&y2)[kMaxTail]) {
  const int ncol = c.n0 + cb;            // global N index of v[0] (selects the V^T path)
  const int lcol = ocol0 + cb;           // logical output channel of v[0] (bias / gamma / residual index)
  const int nvalid = min(W, d.n_logical - lcol);   // columns of this chunk that exist
  // x += gamma * v: fetch the fp32 residual-stream segment (one full sector per access) BEFORE waiting for the
  // accumulator, so its L2 latency overlaps the TMEM load instead of following it
  uint32_t xpre[TAILN == 0 ? W / 8 : 1][8];
  const bool pre = TAILN == 0 && d.gamma != nullptr && row_ok && nvalid == W && !(d.vt != nullptr && ncol >= d.vt_col0);
  if (pre) {
    const float* xp = reinterpret_cast<const float*>(d.out) + orow * d.out_ld + lcol + d.out_col0;
#pragma unroll
    for (int j = 0; j < (TAILN == 0 ? W / 8 : 1); ++j) ld_global_256(xp + 8 * j, xpre[j]);
  }
  uint32_t v[W];
  if (W == 32) tmem_ld32(taddr + cb, reinterpret_cast<uint32_t (&)[32]>(v));
  else tmem_ld16(taddr + cb, reinterpret_cast<uint32_t (&)[16]>(v));
  tmem_ld_wait();
#ifdef PF_GEMM_EXP_NOEPI
  return;
#endif
  if (!row_ok) return;
  if (nvalid <= 0) return;
  float f[W];
#pragma unroll
  for (int j = 0; j < W; ++j) f[j] = __uint_as_float(v[j]);
  if (nvalid == W) epilogue_chunk<true, W, TAILN>(d, f, orow, ncol, lcol, W, y2, xpre);
  else epilogue_chunk<false, W, TAILN>(d, f, orow, ncol, lcol, nvalid, y2, xpre);
}

// All column chunks of one accumulator row pair of warps: the two warps of a TMEM lane quadrant take alternate
// 32-column chunks; an unpaired last chunk (block_n = 32, 96, 160, 224) is split 16 / 16 so both warps carry the
// same load (the activation + fused-tail work of the clb / N = 32 layers is epilogue-bound).
template <int TAILN>
__device__ __forceinline__ void epilogue_row(const GemmDesc& d, uint32_t taddr, int half, const TileCoord& c, int ocol0,
                                             long long orow, bool row_ok, float (&y2)[kMaxTail]) {
  const int nchunks = d.block_n >> 5;
  const int paired = nchunks & ~1;
  for (int ch = half; ch < paired; ch += 2) epilogue_cols<32, TAILN>(d, taddr, ch * 32, c, ocol0, orow, row_ok, y2);
  if (nchunks & 1) epilogue_cols<16, TAILN>(d, taddr, paired * 32 + half * 16, c, ocol0, orow, row_ok, y2);
}

// ------------------------------------------------------------------------------------------------------------
// Epilogue through shared memory + TMA (pf_gemm_kernel, d.tma_out != 0).
// A thread owns one accumulator row, so direct global accesses are 32-byte pieces of 32 different rows per instruction:
// one L2 request per sector.  The in-kernel timeline (tools/gemm_trace.py) showed what that costs: a 128 x 256 tile
// of the ViT linears is 4096 128-byte operand requests, its epilogue another 2048 (bf16) to 8192 (fp32 residual stream,
// read + write) sector requests, and the TMA load latency of the NEXT tile's mainloop rose from 1.6k to 2.6-4.7k clk
// while an epilogue was draining (mainloop 8.4k clk alone, 13.8k-21k clk with a concurrent epilogue).  Here each warp
// stages its 32 rows in a swizzled 4 KB tile and ONE elected lane moves it with a bulk tensor copy: 128-byte requests,
// no LSU work; the fp32 residual stream is updated by a bulk reduce-add (no read at all).
struct EpiTma {
  const CUtensorMap* tm;
  uint32_t stg;        // shared address of this warp's staging tile (1024-B aligned)
  uint8_t* stg_ptr;
};

// store this warp's 32 rows x 64 bf16 columns starting at column col
__device__ __forceinline__ void epi_tma_store(const GemmDesc& d, const EpiTma& e, const TileCoord& c, int q, int col) {
  if (d.a_mode == 1) {
    const int r0 = q * 32;
    const int yy = r0 / d.bw, xx = r0 - yy * d.bw;
    tma_store_4d(e.tm, e.stg_ptr, col, c.x0 + xx, c.y0 + yy, c.img);
  } else {
    tma_store_2d(e.tm, e.stg_ptr, col, c.m0 + q * 32);
  }
}

__device__ __forceinline__ void epi_act(const GemmDesc& d, float (&f)[32]) {
  if (d.act == PF_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
  } else if (d.act == PF_ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 32; j += 2) gelu_erf2(f[j], f[j + 1]);
  } else if (d.act == PF_ACT_SOFTPLUS) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = softplus(f[j]);
  }
}
// acc + bias for 32 columns starting at logical column lcol (columns >= n_logical read no bias; TMA clips them)
__device__ __forceinline__ void epi_bias(const GemmDesc& d, const uint32_t (&v)[32], int lcol, float (&f)[32]) {
#pragma unroll
  for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
  if (d.bias == nullptr) return;
  if (lcol + 32 <= d.n_logical) {
    const float4* bp = reinterpret_cast<const float4*>(d.bias + lcol);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = __ldg(bp + j);
      f[4 * j] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) if (lcol + j < d.n_logical) f[j] += __ldg(d.bias + lcol + j);
  }
}

// bf16 output: groups of 64 columns (128-byte row segments); the two warps of a quadrant take alternate groups
__device__ __forceinline__ void epilogue_tile_tma_bf16(const GemmDesc& d, EpiTma& e, uint32_t taddr, int half,
                                                       const TileCoord& c, int q, int lane) {
  const int ng = d.block_n >> 6;
  for (int g = half; g < ng; g += 2) {
    const int lcol = c.n0 + g * 64;
    if (lcol >= d.n_logical) break;
    uint32_t v0[32], v1[32];
    tmem_ld32(taddr + g * 64, v0);
    tmem_ld32(taddr + g * 64 + 32, v1);
    tmem_ld_wait();
    float f0[32], f1[32];
    epi_bias(d, v0, lcol, f0);
    epi_bias(d, v1, lcol + 32, f1);
    epi_act(d, f0);
    epi_act(d, f1);
    if (lane == 0) bulk_wait_read0();                      // the previous copy has finished reading the staging tile
    __syncwarp();
    const uint32_t row = e.stg + lane * 128;
    const int sw = lane & 7;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      st_shared_v4(row + ((j ^ sw) << 4), pack_bf16(f0[8 * j], f0[8 * j + 1]), pack_bf16(f0[8 * j + 2], f0[8 * j + 3]),
                   pack_bf16(f0[8 * j + 4], f0[8 * j + 5]), pack_bf16(f0[8 * j + 6], f0[8 * j + 7]));
      st_shared_v4(row + (((j + 4) ^ sw) << 4), pack_bf16(f1[8 * j], f1[8 * j + 1]), pack_bf16(f1[8 * j + 2], f1[8 * j + 3]),
                   pack_bf16(f1[8 * j + 4], f1[8 * j + 5]), pack_bf16(f1[8 * j + 6], f1[8 * j + 7]));
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) { epi_tma_store(d, e, c, q, d.out_col0 + lcol); bulk_commit(); }
  }
}

// fp32 output (a_mode 0): chunks of 32 columns (128-byte row segments).  The residual-stream update x += gamma * (acc +
// bias) stages gamma * (acc + bias) and lets the copy engine ADD it into x (cp.reduce.async.bulk.tensor .add.f32,
// performed in the L2): the SM never reads x.  Each element is updated by exactly one tile: deterministic.
__device__ __forceinline__ void epilogue_tile_tma_f32(const GemmDesc& d, EpiTma& e, uint32_t taddr, int half,
                                                      const TileCoord& c, int q, int lane) {
  const int nchunks = d.block_n >> 5;
  for (int ch = half; ch < nchunks; ch += 2) {
    const int lcol = c.n0 + ch * 32;
    if (lcol >= d.n_logical) break;
    uint32_t v[32];
    tmem_ld32(taddr + ch * 32, v);
    tmem_ld_wait();
    float f[32];
    epi_bias(d, v, lcol, f);
    epi_act(d, f);
    if (d.gamma != nullptr) {
      if (lcol + 32 <= d.n_logical) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(d.gamma + lcol) + j);
          f[4 * j] *= g.x; f[4 * j + 1] *= g.y; f[4 * j + 2] *= g.z; f[4 * j + 3] *= g.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = lcol + j < d.n_logical ? f[j] * __ldg(d.gamma + lcol + j) : 0.f;
      }
    }
    if (lane == 0) bulk_wait_read0();                      // the previous copy has finished reading the staging tile
    __syncwarp();
    const uint32_t row = e.stg + lane * 128;
    const int sw = lane & 7;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      st_shared_v4(row + ((j ^ sw) << 4), __float_as_uint(f[4 * j]), __float_as_uint(f[4 * j + 1]),
                   __float_as_uint(f[4 * j + 2]), __float_as_uint(f[4 * j + 3]));
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      if (d.gamma != nullptr) tma_reduce_add_2d(e.tm, e.stg_ptr, d.out_col0 + lcol, c.m0 + q * 32);
      else tma_store_2d(e.tm, e.stg_ptr, d.out_col0 + lcol, c.m0 + q * 32);
      bulk_commit();
    }
  }
}

// Epilogue warps (4 warps, TMEM lane quadrant = warp & 3): drain accumulator stage `acc` of each tile this CTA owns.
// Tile schedule.  Plain: CTA b takes tiles b, b + grid, ...  Multicast pairs (MC): cluster c takes tile PAIRS c, c + clusters, ...
// where pair p = (m-tile pair p / n_tiles, n-tile p % n_tiles) and the CTA of cluster rank r owns m-tile 2 * mp + r
// (an odd last m-tile pairs with an all-out-of-range one: TMA zero-fills it, the epilogue drops its rows).
struct TileIter {
  int first, step, count, rank, n_tiles, cl;
  __device__ __forceinline__ int tile(int i) const {
    if (cl == 1) return i;
    const int nt = i % n_tiles, mp = i / n_tiles;
    return (cl * mp + rank) * n_tiles + nt;
  }
};
// cl = CTAs per cluster sharing the weight tiles (1 = plain schedule): cluster g takes work items g, g + clusters, ...
// where item p = (m-tile group p / n_tiles, n-tile p % n_tiles) and the CTA of cluster rank r owns m-tile cl * group + r.
__device__ __forceinline__ TileIter make_iter(const GemmDesc& d, int total_tiles, int cl) {
  TileIter it;
  it.cl = cl; it.n_tiles = d.n_tiles;
  if (cl == 1) { it.first = blockIdx.x; it.step = gridDim.x; it.count = total_tiles; it.rank = 0; }
  else { it.first = blockIdx.x / cl; it.step = gridDim.x / cl; it.count = ((d.m_tiles + cl - 1) / cl) * d.n_tiles; it.rank = cluster_ctarank(); }
  return it;
}

__device__ __forceinline__ void epilogue_loop(const GemmDesc& d, const TileIter& it, uint64_t* tmem_full,
                                              uint64_t* tmem_empty, uint32_t tmem_base, int warp, int lane,
                                              float* tail_smem, EpiTma* et = nullptr) {
  const int q = warp & 3;                 // TMEM lane quadrant this warp may access
  const int half = warp >> 2;             // which of the quadrant's two warps: takes every other 32-column chunk
  const int r = q * 32 + lane;            // accumulator row owned by this thread
  int acc = 0; uint32_t acc_phase = 0;
  GEMM_TRACE_DECL
  const bool tr = lane == 0 && (warp == 0 || warp == 4);
  for (int ti = it.first; ti < it.count; ti += it.step) {
    TileCoord c = decode_tile(d, it.tile(ti));
    // ---- row mapping
    bool row_ok;
    long long orow;       // output row (pixel / token) index
    if (d.a_mode == 1) {
      int yy = r / d.bw, xx = r - yy * d.bw;
      int y = c.y0 + yy, x = c.x0 + xx;
      row_ok = (y < d.H) && (x < d.W) && (c.img < d.NB);     // img >= NB: the phantom tile that pairs an odd last m-tile
      orow = (static_cast<long long>(c.img) * d.H + y) * d.W + x;
    } else {
      int m = c.m0 + r;
      row_ok = m < d.M;
      orow = m;
    }
    int ocol0 = c.n0;     // output column of accumulator column 0
    if (d.ps > 1 && d.a_mode == 0) {
      // ConvTranspose k==s: columns are ordered (ky, kx, cout); this N tile belongs to one (ky,kx).
      int tap = c.n0 / d.ps_cout_pad;
      ocol0 = c.n0 - tap * d.ps_cout_pad;
      int ky = tap / d.ps, kx = tap - ky * d.ps;
      int m = c.m0 + r;
      int img = m / (d.H * d.W);
      int rem = m - img * d.H * d.W;
      int y = rem / d.W, x = rem - y * d.W;
      orow = (static_cast<long long>(img) * d.H * d.ps + y * d.ps + ky) * (d.W * d.ps) + x * d.ps + kx;
    }
    if (tr) GEMM_TRACE(2 + half, ti, 0);
    // TMA epilogue for this tile?  (the V^T tiles of the fused qkv projection keep the transposing direct store)
    const bool tma_tile = et != nullptr && d.tma_out != 0 && !(d.vt != nullptr && c.n0 >= d.vt_col0);
    mbar_wait(&tmem_full[acc], acc_phase);
    tc_fence_after();
    if (tr) GEMM_TRACE(2 + half, ti, 1);
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * d.block_n;
    if (tma_tile) {
      if (d.tma_out == 1) epilogue_tile_tma_bf16(d, *et, taddr, half, c, q, lane);
      else epilogue_tile_tma_f32(d, *et, taddr, half, c, q, lane);
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);
      if (tr) GEMM_TRACE(2 + half, ti, 2);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      continue;
    }
    float y2[kMaxTail];
#pragma unroll
    for (int i = 0; i < kMaxTail; ++i) y2[i] = 0.f;
    if (d.w2 == nullptr) epilogue_row<0>(d, taddr, half, c, ocol0, orow, row_ok, y2);
    else if (d.n2 <= 1) epilogue_row<1>(d, taddr, half, c, ocol0, orow, row_ok, y2);
    else if (d.n2 <= 4) epilogue_row<4>(d, taddr, half, c, ocol0, orow, row_ok, y2);
    else epilogue_row<kMaxTail>(d, taddr, half, c, ocol0, orow, row_ok, y2);
    if (d.w2 != nullptr) {
      // combine the two half-row partial sums of the fused trailing layer through shared memory
      float* ts = tail_smem + r * kMaxTail;
      if (half == 1) {
#pragma unroll
        for (int i = 0; i < kMaxTail; ++i) ts[i] = y2[i];
      }
      asm volatile("bar.sync %0, 64;" ::"r"(q + 1) : "memory");
      if (half == 0) {
#pragma unroll
        for (int i = 0; i < kMaxTail; ++i) y2[i] += ts[i];
      }
      asm volatile("bar.sync %0, 64;" ::"r"(q + 1) : "memory");
    }
    if (d.w2 != nullptr && row_ok && half == 0) {
      // trailing layer output: fp32 [rows, out3_ld], bias + activation
      float* op = d.out3 + orow * d.out3_ld;
#pragma unroll
      for (int i = 0; i < kMaxTail; ++i) {
        if (i < d.n2) {
          float v2 = y2[i] + (d.b2 != nullptr ? __ldg(d.b2 + i) : 0.f);
          if (d.act2 == PF_ACT_RELU) v2 = fmaxf(v2, 0.f);
          else if (d.act2 == PF_ACT_SOFTPLUS) v2 = softplus(v2);
          else if (d.act2 == PF_ACT_GELU) v2 = gelu_erf(v2);
          op[i] = v2;
        }
      }
    }
    tc_fence_before();
    mbar_arrive(&tmem_empty[acc]);
    if (tr) GEMM_TRACE(2 + half, ti, 2);
    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
  }
  if (et != nullptr && lane == 0) bulk_wait0();    // the staging tiles are read (and the writes performed) before the CTA retires
}

// MC = true: launched as clusters of 2 CTAs that take two m-tiles of the SAME n-tile; each CTA fetches half of the
// weight tile and multicasts it into both CTAs' shared memory (the linear layers are bound by L2 -> SM bandwidth:
// A 16 KB + B 32 KB per 512 MMA clocks is ~94 B/clk/SM against a chip-wide ~43 B/clk/SM; halving B cuts it by a third).
template <bool MC>
__global__ void __launch_bounds__(kGemmThreads, 1) pf_gemm_kernel(const __grid_constant__ GemmKernelParams P) {
  extern __shared__ uint8_t smem_raw[];
  const GemmDesc& d = P.d;
  // 1024-B alignment is required by SWIZZLE_128B operand tiles.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stages = P.stages;
  const int b_tile_bytes = d.block_n * kBlockK * 2;
  const int stage_bytes = kATileBytes + b_tile_bytes;
  uint8_t* epi_stage = smem + stages * stage_bytes;            // [kEpiWarps][4 KB], 1024-B aligned (stage sizes are 4 KB multiples)
  float* tail_smem = reinterpret_cast<float*>(epi_stage);      // [128][kMaxTail]: fused-tail launches never stage
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_stage + kEpiStageBytes);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tmem_full = empty_bar + stages;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();          // persistent kernel, all CTAs resident: let the next kernel's prologue start

  if (warp == kTmaWarp && lane == 0) {
    for (int s = 0; s < d.num_src; ++s) prefetch_tmap(&P.tmA[s]);
    prefetch_tmap(&P.tmB);
    // a multicast stage is refilled only after BOTH CTAs' MMAs released it: two arrivals per phase
    for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], MC ? 2 : 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], kEpiWarps * 32); }
    if (d.tma_out) prefetch_tmap(&P.tmOut);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  if (MC) cluster_sync_all();       // the peer's barriers exist before anything is multicast into its shared memory
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const TileIter it = make_iter(d, P.total_tiles, MC ? 2 : 1);
  pdl_wait();                       // predecessor's results are visible from here on

  if (warp == kTmaWarp) {
    // ===================== TMA producer =====================
    // whole warp in uniform control flow (coordinates and stage counters in uniform registers), one elected lane issues
    {
      int stage = 0; uint32_t phase = 0;
      GEMM_TRACE_DECL
      if (lane == 0) GEMM_TRACE(0, 0, 9);
      for (int ti = it.first; ti < it.count; ti += it.step) {
        TileCoord c = decode_tile(d, it.tile(ti));
        if (lane == 0) GEMM_TRACE(0, ti, 0);
        int kb = 0;  // running 64-wide K block index into the packed weights
        for (int s = 0; s < d.num_src; ++s) {
          for (int tap = 0; tap < d.taps; ++tap) {
            int dy = d.taps == 9 ? tap / 3 - 1 : 0;
            int dx = d.taps == 9 ? tap % 3 - 1 : 0;
            for (int ch = 0; ch < d.chunks[s]; ++ch, ++kb) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              if (lane == 0 && ti == it.first + it.step) GEMM_TRACE(0, ti, 100 + kb);
              uint8_t* sa = smem + stage * stage_bytes;
              uint8_t* sb = sa + kATileBytes;
              if (elect_one()) {
                mbar_expect_tx(&full_bar[stage], stage_bytes);
                if (d.a_mode == 0) tma_load_2d(sa, &P.tmA[s], &full_bar[stage], ch * kBlockK, c.m0);
                else tma_load_4d(sa, &P.tmA[s], &full_bar[stage], ch * kBlockK, c.x0 + dx, c.y0 + dy, c.img);
                if (MC) {
                  const int half_rows = d.block_n >> 1;
                  tma_load_2d_mc(sb + it.rank * half_rows * 128, &P.tmBh, &full_bar[stage], kb * kBlockK,
                                 c.n0 + it.rank * half_rows, static_cast<uint16_t>(3));
                } else {
                  tma_load_2d(sb, &P.tmB, &full_bar[stage], kb * kBlockK, c.n0);
                }
              }
              if (++stage == stages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    // whole warp in uniform control flow, one elected lane issues (see pf_conv3_halo_kernel)
    {
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      GEMM_TRACE_DECL
      for (int ti = it.first; ti < it.count; ti += it.step) {
        if (lane == 0) GEMM_TRACE(1, ti, 0);
        // the last n-tile issues only the columns that exist (rounded to the MMA granule of 16): N = 544 runs as
        // 192 + 192 + 160 instead of 3 x 192
        const uint32_t idesc = umma_idesc_bf16(kBlockM, tile_n_eff(d, it.tile(ti)));
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        if (lane == 0) GEMM_TRACE(1, ti, 1);
        const uint32_t tmem_d = tmem_base + acc * d.block_n;
        uint32_t accum = 0;
        for (int s = 0; s < d.num_src; ++s) {
          const int nch = d.chunks[s];
          const int nk_last = last_chunk_k16(d, s);        // K16 steps of the zero-padded last 64-channel chunk
          for (int tap = 0; tap < d.taps; ++tap) {
            for (int ch = 0; ch < nch; ++ch) {
              const int nk = ch == nch - 1 ? nk_last : kBlockK / 16;
              mbar_wait(&full_bar[stage], phase);
              tc_fence_after();
              if (lane == 0 && !accum) GEMM_TRACE(1, ti, 2);
              if (lane == 0 && ti == it.first + it.step) GEMM_TRACE(1, ti, 100 + stage);
              const uint32_t sa = smem_u32(smem + stage * stage_bytes);
              const uint64_t adesc = umma_desc_k128(sa);
              const uint64_t bdesc = umma_desc_k128(sa + kATileBytes);
              if (elect_one()) {
                if (nk == kBlockK / 16) {
#pragma unroll
                  for (int k = 0; k < kBlockK / 16; ++k)
                    // advance 16 bf16 = 32 B along K inside the 128-B swizzle row: +2 in the (addr>>4) field
                    umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, accum | k);
                } else {
#pragma unroll
                  for (int k = 0; k < kBlockK / 16; ++k)
                    if (k < nk) umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, accum | k);
                }
                if (MC) umma_commit_mc(&empty_bar[stage], static_cast<uint16_t>(3));   // release the stage in both CTAs
                else umma_commit(&empty_bar[stage]);       // frees this smem stage once the MMAs above retire
              }
              accum = 1;
              if (++stage == stages) { stage = 0; phase ^= 1; }
            }
          }
        }
        if (elect_one()) umma_commit(&tmem_full[acc]);     // accumulator complete -> epilogue
        if (lane == 0) GEMM_TRACE(1, ti, 3);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp < kEpiWarps) {
    // ===================== epilogue (warps 0..7) =====================
    EpiTma et;
    et.tm = &P.tmOut; et.stg_ptr = epi_stage + warp * kStageTile; et.stg = smem_u32(et.stg_ptr);
    epilogue_loop(d, it, tmem_full, tmem_empty, tmem_base, warp, lane, tail_smem, &et);
  }

  tc_fence_before();
  __syncthreads();
  if (MC) cluster_sync_all();       // no CTA exits while its peer may still multicast into it / arrive on its barriers
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------
// 3x3 convolution with a shared-memory HALO tile (stride 1, zero pad 1).
//
// The generic kernel above fetches the A operand once per tap: 9 TMA boxes of 16 KB per 64-channel chunk, all hitting
// the same pixels.  Here one TMA box {64 ch, 10, 18, 1} brings the (16+2) x (8+2) pixel neighbourhood of a 16x8
// output tile (23 KB) and all nine taps are issued from it: for tap (dy,dx) the 128 accumulator rows are the 16
// image rows of 8 pixels starting at halo row (dy*10 + dx), i.e. a K-major SWIZZLE_128B operand whose 8-row groups
// are 1280 B apart - expressed purely through the UMMA descriptor's start address and stride-byte-offset (the
// 128B-swizzle XOR is a function of the shared-memory address bits, so 128-B-granular shifted views of a TMA-written
// tile stay consistent).  A traffic from L2 drops 6.3x; the weights stream per (chunk, tap) as before.
// Warp roles and the TMEM double-buffered epilogue are those of pf_gemm_kernel.
constexpr int kHaloW = 10, kHaloH = 18;
constexpr int kHaloBytes = kHaloW * kHaloH * 128;          // 23040
constexpr int kHaloSlot = 24 * 1024;                        // 1024-B aligned slot
constexpr int kHaloSlots = 3;

// KC consecutive taps (tap0 .. tap0+KC-1) of one 64-channel chunk: 4 MMAs (K = 16 each) per tap.  `tap0` is a
// literal at every call site, so all A offsets fold to immediates.
// full 64-channel chunks take the lean fully unrolled path; only a source's zero-padded last chunk takes the counted one
#define PF_ISSUE(KC, ...) do { if (nk == kBlockK / 16) issue_taps<KC, true>(__VA_ARGS__); else issue_taps<KC, false>(__VA_ARGS__); } while (0)
template <int CL>
__device__ __forceinline__ void halo_release_b(uint64_t* bar) {
  if (CL > 1) umma_commit_mc(bar, static_cast<uint16_t>((1u << CL) - 1));   // the weight stage is free in every CTA of the cluster
  else umma_commit(bar);
}
template <int KC, bool FULLK>
__device__ __forceinline__ void issue_taps(uint32_t tmem_d, uint64_t a_hi, uint32_t a_lo, uint64_t b_hi, uint32_t b_lo,
                                           uint32_t b_tile16, int tap0, uint32_t idesc, uint32_t first, int nk) {
#pragma unroll
  for (int j = 0; j < KC; ++j) {
    const int tap = tap0 + j;
    const int dy = tap / 3, dx = tap - dy * 3;
    const uint32_t al = a_lo + ((dy * kHaloW + dx) * 128 >> 4);
    const uint32_t bl = b_lo + j * b_tile16;
#pragma unroll
    for (int k = 0; k < kBlockK / 16; ++k) {
      if (FULLK || k < nk) {
        umma_bf16(tmem_d, a_hi | (al + 2 * k), b_hi | (bl + 2 * k), idesc, first ? 0u : 1u);
        first = 0;
      }
    }
  }
}

// ---- fused bilinear resample (align_corners=True) of a halo-kernel source -------------------------------------------
// F.interpolate(x, (H, W), mode='bilinear', align_corners=True) feeding a 3x3 conv (guided_fusion_model.py:98-99,
// 191-203) used to be materialised by resize_bilinear_tiled_kernel (5 % of the step, written once and read back by the
// conv).  Instead a producer warp builds the 18 x 10 pixel halo of one 64-channel chunk directly in the swizzled
// operand tile TMA would have written: lane -> (halo pixel, 8-channel piece), four 16-byte taps from the low-resolution
// map (L1/L2 resident: the taps of neighbouring pixels overlap), the same FFMA2 blend and bf16 rounding as the
// stand-alone kernel, zeros for the conv padding ring / pad channels.  Same source coordinate as ATen: scale * dst.
constexpr int kRsWarp0 = kEpiWarps;      // warps 8, 9: resample producers (alternate chunks)
__device__ __forceinline__ void rs_coord(int dst, int in, float scale, int& lo, int& hi, float& frac) {
  const float src = scale * dst;
  lo = static_cast<int>(src);
  if (lo > in - 1) lo = in - 1;
  hi = lo + (lo < in - 1 ? 1 : 0);
  frac = src - lo;
}
__device__ __forceinline__ void halo_fill_bilinear(uint32_t slot, const GemmDesc& d, int s, int ch, const TileCoord& c, int lane) {
  const int ih = d.rs_h[s], iw = d.rs_w[s], ld = d.rs_ld[s];
  const __nv_bfloat16* src = d.rs_ptr[s] + static_cast<size_t>(c.img) * ih * iw * ld + ch * kBlockK;
  const int cvalid = d.k_true[s] - ch * kBlockK;            // channels of this chunk that exist (multiple of 8)
  const float sy = d.rs_sy[s], sx = d.rs_sx[s];
  const int j = lane & 7;                                    // 16-byte piece (8 channels) of the pixel's 128-byte row
  const bool jok = j * 8 < cvalid && c.img < d.NB;
#pragma unroll 5
  for (int p = lane >> 3; p < kHaloW * kHaloH; p += 4) {
    const int hy = p / kHaloW, hx = p - hy * kHaloW;
    const int Y = c.y0 - 1 + hy, X = c.x0 - 1 + hx;
    uint32_t o0 = 0, o1 = 0, o2 = 0, o3 = 0;
    if (jok && Y >= 0 && Y < d.H && X >= 0 && X < d.W) {
      int y0, y1, x0, x1; float fy, fx;
      rs_coord(Y, ih, sy, y0, y1, fy);
      rs_coord(X, iw, sx, x0, x1, fx);
      const __nv_bfloat16* r0 = src + static_cast<size_t>(y0) * iw * ld + j * 8;
      const __nv_bfloat16* r1 = src + static_cast<size_t>(y1) * iw * ld + j * 8;
      const uint4 ua = __ldg(reinterpret_cast<const uint4*>(r0 + static_cast<size_t>(x0) * ld));
      const uint4 ub = __ldg(reinterpret_cast<const uint4*>(r0 + static_cast<size_t>(x1) * ld));
      const uint4 uc = __ldg(reinterpret_cast<const uint4*>(r1 + static_cast<size_t>(x0) * ld));
      const uint4 ud = __ldg(reinterpret_cast<const uint4*>(r1 + static_cast<size_t>(x1) * ld));
      const float w00 = (1.f - fy) * (1.f - fx), w01 = (1.f - fy) * fx, w10 = fy * (1.f - fx), w11 = fy * fx;
      const uint64_t p00 = pack2f(w00, w00), p01 = pack2f(w01, w01), p10 = pack2f(w10, w10), p11 = pack2f(w11, w11);
      const uint32_t wa[4] = {ua.x, ua.y, ua.z, ua.w}, wb[4] = {ub.x, ub.y, ub.z, ub.w};
      const uint32_t wc[4] = {uc.x, uc.y, uc.z, uc.w}, wd[4] = {ud.x, ud.y, ud.z, ud.w};
      uint32_t ow[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint64_t r = mul2(p00, pack2(wa[k] << 16, wa[k] & 0xffff0000u));
        r = fma2(p01, pack2(wb[k] << 16, wb[k] & 0xffff0000u), r);
        r = fma2(p10, pack2(wc[k] << 16, wc[k] & 0xffff0000u), r);
        r = fma2(p11, pack2(wd[k] << 16, wd[k] & 0xffff0000u), r);
        float lo, hi;
        unpack2f(r, lo, hi);
        ow[k] = pack_bf16(lo, hi);
      }
      o0 = ow[0]; o1 = ow[1]; o2 = ow[2]; o3 = ow[3];
    }
    st_shared_v4(slot + p * 128 + ((j ^ (p & 7)) << 4), o0, o1, o2, o3);
  }
}

// CL > 1: clusters of CL (2 or 4) CTAs take CL m-tiles (pixel tiles) of the SAME n-tile; each CTA fetches 1/CL of the rows
// of every weight tile and multicasts it into all of them.  The weights are ~90 % of this kernel's L2 -> SM traffic (one
// 23 KB halo against nine 16-32 KB tap tiles per 64-channel chunk) and that traffic sits at the fabric's ceiling
// (~46 B/clk/SM measured on the linear layers), so halving it is what the N = 32 layers and the partial chunks need.
template <int CL>
__global__ void __launch_bounds__(kGemmThreads, 1) pf_conv3_halo_kernel(const __grid_constant__ GemmKernelParams P) {
  constexpr bool MC = CL > 1;
  constexpr uint16_t kMask = static_cast<uint16_t>((1u << CL) - 1);
  extern __shared__ uint8_t smem_raw[];
  const GemmDesc& d = P.d;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stages = P.stages;                              // B ring
  const int kc = P.kc;                                      // taps per B stage
  const int b_tile_bytes = d.block_n * kBlockK * 2;
  const int b_stage_bytes = kc * b_tile_bytes;
  uint8_t* smem_b = smem + kHaloSlots * kHaloSlot;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + stages * b_stage_bytes);
  uint64_t* a_full = bars;                                  // [kHaloSlots]
  uint64_t* a_empty = a_full + kHaloSlots;
  uint64_t* b_full = a_empty + kHaloSlots;                  // [stages]
  uint64_t* b_empty = b_full + stages;
  uint64_t* tmem_full = b_empty + stages;                   // [2]
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* tail_smem = reinterpret_cast<float*>(tmem_slot + 4);   // [128][kMaxTail]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();
  if (warp == kTmaWarp && lane == 0) {
    for (int s = 0; s < d.num_src; ++s) if (d.rs_h[s] == 0) prefetch_tmap(&P.tmA[s]);
    prefetch_tmap(&P.tmB);
    for (int s = 0; s < kHaloSlots; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    // a multicast weight stage is refilled only after BOTH CTAs' MMAs released it: two arrivals per phase
    for (int s = 0; s < stages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], CL); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], kEpiWarps * 32); }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  if (MC) cluster_sync_all();       // the peers' barriers exist before anything is multicast into their shared memory
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const TileIter it = make_iter(d, P.total_tiles, CL);
  pdl_wait();                       // predecessor's results are visible from here on

  if (warp == kTmaWarp) {
    // whole warp in uniform control flow, one elected lane issues the copies
    int as = 0; uint32_t aph = 0;
    int bs = 0; uint32_t bph = 0;
    for (int ti = it.first; ti < it.count; ti += it.step) {
      TileCoord c = decode_tile(d, it.tile(ti));
      int kbase = 0;                                       // first 64-wide K block of this source in the weights
      for (int s = 0; s < d.num_src; ++s) {
        const int nch = d.chunks[s];
        for (int ch = 0; ch < nch; ++ch) {
          if (d.rs_h[s] == 0) {                            // (resampled sources: the producer warps own the slot)
            mbar_wait(&a_empty[as], aph ^ 1);
            if (elect_one()) {
              mbar_expect_tx(&a_full[as], kHaloBytes);
              tma_load_4d(smem + as * kHaloSlot, &P.tmA[s], &a_full[as], ch * kBlockK, c.x0 - 1, c.y0 - 1, c.img);
            }
          }
          if (++as == kHaloSlots) { as = 0; aph ^= 1; }
          for (int tap0 = 0; tap0 < 9; tap0 += kc) {
            mbar_wait(&b_empty[bs], bph ^ 1);
            if (elect_one()) {
              mbar_expect_tx(&b_full[bs], b_stage_bytes);
              if (MC) {
                const int part_rows = d.block_n / CL;
                for (int j = 0; j < kc; ++j)
                  tma_load_2d_mc(smem_b + bs * b_stage_bytes + j * b_tile_bytes + it.rank * part_rows * 128, &P.tmBh,
                                 &b_full[bs], (kbase + (tap0 + j) * nch + ch) * kBlockK, c.n0 + it.rank * part_rows,
                                 kMask);
              } else {
                for (int j = 0; j < kc; ++j)
                  tma_load_2d(smem_b + bs * b_stage_bytes + j * b_tile_bytes, &P.tmB, &b_full[bs],
                              (kbase + (tap0 + j) * nch + ch) * kBlockK, c.n0);
              }
            }
            if (++bs == stages) { bs = 0; bph ^= 1; }
          }
        }
        kbase += 9 * nch;
      }
    }
  } else if (warp == kMmaWarp) {
    // The WHOLE warp walks the loop (uniform control flow: barrier waits, descriptor arithmetic and stage counters stay
    // in uniform registers) and one elected lane issues.  Wrapping the loop in `if (lane == 0)` instead made every
    // tcgen05.mma pay an ELECT / R2UR.BROADCAST / BRA.U.ANY round trip (~13 instructions): with N = 32 the MMA
    // itself is 16-32 clk, so the issue loop - not the tensor pipe - bounded those convolutions.
    int as = 0; uint32_t aph = 0;
    int bs = 0; uint32_t bph = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int ti = it.first; ti < it.count; ti += it.step) {
      const uint32_t idesc = umma_idesc_bf16(kBlockM, tile_n_eff(d, it.tile(ti)));
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * d.block_n;
      uint32_t first = 1;
      for (int s = 0; s < d.num_src; ++s) {
        const int nk_last = last_chunk_k16(d, s);
        for (int ch = 0; ch < d.chunks[s]; ++ch) {
          const int nk = ch == d.chunks[s] - 1 ? nk_last : kBlockK / 16;
          mbar_wait(&a_full[as], aph);
          tc_fence_after();
          const uint32_t halo = smem_u32(smem + as * kHaloSlot);
          // descriptors are (constant high word, 32-bit low word); per-tap start offsets are compile-time immediates
          // in the fully unrolled KC variants
          const uint64_t a_hi = (static_cast<uint64_t>(1) << 46) | (static_cast<uint64_t>(2) << 61) |
                                (static_cast<uint64_t>((kHaloW * 128) >> 4) << 32) | (static_cast<uint64_t>(1) << 16);
          const uint64_t b_hi = (static_cast<uint64_t>(1) << 46) | (static_cast<uint64_t>(2) << 61) |
                                (static_cast<uint64_t>(1024 >> 4) << 32) | (static_cast<uint64_t>(1) << 16);
          const uint32_t a_lo = (halo & 0x3FFFF) >> 4;
          if (kc == 9) {
            mbar_wait(&b_full[bs], bph);
            tc_fence_after();
            const uint32_t b_lo = (smem_u32(smem_b + bs * b_stage_bytes) & 0x3FFFF) >> 4;
            if (elect_one()) {
              PF_ISSUE(9, tmem_d, a_hi, a_lo, b_hi, b_lo, b_tile_bytes >> 4, 0, idesc, first, nk);
              halo_release_b<CL>(&b_empty[bs]);
              umma_commit(&a_empty[as]);                     // halo slot reusable once its 36 MMAs retire
            }
            first = 0;
            if (++bs == stages) { bs = 0; bph ^= 1; }
          } else if (kc == 3) {
#pragma unroll
            for (int g = 0; g < 3; ++g) {
              mbar_wait(&b_full[bs], bph);
              tc_fence_after();
              const uint32_t b_lo = (smem_u32(smem_b + bs * b_stage_bytes) & 0x3FFFF) >> 4;
              if (elect_one()) {
                if (g == 0) PF_ISSUE(3, tmem_d, a_hi, a_lo, b_hi, b_lo, b_tile_bytes >> 4, 0, idesc, first, nk);
                else if (g == 1) PF_ISSUE(3, tmem_d, a_hi, a_lo, b_hi, b_lo, b_tile_bytes >> 4, 3, idesc, first, nk);
                else PF_ISSUE(3, tmem_d, a_hi, a_lo, b_hi, b_lo, b_tile_bytes >> 4, 6, idesc, first, nk);
                halo_release_b<CL>(&b_empty[bs]);
                if (g == 2) umma_commit(&a_empty[as]);
              }
              first = 0;
              if (++bs == stages) { bs = 0; bph ^= 1; }
            }
          } else {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait(&b_full[bs], bph);
              tc_fence_after();
              const uint32_t b_lo = (smem_u32(smem_b + bs * b_stage_bytes) & 0x3FFFF) >> 4;
              if (elect_one()) {
                PF_ISSUE(1, tmem_d, a_hi, a_lo, b_hi, b_lo, 0, tap, idesc, first, nk);
                halo_release_b<CL>(&b_empty[bs]);
                if (tap == 8) umma_commit(&a_empty[as]);
              }
              first = 0;
              if (++bs == stages) { bs = 0; bph ^= 1; }
            }
          }
          if (++as == kHaloSlots) { as = 0; aph ^= 1; }
        }
      }
      if (elect_one()) umma_commit(&tmem_full[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= kRsWarp0 && warp < kRsWarp0 + 2) {
    // ===================== resample producers: halo tiles of the sources read through a bilinear resample =====================
    if (d.rs_any) {
      int as = 0; uint32_t aph = 0;
      int turn = 0;                                          // the two warps take alternate resampled chunks
      for (int ti = it.first; ti < it.count; ti += it.step) {
        TileCoord c = decode_tile(d, it.tile(ti));
        for (int s = 0; s < d.num_src; ++s) {
          for (int ch = 0; ch < d.chunks[s]; ++ch) {
            if (d.rs_h[s] != 0) {
              if ((turn & 1) == warp - kRsWarp0) {
                mbar_wait(&a_empty[as], aph ^ 1);
                halo_fill_bilinear(smem_u32(smem + as * kHaloSlot), d, s, ch, c, lane);
                fence_proxy_async_smem();                    // generic-proxy writes -> visible to the tensor core's reads
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_full[as]);
              }
              ++turn;
            }
            if (++as == kHaloSlots) { as = 0; aph ^= 1; }
          }
        }
      }
    }
  } else if (warp < kEpiWarps) {
    epilogue_loop(d, it, tmem_full, tmem_empty, tmem_base, warp, lane, tail_smem);
  }
  tc_fence_before();
  __syncthreads();
  if (MC) cluster_sync_all();       // no CTA exits while its peer may still multicast into it / arrive on its barriers
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
static int g_sm_counts[kMaxDevices] = {0};

int gemm_launch(const GemmDesc& d, const CUtensorMap* tmA, const CUtensorMap& tmB, const CUtensorMap* tmBh,
                const CUtensorMap* tmOut, cudaStream_t stream) {
  static bool attr_done[kMaxDevices] = {false};
  const int kMaxSmem = 227 * 1024;
  const int dev = current_device();
  if (!attr_done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(pf_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(pf_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e != cudaSuccess) return set_error("cudaFuncSetAttribute(pf_gemm_kernel): %s", cudaGetErrorString(e));
    e = cudaFuncSetAttribute(pf_conv3_halo_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(pf_conv3_halo_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(pf_conv3_halo_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e != cudaSuccess) return set_error("cudaFuncSetAttribute(pf_conv3_halo_kernel): %s", cudaGetErrorString(e));
    cudaDeviceGetAttribute(&g_sm_counts[dev], cudaDevAttrMultiProcessorCount, dev);
    attr_done[dev] = true;
  }
  const int g_sm_count = g_sm_counts[dev];
  if (d.block_n % 32 != 0 || d.block_n < 32 || d.block_n > 256) return set_error("gemm: bad block_n %d", d.block_n);
  {
    const long long rows = d.a_mode == 1 ? static_cast<long long>(d.NB) * d.H * d.W : d.M;
    int ktrue = 0;
    for (int s = 0; s < d.num_src; ++s) ktrue += d.k_true[s];
    const int n_true = d.ps > 1 ? d.n_logical * d.ps * d.ps : d.N;
    note_work(2.0 * rows * ktrue * d.taps * n_true, "%s rows%lld K%dx%d N%d%s%s%s", d.taps == 9 ? "conv3x3" : (d.a_mode == 1 ? "conv1x1" : (d.ps > 1 ? "convT" : "linear")),
              rows, d.taps, ktrue, n_true, d.act ? (d.act == PF_ACT_GELU ? " gelu" : (d.act == PF_ACT_RELU ? " relu" : " softplus")) : "",
              d.gamma ? " gamma" : (d.vt ? " vt" : ""), d.w2 ? " tail" : "");
  }
  GemmKernelParams P;
  for (int s = 0; s < d.num_src; ++s) P.tmA[s] = tmA[s];
  for (int s = d.num_src; s < 3; ++s) P.tmA[s] = tmA[0];
  P.tmB = tmB;
  P.tmBh = tmBh ? *tmBh : tmB;
  P.tmOut = tmOut ? *tmOut : tmB;
  P.d = d;
  if (d.tma_out && !tmOut) return set_error("gemm: tma_out without an output tensor map");
  P.kc = 1;
  int stage_bytes = kATileBytes + d.block_n * kBlockK * 2;
  int budget = kMaxSmem - 1024 /*align*/ - kEpiSmemBytes /*barriers + epilogue staging (aliases the fused-tail scratch)*/;
  int stages = budget / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) return set_error("gemm: not enough shared memory for 2 stages");
  P.stages = stages;
  int ks = 0;
  for (int s = 0; s < d.num_src; ++s) ks += d.chunks[s] * d.taps;
  P.k_steps = ks;
  P.total_tiles = d.m_tiles * d.n_tiles;
  if (P.total_tiles <= 0 || ks <= 0) return set_error("gemm: empty problem");
  int grid = P.total_tiles < g_sm_count ? P.total_tiles : g_sm_count;
  size_t smem = 1024 + static_cast<size_t>(stages) * stage_bytes + kEpiSmemBytes;
  if (d.halo) {
    int b_bytes = d.block_n * kBlockK * 2;
    int hb = kMaxSmem - 1024 - kTailBytes - kHaloSlots * kHaloSlot;
    // taps per weight stage: amortise the per-stage barrier round trip (~500 clk) over >= ~512 clk of MMA work
    int kc = 1;
    if (9 * b_bytes * 2 <= hb) kc = 9;
    else if (3 * b_bytes * 2 <= hb && d.block_n < 256) kc = 3;
    static const int kc_force = getenv("PF_B200_HALO_KC") ? atoi(getenv("PF_B200_HALO_KC")) : 0;   // tuning hook
    if ((kc_force == 1 || kc_force == 3 || kc_force == 9) && kc_force * b_bytes * 2 <= hb) kc = kc_force;
    P.kc = kc;
    b_bytes *= kc;
    int hstages = hb / b_bytes;
    if (hstages > 6) hstages = 6;
    if (hstages < 2) return set_error("conv3 halo: not enough shared memory");
    P.stages = hstages;
    size_t hsmem = 1024 + static_cast<size_t>(kHaloSlots) * kHaloSlot + static_cast<size_t>(hstages) * b_bytes + kTailBytes;
    cudaError_t le;
    if (tmBh != nullptr) {
      // weight-multicast clusters of cl CTAs over (m-tile group, n-tile) work items.  Every CTA is persistent, so the grid
      // must not exceed what is co-resident: ask the driver how many clusters of this shape fit (GPC boundaries).
      const int cl = d.halo_cl;
      const int groups = ((d.m_tiles + cl - 1) / cl) * d.n_tiles;
      cudaLaunchConfig_t cfg = {};
      cfg.blockDim = dim3(kGemmThreads); cfg.dynamicSmemBytes = hsmem; cfg.stream = stream;
      cudaLaunchAttribute at[2];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[1].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      static int max_clusters[kMaxDevices][5] = {};
      if (max_clusters[dev][cl] == 0) {
        cfg.gridDim = dim3(cl * (g_sm_count / cl));
        int n = 0;
        cudaError_t qe = cl == 2 ? cudaOccupancyMaxActiveClusters(&n, pf_conv3_halo_kernel<2>, &cfg)
                                 : cudaOccupancyMaxActiveClusters(&n, pf_conv3_halo_kernel<4>, &cfg);
        if (qe != cudaSuccess || n < 1) { cudaGetLastError(); n = g_sm_count / cl; }
        max_clusters[dev][cl] = n < g_sm_count / cl ? n : g_sm_count / cl;
      }
      const int clusters = groups < max_clusters[dev][cl] ? groups : max_clusters[dev][cl];
      cfg.gridDim = dim3(cl * clusters);
      cfg.numAttrs = pdl_enabled() ? 2 : 1;
      le = cl == 2 ? cudaLaunchKernelEx(&cfg, pf_conv3_halo_kernel<2>, P) : cudaLaunchKernelEx(&cfg, pf_conv3_halo_kernel<4>, P);
    } else {
      le = launch_pdl(pf_conv3_halo_kernel<1>, dim3(grid), dim3(kGemmThreads), hsmem, stream, P);
    }
    if (le != cudaSuccess) return set_error("pf_conv3_halo_kernel launch: %s", cudaGetErrorString(le));
  } else if (tmBh != nullptr) {
    // weight-multicast pairs: clusters of 2 CTAs over (m-tile pair, n-tile) work items
    const int pairs = ((d.m_tiles + 1) / 2) * d.n_tiles;
    const int clusters = pairs < g_sm_count / 2 ? pairs : g_sm_count / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters); cfg.blockDim = dim3(kGemmThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    cudaError_t le = cudaLaunchKernelEx(&cfg, pf_gemm_kernel<true>, P);
    if (le != cudaSuccess) return set_error("pf_gemm_kernel<multicast> launch: %s", cudaGetErrorString(le));
  } else {
    cudaError_t le = launch_pdl(pf_gemm_kernel<false>, dim3(grid), dim3(kGemmThreads), smem, stream, P);
    if (le != cudaSuccess) return set_error("pf_gemm_kernel launch: %s", cudaGetErrorString(le));
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("pf_gemm_kernel launch: %s", cudaGetErrorString(e));
  count_launch(d.halo ? "pf_conv3_halo_kernel" : "pf_gemm_kernel");
  return 0;
}

#ifdef PF_GEMM_TRACE
extern "C" int pf_gemm_trace_read(unsigned long long* out) {
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(out, g_gemm_trace, sizeof(g_gemm_trace)) == cudaSuccess ? 0 : 1;
}
extern "C" int pf_gemm_trace_clear() {
  static unsigned long long z[4 * 256 * 2] = {0};
  return cudaMemcpyToSymbol(g_gemm_trace, z, sizeof(z)) == cudaSuccess ? 0 : 1;
}
#endif

}  // namespace pf
