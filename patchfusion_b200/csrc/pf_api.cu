// C-ABI glue: error reporting, TMA tensor-map construction (+cache), pf_gemm host wrapper, weight packing.
#include <cuda_bf16.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "pf_kernels.h"

namespace pf {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}
// ---- per-launch profiler (bench.py's roofline pass): one CUDA event after every launch of this library on the
// profiled stream; the duration of launch i is event[i] - event[i-1] (launches are back to back on one stream).
struct ProfRec { const char* name; double flops; char label[80]; cudaEvent_t ev; };
static std::vector<ProfRec> g_prof;
static bool g_prof_on = false;
static cudaStream_t g_prof_stream = nullptr;
static cudaEvent_t g_prof_ev0 = nullptr;
static thread_local double g_next_flops = 0.0;
static thread_local char g_next_label[80] = "";

void note_work(double flops, const char* fmt, ...) {
  if (!g_prof_on) return;
  g_next_flops = flops;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_next_label, sizeof(g_next_label), fmt, ap);
  va_end(ap);
}
void count_launch(const char* name) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (g_prof_on) {
    ProfRec r;
    r.name = name; r.flops = g_next_flops;
    snprintf(r.label, sizeof(r.label), "%s", g_next_label[0] ? g_next_label : name);
    cudaEventCreate(&r.ev);
    cudaEventRecord(r.ev, g_prof_stream);
    g_prof.push_back(r);
    g_next_flops = 0.0; g_next_label[0] = 0;
  }
}
// tuning switches: -1 = not read yet (first use reads the environment variable of the same name)
static std::atomic<int> g_opt[6] = {{-1}, {-1}, {-1}, {-1}, {-1}, {-1}};
int option(int which) {
  static const char* names[6] = {"PF_OPT_TMA_EPILOGUE", "PF_OPT_HALO_MULTICAST", "PF_OPT_GEMM_MULTICAST", "PF_OPT_FUSED_RESAMPLE", "PF_OPT_PDL", "PF_OPT_RESIZE_SEPARABLE"};
  static const int defaults[6] = {1, 1, 1, 0, 0, 0};
  if (which < 0 || which > 5) return 0;
  int v = g_opt[which].load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv(names[which]);
    if (!e && which == PF_OPT_PDL) e = getenv("PF_B200_PDL");
    v = e ? atoi(e) : defaults[which];
    g_opt[which].store(v, std::memory_order_relaxed);
  }
  return v;
}
bool pdl_enabled() {
  // opt-in (PF_OPT_PDL / legacy PF_B200_PDL=1): measured neutral inside CUDA graphs on B200 (every tensor-core kernel is a
  // 1-CTA/SM persistent kernel, so a dependent cannot become resident before its predecessor's CTAs retire)
  return option(PF_OPT_PDL) != 0;
}
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("%s launch: %s", what, cudaGetErrorString(e));
  count_launch(what);
  return 0;
}

// ---------------------------------------------------------------------------------------------- tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  uint64_t v[12];
  bool operator==(const MapKey& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < 12; ++i) { h ^= k.v[i]; h *= 1099511628211ull; }
    return static_cast<size_t>(h);
  }
};
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;
static std::mutex g_maps_mu;

static int encode(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16) {
  MapKey key;
  memset(&key, 0, sizeof(key));
  key.v[0] = reinterpret_cast<uint64_t>(ptr);
  key.v[1] = static_cast<uint64_t>(rank) | (static_cast<uint64_t>(dt) << 8);
  for (int i = 0; i < rank; ++i) {
    key.v[2 + i] = dims[i];
    key.v[6 + i] = (i + 1 < rank ? strides_bytes[i] : 0) ^ (static_cast<uint64_t>(box[i]) << 48);
  }
  {
    std::lock_guard<std::mutex> g(g_maps_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return 0; }
  }
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  if (reinterpret_cast<uintptr_t>(ptr) & 15) return set_error("tensor map: base pointer not 16-byte aligned");
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) {
    gstr[i] = strides_bytes[i];
    if (gstr[i] & 15) return set_error("tensor map: stride %d (%llu B) not a multiple of 16", i, (unsigned long long)gstr[i]);
  }
  CUresult r = fn(out, dt, rank, const_cast<void*>(ptr), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu %llu %llu box %u %u %u %u", (int)r, rank,
                     (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                     (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
                     rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
  }
  std::lock_guard<std::mutex> g(g_maps_mu);
  if (g_maps.size() > 65536) g_maps.clear();      // callers with ever-changing pointers: bound the cache
  g_maps.emplace(key, *out);
  return 0;
}

int tmap_2d_bf16(CUtensorMap* out, const void* ptr, uint64_t cols, uint64_t rows, uint64_t ld, uint32_t box_cols,
                 uint32_t box_rows) {
  uint64_t dims[2] = {cols, rows};
  uint64_t str[1] = {ld * 2};
  uint32_t box[2] = {box_cols, box_rows};
  return encode(out, ptr, 2, dims, str, box);
}
int tmap_2d_f32(CUtensorMap* out, const void* ptr, uint64_t cols, uint64_t rows, uint64_t ld, uint32_t box_cols,
                uint32_t box_rows) {
  uint64_t dims[2] = {cols, rows};
  uint64_t str[1] = {ld * 4};
  uint32_t box[2] = {box_cols, box_rows};
  return encode(out, ptr, 2, dims, str, box, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
}
int tmap_3d_bf16(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t ld1, uint64_t ld2,
                 uint32_t b0, uint32_t b1, uint32_t b2) {
  uint64_t dims[3] = {d0, d1, d2};
  uint64_t str[2] = {ld1 * 2, ld2 * 2};
  uint32_t box[3] = {b0, b1, b2};
  return encode(out, ptr, 3, dims, str, box);
}
int tmap_4d_nhwc_bf16(CUtensorMap* out, const void* ptr, uint64_t C, uint64_t W, uint64_t H, uint64_t N, uint64_t ld,
                      uint32_t box_c, uint32_t box_w, uint32_t box_h) {
  uint64_t dims[4] = {C, W, H, N};
  uint64_t str[3] = {ld * 2, ld * 2 * W, ld * 2 * W * H};
  uint32_t box[4] = {box_c, box_w, box_h, 1};
  return encode(out, ptr, 4, dims, str, box);
}

// ---------------------------------------------------------------------------------------------- weight packing
__global__ void pack_weight_kernel(const float* __restrict__ w, int N, int N_pad, int num_src, int c0, int c1, int c2,
                                   int taps, const float* __restrict__ scale, __nv_bfloat16* __restrict__ dst,
                                   int Ktot) {
  long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  long long total = static_cast<long long>(N_pad) * Ktot;
  if (idx >= total) return;
  int n = static_cast<int>(idx / Ktot);
  int k = static_cast<int>(idx - static_cast<long long>(n) * Ktot);
  int cs[3] = {c0, c1, c2};
  int ctot = c0 + (num_src > 1 ? c1 : 0) + (num_src > 2 ? c2 : 0);
  float v = 0.0f;
  int cbase = 0, kbase = 0;
  for (int s = 0; s < num_src; ++s) {
    int cp = (cs[s] + 63) / 64 * 64;
    int seg = taps * cp;
    if (k < kbase + seg) {
      int kk = k - kbase;
      int tap = kk / cp, c = kk - tap * cp;
      if (n < N && c < cs[s]) {
        // PyTorch layout [N][Ctot][kh][kw], tap = ky*3+kx
        v = w[(static_cast<long long>(n) * ctot + cbase + c) * taps + tap];
        if (scale) v *= scale[n];
      }
      break;
    }
    kbase += seg;
    cbase += cs[s];
  }
  dst[idx] = __float2bfloat16(v);
}

__global__ void pack_convT_kernel(const float* __restrict__ w, int Cin, int Cout, int k, __nv_bfloat16* __restrict__ dst,
                                  int Kp) {
  long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  int Cp = (Cout + 31) / 32 * 32;
  long long total = static_cast<long long>(k) * k * Cp * Kp;
  if (idx >= total) return;
  int row = static_cast<int>(idx / Kp);
  int ci = static_cast<int>(idx - static_cast<long long>(row) * Kp);
  int tap = row / Cp, co = row - tap * Cp;
  float v = 0.0f;
  if (ci < Cin && co < Cout) v = w[(static_cast<long long>(ci) * Cout + co) * k * k + tap];   // [Cin][Cout][ky][kx]
  dst[idx] = __float2bfloat16(v);
}

}  // namespace pf

using namespace pf;

extern "C" {

const char* pf_last_error(void) { return g_err; }
int pf_version(void) { return 100; }
long long pf_launch_count(void) { return g_launches.load(); }

int pf_set_option(int32_t which, int32_t value) {
  if (which < 0 || which > 5) return set_error("pf_set_option: unknown option %d", which);
  g_opt[which].store(value < 0 ? 0 : value, std::memory_order_relaxed);
  return 0;
}

int pf_profile_start(void* stream) {
  for (auto& r : g_prof) cudaEventDestroy(r.ev);
  g_prof.clear();
  if (g_prof_ev0) cudaEventDestroy(g_prof_ev0);
  g_prof_stream = static_cast<cudaStream_t>(stream);
  cudaEventCreate(&g_prof_ev0);
  cudaEventRecord(g_prof_ev0, g_prof_stream);
  g_prof_on = true;
  return 0;
}
int pf_profile_stop(void) {
  g_prof_on = false;
  if (!g_prof.empty()) cudaEventSynchronize(g_prof.back().ev);
  return static_cast<int>(g_prof.size());
}
int pf_profile_get(int32_t i, const char** name, const char** label, double* flops, float* ms) {
  if (i < 0 || i >= static_cast<int>(g_prof.size())) return set_error("pf_profile_get: index %d out of range", i);
  const ProfRec& r = g_prof[i];
  *name = r.name; *label = r.label; *flops = r.flops;
  cudaError_t e = cudaEventElapsedTime(ms, i == 0 ? g_prof_ev0 : g_prof[i - 1].ev, r.ev);
  if (e != cudaSuccess) return set_error("cudaEventElapsedTime: %s", cudaGetErrorString(e));
  return 0;
}

int pf_pack_weight(const float* w, int32_t N, int32_t N_pad, int32_t num_src, const int32_t* src_c, int32_t taps,
                   const float* scale, void* dst, void* stream) {
  if (num_src < 1 || num_src > 3 || (taps != 1 && taps != 9)) return set_error("pf_pack_weight: bad arguments");
  int c[3] = {src_c[0], num_src > 1 ? src_c[1] : 0, num_src > 2 ? src_c[2] : 0};
  int Ktot = 0;
  for (int s = 0; s < num_src; ++s) Ktot += taps * ((c[s] + 63) / 64 * 64);
  long long total = static_cast<long long>(N_pad) * Ktot;
  int threads = 256;
  long long blocks = (total + threads - 1) / threads;
  pack_weight_kernel<<<static_cast<unsigned>(blocks), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      w, N, N_pad, num_src, c[0], c[1], c[2], taps, scale, static_cast<__nv_bfloat16*>(dst), Ktot);
  return check_launch("pack_weight_kernel");
}

int pf_pack_weight_convT(const float* w, int32_t Cin, int32_t Cout, int32_t k, void* dst, void* stream) {
  int Kp = (Cin + 63) / 64 * 64;
  long long total = static_cast<long long>(k) * k * ((Cout + 31) / 32 * 32) * Kp;
  int threads = 256;
  long long blocks = (total + threads - 1) / threads;
  pack_convT_kernel<<<static_cast<unsigned>(blocks), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      w, Cin, Cout, k, static_cast<__nv_bfloat16*>(dst), Kp);
  return check_launch("pack_convT_kernel");
}

static void choose_tile(int H, int W, int* bh, int* bw) {
  const int cand[6][2] = {{8, 16}, {4, 32}, {16, 8}, {2, 64}, {1, 128}, {32, 4}};
  long long best = -1;
  for (int i = 0; i < 6; ++i) {
    long long ty = (H + cand[i][0] - 1) / cand[i][0], tx = (W + cand[i][1] - 1) / cand[i][1];
    long long cost = ty * tx;
    if (best < 0 || cost < best) { best = cost; *bh = cand[i][0]; *bw = cand[i][1]; }
  }
}

int pf_gemm(pf_gemm_desc* u, void* stream) {
  if (!u) return set_error("pf_gemm: null descriptor");
  if (u->num_src < 1 || u->num_src > 3) return set_error("pf_gemm: num_src %d", u->num_src);
  if (u->taps != 1 && u->taps != 9) return set_error("pf_gemm: taps %d", u->taps);
  if (u->taps == 9 && u->a_mode != 1) return set_error("pf_gemm: 3x3 window needs a_mode 1");
  if (u->N <= 0) return set_error("pf_gemm: N %d", u->N);
  if (u->out_ld % 8 || u->out_col0 % 8) return set_error("pf_gemm: out_ld/out_col0 must be multiples of 8");
  GemmDesc d;
  memset(&d, 0, sizeof(d));
  d.num_src = u->num_src; d.a_mode = u->a_mode; d.taps = u->taps;
  int ksteps = 0;
  for (int s = 0; s < u->num_src; ++s) {
    d.chunks[s] = (u->a_c[s] + 63) / 64;
    d.k_true[s] = u->a_c[s];
    u->chunks[s] = d.chunks[s];
    ksteps += d.chunks[s] * u->taps;
    if (u->a_c[s] % 8 || u->a_ld[s] % 8) return set_error("pf_gemm: source %d channels/ld must be multiples of 8", s);
  }
  if (u->Ktot != ksteps * 64) return set_error("pf_gemm: Ktot %d != %d expected from sources", u->Ktot, ksteps * 64);
  // N tiling
  int bn = u->block_n;
  if (bn == 0 && u->ps > 1) {
    // pixel shuffle: an N tile must not straddle two (ky,kx) taps
    int cpad = (u->ps_cout + 31) / 32 * 32;
    for (int c = 256; c >= 32; c -= 32)
      if (cpad % c == 0) { bn = c; break; }
  }
  if (bn == 0) {
    int n32 = (u->N + 31) / 32 * 32;
    if (n32 <= 256) bn = n32;
    else {
      // largest multiple of 32 in [128,256] minimising padded columns
      int bestpad = 1 << 30;
      for (int c = 256; c >= 128; c -= 32) {
        int pad = (u->N + c - 1) / c * c - u->N;
        if (pad < bestpad) { bestpad = pad; bn = c; }
      }
    }
  }
  d.block_n = bn; d.N = u->N; d.n_tiles = (u->N + bn - 1) / bn;
  u->block_n = bn; u->n_tiles = d.n_tiles;
  d.M = u->M; d.NB = u->NB; d.H = u->H; d.W = u->W;
  CUtensorMap tmA[3], tmB;
  // 3x3 convs go through the halo-tile kernel (one A fetch per 64-channel chunk instead of nine) unless the caller
  // pins a tile shape or PF_B200_NO_HALO is set.
  static const bool no_halo = getenv("PF_B200_NO_HALO") != nullptr;
  const bool halo = u->a_mode == 1 && u->taps == 9 && u->bh == 0 && u->bw == 0 && !no_halo;
  d.halo = halo ? 1 : 0;
  bool any_rs = false;
  for (int s = 0; s < u->num_src; ++s) {
    if (u->rs_h[s] < 0 || u->rs_w[s] < 0 || (u->rs_h[s] > 0) != (u->rs_w[s] > 0)) return set_error("pf_gemm: bad rs_h/rs_w of source %d", s);
    any_rs = any_rs || u->rs_h[s] > 0;
  }
  if (any_rs && !halo) return set_error("pf_gemm: resampled sources (rs_h > 0) need the 3x3 halo-tile path");
  if (halo) {
    d.bh = 16; d.bw = 8;
    d.tiles_y = (u->H + 15) / 16; d.tiles_x = (u->W + 7) / 8;
    d.m_tiles = u->NB * d.tiles_y * d.tiles_x;
    int first_plain = -1;
    for (int s = 0; s < u->num_src; ++s) {
      if (u->rs_h[s] > 0) {
        // read through a fused bilinear resample: no tensor map, the producer warps gather from the low-resolution map
        d.rs_ptr[s] = static_cast<const __nv_bfloat16*>(u->a_ptr[s]);
        d.rs_h[s] = u->rs_h[s]; d.rs_w[s] = u->rs_w[s]; d.rs_ld[s] = u->a_ld[s];
        d.rs_sy[s] = u->H > 1 ? static_cast<float>(u->rs_h[s] - 1) / static_cast<float>(u->H - 1) : 0.f;
        d.rs_sx[s] = u->W > 1 ? static_cast<float>(u->rs_w[s] - 1) / static_cast<float>(u->W - 1) : 0.f;
        d.rs_any = 1;
        if (reinterpret_cast<uintptr_t>(u->a_ptr[s]) & 15) return set_error("pf_gemm: resampled source %d not 16-byte aligned", s);
        continue;
      }
      if (tmap_4d_nhwc_bf16(&tmA[s], u->a_ptr[s], u->a_c[s], u->W, u->H, u->NB, u->a_ld[s], 64, 10, 18)) return 1;
      if (first_plain < 0) first_plain = s;
    }
    for (int s = 0; s < u->num_src; ++s)       // placeholder maps for the resampled sources (never dereferenced)
      if (u->rs_h[s] > 0) {
        if (first_plain >= 0) tmA[s] = tmA[first_plain];
        else memset(&tmA[s], 0, sizeof(CUtensorMap));
      }
  } else if (u->a_mode == 1) {
    int bh = u->bh, bw = u->bw;
    if (bh == 0 || bw == 0) choose_tile(u->H, u->W, &bh, &bw);
    if (bh * bw != 128) return set_error("pf_gemm: bh*bw must be 128");
    d.bh = bh; d.bw = bw;
    d.tiles_y = (u->H + bh - 1) / bh; d.tiles_x = (u->W + bw - 1) / bw;
    d.m_tiles = u->NB * d.tiles_y * d.tiles_x;
    for (int s = 0; s < u->num_src; ++s)
      if (tmap_4d_nhwc_bf16(&tmA[s], u->a_ptr[s], u->a_c[s], u->W, u->H, u->NB, u->a_ld[s], 64, bw, bh)) return 1;
  } else {
    d.m_tiles = (u->M + 127) / 128;
    for (int s = 0; s < u->num_src; ++s)
      if (tmap_2d_bf16(&tmA[s], u->a_ptr[s], u->a_c[s], u->M, u->a_ld[s], 64, 128)) return 1;
  }
  u->bh = d.bh; u->bw = d.bw; u->tiles_y = d.tiles_y; u->tiles_x = d.tiles_x; u->m_tiles = d.m_tiles;
  int n_pad = d.n_tiles * bn;
  if (tmap_2d_bf16(&tmB, u->w_ptr, u->Ktot, n_pad, u->Ktot, 64, bn)) return 1;
  d.bias = u->bias; d.act = u->act;
  d.res1 = static_cast<const __nv_bfloat16*>(u->res1);
  d.res2 = static_cast<const __nv_bfloat16*>(u->res2);
  d.res_ld = u->res_ld;
  d.gamma = u->gamma;
  d.out = u->out; d.out_f32 = u->out_f32 || u->gamma != nullptr; d.out_ld = u->out_ld; d.out_col0 = u->out_col0;
  d.out2 = static_cast<__nv_bfloat16*>(u->out2); d.out2_ld = u->out2_ld;
  if ((u->out_f32 || u->gamma) && reinterpret_cast<uintptr_t>(u->out) % 32 != 0)
    return set_error("pf_gemm: fp32 outputs must be 32-byte aligned (256-bit epilogue stores)");
  // 256-bit epilogue accesses need 32-byte aligned rows: bf16 pitches / offsets in multiples of 16 elements
  d.wide = ((reinterpret_cast<uintptr_t>(u->out) | reinterpret_cast<uintptr_t>(u->out2) | reinterpret_cast<uintptr_t>(u->res1) |
             reinterpret_cast<uintptr_t>(u->res2)) % 32 == 0) &&
           u->out_col0 % 16 == 0 && (d.out_f32 || u->out_ld % 16 == 0) && (!u->out2 || u->out2_ld % 16 == 0) &&
           (!(u->res1 || u->res2) || u->res_ld % 16 == 0);
  d.ps = u->ps > 1 ? u->ps : 1;
  d.n_logical = u->N;
  if (d.ps > 1) {
    if (u->a_mode != 0) return set_error("pf_gemm: pixel shuffle needs a_mode 0");
    int cpad = (u->ps_cout + 31) / 32 * 32;           // per-tap column stride of pf_pack_weight_convT
    if (cpad % bn != 0) return set_error("pf_gemm: padded ps_cout %d must be a multiple of block_n %d", cpad, bn);
    if (u->N != d.ps * d.ps * cpad) return set_error("pf_gemm: pixel shuffle N %d != k*k*pad32(Cout) %d", u->N, d.ps * d.ps * cpad);
    d.ps_cout_pad = cpad;
    d.n_logical = u->ps_cout;
  }
  d.w2 = u->w2; d.b2 = u->b2; d.n2 = u->n2; d.act2 = u->act2; d.skip_main = u->skip_main;
  d.out3 = u->out3; d.out3_ld = u->out3_ld;
  if (d.w2) {
    if (d.n_tiles != 1) return set_error("pf_gemm: fused trailing layer needs the whole row in one N tile (N %d)", u->N);
    if (d.n2 < 1 || d.n2 > 16 || !d.out3) return set_error("pf_gemm: fused trailing layer n2 %d (1..16), out3 required", d.n2);
    if (u->N % 4) return set_error("pf_gemm: fused trailing layer needs N % 4 == 0");
  }
  d.vt = static_cast<__nv_bfloat16*>(u->vt);
  d.vt_col0 = u->vt_col0; d.vt_seq = u->vt_seq; d.vt_seq_pad = u->vt_seq_pad; d.vt_dim = u->vt_dim;
  if (d.vt && (d.vt_col0 % bn) != 0) return set_error("pf_gemm: vt_col0 must be a multiple of block_n");
  // Linear layers with several m-tiles per n-tile are L2 -> SM bandwidth bound: pairs of CTAs share the weight tile by
  // TMA multicast (PF_OPT_GEMM_MULTICAST / PF_OPT_HALO_MULTICAST).  Needs an even split of the n-tile into 1024-B aligned halves.
  // PF_OPT_HALO_MULTICAST: 0 off, 1 clusters of 2, 2 clusters of 4
  int cl = 1;
  if (d.m_tiles >= 4 && static_cast<long long>(d.m_tiles) * d.n_tiles >= 148) {
    if (halo) cl = option(PF_OPT_HALO_MULTICAST) >= 2 ? 4 : (option(PF_OPT_HALO_MULTICAST) == 1 ? 2 : 1);
    else if (option(PF_OPT_GEMM_MULTICAST) != 0 && u->a_mode == 0 && d.ps == 1 && bn % 16 == 0) cl = 2;
  }
  d.halo_cl = halo ? cl : 1;
  const bool mc = cl > 1;
  CUtensorMap tmBh;
  if (mc && tmap_2d_bf16(&tmBh, u->w_ptr, u->Ktot, n_pad, u->Ktot, 64, bn / cl)) return 1;
  // Epilogue through shared memory + TMA (pf_gemm_kernel only): plain bf16 outputs in 64-column groups, fp32 outputs
  // and the fp32 residual stream (x += gamma * v) in 32-column chunks.  PF_OPT_TMA_EPILOGUE = 0 keeps the direct stores.
  const bool no_tma_epi = option(PF_OPT_TMA_EPILOGUE) == 0;
  CUtensorMap tmOut;
  d.tma_out = 0;
  if (!no_tma_epi && !halo && d.ps == 1 && !d.w2 && !d.res1 && !d.res2 && !d.out2) {
    const uint64_t ocols = static_cast<uint64_t>(u->out_col0) + u->N;     // columns >= N are clipped by the copy
    if (!d.out_f32 && bn % 64 == 0 && reinterpret_cast<uintptr_t>(u->out) % 16 == 0) {
      if (u->a_mode == 0) {
        if (tmap_2d_bf16(&tmOut, u->out, ocols, u->M, u->out_ld, 64, 32)) return 1;
      } else {
        const uint32_t bwx = d.bw < 32 ? d.bw : 32;
        if (tmap_4d_nhwc_bf16(&tmOut, u->out, ocols, u->W, u->H, u->NB, u->out_ld, 64, bwx, 32 / bwx)) return 1;
      }
      d.tma_out = 1;
    } else if (d.out_f32 && u->a_mode == 0 && u->out_ld % 4 == 0 && reinterpret_cast<uintptr_t>(u->out) % 16 == 0) {
      if (tmap_2d_f32(&tmOut, u->out, ocols, u->M, u->out_ld, 32, 32)) return 1;
      d.tma_out = 2;
    }
  }
  return gemm_launch(d, tmA, tmB, mc ? &tmBh : nullptr, d.tma_out ? &tmOut : nullptr, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
