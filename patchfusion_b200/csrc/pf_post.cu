// Callers downstream of the hot path: fused depth-metric reduction and colour mapping.
//
//   pf_depth_metrics  `estimator/utils/metric.py:10-50` (compute_errors), `:67-72` (soft_edge_error), `:97-148`
//                     (compute_metrics: resample, clamp, validity mask) as ONE pass over the ground-truth grid and a
//                     fixed-order second stage - the reference does ~25 full-image numpy passes on the host.
//   pf_colorize_u8    `estimator/utils/color.py:95-140` after the percentile normalisation: LUT lookup to 8-bit BGR/RGB.
#include "pf_common.cuh"
#include "pf_kernels.h"

namespace pf {

constexpr int kMetricSums = 12;

// F.interpolate(mode='bilinear', align_corners=False) sample of a [PH, PW] map at gt-grid pixel (y, x).
__device__ __forceinline__ float sample_pred(const float* __restrict__ pred, int PH, int PW, int H, int W, int y, int x) {
  if (PH == H && PW == W) return pred[static_cast<long long>(y) * W + x];
  const float sy = fmaxf((y + 0.5f) * (static_cast<float>(PH) / H) - 0.5f, 0.0f);
  const float sx = fmaxf((x + 0.5f) * (static_cast<float>(PW) / W) - 0.5f, 0.0f);
  const int y0 = min(static_cast<int>(sy), PH - 1), x0 = min(static_cast<int>(sx), PW - 1);
  const int y1 = min(y0 + 1, PH - 1), x1 = min(x0 + 1, PW - 1);
  const float fy = sy - y0, fx = sx - x0;
  const float* r0 = pred + static_cast<long long>(y0) * PW;
  const float* r1 = pred + static_cast<long long>(y1) * PW;
  return (1.f - fy) * ((1.f - fx) * r0[x0] + fx * r0[x1]) + fy * ((1.f - fx) * r1[x0] + fx * r1[x1]);
}

__device__ __forceinline__ float clamp_pred(float p, float lo, float hi) {
  // metric.py:107-110: < min -> min, > max -> max, inf -> max, nan -> min
  if (isnan(p)) return lo;
  return fminf(fmaxf(p, lo), hi);
}

__global__ void __launch_bounds__(256) depth_metrics_kernel(const float* __restrict__ pred, int PH, int PW,
                                                            const float* __restrict__ gt, int H, int W, float lo, float hi,
                                                            const uint8_t* __restrict__ edges,
                                                            const uint8_t* __restrict__ extra, double* __restrict__ partials) {
  double acc[kMetricSums];
#pragma unroll
  for (int i = 0; i < kMetricSums; ++i) acc[i] = 0.0;
  const long long total = static_cast<long long>(H) * W;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int y = static_cast<int>(idx / W), x = static_cast<int>(idx - static_cast<long long>(y) * W);
    const float g = gt[idx];
    bool valid = g > lo && g < hi;
    if (extra != nullptr) valid = valid && extra[idx] != 0;
    if (!valid) continue;
    const float p = clamp_pred(sample_pred(pred, PH, PW, H, W, y, x), lo, hi);
    const float th = fmaxf(g / p, p / g);
    const float d = g - p;
    const float lg = logf(g), lp = logf(p);
    acc[0] += 1.0;
    acc[1] += th < 1.25f ? 1.0 : 0.0;
    acc[2] += th < 1.25f * 1.25f ? 1.0 : 0.0;
    acc[3] += th < 1.25f * 1.25f * 1.25f ? 1.0 : 0.0;
    acc[4] += fabsf(d) / g;
    acc[5] += d * d / g;
    acc[6] += static_cast<double>(d) * d;
    acc[7] += static_cast<double>(lg - lp) * (lg - lp);
    acc[8] += lp - lg;
    acc[9] += fabsf(log10f(g) - log10f(p));
    if (edges != nullptr && edges[idx] != 0) {
      // soft edge error: min over the 3x3 neighbourhood of |gt_shifted - pred|, out-of-image gt = 0
      float best = INFINITY;
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          const int yy = y - dy, xx = x - dx;
          const float gs = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? gt[static_cast<long long>(yy) * W + xx] : 0.0f;
          best = fminf(best, fabsf(gs - p));
        }
      acc[10] += 1.0;
      acc[11] += best;
    }
  }
  // block reduction in a fixed order: lanes by shuffle tree, warps sequentially
  __shared__ double sh[8][kMetricSums];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < kMetricSums; ++i) {
    double v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < kMetricSums) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += sh[w][threadIdx.x];
    partials[static_cast<long long>(blockIdx.x) * kMetricSums + threadIdx.x] = v;
  }
}

__global__ void depth_metrics_final_kernel(const double* __restrict__ partials, int nblocks, double* __restrict__ out) {
  const int i = threadIdx.x;
  if (i >= kMetricSums) return;
  double v = 0.0;
  for (int b = 0; b < nblocks; ++b) v += partials[static_cast<long long>(b) * kMetricSums + i];
  out[i] = v;
}

__global__ void colorize_kernel(const float* __restrict__ d, long long n, float vmin, float vmax, float invalid_val,
                                const uint8_t* __restrict__ lut, int bgr, uint8_t bg0, uint8_t bg1, uint8_t bg2,
                                uint8_t* __restrict__ out) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const float v = d[i];
  uint8_t c0, c1, c2;
  if (v == invalid_val || isnan(v)) {
    c0 = bg0; c1 = bg1; c2 = bg2;
  } else {
    // color.py:121-125 then matplotlib Colormap.__call__: index = int(x * N), x == 1 -> N - 1, clipped to [0, N-1]
    const float t = vmin != vmax ? (v - vmin) / (vmax - vmin) : 0.0f;
    int k = static_cast<int>(t * 256.0f);
    if (t < 0.0f) k = 0;                        // under -> lowest LUT entry
    k = min(max(k, 0), 255);
    c0 = lut[3 * k]; c1 = lut[3 * k + 1]; c2 = lut[3 * k + 2];
  }
  uint8_t* o = out + 3 * i;
  if (bgr) { o[0] = c2; o[1] = c1; o[2] = c0; } else { o[0] = c0; o[1] = c1; o[2] = c2; }
}

}  // namespace pf

using namespace pf;

extern "C" {

int pf_depth_metrics(const float* pred, int32_t PH, int32_t PW, const float* gt, int32_t H, int32_t W, float min_eval,
                     float max_eval, const uint8_t* edges, const uint8_t* extra_mask, double* partials, int32_t nblocks,
                     double* out, void* stream) {
  if (nblocks < 1 || H < 1 || W < 1 || PH < 1 || PW < 1) return set_error("pf_depth_metrics: bad shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  depth_metrics_kernel<<<nblocks, 256, 0, st>>>(pred, PH, PW, gt, H, W, min_eval, max_eval, edges, extra_mask, partials);
  if (check_launch("depth_metrics_kernel")) return 1;
  depth_metrics_final_kernel<<<1, 32, 0, st>>>(partials, nblocks, out);
  return check_launch("depth_metrics_final_kernel");
}

int pf_colorize_u8(const float* depth, int64_t n, float vmin, float vmax, float invalid_val, const uint8_t* lut_rgb,
                   int32_t bgr, uint8_t* out, void* stream) {
  if (n < 0) return set_error("pf_colorize_u8: n < 0");
  if (n == 0) return 0;
  colorize_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      depth, n, vmin, vmax, invalid_val, lut_rgb, bgr, 128, 128, 128, out);
  return check_launch("colorize_kernel");
}

}  // extern "C"
