// modules.ConvFlow.forward (reference modules.py:346-390) as a standalone operator: the spline coupling flow the
// north_star names.  The reference defines it but never instantiates it, so it takes its weights explicitly.
//
//     x0, x1 = split(x)                       half = C / 2 channels each
//     h = pre(x0)                             Conv1d(half -> F, 1)                                  modules.py:365
//     h = DDSConv(h, mask)                    n_layers x { depthwise dilated k-tap conv of h * mask (dilation k^i),
//                                             LayerNorm over channels, GELU, Conv1d(F -> F, 1), LayerNorm, GELU,
//                                             h = h + y } ; * mask                                  modules.py:96-108, 20-32
//     h = proj(h) * mask                      Conv1d(F -> half * (3 nb - 1), 1)                     modules.py:367
//     widths, heights = h[.., :nb], h[.., nb:2nb] / sqrt(F) ; derivatives = h[.., 2nb:]             modules.py:372-374
//     x1, logabsdet = rq_spline(x1, ..., inverse = reverse, tails = linear)                         transforms.py:12-193
//     y = cat(x0, x1) * mask ; logdet = sum(logabsdet * mask)  (forward only)                       modules.py:385-390
//
// Kernels (fp32 FFMA; the channel widths of this module -- 2..192 -- are far below a tensor-core tile and the module is
// not on the infer path): one thread owns one time step, channels are walked in shared memory, so every global access
// is coalesced along time and the weights are warp-uniform (broadcast) loads.
//   cf_pre_kernel          h = pre(x0)
//   cf_dds_layer_kernel    one DDSConv layer, both LayerNorms and the residual fused; tile of 64 steps x F channels
//   cf_proj_spline_kernel  proj for the 3 nb - 1 parameters of ONE (b, c, t) element + the spline + cat / mask;
//                          per-block partial sums of logabsdet * mask, reduced in a fixed order by cf_logdet_kernel
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <string>

#include "../../include/svk.h"
#include "spline.cuh"

extern "C" int svk__set_error(int code, const char* msg);

namespace svk {
namespace {

constexpr int CF_T = 64;  // time steps per block

__global__ void __launch_bounds__(CF_T) cf_pre_kernel(const float* __restrict__ x, int C, int half, int T, const float* __restrict__ w,
                                                      const float* __restrict__ bias, int F, float* __restrict__ h) {
  const int b = blockIdx.y, t = blockIdx.x * CF_T + threadIdx.x;
  if (t >= T) return;
  float xv[8];
  for (int c0 = 0; c0 < half; c0 += 8) {
    const int nc = min(8, half - c0);
    for (int c = 0; c < nc; ++c) xv[c] = x[((size_t)b * C + c0 + c) * T + t];
    for (int f = 0; f < F; ++f) {
      float acc = c0 == 0 ? bias[f] : h[((size_t)b * F + f) * T + t];
      for (int c = 0; c < nc; ++c) acc = fmaf(w[(size_t)f * half + c0 + c], xv[c], acc);
      h[((size_t)b * F + f) * T + t] = acc;
    }
  }
}

__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)); }

// F.layer_norm over the channel column of this thread (biased variance, eps inside the sqrt), then GELU, in place.
__device__ __forceinline__ void ln_gelu_column(float* col, int F, int stride, const float* __restrict__ gamma,
                                               const float* __restrict__ beta, float eps) {
  float mean = 0.f;
  for (int c = 0; c < F; ++c) mean += col[c * stride];
  mean /= (float)F;
  float var = 0.f;
  for (int c = 0; c < F; ++c) {
    const float d = col[c * stride] - mean;
    var = fmaf(d, d, var);
  }
  const float rstd = rsqrtf(var / (float)F + eps);
  for (int c = 0; c < F; ++c) col[c * stride] = gelu_erf((col[c * stride] - mean) * rstd * gamma[c] + beta[c]);
}

__global__ void __launch_bounds__(CF_T) cf_dds_layer_kernel(const float* __restrict__ hin, const float* __restrict__ mask, int F, int T,
                                                            int K, int dil, const float* __restrict__ sep_w, const float* __restrict__ sep_b,
                                                            const float* __restrict__ g1, const float* __restrict__ b1,
                                                            const float* __restrict__ pw_w, const float* __restrict__ pw_b,
                                                            const float* __restrict__ g2, const float* __restrict__ b2, float eps,
                                                            float* __restrict__ hout) {
  extern __shared__ float sm[];
  float* ys = sm;                 // [F][CF_T]
  float* zs = sm + (size_t)F * CF_T;
  const int b = blockIdx.y, tid = threadIdx.x, t = blockIdx.x * CF_T + tid;
  const bool tin = t < T;
  const float* mrow = mask + (size_t)b * T;
  const int pad = (K * dil - dil) / 2;
  // y = convs_sep(x * mask): depthwise, dilation k^i, zero padding (modules.py:98)
  for (int c = 0; c < F; ++c) {
    float acc = sep_b[c];
    const float* xr = hin + ((size_t)b * F + c) * T;
    for (int j = 0; j < K; ++j) {
      const int tt = t - pad + j * dil;
      if (tin && tt >= 0 && tt < T) acc = fmaf(sep_w[c * K + j], xr[tt] * mrow[tt], acc);
    }
    ys[c * CF_T + tid] = acc;
  }
  ln_gelu_column(ys + tid, F, CF_T, g1, b1, eps);  // norms_1 + gelu
  // y = convs_1x1(y)
  for (int o = 0; o < F; ++o) {
    float acc = pw_b[o];
    const float* wr = pw_w + (size_t)o * F;
    for (int c = 0; c < F; ++c) acc = fmaf(wr[c], ys[c * CF_T + tid], acc);
    zs[o * CF_T + tid] = acc;
  }
  ln_gelu_column(zs + tid, F, CF_T, g2, b2, eps);  // norms_2 + gelu (dropout p = 0)
  if (tin)
    for (int c = 0; c < F; ++c) hout[((size_t)b * F + c) * T + t] = hin[((size_t)b * F + c) * T + t] + zs[c * CF_T + tid];  // x = x + y
}

__global__ void __launch_bounds__(CF_T) cf_proj_spline_kernel(const float* __restrict__ x, const float* __restrict__ h,
                                                              const float* __restrict__ mask, int C, int half, int F, int T, int nb,
                                                              const float* __restrict__ pw, const float* __restrict__ pb, float sqrt_f,
                                                              float tail_bound, int reverse, float* __restrict__ y,
                                                              float* __restrict__ partial, int32_t* __restrict__ bins) {
  extern __shared__ float sm[];
  float* hs = sm;                         // [F][CF_T]: (DDSConv output * mask) of this tile
  float* ps = sm + (size_t)F * CF_T;      // [3 nb - 1][CF_T]: the parameters of the current channel
  __shared__ float red[CF_T];
  const int b = blockIdx.y, tid = threadIdx.x, t = blockIdx.x * CF_T + tid;
  const bool tin = t < T;
  const float mv = tin ? mask[(size_t)b * T + t] : 0.f;
  for (int c = 0; c < F; ++c) hs[c * CF_T + tid] = tin ? h[((size_t)b * F + c) * T + t] * mv : 0.f;  // DDSConv returns x * mask
  const int P = 3 * nb - 1;
  float lsum = 0.f;
  for (int c = 0; c < half; ++c) {
    // h = proj(h) * mask, reshaped [b, c, 3nb-1, t] (modules.py:367-370)
    for (int p = 0; p < P; ++p) {
      const float* wr = pw + (size_t)(c * P + p) * F;
      float acc = pb[c * P + p];
      for (int f = 0; f < F; ++f) acc = fmaf(wr[f], hs[f * CF_T + tid], acc);
      ps[p * CF_T + tid] = acc * mv;
    }
    if (tin) {
      const float x1 = x[((size_t)b * C + half + c) * T + t];
      float yo, lo;
      int bin;
      rq_spline_element(
          x1, [&](int i) { return __fdiv_rn(ps[i * CF_T + tid], sqrt_f); }, [&](int i) { return __fdiv_rn(ps[(nb + i) * CF_T + tid], sqrt_f); },
          [&](int i) { return ps[(2 * nb + i) * CF_T + tid]; }, nb, reverse, tail_bound, 1e-3f, 1e-3f, 1e-3f, yo, lo, bin);
      y[((size_t)b * C + half + c) * T + t] = yo * mv;                       // cat([x0, x1], 1) * x_mask
      y[((size_t)b * C + c) * T + t] = x[((size_t)b * C + c) * T + t] * mv;
      if (bins) bins[((size_t)b * half + c) * T + t] = bin;
      lsum += lo * mv;
    }
  }
  // sum(logabsdet * x_mask, [1, 2]): fixed-order tree inside the block, blocks added in order by cf_logdet_kernel
  red[tid] = lsum;
  __syncthreads();
  for (int s = CF_T / 2; s > 0; s >>= 1) {
    if (tid < s) red[tid] += red[tid + s];
    __syncthreads();
  }
  if (tid == 0 && partial) partial[(size_t)b * gridDim.x + blockIdx.x] = red[0];
}

__global__ void cf_logdet_kernel(const float* __restrict__ partial, int nblk, float* __restrict__ logdet) {
  const int b = blockIdx.x;
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < nblk; ++i) s += partial[(size_t)b * nblk + i];
    logdet[b] = s;
  }
}

int cf_fail(int code, const std::string& m) { return svk__set_error(code, m.c_str()); }

}  // namespace
}  // namespace svk

using namespace svk;

extern "C" size_t svk_convflow_workspace_bytes(int B, int C, int T, int filter_channels, int num_bins) {
  if (B <= 0 || C <= 0 || T <= 0 || filter_channels <= 0) return 0;
  const size_t nblk = (size_t)(T + CF_T - 1) / CF_T;
  (void)C, (void)num_bins;
  return (2 * (size_t)B * filter_channels * T + (size_t)B * nblk) * sizeof(float) + 512;
}

extern "C" int svk_convflow(const float* x, const float* mask, int B, int C, int T, int F, int kernel_size, int n_layers, int num_bins,
                            float tail_bound, const svk_convflow_weights* w, int reverse, float* y, float* logdet, int32_t* bins,
                            void* workspace, size_t workspace_bytes, void* stream) {
  if (!x || !mask || !w || !y || !workspace) return cf_fail(SVK_ERR_INVALID, "svk_convflow: null argument");
  if (B <= 0 || T <= 0) return cf_fail(SVK_ERR_INVALID, "svk_convflow: B and T must be positive");
  if (C < 2 || C % 2) return cf_fail(SVK_ERR_INVALID, "svk_convflow: in_channels must be even");
  if (F < 1 || kernel_size < 1 || !(kernel_size & 1) || n_layers < 0) return cf_fail(SVK_ERR_INVALID, "svk_convflow: bad filter_channels / kernel_size / n_layers");
  if (num_bins < 1 || num_bins > SPLINE_MAX_BINS) return cf_fail(SVK_ERR_INVALID, "svk_convflow: num_bins out of range");
  if (!w->pre_w || !w->pre_b || !w->proj_w || !w->proj_b || (n_layers > 0 && (!w->sep_w || !w->sep_b || !w->pw_w || !w->pw_b || !w->norm1_g ||
                                                                               !w->norm1_b || !w->norm2_g || !w->norm2_b)))
    return cf_fail(SVK_ERR_INVALID, "svk_convflow: null weight pointer");
  // Minimal bin width / height x bins must stay below 1 (transforms.py:103-106)
  if (1e-3 * num_bins > 1.0) return cf_fail(SVK_ERR_INVALID, "Minimal bin width too large for the number of bins");
  if (workspace_bytes < svk_convflow_workspace_bytes(B, C, T, F, num_bins)) return cf_fail(SVK_ERR_WORKSPACE, "svk_convflow: workspace too small");
  const size_t smem_dds = 2 * (size_t)F * CF_T * sizeof(float), smem_proj = ((size_t)F + 3 * num_bins - 1) * CF_T * sizeof(float);
  if (smem_dds > 200 * 1024 || smem_proj > 200 * 1024) return cf_fail(SVK_ERR_INVALID, "svk_convflow: filter_channels too large (<= 400)");
  cudaStream_t s = (cudaStream_t)stream;
  const int half = C / 2, nblk = (T + CF_T - 1) / CF_T;
  float* ha = (float*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  float* hb = ha + (size_t)B * F * T;
  float* partial = hb + (size_t)B * F * T;
  const dim3 grid(nblk, B);
  cudaError_t e = cudaSuccess;
  if (smem_dds > 48 * 1024) e = cudaFuncSetAttribute(cf_dds_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dds);
  if (e == cudaSuccess && smem_proj > 48 * 1024) e = cudaFuncSetAttribute(cf_proj_spline_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_proj);
  if (e != cudaSuccess) return cf_fail(SVK_ERR_CUDA, std::string("svk_convflow: ") + cudaGetErrorString(e));
  cf_pre_kernel<<<grid, CF_T, 0, s>>>(x, C, half, T, w->pre_w, w->pre_b, F, ha);
  int dil = 1;
  for (int i = 0; i < n_layers; ++i) {
    cf_dds_layer_kernel<<<grid, CF_T, smem_dds, s>>>(ha, mask, F, T, kernel_size, dil, w->sep_w + (size_t)i * F * kernel_size, w->sep_b + (size_t)i * F,
                                                    w->norm1_g + (size_t)i * F, w->norm1_b + (size_t)i * F, w->pw_w + (size_t)i * F * F,
                                                    w->pw_b + (size_t)i * F, w->norm2_g + (size_t)i * F, w->norm2_b + (size_t)i * F, 1e-5f, hb);
    float* tmp = ha;
    ha = hb, hb = tmp;
    dil *= kernel_size;  // dilation = kernel_size ** i (modules.py:86)
  }
  cf_proj_spline_kernel<<<grid, CF_T, smem_proj, s>>>(x, ha, mask, C, half, F, T, num_bins, w->proj_w, w->proj_b, (float)sqrt((double)F),
                                                     tail_bound, reverse ? 1 : 0, y, logdet ? partial : nullptr, bins);
  if (logdet) cf_logdet_kernel<<<B, 32, 0, s>>>(partial, nblk, logdet);
  e = cudaGetLastError();
  if (e != cudaSuccess) return cf_fail(SVK_ERR_CUDA, std::string("svk_convflow: ") + cudaGetErrorString(e));
  return SVK_OK;
}
