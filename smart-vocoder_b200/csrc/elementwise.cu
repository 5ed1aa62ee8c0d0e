// Small memory-bound kernels of the path: sequence mask, latent sampling, Flip, weight-norm fold,
// and the rational-quadratic spline operator.  All are one pass over their operands, coalesced
// along time, grid sized in multiples of the SM count where the problem is large enough.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "spline.cuh"
#include "svk_kernels.cuh"

namespace svk {
namespace {

constexpr int EW_THREADS = 256;

inline int ew_blocks(int64_t n) {
  int64_t b = (n + EW_THREADS - 1) / EW_THREADS;
  const int64_t cap = 148 * 16;  // grid-stride beyond 16 CTAs per SM
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// commons.sequence_mask (commons.py:121-125) + .to(x.dtype) (models.py:40)
__global__ void sequence_mask_kernel(const int64_t* __restrict__ lengths, int B, int T,
                                     float* __restrict__ mask) {
  const int64_t n = (int64_t)B * T;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / T);
    const int64_t t = i - (int64_t)b * T;
    mask[i] = t < lengths[b] ? 1.0f : 0.0f;
  }
}

// Lengths seen by a time window [a, a + w) of the batch: clamp(length - a, 0, w) (integer; svk_infer_window).
__global__ void window_lengths_kernel(const int64_t* __restrict__ lengths, int B, int64_t a, int64_t w,
                                      int64_t* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) {
    const int64_t v = lengths[b] - a;
    out[b] = v < 0 ? 0 : (v > w ? w : v);
  }
}

// modules.Flip (modules.py:272): torch.flip(x, [1])
__global__ void flip_kernel(const float* __restrict__ x, int B, int C, int T, float* __restrict__ y) {
  const int64_t n = (int64_t)B * C * T;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = i % T;
    const int64_t bc = i / T;
    const int64_t c = bc % C, b = bc / C;
    y[i] = x[(b * C + (C - 1 - c)) * T + t];
  }
}

// models.py:336: z_p = m_p + randn * exp(logs_p) * noise_scale  (randn injected as eps).
// Evaluated in the reference's order: ((eps * exp(logs)) * noise_scale) + m.
__global__ void sample_kernel(const float* __restrict__ m, const float* __restrict__ logs,
                              const float* __restrict__ eps, float noise_scale, float* __restrict__ z_p,
                              float* __restrict__ z, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float v = __fadd_rn(m[i], __fmul_rn(__fmul_rn(eps[i], expf(logs[i])), noise_scale));
    if (z_p) z_p[i] = v;
    if (z) z[i] = v;
  }
}

// PosteriorEncoder (models.py:109): z = (m + randn * exp(logs)) * x_mask, in the reference's evaluation order.
__global__ void posterior_sample_kernel(const float* __restrict__ m, const float* __restrict__ logs,
                                        const float* __restrict__ eps, const float* __restrict__ mask, int C, int T,
                                        float* __restrict__ z, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = i % T, b = i / ((int64_t)C * T);
    z[i] = __fmul_rn(__fadd_rn(m[i], __fmul_rn(eps[i], expf(logs[i]))), mask[b * T + t]);
  }
}

// weight_norm fold: one CTA per dim-0 slice.
__global__ void weight_norm_kernel(const float* __restrict__ v, const float* __restrict__ g,
                                   int64_t inner, float* __restrict__ w) {
  __shared__ float red[32];
  const int64_t i = blockIdx.x;
  const float* vr = v + i * inner;
  float s = 0.f;
  for (int64_t j = threadIdx.x; j < inner; j += blockDim.x) s = fmaf(vr[j], vr[j], s);
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
    if (threadIdx.x == 0) red[0] = t;
  }
  __syncthreads();
  const float scale = g[i] / sqrtf(red[0]);
  for (int64_t j = threadIdx.x; j < inner; j += blockDim.x) w[i * inner + j] = vr[j] * scale;
}

// transforms.py:12-193, one thread per element; parameters are read once (29 floats / element); spline.cuh holds the math.
__global__ void rq_spline_kernel(const float* __restrict__ x, const float* __restrict__ uw,
                                 const float* __restrict__ uh, const float* __restrict__ ud, int64_t n,
                                 int nb, int inverse, float tail_bound, float min_bw, float min_bh,
                                 float min_d, float* __restrict__ y, float* __restrict__ lad,
                                 int32_t* __restrict__ bins) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n;
       e += (int64_t)gridDim.x * blockDim.x) {
    const float* pw = uw + e * nb;
    const float* ph = uh + e * nb;
    const float* pd = ud + e * (nb - 1);
    float yo, lo;
    int bin;
    rq_spline_element(
        x[e], [&](int i) { return pw[i]; }, [&](int i) { return ph[i]; }, [&](int i) { return pd[i]; }, nb, inverse,
        tail_bound, min_bw, min_bh, min_d, yo, lo, bin);
    y[e] = yo;
    lad[e] = lo;
    if (bins) bins[e] = bin;
  }
}

// PCM egress: float waveform -> int16 (max_wav_value 32768, configs/iitp_base.json:23; the inverse of the
// `audio / 32768.0` of inference.ipynb cell 4): round to nearest even, saturate.  8 samples per thread, one 16 B store.
__global__ void __launch_bounds__(256) pcm_to_int16_kernel(const float* __restrict__ x, int64_t n, float scale, int16_t* __restrict__ y) {
  const int64_t n8 = n >> 3;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(x)[2 * i], b = reinterpret_cast<const float4*>(x)[2 * i + 1];
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int lo = max(-32768, min(32767, __float2int_rn(v[2 * e] * scale)));
      const int hi = max(-32768, min(32767, __float2int_rn(v[2 * e + 1] * scale)));
      w[e] = ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16);
    }
    reinterpret_cast<uint4*>(y)[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
    const int64_t i = (n8 << 3) + threadIdx.x;
    y[i] = (int16_t)max(-32768, min(32767, __float2int_rn(x[i] * scale)));
  }
}

}  // namespace

cudaError_t launch_pcm_to_int16(const float* x, int64_t n, float scale, int16_t* y, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15)) return cudaErrorInvalidValue;
  pcm_to_int16_kernel<<<ew_blocks((n + 7) / 8), 256, 0, s>>>(x, n, scale, y);
  return cudaGetLastError();
}

cudaError_t launch_sequence_mask(const int64_t* lengths, int B, int T, float* mask, cudaStream_t s) {
  if ((int64_t)B * T == 0) return cudaSuccess;
  sequence_mask_kernel<<<ew_blocks((int64_t)B * T), EW_THREADS, 0, s>>>(lengths, B, T, mask);
  return cudaGetLastError();
}

cudaError_t launch_window_lengths(const int64_t* lengths, int B, int64_t a, int64_t w, int64_t* out, cudaStream_t s) {
  if (B == 0) return cudaSuccess;
  window_lengths_kernel<<<(B + EW_THREADS - 1) / EW_THREADS, EW_THREADS, 0, s>>>(lengths, B, a, w, out);
  return cudaGetLastError();
}

cudaError_t launch_flip(const float* x, int B, int C, int T, float* y, cudaStream_t s) {
  const int64_t n = (int64_t)B * C * T;
  if (n == 0) return cudaSuccess;
  flip_kernel<<<ew_blocks(n), EW_THREADS, 0, s>>>(x, B, C, T, y);
  return cudaGetLastError();
}

cudaError_t launch_sample(const float* m, const float* logs, const float* eps, float noise_scale,
                          float* z_p, float* z, int64_t n, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  sample_kernel<<<ew_blocks(n), EW_THREADS, 0, s>>>(m, logs, eps, noise_scale, z_p, z, n);
  return cudaGetLastError();
}

cudaError_t launch_posterior_sample(const float* m, const float* logs, const float* eps, const float* mask, int B, int C,
                                    int T, float* z, cudaStream_t s) {
  const int64_t n = (int64_t)B * C * T;
  if (n == 0) return cudaSuccess;
  posterior_sample_kernel<<<ew_blocks(n), EW_THREADS, 0, s>>>(m, logs, eps, mask, C, T, z, n);
  return cudaGetLastError();
}

cudaError_t launch_weight_norm(const float* v, const float* g, int64_t dim0, int64_t inner, float* w,
                               cudaStream_t s) {
  if (dim0 == 0) return cudaSuccess;
  weight_norm_kernel<<<(unsigned)dim0, 256, 0, s>>>(v, g, inner, w);
  return cudaGetLastError();
}

cudaError_t launch_rq_spline(const float* x, const float* uw, const float* uh, const float* ud,
                             int64_t n, int nb, int inverse, float tail_bound, float min_bw,
                             float min_bh, float min_d, float* y, float* lad, int32_t* bins,
                             cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  if (nb < 1 || nb > SPLINE_MAX_BINS) return cudaErrorInvalidValue;
  rq_spline_kernel<<<ew_blocks(n), 128, 0, s>>>(x, uw, uh, ud, n, nb, inverse, tail_bound, min_bw,
                                                min_bh, min_d, y, lad, bins);
  return cudaGetLastError();
}

}  // namespace svk
