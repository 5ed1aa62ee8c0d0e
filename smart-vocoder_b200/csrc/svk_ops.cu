// Stateless operator entry points of the C ABI (weights passed explicitly in reference layouts).
// They exist so each reference operator can be parity-tested in isolation; the hot path
// (svk_infer) uses weights packed once at load time instead.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/svk.h"
#include "svk_kernels.cuh"

using namespace svk;

namespace {

int op_fail(int code, const char* msg);

// w [Cout][Cin][K] -> wp [Cin][K][CoutPad], bias -> [CoutPad]
__global__ void pack_conv_kernel(const float* __restrict__ w, const float* __restrict__ bias, int Cout, int Cin,
                                 int K, int CoutPad, float* __restrict__ wp, float* __restrict__ bp) {
  const int64_t n = (int64_t)Cin * K * CoutPad;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(i % CoutPad);
    const int64_t cj = i / CoutPad;
    const int c = (int)(cj / K), j = (int)(cj % K);
    wp[i] = o < Cout ? w[((int64_t)o * Cin + c) * K + j] : 0.f;
  }
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < CoutPad; o += gridDim.x * blockDim.x)
    bp[o] = (bias && o < Cout) ? bias[o] : 0.f;
}

// ConvTranspose1d w [Cin][Cout][k] -> polyphase virtual conv [Cin][Kv][CoutPad], o' = co*s + r
__global__ void pack_convt_kernel(const float* __restrict__ w, const float* __restrict__ bias, int Cin, int Cout,
                                  int k, int s, int Kv, int CoutPad, float* __restrict__ wp, float* __restrict__ bp) {
  const int64_t n = (int64_t)Cin * Kv * CoutPad;
  const int CoutV = Cout * s;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(i % CoutPad);
    const int64_t cj = i / CoutPad;
    const int c = (int)(cj / Kv), jj = (int)(cj % Kv);
    float v = 0.f;
    if (o < CoutV) {
      const int co = o / s, r = o % s;
      const int j = r + (Kv - 1 - jj) * s;
      if (j < k) v = w[((int64_t)c * Cout + co) * k + j];
    }
    wp[i] = v;
  }
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < CoutPad; o += gridDim.x * blockDim.x)
    bp[o] = (bias && o < CoutV) ? bias[o / s] : 0.f;
}

}  // namespace

// svk_model.cu owns the thread-local error string; reuse it through a tiny setter.
extern "C" int svk__set_error(int code, const char* msg);
namespace {
int op_fail(int code, const char* msg) { return svk__set_error(code, msg); }
int cuda_fail(cudaError_t e, const char* who) {
  std::string m = std::string(who) + ": " + cudaGetErrorString(e);
  return svk__set_error(SVK_ERR_CUDA, m.c_str());
}
}  // namespace

extern "C" int svk_conv1d(const float* x, int B, int Cin, int L, const float* w, const float* bias, int Cout, int k,
                          int dilation, int padding, float pre_slope, float* y, void* stream) {
  if (!x || !w || !y || B <= 0 || Cin <= 0 || L <= 0 || Cout <= 0) return op_fail(SVK_ERR_INVALID, "svk_conv1d: bad argument");
  if (Cin % 8) return op_fail(SVK_ERR_INVALID, "svk_conv1d: Cin must be a multiple of 8");
  if (!conv_ffma_supports_k(k)) return op_fail(SVK_ERR_INVALID, "svk_conv1d: unsupported kernel size");
  if (dilation < 1 || dilation > 8) return op_fail(SVK_ERR_INVALID, "svk_conv1d: dilation must be in [1,8]");
  const int Lout = L + 2 * padding - dilation * (k - 1);
  if (Lout <= 0) return op_fail(SVK_ERR_INVALID, "svk_conv1d: empty output");
  cudaStream_t s = (cudaStream_t)stream;
  const int ot = conv_ffma_channel_tile(Cout);
  const int CoutPad = (Cout + ot - 1) / ot * ot;
  float* wp = nullptr;
  const size_t nw = (size_t)Cin * k * CoutPad;
  cudaError_t e = cudaMallocAsync((void**)&wp, (nw + CoutPad) * sizeof(float), s);
  if (e != cudaSuccess) return cuda_fail(e, "svk_conv1d: cudaMallocAsync");
  float* bp = wp + nw;
  pack_conv_kernel<<<256, 256, 0, s>>>(w, bias, Cout, Cin, k, CoutPad, wp, bp);
  ConvArgs a;
  memset(&a, 0, sizeof(a));
  a.x = x, a.x_C = Cin, a.x_stride = L, a.Lin = L, a.pre_slope = pre_slope;
  a.wp = wp, a.bias = bp, a.Cin = Cin, a.Cout = Cout, a.CoutPad = CoutPad, a.K = k, a.dil = dilation, a.pad = padding;
  a.Lout = Lout, a.y_stride = Lout, a.mode = MODE_STORE, a.split = 1 << 30, a.post_div = 1.0f, a.B = B;
  a.e[0].y = y, a.e[0].C = Cout, a.e[0].ch_sign = 1, a.e[1].ch_sign = 1;
  e = launch_conv_ffma(a, s);
  cudaFreeAsync(wp, s);
  if (e != cudaSuccess) return cuda_fail(e, "svk_conv1d");
  return SVK_OK;
}

extern "C" int svk_conv_transpose1d(const float* x, int B, int Cin, int L, const float* w, const float* bias, int Cout,
                                    int k, int stride, int padding, float pre_slope, float* y, void* stream) {
  if (!x || !w || !y || B <= 0 || Cin <= 0 || L <= 0 || Cout <= 0 || stride < 1 || k < 1)
    return op_fail(SVK_ERR_INVALID, "svk_conv_transpose1d: bad argument");
  if (Cin % 8) return op_fail(SVK_ERR_INVALID, "svk_conv_transpose1d: Cin must be a multiple of 8");
  const int Kv = (k + stride - 1) / stride;
  if (!conv_ffma_supports_k(Kv)) return op_fail(SVK_ERR_INVALID, "svk_conv_transpose1d: unsupported k/stride ratio");
  const int Lout = (L - 1) * stride - 2 * padding + k;
  if (Lout <= 0 || padding < 0) return op_fail(SVK_ERR_INVALID, "svk_conv_transpose1d: empty output");
  cudaStream_t s = (cudaStream_t)stream;
  const int CoutV = Cout * stride;
  const int ot = conv_ffma_channel_tile(CoutV);
  const int CoutPad = (CoutV + ot - 1) / ot * ot;
  float* wp = nullptr;
  const size_t nw = (size_t)Cin * Kv * CoutPad;
  cudaError_t e = cudaMallocAsync((void**)&wp, (nw + CoutPad) * sizeof(float), s);
  if (e != cudaSuccess) return cuda_fail(e, "svk_conv_transpose1d: cudaMallocAsync");
  float* bp = wp + nw;
  pack_convt_kernel<<<256, 256, 0, s>>>(w, bias, Cin, Cout, k, stride, Kv, CoutPad, wp, bp);
  ConvArgs a;
  memset(&a, 0, sizeof(a));
  a.x = x, a.x_C = Cin, a.x_stride = L, a.Lin = L, a.pre_slope = pre_slope;
  a.wp = wp, a.bias = bp, a.Cin = Cin, a.Cout = CoutV, a.CoutPad = CoutPad, a.K = Kv, a.dil = 1, a.pad = Kv - 1;
  a.Lout = (Lout - 1 + padding) / stride + 1;
  a.y_stride = Lout, a.mode = MODE_SHUFFLE, a.split = 1 << 30, a.post_div = 1.0f, a.B = B;
  a.shuf_s = stride, a.shuf_p = padding, a.shuf_Lout = Lout;
  a.e[0].y = y, a.e[0].C = Cout, a.e[0].ch_sign = 1, a.e[1].ch_sign = 1;
  e = launch_conv_ffma(a, s);
  cudaFreeAsync(wp, s);
  if (e != cudaSuccess) return cuda_fail(e, "svk_conv_transpose1d");
  return SVK_OK;
}

// ---- the same two operators on the tcgen05 engine (conv_tc.cu).  Weight images are built on the
// host per call (device -> host copy, fp16 hi/lo split, upload): these entry points exist for
// operator-level parity tests; the hot path packs once in svk_finalize_weights.
namespace {

int tc_tile(int Cout) {
  const int ntiles = (Cout + 127) / 128;
  const int per = (Cout + ntiles - 1) / ntiles;
  return (per + 15) / 16 * 16;
}

// w_ock: logical [CoutV][Cin][K] on the host; bias_v: [CoutV] on the host.
int run_tc(const float* x, int B, int Cin, int L, const std::vector<float>& w_ock, const std::vector<float>& bias_v,
           int CoutV, int K, int dil, int pad, float pre_slope, ConvArgs a, cudaStream_t s, const char* who) {
  const int N = tc_tile(CoutV);
  const int ntiles = (CoutV + N - 1) / N;
  const int CoutP = ntiles * N;
  std::vector<float> wpad((size_t)CoutP * Cin * K, 0.f), bpad(CoutP, 0.f);
  memcpy(wpad.data(), w_ock.data(), w_ock.size() * sizeof(float));
  memcpy(bpad.data(), bias_v.data(), bias_v.size() * sizeof(float));
  const float scale = conv_tc_weight_scale(wpad.data(), wpad.size());
  std::vector<uint16_t> img(conv_tc_packed_halves(Cin, CoutP, K, N));
  conv_tc_pack(wpad.data(), CoutP, Cin, K, N, scale, img.data());
  uint16_t* dimg = nullptr;
  float* dbias = nullptr;
  cudaError_t e = cudaMalloc((void**)&dimg, img.size() * 2);
  if (e == cudaSuccess) e = cudaMalloc((void**)&dbias, bpad.size() * 4);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dimg, img.data(), img.size() * 2, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dbias, bpad.data(), bpad.size() * 4, cudaMemcpyHostToDevice, s);
  // the kernel reads weights / biases before its griddep_wait() (programmatic dependent launch): they must be in place
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e == cudaSuccess) {
    ConvTcArgs ta;
    memset(&ta, 0, sizeof(ta));
    ta.c = a;
    ta.c.x = x, ta.c.x_C = Cin, ta.c.x_stride = L, ta.c.Lin = L, ta.c.pre_slope = pre_slope;
    ta.c.bias = dbias, ta.c.Cin = Cin, ta.c.Cout = CoutV, ta.c.CoutPad = CoutP, ta.c.K = K, ta.c.dil = dil, ta.c.pad = pad;
    ta.c.split = 1 << 30, ta.c.post_div = 1.0f, ta.c.B = B;
    ta.wtc = dimg, ta.unscale = 1.0f / scale, ta.N = N;
    e = launch_conv_tc(ta, s);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);  // the staging buffers are freed below
  cudaFree(dimg);
  cudaFree(dbias);
  if (e != cudaSuccess) return cuda_fail(e, who);
  return SVK_OK;
}

}  // namespace

extern "C" int svk_conv1d_tc(const float* x, int B, int Cin, int L, const float* w, const float* bias, int Cout, int k,
                             int dilation, int padding, float pre_slope, float* y, void* stream) {
  if (!x || !w || !y || B <= 0 || Cin <= 0 || L <= 0 || Cout <= 0 || k < 1) return op_fail(SVK_ERR_INVALID, "svk_conv1d_tc: bad argument");
  if (Cin % TC_KC) return op_fail(SVK_ERR_INVALID, "svk_conv1d_tc: Cin must be a multiple of 32");
  if (dilation < 1) return op_fail(SVK_ERR_INVALID, "svk_conv1d_tc: dilation must be >= 1");
  const int Lout = L + 2 * padding - dilation * (k - 1);
  if (Lout <= 0) return op_fail(SVK_ERR_INVALID, "svk_conv1d_tc: empty output");
  cudaStream_t s = (cudaStream_t)stream;
  std::vector<float> hw((size_t)Cout * Cin * k), hb(Cout, 0.f);
  cudaError_t e = cudaMemcpyAsync(hw.data(), w, hw.size() * 4, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess && bias) e = cudaMemcpyAsync(hb.data(), bias, hb.size() * 4, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return cuda_fail(e, "svk_conv1d_tc: weight staging");
  ConvArgs a;
  memset(&a, 0, sizeof(a));
  a.Lout = Lout, a.y_stride = Lout, a.mode = MODE_STORE;
  a.e[0].y = y, a.e[0].C = Cout, a.e[0].ch_sign = 1, a.e[1].ch_sign = 1;
  return run_tc(x, B, Cin, L, hw, hb, Cout, k, dilation, padding, pre_slope, a, s, "svk_conv1d_tc");
}

extern "C" int svk_conv_transpose1d_tc(const float* x, int B, int Cin, int L, const float* w, const float* bias,
                                       int Cout, int k, int stride, int padding, float pre_slope, float* y,
                                       void* stream) {
  if (!x || !w || !y || B <= 0 || Cin <= 0 || L <= 0 || Cout <= 0 || stride < 1 || k < 1)
    return op_fail(SVK_ERR_INVALID, "svk_conv_transpose1d_tc: bad argument");
  if (Cin % TC_KC) return op_fail(SVK_ERR_INVALID, "svk_conv_transpose1d_tc: Cin must be a multiple of 32");
  const int Kv = (k + stride - 1) / stride;
  const int Lout = (L - 1) * stride - 2 * padding + k;
  if (Lout <= 0 || padding < 0) return op_fail(SVK_ERR_INVALID, "svk_conv_transpose1d_tc: empty output");
  cudaStream_t s = (cudaStream_t)stream;
  std::vector<float> hw((size_t)Cin * Cout * k), hb(Cout, 0.f);
  cudaError_t e = cudaMemcpyAsync(hw.data(), w, hw.size() * 4, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess && bias) e = cudaMemcpyAsync(hb.data(), bias, hb.size() * 4, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return cuda_fail(e, "svk_conv_transpose1d_tc: weight staging");
  // polyphase virtual conv: o' = co*s + r, tap jj <-> original tap r + (Kv-1-jj)*s
  const int CoutV = Cout * stride;
  std::vector<float> wv((size_t)CoutV * Cin * Kv, 0.f), bv(CoutV, 0.f);
  for (int o = 0; o < CoutV; ++o) {
    const int co = o / stride, r = o % stride;
    bv[o] = hb[co];
    for (int c = 0; c < Cin; ++c)
      for (int jj = 0; jj < Kv; ++jj) {
        const int j = r + (Kv - 1 - jj) * stride;
        if (j < k) wv[((size_t)o * Cin + c) * Kv + jj] = hw[((size_t)c * Cout + co) * k + j];
      }
  }
  ConvArgs a;
  memset(&a, 0, sizeof(a));
  a.Lout = (Lout - 1 + padding) / stride + 1;
  a.y_stride = Lout, a.mode = MODE_SHUFFLE;
  a.shuf_s = stride, a.shuf_p = padding, a.shuf_Lout = Lout;
  a.e[0].y = y, a.e[0].C = Cout, a.e[0].ch_sign = 1, a.e[1].ch_sign = 1;
  return run_tc(x, B, Cin, L, wv, bv, CoutV, Kv, 1, Kv - 1, pre_slope, a, s, "svk_conv_transpose1d_tc");
}

extern "C" int svk_sequence_mask(const int64_t* lengths, int B, int T, float* mask, void* stream) {
  if (!lengths || !mask || B < 0 || T < 0) return op_fail(SVK_ERR_INVALID, "svk_sequence_mask: bad argument");
  cudaError_t e = launch_sequence_mask(lengths, B, T, mask, (cudaStream_t)stream);
  return e == cudaSuccess ? SVK_OK : cuda_fail(e, "svk_sequence_mask");
}

extern "C" int svk_pcm_to_int16(const float* pcm, int64_t n, float max_wav_value, int16_t* out, void* stream) {
  if (!pcm || !out || n < 0 || !(max_wav_value > 0.f)) return op_fail(SVK_ERR_INVALID, "svk_pcm_to_int16: bad argument");
  cudaError_t e = launch_pcm_to_int16(pcm, n, max_wav_value, out, (cudaStream_t)stream);
  return e == cudaSuccess ? SVK_OK : cuda_fail(e, "svk_pcm_to_int16 (buffers must be 16 B-aligned)");
}

extern "C" int svk_flip(const float* x, int B, int C, int T, float* y, void* stream) {
  if (!x || !y || x == y || B < 0 || C < 0 || T < 0) return op_fail(SVK_ERR_INVALID, "svk_flip: bad argument (in-place not supported)");
  cudaError_t e = launch_flip(x, B, C, T, y, (cudaStream_t)stream);
  return e == cudaSuccess ? SVK_OK : cuda_fail(e, "svk_flip");
}

extern "C" int svk_weight_norm(const float* v, const float* g, int64_t dim0, int64_t inner, float* w, void* stream) {
  if (!v || !g || !w || dim0 < 0 || inner <= 0) return op_fail(SVK_ERR_INVALID, "svk_weight_norm: bad argument");
  cudaError_t e = launch_weight_norm(v, g, dim0, inner, w, (cudaStream_t)stream);
  return e == cudaSuccess ? SVK_OK : cuda_fail(e, "svk_weight_norm");
}

extern "C" int svk_rq_spline(const float* x, const float* uw, const float* uh, const float* ud, int64_t n, int num_bins,
                             int inverse, float tail_bound, float min_bin_width, float min_bin_height,
                             float min_derivative, float* y, float* logabsdet, int32_t* bin, void* stream) {
  if (!x || !uw || !uh || !ud || !y || !logabsdet || n < 0) return op_fail(SVK_ERR_INVALID, "svk_rq_spline: bad argument");
  if (num_bins < 1 || num_bins > 32) return op_fail(SVK_ERR_INVALID, "svk_rq_spline: num_bins must be in [1,32]");
  // reference transforms.py:108-111
  if (min_bin_width * num_bins > 1.0f) return op_fail(SVK_ERR_INVALID, "Minimal bin width too large for the number of bins");
  if (min_bin_height * num_bins > 1.0f) return op_fail(SVK_ERR_INVALID, "Minimal bin height too large for the number of bins");
  cudaError_t e = launch_rq_spline(x, uw, uh, ud, n, num_bins, inverse, tail_bound, min_bin_width, min_bin_height,
                                   min_derivative, y, logabsdet, bin, (cudaStream_t)stream);
  return e == cudaSuccess ? SVK_OK : cuda_fail(e, "svk_rq_spline");
}
