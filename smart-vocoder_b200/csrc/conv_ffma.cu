// Generic stride-1 Conv1d (dilated, any odd/even tap count) as an implicit GEMM on the fp32 FFMA
// pipe, with every element-wise neighbour of the reference graph fused into prologue/epilogue:
//   prologue : leaky_relu on the input (modules.py:212,216; models.py:147,156), input mask and
//              max_len truncation (models.py:338), conv zero padding;
//   epilogue : bias, gated tanh*sigmoid (commons.py:100-107), residual add (modules.py:220),
//              WN's x=(x+res)*mask / out+=skip split (modules.py:169-175), ResBlock sum and /3
//              (models.py:150-155), tanh (models.py:158), ConvTranspose1d polyphase store
//              (models.py:149), channel-reversed store (folded Flip).
//
// Tiling: one CTA = 8 warps = (WO channel-warps) x (WT time-warps).  A warp owns 8 output channels
// x 32*TT time steps; a thread owns 8 channels x TT time steps, time strided by 32 so that every
// shared-memory read of activations is lane-consecutive (conflict free) and every weight read is a
// warp-wide broadcast LDS.128.  Input channels are consumed in chunks of 8, double buffered:
// weights by cp.async (LDGSTS), activations through registers (the prologue is applied there).
//
// Roofline: FFMA-bound.  Per (channel, tap) a thread issues 2 LDS.128 + TT LDS.32 for 8*TT FFMA.
#include <cuda_runtime.h>
#include <stdint.h>

#include "svk_kernels.cuh"

namespace svk {

namespace {

constexpr int CC = 8;       // input channels per smem chunk (== warps per CTA: one staging warp per row)
constexpr int MAXDIL = 8;   // largest dilation an instance can stage (reference uses 1,3,5)

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

__device__ __forceinline__ float sigmoidf_(float v) { return 1.0f / (1.0f + expf(-v)); }

template <int K, int WO, int WT, int TT>
__global__ void __launch_bounds__(WO* WT * 32, 2) conv_ffma_kernel(const ConvArgs a) {
  static_assert(WO * WT == CC, "one staging warp per input-channel row of a chunk");
  constexpr int NT = WO * WT * 32;
  constexpr int OT = WO * 8;
  constexpr int TILE_T = WT * 32 * TT;
  constexpr int NX = (TILE_T + (K - 1) * MAXDIL + 31) / 32;
  extern __shared__ __align__(16) float smem[];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wo = warp % WO, wt = warp / WO;
  const int XW = TILE_T + (K - 1) * a.dil;
  const int XWp = (XW + 3) & ~3;
  float* xs = smem;                 // [2][CC][XWp]
  float* ws = smem + 2 * CC * XWp;  // [2][CC*K*OT]
  const int t0 = blockIdx.x * TILE_T, o0 = blockIdx.y * OT, b = blockIdx.z;
  const int nchunks = a.Cin / CC;

  float acc[8][TT];
#pragma unroll
  for (int e = 0; e < 8; ++e)
#pragma unroll
    for (int i = 0; i < TT; ++i) acc[e][i] = 0.f;

  float xr[NX];
  const float* mrow = a.in_mask ? a.in_mask + (size_t)b * a.mask_stride : nullptr;
  const float slope = a.pre_slope;

  auto load_x = [&](int chunk) {
    const int gch = a.x_ch_off + chunk * CC + warp;
    const float* xrow = a.x + ((size_t)b * a.x_C + gch) * a.x_stride;
#pragma unroll
    for (int n = 0; n < NX; ++n) {
      const int tt = lane + 32 * n;
      const int gpos = t0 - a.pad + tt;
      float v = 0.f;
      if (tt < XW && gpos >= 0 && gpos < a.Lin && gch < a.x_C) {  // gch >= x_C: zero-padded input channels (Cin % 8 != 0)
        v = __ldg(xrow + gpos);
        v = v > 0.f ? v : v * slope;
        if (mrow) v *= __ldg(mrow + gpos);
      }
      xr[n] = v;
    }
  };
  auto store_x = [&](int buf) {
    float* row = xs + (buf * CC + warp) * XWp;
#pragma unroll
    for (int n = 0; n < NX; ++n) {
      const int tt = lane + 32 * n;
      if (tt < XWp) row[tt] = xr[n];
    }
  };
  auto load_w = [&](int chunk, int buf) {
    constexpr int V4 = OT / 4;
    constexpr int TOTAL = CC * K * V4;
    const float* src = a.wp + (size_t)chunk * CC * K * a.CoutPad + o0;
    float* dst = ws + buf * (CC * K * OT);
    for (int idx = tid; idx < TOTAL; idx += NT) {
      const int row = idx / V4, v4 = idx % V4;
      cp_async16(dst + row * OT + 4 * v4, src + (size_t)row * a.CoutPad + 4 * v4);
    }
  };

  load_x(0);
  load_w(0, 0);
  store_x(0);
  cp_async_commit_wait_all();
  __syncthreads();

  for (int ch = 0; ch < nchunks; ++ch) {
    const int cur = ch & 1;
    const bool more = ch + 1 < nchunks;
    if (more) {
      load_x(ch + 1);
      load_w(ch + 1, cur ^ 1);
    }
    const float* xb = xs + cur * CC * XWp + wt * (32 * TT) + lane;
    const float* wb = ws + cur * (CC * K * OT) + wo * 8;
#pragma unroll 1
    for (int c = 0; c < CC; ++c) {
#pragma unroll
      for (int j = 0; j < K; ++j) {
        const float4 w0 = *reinterpret_cast<const float4*>(wb + (c * K + j) * OT);
        const float4 w1 = *reinterpret_cast<const float4*>(wb + (c * K + j) * OT + 4);
        const float* xp = xb + c * XWp + j * a.dil;
        float xv[TT];
#pragma unroll
        for (int i = 0; i < TT; ++i) xv[i] = xp[32 * i];
#pragma unroll
        for (int i = 0; i < TT; ++i) {
          acc[0][i] = fmaf(w0.x, xv[i], acc[0][i]);
          acc[1][i] = fmaf(w0.y, xv[i], acc[1][i]);
          acc[2][i] = fmaf(w0.z, xv[i], acc[2][i]);
          acc[3][i] = fmaf(w0.w, xv[i], acc[3][i]);
          acc[4][i] = fmaf(w1.x, xv[i], acc[4][i]);
          acc[5][i] = fmaf(w1.y, xv[i], acc[5][i]);
          acc[6][i] = fmaf(w1.z, xv[i], acc[6][i]);
          acc[7][i] = fmaf(w1.w, xv[i], acc[7][i]);
        }
      }
    }
    if (more) {
      store_x(cur ^ 1);
      cp_async_commit_wait_all();
    }
    __syncthreads();
  }

  // ------------------------------------------------------------------ epilogue
  const int ob = o0 + wo * 8;  // first (virtual) output channel of this thread
  const int tb = t0 + wt * (32 * TT) + lane;
  if (ob >= a.Cout) return;
  const float* omask = a.out_mask ? a.out_mask + (size_t)b * a.mask_stride : nullptr;

  if (a.mode == MODE_GATE) {
    // packed pairs: e<4 -> tanh half of channel 4*grp+e, e>=4 -> sigmoid half of the same channel
    const int cbase = (ob >> 3) * 4;
    const EpiDesc& d = a.e[0];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float bt = __ldg(a.bias + ob + e), bs = __ldg(a.bias + ob + e + 4);
      float* yrow = d.y + ((size_t)b * d.C + d.ch_off + cbase + e) * a.y_stride;
#pragma unroll
      for (int i = 0; i < TT; ++i) {
        const int t = tb + 32 * i;
        if (t < a.Lout) yrow[t] = tanhf(acc[e][i] + bt) * sigmoidf_(acc[e + 4][i] + bs);
      }
    }
    return;
  }

  if (a.mode == MODE_SHUFFLE) {
    const EpiDesc& d = a.e[0];
    const int s = a.shuf_s;
    if ((s & 7) == 0 && (a.shuf_p & 3) == 0) {
      // 8 consecutive virtual channels = 8 consecutive output samples of one channel
      const int co = ob / s, r0 = ob % s;
      float* yrow = d.y + ((size_t)b * d.C + d.ch_off + co) * a.y_stride;
      float bv[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) bv[e] = __ldg(a.bias + ob + e);
#pragma unroll
      for (int i = 0; i < TT; ++i) {
        const int q = tb + 32 * i;
        if (q >= a.Lout) continue;
        const int t = s * q + r0 - a.shuf_p;
        if (t >= 0 && t + 3 < a.shuf_Lout)
          *reinterpret_cast<float4*>(yrow + t) =
              make_float4(acc[0][i] + bv[0], acc[1][i] + bv[1], acc[2][i] + bv[2], acc[3][i] + bv[3]);
        if (t + 4 >= 0 && t + 7 < a.shuf_Lout)
          *reinterpret_cast<float4*>(yrow + t + 4) =
              make_float4(acc[4][i] + bv[4], acc[5][i] + bv[5], acc[6][i] + bv[6], acc[7][i] + bv[7]);
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int o = ob + e;
        if (o >= a.Cout) break;
        const int co = o / s, r = o % s;
        const float bv = __ldg(a.bias + o);
        float* yrow = d.y + ((size_t)b * d.C + d.ch_off + co) * a.y_stride;
#pragma unroll
        for (int i = 0; i < TT; ++i) {
          const int q = tb + 32 * i;
          const int t = s * q + r - a.shuf_p;
          if (q < a.Lout && t >= 0 && t < a.shuf_Lout) yrow[t] = acc[e][i] + bv;
        }
      }
    }
    return;
  }

  // MODE_STORE (split is a multiple of 8, so the side is uniform per thread)
  const int side = ob < a.split ? 0 : 1;
  const EpiDesc d = a.e[side];
  const int rel0 = ob - (side ? a.split : 0);
  const bool use_mask = d.use_mask && omask;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int o = ob + e;
    if (o >= a.Cout) break;
    const float bv = __ldg(a.bias + o);
    const size_t rowoff = ((size_t)b * d.C + d.ch_off + d.ch_sign * (rel0 + e)) * a.y_stride;
    const float* rrow = d.res ? d.res + rowoff : nullptr;
    const float* arow = d.acc_in ? d.acc_in + rowoff : nullptr;
    float* yrow = d.y + rowoff;
#pragma unroll
    for (int i = 0; i < TT; ++i) {
      const int t = tb + 32 * i;
      if (t >= a.Lout) continue;
      float v = acc[e][i] + bv;
      if (rrow) v += rrow[t];
      if (arow) v += arow[t];
      if (a.post_div != 1.0f) v = v / a.post_div;
      if (use_mask) v *= omask[t];
      if (a.act_tanh) {
        v = tanhf(v);
        if (a.range_flag && !(fabsf(v) <= 1.0f)) *a.range_flag = 1;
      }
      yrow[t] = v;
    }
  }
}

template <int K, int WO, int WT, int TT>
cudaError_t launch_instance(const ConvArgs& a, cudaStream_t stream) {
  constexpr int OT = WO * 8;
  constexpr int TILE_T = WT * 32 * TT;
  if (a.dil > MAXDIL || a.dil < 1) return cudaErrorInvalidValue;
  if (a.Cin % CC != 0 || a.CoutPad % OT != 0) return cudaErrorInvalidValue;
  const int XWp = (TILE_T + (K - 1) * a.dil + 3) & ~3;
  const size_t smem = sizeof(float) * 2 * (size_t)(CC * XWp + CC * K * OT);
  static size_t configured[64] = {0};  // per device: opt-in dynamic smem already granted
  auto kern = conv_ffma_kernel<K, WO, WT, TT>;
  int dev = 0;
  cudaGetDevice(&dev);
  if (smem > configured[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured[dev & 63] = smem;
  }
  dim3 grid((a.Lout + TILE_T - 1) / TILE_T, (a.Cout + OT - 1) / OT, a.B);
  if (grid.x == 0 || grid.y == 0 || grid.z == 0) return cudaSuccess;
  kern<<<grid, WO * WT * 32, smem, stream>>>(a);
  return cudaGetLastError();
}

template <int K>
cudaError_t launch_k(const ConvArgs& a, cudaStream_t stream) {
  const int ot = conv_ffma_channel_tile(a.Cout);
  if (ot == 64) return launch_instance<K, 8, 1, 8>(a, stream);
  if (ot == 32) return launch_instance<K, 4, 2, 8>(a, stream);
  return launch_instance<K, 1, 8, 4>(a, stream);
}

}  // namespace

int conv_ffma_channel_tile(int cout) { return cout >= 64 ? 64 : (cout >= 32 ? 32 : 8); }

bool conv_ffma_supports_k(int k) {
  return k == 1 || k == 2 || k == 3 || k == 5 || k == 7 || k == 9 || k == 11;
}

// ---- Cout == 1 (dec.conv_post, models.py:156-158): a pure HBM stream -- every input element is
// read once, 2*Cin*K flops per output.  One CTA = 1024 outputs of one utterance, 4 per thread;
// input channels are staged 8 at a time through shared memory (prologue applied while staging),
// read back as conflict-free LDS.128.
constexpr int C1_TILE = 1024, C1_THREADS = 256, C1_CC = 8, C1_MAXK = 11;
constexpr int C1_PITCH = C1_TILE + 16;  // >= C1_TILE + C1_MAXK - 1, multiple of 4

static __global__ void __launch_bounds__(C1_THREADS) conv_cout1_kernel(const ConvArgs a) {
  __shared__ __align__(16) float xs[C1_CC][C1_PITCH];
  __shared__ float ws[C1_CC * C1_MAXK];
  const int tid = threadIdx.x, b = blockIdx.y;
  const int t0 = blockIdx.x * C1_TILE;
  const int K = a.K, span = C1_TILE + K - 1;
  const float slope = a.pre_slope;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c0 = 0; c0 < a.Cin; c0 += C1_CC) {
    __syncthreads();
    const float* xb = a.x + ((size_t)b * a.x_C + a.x_ch_off + c0) * a.x_stride;
    for (int idx = tid; idx < C1_CC * span; idx += C1_THREADS) {
      const int c = idx / span, r = idx - c * span;
      const int t = t0 - a.pad + r;
      float v = (t >= 0 && t < a.Lin) ? __ldg(xb + (size_t)c * a.x_stride + t) : 0.f;
      xs[c][r] = v > 0.f ? v : v * slope;
    }
    if (tid < C1_CC * K) ws[tid] = a.wp[(size_t)(c0 * K + tid) * a.CoutPad];
    __syncthreads();
#pragma unroll
    for (int c = 0; c < C1_CC; ++c) {
      float v[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 f = *reinterpret_cast<const float4*>(&xs[c][4 * tid + 4 * q]);
        v[4 * q] = f.x, v[4 * q + 1] = f.y, v[4 * q + 2] = f.z, v[4 * q + 3] = f.w;
      }
#pragma unroll
      for (int j = 0; j < C1_MAXK; ++j) {
        if (j < K) {
          const float w = ws[c * K + j];
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i] = fmaf(w, v[i + j], acc[i]);
        }
      }
    }
  }
  const float bias = a.bias[0];
  float* y = a.e[0].y + ((size_t)b * a.e[0].C + a.e[0].ch_off) * a.y_stride;
  const float* om = a.out_mask ? a.out_mask + (size_t)b * a.mask_stride : nullptr;
  const int t = t0 + 4 * tid;
  float o[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float v = acc[i] + bias;
    if (a.post_div != 1.0f) v = v / a.post_div;
    if (om && a.e[0].use_mask && t + i < a.Lout) v *= om[t + i];
    o[i] = a.act_tanh ? tanhf(v) : v;
  }
  if (a.act_tanh && a.range_flag) {
    bool bad = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) bad = bad || (t + i < a.Lout && !(fabsf(o[i]) <= 1.0f));
    if (bad) *a.range_flag = 1;
  }
  if (t + 3 < a.Lout && (a.y_stride & 3) == 0) {
    *reinterpret_cast<float4*>(y + t) = make_float4(o[0], o[1], o[2], o[3]);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (t + i < a.Lout) y[t + i] = o[i];
  }
}

// Streaming variant for the shape dec.conv_post actually has (k = 7, pad = 3, long rows): no shared-memory staging
// and no block-wide barriers -- a thread owns 4 consecutive outputs and reads, per input channel, the three aligned
// float4 covering t-4 .. t+7 (neighbouring threads share them through L1; HBM sees every element once), so the
// kernel is a pure stream with enough independent loads in flight to approach the HBM rate.
constexpr int CS_THREADS = 256;
template <int K, int PAD>
static __global__ void __launch_bounds__(CS_THREADS) conv_cout1_stream_kernel(const ConvArgs a) {
  static_assert(PAD <= 4 && K - PAD <= 5, "taps must lie inside t-4 .. t+7");
  __shared__ float ws[64 * K];
  const int Cin = a.Cin;
  for (int i = threadIdx.x; i < Cin * K; i += CS_THREADS) ws[i] = a.wp[(size_t)i * a.CoutPad];
  __syncthreads();
  const int b = blockIdx.y;
  const int t = (blockIdx.x * CS_THREADS + threadIdx.x) * 4;
  if (t >= a.Lout) return;
  const float slope = a.pre_slope;
  const float* xb = a.x + ((size_t)b * a.x_C + a.x_ch_off) * a.x_stride;
  const bool interior = t >= 4 && t + 8 <= a.Lin;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (int c = 0; c < Cin; ++c) {
    const float* xr = xb + (size_t)c * a.x_stride;
    float v[12];
    if (interior) {
      const float4 f0 = __ldg(reinterpret_cast<const float4*>(xr + t - 4));
      const float4 f1 = __ldg(reinterpret_cast<const float4*>(xr + t));
      const float4 f2 = __ldg(reinterpret_cast<const float4*>(xr + t + 4));
      v[0] = f0.x, v[1] = f0.y, v[2] = f0.z, v[3] = f0.w, v[4] = f1.x, v[5] = f1.y, v[6] = f1.z, v[7] = f1.w;
      v[8] = f2.x, v[9] = f2.y, v[10] = f2.z, v[11] = f2.w;
    } else {
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        const int tt = t - 4 + i;
        v[i] = (tt >= 0 && tt < a.Lin) ? __ldg(xr + tt) : 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) v[i] = v[i] > 0.f ? v[i] : v[i] * slope;
    const float* w = ws + c * K;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const float wj = w[j];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(wj, v[4 - PAD + i + j], acc[i]);  // x[t + i + j - pad]
    }
  }
  const float bias = a.bias[0];
  float* y = a.e[0].y + ((size_t)b * a.e[0].C + a.e[0].ch_off) * a.y_stride;
  const float* om = a.out_mask ? a.out_mask + (size_t)b * a.mask_stride : nullptr;
  float o[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float q = acc[i] + bias;
    if (a.post_div != 1.0f) q = q / a.post_div;
    if (om && a.e[0].use_mask && t + i < a.Lout) q *= om[t + i];
    o[i] = a.act_tanh ? tanhf(q) : q;
  }
  if (a.act_tanh && a.range_flag) {
    bool bad = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) bad = bad || (t + i < a.Lout && !(fabsf(o[i]) <= 1.0f));
    if (bad) *a.range_flag = 1;
  }
  if (t + 3 < a.Lout && (a.y_stride & 3) == 0) {
    *reinterpret_cast<float4*>(y + t) = make_float4(o[0], o[1], o[2], o[3]);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (t + i < a.Lout) y[t + i] = o[i];
  }
}

static bool conv_cout1_applicable(const ConvArgs& a) {
  return a.Cout == 1 && a.mode == MODE_STORE && a.dil == 1 && a.K <= C1_MAXK && a.Cin % C1_CC == 0 && !a.in_mask &&
         !a.e[0].res && !a.e[0].acc_in && a.e[0].ch_sign == 1 && a.split > 0;
}

cudaError_t launch_conv_ffma(const ConvArgs& a, cudaStream_t stream) {
  if (a.mode == MODE_SHUFFLE && a.shuf_rmajor) return cudaErrorInvalidValue;  // a tcgen05-engine channel order
  if (conv_cout1_applicable(a)) {
    if (a.B <= 0 || a.Lout <= 0) return cudaSuccess;
    // taps must lie inside the three float4 a thread reads (t-4 .. t+7), rows must keep float4 loads aligned
    const bool stream_ok = a.K == 7 && a.pad == 3 && a.Cin <= 64 && (a.x_stride & 3) == 0 && a.Lout == a.Lin &&
                           (reinterpret_cast<uintptr_t>(a.x) & 15) == 0;
    if (stream_ok) {  // dec.conv_post (models.py:135: Conv1d(ch, 1, 7, 1, padding=3))
      conv_cout1_stream_kernel<7, 3><<<dim3((a.Lout + 4 * CS_THREADS - 1) / (4 * CS_THREADS), a.B), CS_THREADS, 0, stream>>>(a);
      return cudaGetLastError();
    }
    conv_cout1_kernel<<<dim3((a.Lout + C1_TILE - 1) / C1_TILE, a.B), C1_THREADS, 0, stream>>>(a);
    return cudaGetLastError();
  }
  switch (a.K) {
    case 1: return launch_k<1>(a, stream);
    case 2: return launch_k<2>(a, stream);
    case 3: return launch_k<3>(a, stream);
    case 5: return launch_k<5>(a, stream);
    case 7: return launch_k<7>(a, stream);
    case 9: return launch_k<9>(a, stream);
    case 11: return launch_k<11>(a, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace svk
