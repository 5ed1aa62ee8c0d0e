// Mel front-end of the path's callers: waveform -> linear spectrogram -> log-mel,
// the step immediately before SynthesizerTrn.infer in inference.ipynb:100-111 and train.py:266-272.
// Reference: mel_processing.py:51-69 (spectrogram_torch), :72-81 (spec_to_mel_torch), :84-112
// (mel_spectrogram_torch), :16-22 (log(clamp(x, 1e-5))).
//
// One kernel, one pass: a CTA owns FT consecutive frames of one utterance.  Frames are transformed two at a
// time as the real and imaginary part of one complex FFT (Stockham radix-2 autosort in shared memory, fp32,
// twiddles from a double-precision table), magnitudes sqrt(re^2 + im^2 + 1e-6) land in a shared [bin][frame]
// tile, and the mel projection + log run from that tile, so on the fused entry point the 513-bin spectrogram
// never reaches HBM: traffic is 4 B per sample in (re-read from L2 for the overlapping frames) and
// 4 * n_mels / hop B per sample out.  HBM/latency-bound, < 0.1 % of an infer step; no tensor cores (the DFT as a
// GEMM would need the 3-pass fp16 split and 20x the FLOPs of the FFT).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/svk.h"

extern "C" int svk__set_error(int code, const char* msg);

namespace {

constexpr int FE_THREADS = 256;
constexpr int FE_FT = 16;          // frames per CTA
constexpr int FE_LD = FE_FT + 1;   // padded row of the magnitude tile (conflict-free column writes)

int fe_fail(int code, const std::string& m) { return svk__set_error(code, m.c_str()); }

struct FeArgs {
  const float* y;       // [B, n] waveform, or null when `spec_in` is given
  const float* spec_in; // [B, n_bins, T]
  int64_t n, T;
  int n_fft, log2n, hop, pad, n_bins, n_mels, tiles_t;
  const float* window;   // [n_fft]
  const float2* twiddle; // [n_fft / 2]  exp(-2 pi i k / n_fft)
  const float* basis;    // [n_mels, n_bins]
  const int2* range;     // [n_mels] first / one-past-last non-zero bin
  float* spec_out;       // [B, n_bins, T] or null
  float* mel_out;        // [B, n_mels, T] or null
};

__global__ void __launch_bounds__(FE_THREADS) mel_frontend_kernel(const FeArgs a) {
  extern __shared__ __align__(16) unsigned char fe_smem[];
  float2* bufA = reinterpret_cast<float2*>(fe_smem);
  float2* bufB = bufA + a.n_fft;
  float* mag = reinterpret_cast<float*>(bufB + a.n_fft);  // [n_bins][FE_LD]

  const int tid = threadIdx.x;
  const int b = blockIdx.x / a.tiles_t;
  const int64_t f0 = (int64_t)(blockIdx.x % a.tiles_t) * FE_FT;
  const int nf = (int)min((int64_t)FE_FT, a.T - f0);
  const int N = a.n_fft, half = N >> 1;

  if (a.spec_in) {
    const float* sp = a.spec_in + (int64_t)b * a.n_bins * a.T + f0;
    for (int i = tid; i < a.n_bins * FE_FT; i += FE_THREADS) {
      const int k = i / FE_FT, f = i % FE_FT;
      mag[k * FE_LD + f] = f < nf ? sp[(int64_t)k * a.T + f] : 0.f;
    }
  } else {
    const float* yb = a.y + (int64_t)b * a.n;
    for (int p = 0; p < nf; p += 2) {
      // z = w * (frame_p + i frame_{p+1}); reflect padding of mel_processing.py:62 folded into the index
      const int64_t s0 = (f0 + p) * a.hop - a.pad;
      const bool two = p + 1 < nf;
      for (int i = tid; i < N; i += FE_THREADS) {
        int64_t q0 = s0 + i, q1 = q0 + a.hop;
        q0 = q0 < 0 ? -q0 : (q0 >= a.n ? 2 * (a.n - 1) - q0 : q0);
        q1 = q1 < 0 ? -q1 : (q1 >= a.n ? 2 * (a.n - 1) - q1 : q1);
        const float w = a.window[i];
        bufA[i] = make_float2(w * __ldg(yb + q0), two ? w * __ldg(yb + q1) : 0.f);
      }
      __syncthreads();
      // Stockham radix-2 DIF: stage (n, s): a = x[j], b = x[j + N/2], j = p*s + q;
      // y[q + 2sp] = a + b, y[q + 2sp + s] = (a - b) * W_n^p, W_n^p = table[p * s]
      float2* x = bufA;
      float2* y = bufB;
      for (int st = 0, s = 1; st < a.log2n; ++st, s <<= 1) {
        for (int j = tid; j < half; j += FE_THREADS) {
          const int q = j & (s - 1);
          const int ps = j - q;  // p * s
          const float2 u = x[j], v = x[j + half];
          const float2 w = a.twiddle[ps];
          const float dr = u.x - v.x, di = u.y - v.y;
          y[q + 2 * ps] = make_float2(u.x + v.x, u.y + v.y);
          y[q + 2 * ps + s] = make_float2(dr * w.x - di * w.y, dr * w.y + di * w.x);
        }
        __syncthreads();
        float2* t = x;
        x = y, y = t;
      }
      // split the two real spectra: F1 = (Z[k] + conj Z[N-k]) / 2, F2 = (Z[k] - conj Z[N-k]) / 2i
      for (int k = tid; k <= half; k += FE_THREADS) {
        const float2 zk = x[k], zn = x[(N - k) & (N - 1)];
        const float r1 = 0.5f * (zk.x + zn.x), i1 = 0.5f * (zk.y - zn.y);
        const float r2 = 0.5f * (zk.y + zn.y), i2 = 0.5f * (zn.x - zk.x);
        mag[k * FE_LD + p] = sqrtf(r1 * r1 + i1 * i1 + 1e-6f);       // mel_processing.py:68
        mag[k * FE_LD + p + 1] = two ? sqrtf(r2 * r2 + i2 * i2 + 1e-6f) : 0.f;
      }
      __syncthreads();
    }
    for (int i = tid; i < a.n_bins; i += FE_THREADS)
      for (int f = (nf + 1) & ~1; f < FE_FT; ++f) mag[i * FE_LD + f] = 0.f;
  }
  __syncthreads();

  if (a.spec_out) {
    float* so = a.spec_out + (int64_t)b * a.n_bins * a.T + f0;
    for (int i = tid; i < a.n_bins * FE_FT; i += FE_THREADS) {
      const int k = i / FE_FT, f = i % FE_FT;
      if (f < nf) so[(int64_t)k * a.T + f] = mag[k * FE_LD + f];
    }
  }
  if (a.mel_out) {
    float* mo = a.mel_out + (int64_t)b * a.n_mels * a.T + f0;
    for (int i = tid; i < a.n_mels * FE_FT; i += FE_THREADS) {
      const int m = i / FE_FT, f = i % FE_FT;
      if (f >= nf) continue;
      const int2 r = a.range[m];
      const float* bm = a.basis + (int64_t)m * a.n_bins;
      float acc = 0.f;
      for (int k = r.x; k < r.y; ++k) acc = fmaf(__ldg(bm + k), mag[k * FE_LD + f], acc);  // mel_basis @ spec (:78)
      mo[(int64_t)m * a.T + f] = logf(fmaxf(acc, 1e-5f));                                   // :22
    }
  }
}

// Slaney mel scale of librosa.filters.mel (htk=False)
double hz_to_mel(double f) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
  return f >= min_log_hz ? min_log_mel + log(f / min_log_hz) / logstep : f / f_sp;
}
double mel_to_hz(double m) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
  return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m;
}

}  // namespace

struct svk_frontend {
  int device = 0, n_fft = 0, log2n = 0, hop = 0, win = 0, sr = 0, n_mels = 0, n_bins = 0;
  float *d_window = nullptr, *d_basis = nullptr;
  float2* d_twiddle = nullptr;
  int2* d_range = nullptr;
  size_t smem = 0;
};

// librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) as mel_processing.py:76,95 calls it (htk=False, norm="slaney"):
// third-party code absent from the reference tree (requirements.txt: librosa==0.8.0); published algorithm restated.
extern "C" int svk_mel_basis(int sr, int n_fft, int n_mels, float f_lo, float f_hi, float* out) {
  if (!out || sr <= 0 || n_fft < 2 || n_mels < 1 || f_lo < 0) return fe_fail(SVK_ERR_INVALID, "svk_mel_basis: bad argument");
  const double hi = f_hi > 0 ? (double)f_hi : sr / 2.0;  // fmax None -> sr / 2
  if (hi <= f_lo) return fe_fail(SVK_ERR_INVALID, "svk_mel_basis: fmax must exceed fmin");
  const int nb = n_fft / 2 + 1;
  std::vector<double> mel_f(n_mels + 2);
  const double m_lo = hz_to_mel(f_lo), m_hi = hz_to_mel(hi);
  for (int i = 0; i < n_mels + 2; ++i) mel_f[i] = mel_to_hz(m_lo + (m_hi - m_lo) * i / (n_mels + 1));
  for (int m = 0; m < n_mels; ++m) {
    const double enorm = 2.0 / (mel_f[m + 2] - mel_f[m]);
    for (int k = 0; k < nb; ++k) {
      const double f = (sr / 2.0) * k / (nb - 1);
      const double lower = (f - mel_f[m]) / (mel_f[m + 1] - mel_f[m]);
      const double upper = (mel_f[m + 2] - f) / (mel_f[m + 2] - mel_f[m + 1]);
      const double w = fmax(0.0, fmin(lower, upper));
      out[(size_t)m * nb + k] = (float)(w * enorm);
    }
  }
  return SVK_OK;
}

// torch.hann_window(win_size) in fp32 (mel_processing.py:59-60, cast to the signal dtype afterwards), centred in
// n_fft taps as torch.stft does for win_size < n_fft.  Same roundings as torch's arange * float(2 pi / win) -> cos
// -> * -0.5 + 0.5 recipe.
extern "C" int svk_hann_window(int win, int n_fft, float* out) {
  if (!out || win < 1 || n_fft < win) return fe_fail(SVK_ERR_INVALID, "svk_hann_window: need 1 <= win_size <= n_fft");
  const int off = (n_fft - win) / 2;
  for (int i = 0; i < n_fft; ++i) out[i] = 0.f;
  const float step = (float)(2.0 * M_PI / win);
  for (int i = 0; i < win; ++i) {
    const float arg = (float)i * step;
    out[off + i] = (float)cos((double)arg) * -0.5f + 0.5f;
  }
  return SVK_OK;
}

extern "C" int svk_frontend_create(int n_fft, int hop, int win, int sr, int n_mels, float fmin, float fmax, int device,
                                   svk_frontend** out) {
  if (!out) return fe_fail(SVK_ERR_INVALID, "svk_frontend_create: null argument");
  *out = nullptr;
  int log2n = 0;
  while ((1 << log2n) < n_fft) ++log2n;
  if (n_fft < 64 || n_fft > 4096 || (1 << log2n) != n_fft)
    return fe_fail(SVK_ERR_INVALID, "svk_frontend_create: n_fft must be a power of two in [64, 4096]");
  if (hop < 1 || hop > n_fft || (n_fft - hop) % 2) return fe_fail(SVK_ERR_INVALID, "svk_frontend_create: need 1 <= hop <= n_fft, n_fft - hop even");
  if (win < 1 || win > n_fft) return fe_fail(SVK_ERR_INVALID, "svk_frontend_create: need 1 <= win_size <= n_fft");
  const int nb = n_fft / 2 + 1;
  std::vector<float> basis((size_t)n_mels > 0 ? (size_t)n_mels * nb : 0), window(n_fft);
  if (n_mels < 1) return fe_fail(SVK_ERR_INVALID, "svk_frontend_create: n_mels must be positive");
  int st = svk_mel_basis(sr, n_fft, n_mels, fmin, fmax, basis.data());
  if (st < 0) return st;
  st = svk_hann_window(win, n_fft, window.data());
  if (st < 0) return st;
  std::vector<float2> tw(n_fft / 2);
  for (int k = 0; k < n_fft / 2; ++k) {
    const double ang = -2.0 * M_PI * k / n_fft;
    tw[k] = make_float2((float)cos(ang), (float)sin(ang));
  }
  std::vector<int2> range(n_mels);
  for (int m = 0; m < n_mels; ++m) {
    int lo = nb, hi = 0;
    for (int k = 0; k < nb; ++k)
      if (basis[(size_t)m * nb + k] != 0.f) lo = k < lo ? k : lo, hi = k + 1;
    range[m] = lo < hi ? make_int2(lo, hi) : make_int2(0, 0);
  }

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
    return fe_fail(SVK_ERR_CUDA, "svk_frontend_create: no usable CUDA device (there is no CPU fallback)");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10)
    return fe_fail(SVK_ERR_CUDA, "svk_frontend_create: libsvk is built for sm_100a (B200) only");
  if (cudaSetDevice(device) != cudaSuccess) return fe_fail(SVK_ERR_CUDA, "svk_frontend_create: cudaSetDevice failed");
  svk_frontend* f = new svk_frontend();
  f->device = device, f->n_fft = n_fft, f->log2n = log2n, f->hop = hop, f->win = win, f->sr = sr, f->n_mels = n_mels, f->n_bins = nb;
  f->smem = 2 * (size_t)n_fft * sizeof(float2) + (size_t)nb * FE_LD * sizeof(float);
  cudaError_t e = cudaMalloc(&f->d_window, window.size() * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&f->d_basis, basis.size() * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&f->d_twiddle, tw.size() * sizeof(float2));
  if (e == cudaSuccess) e = cudaMalloc(&f->d_range, range.size() * sizeof(int2));
  if (e == cudaSuccess) e = cudaMemcpy(f->d_window, window.data(), window.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(f->d_basis, basis.data(), basis.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(f->d_twiddle, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(f->d_range, range.data(), range.size() * sizeof(int2), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(mel_frontend_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f->smem);
  if (e != cudaSuccess) {
    svk_frontend_destroy(f);
    return fe_fail(SVK_ERR_CUDA, std::string("svk_frontend_create: ") + cudaGetErrorString(e));
  }
  *out = f;
  return SVK_OK;
}

extern "C" void svk_frontend_destroy(svk_frontend* f) {
  if (!f) return;
  cudaSetDevice(f->device);
  cudaFree(f->d_window), cudaFree(f->d_basis), cudaFree(f->d_twiddle), cudaFree(f->d_range);
  delete f;
}

// Frames torch.stft(center=False) yields over the reflect-padded signal (mel_processing.py:62-66).
extern "C" int64_t svk_frontend_frames(const svk_frontend* f, int64_t n) {
  if (!f || n <= 0) return 0;
  const int64_t padded = n + (f->n_fft - f->hop);
  return padded < f->n_fft ? 0 : 1 + (padded - f->n_fft) / f->hop;
}

namespace {
int fe_launch(svk_frontend* f, const char* who, const float* y, const float* spec_in, int B, int64_t n, int64_t T,
              float* spec_out, float* mel_out, void* stream) {
  if (!f || B <= 0 || T <= 0) return fe_fail(SVK_ERR_INVALID, std::string(who) + ": bad argument (null handle, empty batch or no frames)");
  const int64_t tiles = (T + FE_FT - 1) / FE_FT;
  if (tiles * B > 0x7fffffffLL) return fe_fail(SVK_ERR_INVALID, std::string(who) + ": too many frames for one launch");
  FeArgs a;
  a.y = y, a.spec_in = spec_in, a.n = n, a.T = T;
  a.n_fft = f->n_fft, a.log2n = f->log2n, a.hop = f->hop, a.pad = (f->n_fft - f->hop) / 2, a.n_bins = f->n_bins, a.n_mels = f->n_mels;
  a.tiles_t = (int)tiles;
  a.window = f->d_window, a.twiddle = f->d_twiddle, a.basis = f->d_basis, a.range = f->d_range;
  a.spec_out = spec_out, a.mel_out = mel_out;
  mel_frontend_kernel<<<(unsigned)(tiles * B), FE_THREADS, f->smem, (cudaStream_t)stream>>>(a);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fe_fail(SVK_ERR_CUDA, std::string(who) + ": " + cudaGetErrorString(e));
  return SVK_OK;
}
}  // namespace

extern "C" int svk_spectrogram(svk_frontend* f, const float* y, int B, int64_t n, float* spec, void* stream) {
  if (!f || !y || !spec) return fe_fail(SVK_ERR_INVALID, "svk_spectrogram: null argument");
  if (n <= (f->n_fft - f->hop) / 2) return fe_fail(SVK_ERR_INVALID, "svk_spectrogram: reflect padding needs n_samples > (n_fft - hop) / 2");
  return fe_launch(f, "svk_spectrogram", y, nullptr, B, n, svk_frontend_frames(f, n), spec, nullptr, stream);
}

extern "C" int svk_spec_to_mel(svk_frontend* f, const float* spec, int B, int64_t T, float* mel, void* stream) {
  if (!f || !spec || !mel) return fe_fail(SVK_ERR_INVALID, "svk_spec_to_mel: null argument");
  return fe_launch(f, "svk_spec_to_mel", nullptr, spec, B, 0, T, nullptr, mel, stream);
}

extern "C" int svk_mel_spectrogram(svk_frontend* f, const float* y, int B, int64_t n, float* mel, float* spec, void* stream) {
  if (!f || !y || !mel) return fe_fail(SVK_ERR_INVALID, "svk_mel_spectrogram: null argument");
  if (n <= (f->n_fft - f->hop) / 2) return fe_fail(SVK_ERR_INVALID, "svk_mel_spectrogram: reflect padding needs n_samples > (n_fft - hop) / 2");
  return fe_launch(f, "svk_mel_spectrogram", y, nullptr, B, n, svk_frontend_frames(f, n), spec, mel, stream);
}
