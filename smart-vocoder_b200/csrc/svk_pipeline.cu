// Host-side runtime around svk_infer (include/svk.h): CUDA-graph replay of the launch sequence, and a pipelined
// host-buffer entry that overlaps the H2D of call i+1 and the D2H of call i-1 with the kernels of call i
// ("pinned-memory D2H overlapped with compute").  Built on the public entry points only;
// the one kernel here is the counter-based N(0,1) generator for callers that do not bring their own eps.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/svk.h"

extern "C" int svk__set_error(int code, const char* msg);
extern "C" int svk__device(const svk_handle* h);
extern "C" int* svk__range_flag(svk_handle* h);
extern "C" void svk__set_pdl(int enabled);

namespace {

int pfail(int code, const std::string& m) { return svk__set_error(code, m.c_str()); }

#define P_CUDA(expr)                                                                                     \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess) return pfail(SVK_ERR_CUDA, std::string(#expr " failed: ") + cudaGetErrorString(_e)); \
  } while (0)
#define P_TRY(expr)        \
  do {                     \
    int _s = (expr);       \
    if (_s < 0) return _s; \
  } while (0)

inline size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }

// ------------------------------------------------------------------------------------- N(0,1) generator
// Philox4x32-10 (Salmon et al., SC'11; the generator behind torch.randn on CUDA and curand's default), Box-Muller on
// the four outputs.  Element i of a call is a pure function of (seed, offset + i / 4): the stream does not depend on
// grid shape.  It is NOT torch's element order -- callers that need the reference's exact draw pass eps themselves
// (SURVEY F11); this one serves svk_pipeline_submit(eps_host = NULL).
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
  c[0] = n0, c[1] = lo1, c[2] = n2, c[3] = lo0;
}

__global__ void __launch_bounds__(256) randn_kernel(uint64_t seed, uint64_t offset, int64_t n, float* __restrict__ out) {
  const int64_t quads = (n + 3) / 4;
  for (int64_t qd = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; qd < quads; qd += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t ctr = offset + (uint64_t)qd;
    uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      philox_round(c, k0, k1);
      k0 += 0x9E3779B9u, k1 += 0xBB67AE85u;
    }
    float v[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      // (0, 1] uniform from 32 bits, then Box-Muller
      const float u1 = ((float)c[2 * h] + 1.0f) * 2.3283064365386963e-10f;
      const float u2 = ((float)c[2 * h + 1] + 1.0f) * 2.3283064365386963e-10f;
      const float rad = sqrtf(-2.0f * logf(u1));
      float sn, cs;
      sincospif(2.0f * u2, &sn, &cs);
      v[2 * h] = rad * cs, v[2 * h + 1] = rad * sn;
    }
    const int64_t i0 = 4 * qd;
    if (i0 + 3 < n) {
      *reinterpret_cast<float4*>(out + i0) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
      for (int e = 0; e < 4 && i0 + e < n; ++e) out[i0 + e] = v[e];
    }
  }
}

struct Shapes {
  int B, T, Tp, n_mel, inter, hop;
  size_t n_mel_el, n_lat, n_o;
};

int shapes_of(const svk_handle* h, int B, int T, int max_len, Shapes* s) {
  svk_config c;
  P_TRY(svk_get_config(h, &c));
  if (B <= 0 || T <= 0) return pfail(SVK_ERR_INVALID, "B and T must be positive");
  s->B = B, s->T = T, s->Tp = max_len > 0 && max_len < T ? max_len : T;
  s->n_mel = c.n_mel, s->inter = c.inter_channels;
  s->hop = 1;
  for (int i = 0; i < c.n_upsamples; ++i) s->hop *= c.upsample_rates[i];
  s->n_mel_el = (size_t)B * c.n_mel * T, s->n_lat = (size_t)B * c.inter_channels * T, s->n_o = (size_t)B * s->hop * s->Tp;
  return SVK_OK;
}

}  // namespace

extern "C" int svk_randn(svk_handle* h, uint64_t seed, uint64_t offset, int64_t n, float* out_dev, void* stream) {
  if (!h || !out_dev || n < 0) return pfail(SVK_ERR_INVALID, "svk_randn: bad argument");
  if (n == 0) return SVK_OK;
  if ((reinterpret_cast<uintptr_t>(out_dev) & 15) != 0) return pfail(SVK_ERR_INVALID, "svk_randn: output must be 16 B-aligned");
  P_CUDA(cudaSetDevice(svk__device(h)));
  int64_t blocks = ((n + 3) / 4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  randn_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(seed, offset, n, out_dev);
  P_CUDA(cudaGetLastError());
  return SVK_OK;
}

// ------------------------------------------------------------------------------------------ CUDA graph
struct svk_graph {
  svk_handle* h = nullptr;
  Shapes s{};
  float noise_scale = 0.f;
  void* dev = nullptr;  // one allocation: mel | lengths | eps | o | mask | 4 latents | workspace
  size_t o_mel = 0, o_len = 0, o_eps = 0, o_o = 0, o_mask = 0, o_lat[4] = {0, 0, 0, 0}, o_ws = 0, ws_bytes = 0;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  int64_t launches = 0;
  int pdl = 1;
};

extern "C" int svk_graph_create(svk_handle* h, int B, int T, int max_len, float noise_scale, svk_graph** out) {
  if (!h || !out) return pfail(SVK_ERR_INVALID, "svk_graph_create: null argument");
  *out = nullptr;
  svk_graph* g = new svk_graph();
  g->h = h, g->noise_scale = noise_scale;
  int rc = shapes_of(h, B, T, max_len, &g->s);
  if (rc < 0) {
    delete g;
    return rc;
  }
  const Shapes& s = g->s;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = up256(off + bytes);
    return o;
  };
  g->o_mel = take(s.n_mel_el * 4), g->o_len = take((size_t)B * 8), g->o_eps = take(s.n_lat * 4), g->o_o = take(s.n_o * 4);
  g->o_mask = take((size_t)B * T * 4);
  for (int i = 0; i < 4; ++i) g->o_lat[i] = take(s.n_lat * 4);
  g->ws_bytes = svk_workspace_bytes(h, B, T, max_len);
  g->o_ws = take(g->ws_bytes);
  cudaStream_t cs = nullptr;
  auto cleanup = [&](int code) {
    if (cs) cudaStreamDestroy(cs);
    svk_graph_destroy(g);
    return code;
  };
  if (cudaSetDevice(svk__device(h)) != cudaSuccess || cudaMalloc(&g->dev, off) != cudaSuccess)
    return cleanup(pfail(SVK_ERR_CUDA, "svk_graph_create: device allocation failed"));
  if (cudaMemset(g->dev, 0, g->o_ws) != cudaSuccess || cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking) != cudaSuccess)
    return cleanup(pfail(SVK_ERR_CUDA, "svk_graph_create: stream creation failed"));
  char* d = (char*)g->dev;
  auto run = [&](cudaStream_t st) {
    return svk_infer(h, (const float*)(d + g->o_mel), (const int64_t*)(d + g->o_len), (const float*)(d + g->o_eps), noise_scale, B, T,
                     max_len, (float*)(d + g->o_o), (float*)(d + g->o_mask), (float*)(d + g->o_lat[0]), (float*)(d + g->o_lat[1]),
                     (float*)(d + g->o_lat[2]), (float*)(d + g->o_lat[3]), d + g->o_ws, g->ws_bytes, st);
  };
  // eager once: kernel attributes (opt-in shared memory) are set on first use, which capture does not allow
  rc = run(cs);
  if (rc < 0 || cudaStreamSynchronize(cs) != cudaSuccess) return cleanup(rc < 0 ? rc : pfail(SVK_ERR_CUDA, "svk_graph_create: warm-up run failed"));
  g->launches = svk_last_launch_count(h);
  // Capture.  Programmatic dependent launch edges are kept when the toolkit captures them; if instantiation refuses
  // them, capture again with plain stream-order edges.
  for (int attempt = 0; attempt < 2 && !g->exec; ++attempt) {
    g->pdl = attempt == 0;
    svk__set_pdl(g->pdl);
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
    if (e == cudaSuccess) {
      rc = run(cs);
      e = cudaStreamEndCapture(cs, &graph);
      if (rc >= 0 && e == cudaSuccess && graph) {
        cudaGraphExec_t exec = nullptr;
        if (cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) g->graph = graph, g->exec = exec;
      }
      if (!g->exec && graph) cudaGraphDestroy(graph);
    }
    cudaGetLastError();  // clear a failed attempt
  }
  svk__set_pdl(1);
  if (!g->exec) return cleanup(pfail(SVK_ERR_CUDA, "svk_graph_create: stream capture / instantiation failed"));
  cudaStreamDestroy(cs);
  *out = g;
  return SVK_OK;
}

extern "C" int svk_graph_buffers(const svk_graph* g, svk_graph_io* io) {
  if (!g || !io) return pfail(SVK_ERR_INVALID, "svk_graph_buffers: null argument");
  char* d = (char*)g->dev;
  io->mel = (float*)(d + g->o_mel), io->lengths = (int64_t*)(d + g->o_len), io->eps = (float*)(d + g->o_eps);
  io->o = (float*)(d + g->o_o), io->x_mask = (float*)(d + g->o_mask);
  io->z = (float*)(d + g->o_lat[0]), io->z_p = (float*)(d + g->o_lat[1]), io->m_p = (float*)(d + g->o_lat[2]), io->logs_p = (float*)(d + g->o_lat[3]);
  io->B = g->s.B, io->T = g->s.T, io->T_out = g->s.Tp, io->kernel_nodes = g->launches, io->programmatic_edges = g->pdl;
  return SVK_OK;
}

extern "C" int svk_graph_launch(svk_graph* g, void* stream) {
  if (!g || !g->exec) return pfail(SVK_ERR_INVALID, "svk_graph_launch: null graph");
  P_CUDA(cudaGraphLaunch(g->exec, (cudaStream_t)stream));
  return SVK_OK;
}

extern "C" void svk_graph_destroy(svk_graph* g) {
  if (!g) return;
  if (g->exec) cudaGraphExecDestroy(g->exec);
  if (g->graph) cudaGraphDestroy(g->graph);
  if (g->dev) cudaFree(g->dev);
  delete g;
}

// ------------------------------------------------------------------------------------------ pipeline
struct PipeSlot {
  char* dev = nullptr;  // mel | lengths | eps | o | mask | flag
  cudaEvent_t ev_in = nullptr, ev_c = nullptr, ev_out = nullptr;
  int* flag_host = nullptr;  // pinned
  int64_t ticket = -1;       // ticket occupying the slot, -1 = free
  bool waited = true;
};

struct svk_pipeline {
  svk_handle* h = nullptr;
  Shapes s{};
  int max_len = 0, depth = 0;
  size_t o_mel = 0, o_len = 0, o_eps = 0, o_o = 0, o_mask = 0, o_flag = 0, slot_bytes = 0, ws_bytes = 0;
  void* ws = nullptr;
  cudaStream_t s_in = nullptr, s_c = nullptr, s_out = nullptr;
  std::vector<PipeSlot> slots;
  int64_t next_ticket = 0;
  uint64_t draws = 0;  // Philox offset consumed so far (quads)
};

extern "C" void svk_pipeline_destroy(svk_pipeline* p) {
  if (!p) return;
  cudaSetDevice(svk__device(p->h));
  if (p->s_c) cudaStreamSynchronize(p->s_c);
  if (p->s_out) cudaStreamSynchronize(p->s_out);
  if (p->s_in) cudaStreamSynchronize(p->s_in);
  for (PipeSlot& sl : p->slots) {
    if (sl.dev) cudaFree(sl.dev);
    if (sl.flag_host) cudaFreeHost(sl.flag_host);
    if (sl.ev_in) cudaEventDestroy(sl.ev_in);
    if (sl.ev_c) cudaEventDestroy(sl.ev_c);
    if (sl.ev_out) cudaEventDestroy(sl.ev_out);
  }
  if (p->ws) cudaFree(p->ws);
  if (p->s_in) cudaStreamDestroy(p->s_in);
  if (p->s_c) cudaStreamDestroy(p->s_c);
  if (p->s_out) cudaStreamDestroy(p->s_out);
  delete p;
}

extern "C" int svk_pipeline_create(svk_handle* h, int B, int T, int max_len, int depth, svk_pipeline** out) {
  if (!h || !out) return pfail(SVK_ERR_INVALID, "svk_pipeline_create: null argument");
  *out = nullptr;
  if (depth < 1 || depth > 8) return pfail(SVK_ERR_INVALID, "svk_pipeline_create: depth must be in [1, 8]");
  svk_pipeline* p = new svk_pipeline();
  p->h = h, p->max_len = max_len, p->depth = depth;
  int rc = shapes_of(h, B, T, max_len, &p->s);
  if (rc < 0) {
    delete p;
    return rc;
  }
  const Shapes& s = p->s;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = up256(off + bytes);
    return o;
  };
  p->o_mel = take(s.n_mel_el * 4), p->o_len = take((size_t)B * 8), p->o_eps = take(s.n_lat * 4), p->o_o = take(s.n_o * 4);
  p->o_mask = take((size_t)B * T * 4), p->o_flag = take(4);
  p->slot_bytes = off;
  p->ws_bytes = svk_workspace_bytes(h, B, T, max_len);
  bool ok = cudaSetDevice(svk__device(h)) == cudaSuccess && cudaMalloc(&p->ws, p->ws_bytes) == cudaSuccess &&
            cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&p->s_c, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking) == cudaSuccess;
  p->slots.resize(depth);
  for (int i = 0; ok && i < depth; ++i) {
    PipeSlot& sl = p->slots[i];
    ok = cudaMalloc((void**)&sl.dev, p->slot_bytes) == cudaSuccess && cudaHostAlloc((void**)&sl.flag_host, sizeof(int), cudaHostAllocDefault) == cudaSuccess &&
         cudaEventCreateWithFlags(&sl.ev_in, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&sl.ev_c, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&sl.ev_out, cudaEventDisableTiming) == cudaSuccess;
    if (ok) *sl.flag_host = 0;
  }
  if (!ok) {
    svk_pipeline_destroy(p);
    return pfail(SVK_ERR_CUDA, "svk_pipeline_create: allocation failed");
  }
  *out = p;
  return SVK_OK;
}

extern "C" int svk_pipeline_wait(svk_pipeline* p, int64_t ticket) {
  if (!p || ticket < 0 || ticket >= p->next_ticket) return pfail(SVK_ERR_INVALID, "svk_pipeline_wait: unknown ticket");
  PipeSlot& sl = p->slots[ticket % p->depth];
  if (sl.ticket != ticket) return pfail(SVK_ERR_STATE, "svk_pipeline_wait: the slot of this ticket was reused (wait in submission order)");
  if (sl.waited) return SVK_OK;
  P_CUDA(cudaEventSynchronize(sl.ev_out));
  sl.waited = true;
  if (*sl.flag_host) {
    *sl.flag_host = 0;
    return pfail(SVK_ERR_RANGE, "non-finite waveform sample: an activation left the fp16 operand range of the tensor-core engine");
  }
  return SVK_OK;
}

extern "C" int svk_pipeline_submit(svk_pipeline* p, const float* mel, const int64_t* lengths, const float* eps, uint64_t seed,
                                   float noise_scale, float* o, float* x_mask, int64_t* ticket_out) {
  if (!p || !mel || !lengths || !o) return pfail(SVK_ERR_INVALID, "svk_pipeline_submit: null mel / lengths / o");
  P_CUDA(cudaSetDevice(svk__device(p->h)));
  const int64_t ticket = p->next_ticket;
  PipeSlot& sl = p->slots[ticket % p->depth];
  if (!sl.waited) {
    // the slot still holds an unfinished call: its result must be in host memory before the buffers are reused
    const int rc = svk_pipeline_wait(p, sl.ticket);
    if (rc < 0 && rc != SVK_ERR_RANGE) return rc;
  }
  const Shapes& s = p->s;
  char* d = sl.dev;
  // H2D on its own stream: overlaps the kernels of the previous ticket
  P_CUDA(cudaMemcpyAsync(d + p->o_mel, mel, s.n_mel_el * 4, cudaMemcpyHostToDevice, p->s_in));
  P_CUDA(cudaMemcpyAsync(d + p->o_len, lengths, (size_t)s.B * 8, cudaMemcpyHostToDevice, p->s_in));
  if (eps) P_CUDA(cudaMemcpyAsync(d + p->o_eps, eps, s.n_lat * 4, cudaMemcpyHostToDevice, p->s_in));
  P_CUDA(cudaEventRecord(sl.ev_in, p->s_in));
  // kernels, serialised on the compute stream (one workspace)
  P_CUDA(cudaStreamWaitEvent(p->s_c, sl.ev_in, 0));
  if (!eps) {
    P_TRY(svk_randn(p->h, seed, p->draws, (int64_t)s.n_lat, (float*)(d + p->o_eps), p->s_c));
    p->draws += (s.n_lat + 3) / 4;
  }
  P_TRY(svk_infer(p->h, (const float*)(d + p->o_mel), (const int64_t*)(d + p->o_len), (const float*)(d + p->o_eps), noise_scale, s.B, s.T,
                  p->max_len, (float*)(d + p->o_o), (float*)(d + p->o_mask), nullptr, nullptr, nullptr, nullptr, p->ws, p->ws_bytes, p->s_c));
  int* flag = svk__range_flag(p->h);
  P_CUDA(cudaMemcpyAsync(d + p->o_flag, flag, sizeof(int), cudaMemcpyDeviceToDevice, p->s_c));
  P_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), p->s_c));
  P_CUDA(cudaEventRecord(sl.ev_c, p->s_c));
  // D2H on its own stream: overlaps the kernels of the next ticket
  P_CUDA(cudaStreamWaitEvent(p->s_out, sl.ev_c, 0));
  P_CUDA(cudaMemcpyAsync(o, d + p->o_o, s.n_o * 4, cudaMemcpyDeviceToHost, p->s_out));
  if (x_mask) P_CUDA(cudaMemcpyAsync(x_mask, d + p->o_mask, (size_t)s.B * s.T * 4, cudaMemcpyDeviceToHost, p->s_out));
  P_CUDA(cudaMemcpyAsync(sl.flag_host, d + p->o_flag, sizeof(int), cudaMemcpyDeviceToHost, p->s_out));
  P_CUDA(cudaEventRecord(sl.ev_out, p->s_out));
  // the next use of this slot's INPUT buffers is ordered after these kernels by the wait above (ev_out follows ev_c)
  sl.ticket = ticket, sl.waited = false;
  p->next_ticket = ticket + 1;
  if (ticket_out) *ticket_out = ticket;
  return SVK_OK;
}

extern "C" int svk_pipeline_drain(svk_pipeline* p) {
  if (!p) return pfail(SVK_ERR_INVALID, "svk_pipeline_drain: null pipeline");
  int worst = SVK_OK;
  for (int64_t t = p->next_ticket - p->depth < 0 ? 0 : p->next_ticket - p->depth; t < p->next_ticket; ++t) {
    const int rc = svk_pipeline_wait(p, t);
    if (rc < 0) worst = rc;
  }
  return worst;
}
