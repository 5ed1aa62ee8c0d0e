// Host side of libsvk: checkpoint-key surface, weight folding / packing, workspace planning and
// the launch sequence of SynthesizerTrn.infer (reference models.py:331-339).  Everything numeric
// runs in the CUDA kernels of conv_ffma.cu / elementwise.cu; there is no CPU compute fallback.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <array>
#include <map>
#include <string>
#include <vector>

#include "../../include/svk.h"
#include "svk_kernels.cuh"

using namespace svk;

// ------------------------------------------------------------------------------------- errors
static thread_local std::string g_last_error;

static int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

// used by svk_ops.cu, which shares this translation unit's thread-local error string
extern "C" int svk__set_error(int code, const char* msg) { return fail(code, "%s", msg); }

#define CUDA_TRY(expr)                                                                        \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      return fail(SVK_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                  __LINE__);                                                                  \
  } while (0)

#define SVK_TRY(expr)       \
  do {                      \
    int _s = (expr);        \
    if (_s < 0) return _s;  \
  } while (0)

// ------------------------------------------------------------------------------------ structs
namespace {

struct KeySpec {
  std::vector<int64_t> shape;
  bool dead;      // never read on any built path: accepted and dropped
  bool optional;  // enc_q.*: only the analysis direction (svk_posterior_encoder) reads them; infer works without
};

struct HostTensor {
  std::vector<float> data;
  std::vector<int64_t> shape;
};

// A layer packed for conv_ffma: weights [Cin][K][CoutPad] + bias [CoutPad] inside the device blob.
struct PackedConv {
  size_t w_off = 0, b_off = 0;  // float offsets into the blob
  int Cin = 0, Cout = 0, CoutPad = 0, K = 0;
  // ConvTranspose1d only
  int s = 0, p = 0, k = 0, pad_virtual = 0;
  // tcgen05 image (conv_tc.cu): fp16 hi/lo planes in the half blob, bias in TC channel order
  bool tc = false;
  size_t tc_off = 0, tc_b_off = 0;  // half offset into d_tcblob; float offset into d_blob
  int tc_N = 0;
  float tc_unscale = 1.0f;
  bool rmajor = false;  // ConvTranspose1d, tcgen05 image only: virtual channels ordered r * Cout + co (ConvArgs::shuf_rmajor)
};

struct FlowLayers {
  PackedConv pre, post, post_fwd;  // post: negated (reverse pass subtracts m), post_fwd: as stored (forward adds m)
  std::vector<PackedConv> in, rs;
  int orient = 0;  // 1: this coupling sees the channel-reversed view (odd number of Flips before it)
};

struct ResBlock {
  PackedConv c1[SVK_RESBLOCK_PAIRS], c2[SVK_RESBLOCK_PAIRS];  // type 2: c1[l] = convs[l], c2 unused
  int k = 0, C = 0;
  int dil[SVK_RESBLOCK_PAIRS] = {1, 1, 1};
  int type = 1;  // 1: ResBlock1 (n conv pairs), 2: ResBlock2 (n single convs)
  int n = SVK_RESBLOCK_PAIRS;
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

struct svk_handle {
  svk_config cfg;
  int device = 0;
  std::vector<std::string> key_order;
  std::map<std::string, KeySpec> spec;
  std::map<std::string, HostTensor> raw;
  int n_live = 0;
  bool finalized = false;

  float* d_blob = nullptr;
  size_t blob_floats = 0;
  std::vector<float> h_blob;  // staging while packing
  uint16_t* d_tcblob = nullptr;  // fp16 hi/lo weight images of the tcgen05 path
  std::vector<uint16_t> h_tcblob;

  PackedConv pre_enc, proj, conv_pre, conv_post;
  std::vector<PackedConv> enc_in, enc_rs, ups;
  // PosteriorEncoder (enc_q.*, models.py:83-110): packed when its keys were loaded
  bool has_posterior = false;
  PackedConv encq_pre, encq_proj;
  std::vector<PackedConv> encq_in, encq_rs;
  std::vector<FlowLayers> flows;
  std::vector<ResBlock> resblocks;

  int64_t launches = 0;
  bool fuse_pairs = true;  // fused ResBlock conv pairs on the narrow stages ($SVK_FUSE_PAIRS=0 disables: A/B measurements)
  bool fuse_pairs_all = false;
  bool fuse_pairs_c32 = true;  // every kernel size of the C = 32 stage too (its weights stay resident in the pair kernel):
                               // per launch the k >= 7 pairs only break even with the unfused convs, but they move 6.4 GB
                               // less per step and the power-capped step gains 0.7 % (profiles/r2_ab_fuse_c32.log)
  int img_stream = 2;           // $SVK_IMG_STREAM (A/B; default 2 since the end of round 2: fewest HBM bytes, -0.35 ms on the
                                // power-capped step, profiles/r2_ab_img_stream.log): 3 = the fused pairs image-only, the unfused blocks as in 0;
                                // 0 = per-shape default (Runner::resblock_images), 1 = every unfused ResBlock keeps
                                // its residual stream between pairs as operand images only, 2 = also the first pair and the
                                // fused pairs, -1 = the round-1 rule (C >= 256, C >= 128 with k >= 7)
  int fuse_wn = 1;              // one launch per WN layer (wn_layer.cu): 1 = when the batch fills the GPU (see Runner::wn),
                                // 2 = always, 0 = never ($SVK_FUSE_WN; the two-launch form spreads a layer over 3x more CTAs)
  int wn_stack = 1;             // whole WN stacks in one launch when every tile gets its own SM ($SVK_WN_STACK=0: per layer)

  // svk_profile_begin/end state
  bool profiling = false;
  std::vector<cudaEvent_t> prof_events;   // 2 per record
  std::vector<svk_launch_record> prof_records;
  size_t prof_cap = 0;

  int* d_range_flag = nullptr;  // raised by the final tanh epilogue on a non-finite sample (svk_check_range)
  int* d_wn_flags = nullptr;    // [1024] tile progress counters of the multi-layer WN launches (wn_layer.cu), zeroed per launch

  // svk_infer_host state
  cudaStream_t host_stream = nullptr;
  void* host_dev = nullptr;
  size_t host_dev_bytes = 0;

  bool tensor_engine() const { return cfg.precision == SVK_PRECISION_TC || cfg.precision == SVK_PRECISION_BF16; }
  int planes() const { return cfg.precision == SVK_PRECISION_BF16 ? 1 : 2; }  // operand-image planes (conv_tc.cu)
  int hop() const {
    int h = 1;
    for (int i = 0; i < cfg.n_upsamples; ++i) h *= cfg.upsample_rates[i];
    return h;
  }
  int stage_channels(int i) const { return cfg.upsample_initial_channel >> (i + 1); }
};

// ------------------------------------------------------------------------ checkpoint key surface
namespace {

void add_key(svk_handle* h, const std::string& key, std::vector<int64_t> shape) {
  const bool dead = key.find(".cond_layer.") != std::string::npos || key.rfind("dec.cond.", 0) == 0;
  const bool optional = !dead && key.rfind("enc_q.", 0) == 0;
  h->key_order.push_back(key);
  h->spec[key] = KeySpec{std::move(shape), dead, optional};
  if (!dead && !optional) h->n_live++;
}

void add_wn_keys(svk_handle* h, const std::string& prefix, int hidden, int kernel, int n_layers, int gin) {
  for (int i = 0; i < n_layers; ++i) {
    const std::string p = prefix + ".in_layers." + std::to_string(i);
    add_key(h, p + ".bias", {2 * hidden});
    add_key(h, p + ".weight_g", {2 * hidden, 1, 1});
    add_key(h, p + ".weight_v", {2 * hidden, hidden, kernel});
  }
  for (int i = 0; i < n_layers; ++i) {
    const int rs = i < n_layers - 1 ? 2 * hidden : hidden;
    const std::string p = prefix + ".res_skip_layers." + std::to_string(i);
    add_key(h, p + ".bias", {rs});
    add_key(h, p + ".weight_g", {rs, 1, 1});
    add_key(h, p + ".weight_v", {rs, hidden, 1});
  }
  if (gin != 0) {
    const std::string p = prefix + ".cond_layer";
    const int c = 2 * hidden * n_layers;
    add_key(h, p + ".bias", {c});
    add_key(h, p + ".weight_g", {c, 1, 1});
    add_key(h, p + ".weight_v", {c, gin, 1});
  }
}

// Same ordered surface as SynthesizerTrn(...).state_dict() of the reference.
void build_key_spec(svk_handle* h) {
  const svk_config& c = h->cfg;
  const int H = c.hidden_channels, C = c.inter_channels, gin = c.gin_channels, U = c.upsample_initial_channel;
  add_wn_keys(h, "enc_p.encoder", H, c.wn_kernel, c.enc_layers, gin);
  add_key(h, "enc_p.pre_enc.weight", {H, c.n_mel, 1});
  add_key(h, "enc_p.pre_enc.bias", {H});
  add_key(h, "enc_p.proj.weight", {2 * C, H, 1});
  add_key(h, "enc_p.proj.bias", {2 * C});
  add_key(h, "dec.conv_pre.weight", {U, C, 7});
  add_key(h, "dec.conv_pre.bias", {U});
  for (int i = 0; i < c.n_upsamples; ++i) {
    const int cin = U >> i, cout = U >> (i + 1);
    const std::string p = "dec.ups." + std::to_string(i);
    add_key(h, p + ".bias", {cout});
    add_key(h, p + ".weight_g", {cin, 1, 1});
    add_key(h, p + ".weight_v", {cin, cout, c.upsample_kernel_sizes[i]});
  }
  for (int i = 0; i < c.n_upsamples; ++i) {
    const int ch = U >> (i + 1);
    for (int j = 0; j < c.n_resblock_kernels; ++j) {
      const int n = i * c.n_resblock_kernels + j;
      if (c.resblock_type == 2) {
        for (int l = 0; l < 2; ++l) {
          const std::string p = "dec.resblocks." + std::to_string(n) + ".convs." + std::to_string(l);
          add_key(h, p + ".bias", {ch});
          add_key(h, p + ".weight_g", {ch, 1, 1});
          add_key(h, p + ".weight_v", {ch, ch, c.resblock_kernel_sizes[j]});
        }
        continue;
      }
      for (const char* grp : {"convs1", "convs2"})
        for (int l = 0; l < SVK_RESBLOCK_PAIRS; ++l) {
          const std::string p = "dec.resblocks." + std::to_string(n) + "." + grp + "." + std::to_string(l);
          add_key(h, p + ".bias", {ch});
          add_key(h, p + ".weight_g", {ch, 1, 1});
          add_key(h, p + ".weight_v", {ch, ch, c.resblock_kernel_sizes[j]});
        }
    }
  }
  add_key(h, "dec.conv_post.weight", {1, U >> c.n_upsamples, 7});
  if (gin != 0) {
    add_key(h, "dec.cond.weight", {U, gin, 1});
    add_key(h, "dec.cond.bias", {U});
  }
  add_key(h, "enc_q.pre.weight", {H, c.spec_channels, 1});
  add_key(h, "enc_q.pre.bias", {H});
  add_wn_keys(h, "enc_q.enc", H, 5, 16, gin);
  add_key(h, "enc_q.proj.weight", {2 * C, H, 1});
  add_key(h, "enc_q.proj.bias", {2 * C});
  for (int f = 0; f < c.n_flows; ++f) {
    const std::string p = "flow.flows." + std::to_string(2 * f);
    add_key(h, p + ".pre.weight", {H, C / 2, 1});
    add_key(h, p + ".pre.bias", {H});
    add_wn_keys(h, p + ".enc", H, c.wn_kernel, c.flow_layers, gin);
    add_key(h, p + ".post.weight", {C / 2, H, 1});
    add_key(h, p + ".post.bias", {C / 2});
  }
}

// ------------------------------------------------------------------------------ fold and pack
struct Folded {
  std::vector<float> w;  // [d0][d1][k]
  std::vector<float> b;  // may be empty
  int d0 = 0, d1 = 0, k = 0;
};

// weight_norm fold: norm over dims != 0, accumulated in double, rounded once.
int fold_layer(svk_handle* h, const std::string& prefix, Folded* out) {
  *out = Folded();
  auto itv = h->raw.find(prefix + ".weight_v");
  const HostTensor* wt = nullptr;
  if (itv != h->raw.end()) {
    auto itg = h->raw.find(prefix + ".weight_g");
    if (itg == h->raw.end()) return fail(SVK_ERR_STATE, "missing %s.weight_g", prefix.c_str());
    const HostTensor& v = itv->second;
    out->d0 = (int)v.shape[0], out->d1 = (int)v.shape[1], out->k = (int)v.shape[2];
    const size_t inner = (size_t)out->d1 * out->k;
    out->w.resize(v.data.size());
    for (int i = 0; i < out->d0; ++i) {
      double s = 0;
      for (size_t j = 0; j < inner; ++j) s += (double)v.data[i * inner + j] * (double)v.data[i * inner + j];
      const float scale = (float)((double)itg->second.data[i] / sqrt(s));
      for (size_t j = 0; j < inner; ++j) out->w[i * inner + j] = v.data[i * inner + j] * scale;
    }
  } else {
    auto itw = h->raw.find(prefix + ".weight");
    if (itw == h->raw.end()) return fail(SVK_ERR_STATE, "missing %s.weight", prefix.c_str());
    wt = &itw->second;
    out->d0 = (int)wt->shape[0], out->d1 = (int)wt->shape[1], out->k = (int)wt->shape[2];
    out->w = wt->data;
  }
  auto itb = h->raw.find(prefix + ".bias");
  if (itb != h->raw.end()) out->b = itb->second.data;
  return SVK_OK;
}

// Output channels per CTA of the tcgen05 kernel: as few equal tiles of <= 128 channels as possible.
int tc_tile_n(int Cout, int granule) {
  const int ntiles = (Cout + 127) / 128;
  const int per = (Cout + ntiles - 1) / ntiles;
  return (per + granule - 1) / granule * granule;
}

// tcgen05 image of a virtual conv.  `gate_half` > 0: WN in_layer, each N-tile holds N/2 tanh
// channels followed by the matching N/2 sigmoid channels (real channel + gate_half).
template <class WF, class BF>
void pack_tc(svk_handle* h, PackedConv* pc, WF wv, BF bv, int gate_half) {
  const int Cin = pc->Cin, Cout = pc->Cout, K = pc->K;
  if (!h->tensor_engine() || Cin % TC_KC != 0 || Cout < 16) return;
  const int N = tc_tile_n(Cout, gate_half ? 32 : 16);
  const int ntiles = (Cout + N - 1) / N;
  const int CoutV = ntiles * N;
  auto real = [&](int ov) -> int {  // virtual TC channel -> logical channel (-1: padding)
    if (!gate_half) return ov < Cout ? ov : -1;
    const int nt = ov / N, n = ov % N, hf = N / 2;
    const int c = nt * hf + (n < hf ? n : n - hf);
    if (c >= gate_half) return -1;
    return n < hf ? c : gate_half + c;
  };
  std::vector<float> w((size_t)CoutV * Cin * K, 0.f);
  for (int ov = 0; ov < CoutV; ++ov) {
    const int o = real(ov);
    if (o < 0) continue;
    for (int c = 0; c < Cin; ++c)
      for (int j = 0; j < K; ++j) w[((size_t)ov * Cin + c) * K + j] = wv(o, c, j);
  }
  const int planes = h->planes();
  const float scale = planes == 2 ? conv_tc_weight_scale(w.data(), w.size()) : 1.0f;  // bf16 has fp32's range
  pc->tc_off = align_up(h->h_tcblob.size(), 64);
  h->h_tcblob.resize(pc->tc_off + conv_tc_packed_halves(Cin, CoutV, K, N, planes));
  conv_tc_pack(w.data(), CoutV, Cin, K, N, scale, h->h_tcblob.data() + pc->tc_off, planes);
  pc->tc_b_off = align_up(h->h_blob.size(), 64);
  h->h_blob.resize(pc->tc_b_off + CoutV, 0.f);
  for (int ov = 0; ov < CoutV; ++ov) {
    const int o = real(ov);
    if (o >= 0) h->h_blob[pc->tc_b_off + ov] = bv(o);
  }
  pc->tc = true, pc->tc_N = N, pc->tc_unscale = 1.0f / scale;
}

// Reserve blob space and write a virtual conv: wv(o,c,j), bv(o) -> [Cin][K][CoutPad] + [CoutPad].
template <class WF, class BF>
PackedConv pack_virtual(svk_handle* h, int Cin, int Cout, int K, WF wv, BF bv, bool want_tc = true) {
  PackedConv pc;
  pc.Cin = Cin, pc.Cout = Cout, pc.K = K;
  const int ot = conv_ffma_channel_tile(Cout);
  pc.CoutPad = (Cout + ot - 1) / ot * ot;
  pc.w_off = align_up(h->h_blob.size(), 64);
  h->h_blob.resize(pc.w_off + (size_t)Cin * K * pc.CoutPad, 0.f);
  for (int c = 0; c < Cin; ++c)
    for (int j = 0; j < K; ++j)
      for (int o = 0; o < Cout; ++o) h->h_blob[pc.w_off + ((size_t)c * K + j) * pc.CoutPad + o] = wv(o, c, j);
  pc.b_off = align_up(h->h_blob.size(), 64);
  h->h_blob.resize(pc.b_off + pc.CoutPad, 0.f);
  for (int o = 0; o < Cout; ++o) h->h_blob[pc.b_off + o] = bv(o);
  if (want_tc) pack_tc(h, &pc, wv, bv, 0);
  return pc;
}

PackedConv pack_plain(svk_handle* h, const Folded& f) {
  return pack_virtual(
      h, f.d1, f.d0, f.k, [&](int o, int c, int j) { return f.w[((size_t)o * f.d1 + c) * f.k + j]; },
      [&](int o) { return f.b.empty() ? 0.f : f.b[o]; });
}

// WN in_layer: interleave tanh/sigmoid halves so one thread owns both halves of 4 channels.
PackedConv pack_gate(svk_handle* h, const Folded& f, int H) {
  auto real = [H](int o) {
    const int grp = o >> 3, e = o & 7;
    return e < 4 ? 4 * grp + e : H + 4 * grp + (e - 4);
  };
  PackedConv pc = pack_virtual(
      h, f.d1, f.d0, f.k, [&](int o, int c, int j) { return f.w[((size_t)real(o) * f.d1 + c) * f.k + j]; },
      [&](int o) { return f.b.empty() ? 0.f : f.b[real(o)]; }, /*want_tc=*/false);
  pack_tc(
      h, &pc, [&](int o, int c, int j) { return f.w[((size_t)o * f.d1 + c) * f.k + j]; },
      [&](int o) { return f.b.empty() ? 0.f : f.b[o]; }, H);
  return pc;
}

// ConvTranspose1d as a K'=ceil(k/s)-tap conv over the input producing s*Cout virtual channels
// (polyphase form): o' = co*s + r, tap jj <-> x[q-(K'-1)+jj], j = r + (K'-1-jj)*s.
// rmajor_tc: the tcgen05 image (only) orders the virtual channels o' = r*Cout + co, so that a 16-column epilogue job is 16
// real channels of one output step and the upsampler can write the stage's operand image itself (conv_tc.cu).
PackedConv pack_transposed(svk_handle* h, const Folded& f, int s, int p, bool rmajor_tc = false) {
  const int Cin = f.d0, Cout = f.d1, k = f.k;
  const int Kv = (k + s - 1) / s;
  auto wv = [&](int co, int r, int c, int jj) {
    const int j = r + (Kv - 1 - jj) * s;
    return j < k ? f.w[((size_t)c * Cout + co) * k + j] : 0.f;
  };
  rmajor_tc = rmajor_tc && Cout % 16 == 0;
  PackedConv pc = pack_virtual(
      h, Cin, Cout * s, Kv, [&](int o, int c, int jj) { return wv(o / s, o % s, c, jj); },
      [&](int o) { return f.b.empty() ? 0.f : f.b[o / s]; }, /*want_tc=*/!rmajor_tc);
  if (rmajor_tc) {
    pack_tc(
        h, &pc, [&](int o, int c, int jj) { return wv(o % Cout, o / Cout, c, jj); },
        [&](int o) { return f.b.empty() ? 0.f : f.b[o % Cout]; }, 0);
    pc.rmajor = pc.tc;
  }
  pc.s = s, pc.p = p, pc.k = k, pc.pad_virtual = Kv - 1;
  return pc;
}

int pack_wn(svk_handle* h, const std::string& prefix, int n_layers, std::vector<PackedConv>* in,
            std::vector<PackedConv>* rs) {
  const int H = h->cfg.hidden_channels;
  for (int i = 0; i < n_layers; ++i) {
    Folded f;
    SVK_TRY(fold_layer(h, prefix + ".in_layers." + std::to_string(i), &f));
    in->push_back(pack_gate(h, f, H));
    Folded g;
    SVK_TRY(fold_layer(h, prefix + ".res_skip_layers." + std::to_string(i), &g));
    rs->push_back(pack_plain(h, g));
  }
  return SVK_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------ lifetime
extern "C" {
int svk_g_pdl_enabled = 1;
}
extern "C" void svk__set_pdl(int enabled) { svk_g_pdl_enabled = enabled ? 1 : 0; }
extern "C" int svk__device(const svk_handle* h) { return h ? h->device : 0; }
extern "C" int* svk__range_flag(svk_handle* h) { return h ? h->d_range_flag : nullptr; }
extern "C" int svk_get_config(const svk_handle* h, svk_config* out) {
  if (!h || !out) return fail(SVK_ERR_INVALID, "svk_get_config: null argument");
  *out = h->cfg;
  return SVK_OK;
}
extern "C" int svk_abi_version(void) { return SVK_ABI_VERSION; }
extern "C" const char* svk_last_error(void) { return g_last_error.c_str(); }

extern "C" int svk_create(const svk_config* cfg, int device, svk_handle** out) {
  if (!cfg || !out) return fail(SVK_ERR_INVALID, "svk_create: null argument");
  *out = nullptr;
  const svk_config& c = *cfg;
  // reference asserts: modules.py:308 (channels % 2), modules.py:114 (odd WN kernel)
  if (c.inter_channels <= 0 || c.inter_channels % 2) return fail(SVK_ERR_INVALID, "channels should be divisible by 2");
  if (c.wn_kernel % 2 != 1) return fail(SVK_ERR_INVALID, "WN kernel_size must be odd");
  if (c.n_upsamples < 1 || c.n_upsamples > SVK_MAX_UPSAMPLES) return fail(SVK_ERR_INVALID, "n_upsamples out of range");
  if (c.n_resblock_kernels < 1 || c.n_resblock_kernels > SVK_MAX_RESBLOCK_KERNELS)
    return fail(SVK_ERR_INVALID, "n_resblock_kernels out of range");
  if (c.resblock_type < 0 || c.resblock_type > 2) return fail(SVK_ERR_INVALID, "resblock_type must be 1 (ResBlock1) or 2 (ResBlock2)");
  if (c.precision != SVK_PRECISION_FP32 && c.precision != SVK_PRECISION_TC && c.precision != SVK_PRECISION_BF16)
    return fail(SVK_ERR_INVALID, "unsupported precision %d", c.precision);
  if (c.hidden_channels % 8 || c.inter_channels % 16 || c.n_mel % 8)
    return fail(SVK_ERR_INVALID, "n_mel, hidden_channels must be multiples of 8 and inter_channels of 16");
  if ((c.upsample_initial_channel >> c.n_upsamples) < 8 || (c.upsample_initial_channel >> c.n_upsamples) % 8)
    return fail(SVK_ERR_INVALID, "decoder channel widths must stay multiples of 8");
  if (!conv_ffma_supports_k(c.wn_kernel)) return fail(SVK_ERR_INVALID, "unsupported WN kernel size %d", c.wn_kernel);
  for (int i = 0; i < c.n_upsamples; ++i) {
    const int u = c.upsample_rates[i], k = c.upsample_kernel_sizes[i];
    if (u < 1 || k < u || (k - u) % 2) return fail(SVK_ERR_INVALID, "upsample %d: need k >= u and (k-u) even", i);
    if (!conv_ffma_supports_k((k + u - 1) / u)) return fail(SVK_ERR_INVALID, "upsample %d: unsupported k/u ratio", i);
  }
  for (int j = 0; j < c.n_resblock_kernels; ++j) {
    if (c.resblock_kernel_sizes[j] % 2 != 1 || !conv_ffma_supports_k(c.resblock_kernel_sizes[j]))
      return fail(SVK_ERR_INVALID, "unsupported resblock kernel size %d", c.resblock_kernel_sizes[j]);
    for (int l = 0; l < (c.resblock_type == 2 ? 2 : SVK_RESBLOCK_PAIRS); ++l)
      if (c.resblock_dilations[j][l] < 1 || c.resblock_dilations[j][l] > 8)
        return fail(SVK_ERR_INVALID, "resblock dilation must be in [1,8]");
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
    return fail(SVK_ERR_CUDA, "svk_create: CUDA device %d not available (%d visible); libsvk has no CPU path", device, ndev);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(SVK_ERR_CUDA, "svk_create: device %d is sm_%d%d; libsvk is built for sm_100a only", device, prop.major, prop.minor);
  svk_handle* h = new svk_handle();
  h->cfg = c;
  h->device = device;
  if (const char* e = getenv("SVK_FUSE_PAIRS"))
    h->fuse_pairs = atoi(e) != 0, h->fuse_pairs_all = atoi(e) == 2, h->fuse_pairs_c32 = atoi(e) == 3;  // 1: 3-tap blocks only
  if (const char* e = getenv("SVK_FUSE_WN")) h->fuse_wn = atoi(e);
  if (const char* e = getenv("SVK_WN_STACK")) h->wn_stack = atoi(e);
  if (const char* e = getenv("SVK_IMG_STREAM")) h->img_stream = atoi(e);
  build_key_spec(h);
  if (cudaSetDevice(device) != cudaSuccess || cudaMalloc(&h->d_range_flag, sizeof(int)) != cudaSuccess ||
      cudaMemset(h->d_range_flag, 0, sizeof(int)) != cudaSuccess || cudaMalloc(&h->d_wn_flags, 1024 * sizeof(int)) != cudaSuccess) {
    delete h;
    return fail(SVK_ERR_CUDA, "svk_create: device allocation failed");
  }
  *out = h;
  return SVK_OK;
}

extern "C" void svk_destroy(svk_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->d_blob) cudaFree(h->d_blob);
  if (h->d_tcblob) cudaFree(h->d_tcblob);
  if (h->host_dev) cudaFree(h->host_dev);
  if (h->d_range_flag) cudaFree(h->d_range_flag);
  if (h->d_wn_flags) cudaFree(h->d_wn_flags);
  if (h->host_stream) cudaStreamDestroy(h->host_stream);
  for (cudaEvent_t e : h->prof_events) cudaEventDestroy(e);
  delete h;
}

// ------------------------------------------------------------------------------- weight ingress
extern "C" int svk_load_tensor(svk_handle* h, const char* key, const float* host_data, const int64_t* shape,
                               int ndim) {
  if (!h || !key || !host_data || !shape) return fail(SVK_ERR_INVALID, "svk_load_tensor: null argument");
  auto it = h->spec.find(key);
  if (it == h->spec.end()) return fail(SVK_ERR_UNKNOWN_KEY, "unexpected key '%s'", key);
  const KeySpec& ks = it->second;
  bool same = (int)ks.shape.size() == ndim;
  for (int i = 0; same && i < ndim; ++i) same = ks.shape[i] == shape[i];
  if (!same) {
    std::string want, got;
    for (auto v : ks.shape) want += std::to_string(v) + ",";
    for (int i = 0; i < ndim; ++i) got += std::to_string(shape[i]) + ",";
    return fail(SVK_ERR_INVALID, "size mismatch for %s: checkpoint [%s] vs model [%s]", key, got.c_str(), want.c_str());
  }
  if (ks.dead) return SVK_IGNORED;
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) n *= (size_t)shape[i];
  HostTensor t;
  t.shape.assign(shape, shape + ndim);
  t.data.assign(host_data, host_data + n);
  h->raw[key] = std::move(t);
  h->finalized = false;
  return SVK_OK;
}

extern "C" int svk_weight_status(const svk_handle* h, int* n_live, int* n_loaded) {
  if (!h) return fail(SVK_ERR_INVALID, "null handle");
  if (n_live) *n_live = h->n_live;
  if (n_loaded) {
    int n = 0;
    for (const auto& kv : h->raw) n += h->spec.at(kv.first).optional ? 0 : 1;
    *n_loaded = n;
  }
  return SVK_OK;
}

extern "C" int svk_finalize_weights(svk_handle* h) {
  if (!h) return fail(SVK_ERR_INVALID, "null handle");
  for (const auto& key : h->key_order)
    if (!h->spec[key].dead && !h->spec[key].optional && !h->raw.count(key))
      return fail(SVK_ERR_STATE, "missing key '%s'", key.c_str());
  const svk_config& c = h->cfg;
  const int half = c.inter_channels / 2;
  h->h_blob.clear();
  h->h_tcblob.clear();
  h->enc_in.clear(), h->enc_rs.clear(), h->ups.clear(), h->flows.clear(), h->resblocks.clear();
  h->encq_in.clear(), h->encq_rs.clear(), h->has_posterior = false;

  Folded f;
  SVK_TRY(fold_layer(h, "enc_p.pre_enc", &f));
  h->pre_enc = pack_plain(h, f);
  SVK_TRY(pack_wn(h, "enc_p.encoder", c.enc_layers, &h->enc_in, &h->enc_rs));
  SVK_TRY(fold_layer(h, "enc_p.proj", &f));
  h->proj = pack_plain(h, f);

  for (int fl = 0; fl < c.n_flows; ++fl) {
    FlowLayers L;
    // reverse pass applies [Flip, RCL_{n-1}, Flip, RCL_{n-2}, ...] (models.py:77-79): coupling fl has
    // seen (n_flows - fl) Flips.  Odd -> it operates on the channel-reversed view; we keep storage
    // fixed and reverse the 1x1 weights instead (pure index permutation, bit-exact on x0).
    L.orient = (c.n_flows - fl) & 1;
    const std::string p = "flow.flows." + std::to_string(2 * fl);
    SVK_TRY(fold_layer(h, p + ".pre", &f));
    if (L.orient) {
      const Folded& g = f;
      L.pre = pack_virtual(
          h, g.d1, g.d0, 1, [&](int o, int cc, int) { return g.w[(size_t)o * g.d1 + (half - 1 - cc)]; },
          [&](int o) { return g.b[o]; });
    } else {
      L.pre = pack_plain(h, f);
    }
    SVK_TRY(pack_wn(h, p + ".enc", c.flow_layers, &L.in, &L.rs));
    SVK_TRY(fold_layer(h, p + ".post", &f));
    {
      // x1 = (x1 - m) * mask  ==  (x1 + (-m)) * mask: negate weights/bias (exact), add as residual.
      const Folded& g = f;
      L.post = pack_virtual(
          h, g.d1, g.d0, 1, [&](int o, int cc, int) { return -g.w[(size_t)o * g.d1 + cc]; },
          [&](int o) { return -g.b[o]; });
      L.post_fwd = pack_plain(h, f);  // forward direction: x1 = m + x1 * mask (modules.py:336)
    }
    h->flows.push_back(std::move(L));
  }

  SVK_TRY(fold_layer(h, "dec.conv_pre", &f));
  h->conv_pre = pack_plain(h, f);
  for (int i = 0; i < c.n_upsamples; ++i) {
    SVK_TRY(fold_layer(h, "dec.ups." + std::to_string(i), &f));
    const int u = c.upsample_rates[i], k = c.upsample_kernel_sizes[i];
    // stride-8 upsamplers in front of image-only ResBlock1 stages write the operand image from their own epilogue
    const bool rmajor = h->tensor_engine() && h->planes() == 2 && h->img_stream == 2 && c.resblock_type != 2 && u == 8 &&
                        h->stage_channels(i) % 32 == 0;
    h->ups.push_back(pack_transposed(h, f, u, (k - u) / 2, rmajor));
  }
  for (int i = 0; i < c.n_upsamples; ++i)
    for (int j = 0; j < c.n_resblock_kernels; ++j) {
      ResBlock rb;
      rb.k = c.resblock_kernel_sizes[j];
      rb.C = h->stage_channels(i);
      const std::string p = "dec.resblocks." + std::to_string(i * c.n_resblock_kernels + j);
      if (c.resblock_type == 2) {
        rb.type = 2, rb.n = 2;
        for (int l = 0; l < rb.n; ++l) {
          rb.dil[l] = c.resblock_dilations[j][l];
          SVK_TRY(fold_layer(h, p + ".convs." + std::to_string(l), &f));
          rb.c1[l] = pack_plain(h, f);
        }
        h->resblocks.push_back(rb);
        continue;
      }
      for (int l = 0; l < SVK_RESBLOCK_PAIRS; ++l) {
        rb.dil[l] = c.resblock_dilations[j][l];
        SVK_TRY(fold_layer(h, p + ".convs1." + std::to_string(l), &f));
        rb.c1[l] = pack_plain(h, f);
        SVK_TRY(fold_layer(h, p + ".convs2." + std::to_string(l), &f));
        rb.c2[l] = pack_plain(h, f);
      }
      h->resblocks.push_back(rb);
    }
  SVK_TRY(fold_layer(h, "dec.conv_post", &f));
  h->conv_post = pack_plain(h, f);

  // PosteriorEncoder(spec_channels, inter, hidden, 5, 1, 16) (models.py:312): optional keys, all or nothing
  {
    bool all = c.wn_kernel == 5, any = false;
    for (const auto& key : h->key_order)
      if (h->spec[key].optional) {
        const bool have = h->raw.count(key) != 0;
        all = all && have, any = any || have;
      }
    if (any && !all && c.wn_kernel == 5) return fail(SVK_ERR_STATE, "enc_q.* keys loaded only partially");
    if (all) {
      SVK_TRY(fold_layer(h, "enc_q.pre", &f));
      {
        // Cin = spec_channels (513) is padded to the FFMA kernel's 8-channel chunk with zero weights; the kernel
        // reads input channels >= x_C as zeros
        const Folded& g = f;
        const int cin_pad = (g.d1 + 7) / 8 * 8;
        h->encq_pre = pack_virtual(
            h, cin_pad, g.d0, 1, [&](int o, int cc, int) { return cc < g.d1 ? g.w[(size_t)o * g.d1 + cc] : 0.f; },
            [&](int o) { return g.b[o]; }, /*want_tc=*/false);
      }
      SVK_TRY(pack_wn(h, "enc_q.enc", 16, &h->encq_in, &h->encq_rs));
      SVK_TRY(fold_layer(h, "enc_q.proj", &f));
      h->encq_proj = pack_plain(h, f);
      h->has_posterior = true;
    }
  }

  CUDA_TRY(cudaSetDevice(h->device));
  if (h->d_blob) cudaFree(h->d_blob);
  h->d_blob = nullptr;
  h->blob_floats = h->h_blob.size();
  CUDA_TRY(cudaMalloc(&h->d_blob, h->blob_floats * sizeof(float)));
  CUDA_TRY(cudaMemcpy(h->d_blob, h->h_blob.data(), h->blob_floats * sizeof(float), cudaMemcpyHostToDevice));
  h->h_blob.clear();
  h->h_blob.shrink_to_fit();
  if (h->d_tcblob) cudaFree(h->d_tcblob);
  h->d_tcblob = nullptr;
  if (!h->h_tcblob.empty()) {
    CUDA_TRY(cudaMalloc(&h->d_tcblob, h->h_tcblob.size() * sizeof(uint16_t)));
    CUDA_TRY(cudaMemcpy(h->d_tcblob, h->h_tcblob.data(), h->h_tcblob.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
  }
  h->h_tcblob.clear();
  h->h_tcblob.shrink_to_fit();
  h->finalized = true;
  return SVK_OK;
}

// ------------------------------------------------------------------------------ launch helpers
namespace {

struct Runner {
  svk_handle* h;
  cudaStream_t stream;
  int B;
  cudaError_t err = cudaSuccess;
  const PackedConv* cur = nullptr;  // layer of the ConvArgs most recently built by base()

  ConvArgs base(const PackedConv& pc, const float* x, int x_C, int x_ch_off, int x_stride, int Lin, int dil,
                int pad, int Lout, int y_stride) {
    cur = &pc;
    ConvArgs a;
    memset(&a, 0, sizeof(a));
    a.x = x, a.x_C = x_C, a.x_ch_off = x_ch_off, a.x_stride = x_stride, a.Lin = Lin;
    a.pre_slope = 1.0f;
    a.wp = h->d_blob + pc.w_off, a.bias = h->d_blob + pc.b_off;
    a.Cin = pc.Cin, a.Cout = pc.Cout, a.CoutPad = pc.CoutPad, a.K = pc.K;
    a.dil = dil, a.pad = pad, a.Lout = Lout, a.y_stride = y_stride;
    a.mode = MODE_STORE;
    a.split = 1 << 30;
    a.post_div = 1.0f;
    a.B = B;
    a.e[0].ch_sign = a.e[1].ch_sign = 1;
    return a;
  }
  // algorithmic work of one conv launch (MACs*2 over the convolution only)
  bool prof_open(int layer, const ConvArgs& a, int cout_logical, int64_t out_len, double macs) {
    if (!h->profiling || h->prof_records.size() >= h->prof_cap) return false;
    svk_launch_record r;
    memset(&r, 0, sizeof(r));
    r.layer = layer, r.cin = a.Cin, r.cout = cout_logical, r.k = a.K, r.dilation = a.dil, r.batch = B;
    r.length = out_len;
    r.flops = 2.0 * macs;
    // algorithmic traffic of this launch: input (fp32 tensor or operand image, 4 B/element either way with two planes),
    // weights, and per destination side the fp32 tensor and/or operand image written plus res / acc_in read.
    // dup = the image when the same values are ALSO written as fp32 (the tensor then lives in HBM twice).
    const double img = h->planes() * 0.5;  // an operand image is 2 B per plane per element = img * 4 B
    double bytes = 4.0 * ((double)B * a.Cin * a.Lin + (double)a.Cin * a.K * a.Cout), dup = 0.0;
    if (a.mode != MODE_STORE) {
      const double per = 4.0 * B * (double)cout_logical * out_len;
      const bool y = a.e[0].y != nullptr, sp = a.e[0].split != nullptr;
      bytes += per * ((y ? 1 : 0) + (sp ? img : 0));
      if (y && sp) dup += per * img;
    } else {
      for (int s = 0; s < 2; ++s) {
        const double n_s = s == 0 ? (a.split < a.Cout ? a.split : a.Cout) : (a.split < a.Cout ? a.Cout - a.split : 0);
        const double per = 4.0 * B * (double)out_len * n_s;
        bytes += per * ((a.e[s].y ? 1 : 0) + (a.e[s].split ? img : 0) + (a.e[s].res ? 1 : 0) + (a.e[s].res_img ? img : 0) +
                        (a.e[s].acc_in ? 1 : 0));
        if (a.e[s].y && a.e[s].split) dup += per * img;
      }
    }
    r.dup_bytes = dup;
    r.bytes = bytes;
    h->prof_records.push_back(r);
    cudaEventRecord(h->prof_events[2 * (h->prof_records.size() - 1)], stream);
    return true;
  }
  void prof_close(bool open) {
    if (open) cudaEventRecord(h->prof_events[2 * (h->prof_records.size() - 1) + 1], stream);
  }
  void run(const ConvArgs& a, int layer = SVK_LAYER_OTHER, const uint16_t* x_split = nullptr) {
    if (err != cudaSuccess) return;
    int cout_logical = a.Cout;
    int64_t out_len = a.Lout;
    double macs = (double)B * a.Lout * (double)a.Cout * a.Cin * a.K;
    if (a.mode == MODE_GATE) cout_logical = a.Cout / 2;
    if (a.mode == MODE_SHUFFLE) {  // counted as L_in * Cin * Cout * k
      cout_logical = a.Cout / a.shuf_s, out_len = a.shuf_Lout;
      macs = (double)B * a.Lin * (double)a.Cin * a.Cout * a.K;
    }
    const PackedConv* pc = cur;
    const bool use_tc = pc && pc->tc && a.wp == h->d_blob + pc->w_off;
    const bool open = prof_open(layer, a, cout_logical, out_len, macs);
    if (open) {
      h->prof_records.back().engine = use_tc ? 1 : 0;
      if (x_split) h->prof_records.back().bytes -= (4.0 - 2.0 * h->planes()) * B * (double)a.Cin * a.Lin;  // image input
    }
    if (use_tc) {
      ConvTcArgs ta;
      memset(&ta, 0, sizeof(ta));
      ta.c = a;
      ta.c.bias = h->d_blob + pc->tc_b_off;
      ta.wtc = h->d_tcblob + pc->tc_off;
      ta.unscale = pc->tc_unscale;
      ta.N = pc->tc_N;
      ta.planes = h->planes();
      ta.x_split = x_split;
      err = launch_conv_tc(ta, stream);
    } else if (x_split) {
      err = cudaErrorInvalidValue;  // operand images are a tcgen05-engine format
    } else {
      err = launch_conv_ffma(a, stream);
    }
    prof_close(open);
    h->launches++;
    static const bool sync_each = getenv("SVK_SYNC_LAUNCHES") != nullptr;  // debugging aid: name the launch that faults
    if (sync_each && err == cudaSuccess) {
      err = cudaStreamSynchronize(stream);
      if (err != cudaSuccess)
        fprintf(stderr, "libsvk: launch failed: layer %d Cin %d Cout %d K %d dil %d Lout %d mode %d tc %d image-in %d: %s\n", layer,
                a.Cin, a.Cout, a.K, a.dil, a.Lout, a.mode, (int)use_tc, x_split != nullptr, cudaGetErrorString(err));
    }
  }
  void note(cudaError_t e) {
    if (err == cudaSuccess) err = e;
    h->launches++;
  }
  // fp32 tensor -> operand image (split_image_kernel), recorded like a conv launch: all of its traffic is duplicate
  void split_image(const float* x, int C, int L, float slope, uint16_t* img) {
    if (err != cudaSuccess) return;
    bool open = false;
    if (h->profiling && h->prof_records.size() < h->prof_cap) {
      svk_launch_record r;
      memset(&r, 0, sizeof(r));
      r.layer = SVK_LAYER_SPLIT_IMAGE, r.cin = C, r.cout = C, r.k = 1, r.dilation = 1, r.batch = B, r.length = L, r.engine = 0;
      r.bytes = (4.0 + 2.0 * h->planes()) * B * (double)C * L;
      r.dup_bytes = r.bytes;
      h->prof_records.push_back(r);
      cudaEventRecord(h->prof_events[2 * (h->prof_records.size() - 1)], stream);
      open = true;
    }
    err = launch_split_image(x, B, C, L, slope, img, h->planes(), stream);
    prof_close(open);
    h->launches++;
  }

  // Fused conv pair of a ResBlock1 (conv_tc_pair.cu) with the same per-launch accounting as run().
  bool pair_fusable(const ResBlock& rb, int l) const {
    if (!h->fuse_pairs || h->planes() != 2) return false;
    // Fusing removes HBM traffic but not MMA work (and recomputes a (K-1)-row halo per item).  Measured at
    // 16 x 1024 frames: k = 3 pairs are HBM-bound and gain 25 %, k >= 7 pairs are bound by the per-MMA floor of
    // narrow tiles and lose 0-25 %  ->  the 3-tap blocks are fused, and (round 2: two issuers, eight epi1 warps) every
    // block of the C = 32 stage, where the k >= 7 pairs now break even per launch and save their xt round trip
    // ($SVK_FUSE_PAIRS=1: 3-tap blocks only, =2: every block).
    if (h->fuse_pairs_all == false && rb.k > 3 && !(h->fuse_pairs_c32 && rb.C == 32)) return false;
    return conv_tc_pair_supported(rb.C, rb.k, rb.dil[l]) && rb.c1[l].tc && rb.c2[l].tc && rb.c1[l].tc_N == rb.C &&
           rb.c2[l].tc_N == rb.C;
  }
  void run_pair(const ResBlock& rb, int l, ConvPairArgs p) {
    if (err != cudaSuccess) return;
    const PackedConv &c1 = rb.c1[l], &c2 = rb.c2[l];
    p.B = B, p.C = rb.C, p.K = rb.k, p.dil1 = rb.dil[l];
    p.w1 = h->d_tcblob + c1.tc_off, p.w2 = h->d_tcblob + c2.tc_off;
    p.bias1 = h->d_blob + c1.tc_b_off, p.bias2 = h->d_blob + c2.tc_b_off;
    p.unscale1 = c1.tc_unscale, p.unscale2 = c2.tc_unscale;
    p.xt_slope = 0.1f;
    bool open = false;
    if (h->profiling && h->prof_records.size() < h->prof_cap) {
      svk_launch_record r;
      memset(&r, 0, sizeof(r));
      r.layer = SVK_LAYER_RESBLOCK_PAIR, r.cin = rb.C, r.cout = rb.C, r.k = rb.k, r.dilation = rb.dil[l], r.batch = B;
      r.length = p.L, r.engine = 1;
      const double E = (double)B * rb.C * p.L;
      r.flops = 2.0 * 2.0 * E * rb.C * rb.k;  // both convs
      // the residual read counts unless it IS the input image (image-only stream: the same tensor, fetched once -- ncu on
      // such a pair: 543 MB read + 493 MB written for E = 537 MB, profiles/r2d_ncu_pair_c32_k3.txt)
      const bool res_is_input = p.res_img && p.res_img == p.x_img;
      r.bytes = 4.0 * E * (1 + ((p.res || p.res_img) && !res_is_input ? 1 : 0) + (p.acc_in ? 1 : 0) + (p.y ? 1 : 0) + (p.y_img ? 1 : 0)) +
                2.0 * 4.0 * rb.C * rb.C * rb.k;
      r.dup_bytes = (p.y && p.y_img) ? 4.0 * E : 0.0;
      h->prof_records.push_back(r);
      cudaEventRecord(h->prof_events[2 * (h->prof_records.size() - 1)], stream);
      open = true;
    }
    err = launch_conv_tc_pair(p, stream);
    prof_close(open);
    h->launches++;
  }

  // modules.WN.forward, g=None (modules.py:148-176).  x is updated in place, result in `out`.
  // With x_img / acts_img (tcgen05 engine) the convs read operand images by TMA: x_img must hold the image of
  // the incoming x; every res_skip epilogue rewrites it together with x, the gate epilogue writes acts only
  // as an image (nothing else reads it).
  void wn(const std::vector<PackedConv>& in, const std::vector<PackedConv>& rs, float* x, float* acts,
          float* out, const float* mask, int T, uint16_t* x_img = nullptr, uint16_t* acts_img = nullptr) {
    const int H = h->cfg.hidden_channels, n = (int)in.size(), k = h->cfg.wn_kernel;
    bool images = x_img && acts_img && h->tensor_engine() && H % 32 == 0;
    for (int i = 0; i < n && images; ++i) images = in[i].tc && rs[i].tc;
    // One launch per layer (wn_layer.cu): acts stays in shared memory, the x image ping-pongs between the two buffers.
    // The fused kernel gives one CTA a whole 128-frame tile (all N-tiles of both GEMMs: ~40 us per layer whatever the
    // batch); the two-launch form spreads the same work over 3x as many CTAs and scales down with the batch.  Measured
    // with tools/wn_bench.py (16-layer encoder, 1024 frames): B = 4: 0.64 vs 0.44 ms, B = 6: 0.65 vs 0.47, B = 8: 0.69 vs
    // 0.68, B = 16: 0.73 vs 0.95 -> fused from 64 tiles up.
    const int wn_tiles = B * ((T + 127) / 128);
    bool fused = images && (h->fuse_wn == 2 || (h->fuse_wn == 1 && wn_tiles >= 64));
    for (int i = 0; i < n && fused; ++i)
      fused = wn_layer_supported(H, k, in[i].tc_N, rs[i].tc_N, rs[i].Cout, h->planes()) && rs[i].Cout == (i < n - 1 ? 2 * H : H);
    if (fused) {
      // The whole stack in ONE launch when every tile can have its own SM (tiles synchronise with their neighbours inside
      // the kernel); otherwise one launch per layer.
      const bool stack = h->wn_stack && n > 1 && n <= WN_MAX_LAYERS && wn_tiles <= wn_stack_max_items() && wn_tiles <= 1024 && h->d_wn_flags;
      const int per = stack ? n : 1;
      for (int i0 = 0; i0 < n && err == cudaSuccess; i0 += per) {
        WnLayerArgs w;
        memset(&w, 0, sizeof(w));
        w.B = B, w.T = T, w.H = H, w.K = k, w.planes = h->planes();
        // layer i reads the image in buffer (i & 1) and writes buffer ((i + 1) & 1)
        w.img[0] = (i0 & 1) ? acts_img : x_img, w.img[1] = (i0 & 1) ? x_img : acts_img;
        w.x = x, w.out = out, w.mask = mask;
        w.l0 = i0, w.n_layers = per, w.n_total = n;
        w.N_in = in[i0].tc_N;
        w.flags = h->d_wn_flags;
        double flops = 0.0, bytes = 0.0, dup = 0.0;
        const double E = (double)B * H * T, img = h->planes() * 0.5;
        for (int j = 0; j < per; ++j) {
          const int i = i0 + j;
          WnLayerParams& lp = w.layer[j];
          lp.w_in = h->d_tcblob + in[i].tc_off, lp.bias_in = h->d_blob + in[i].tc_b_off, lp.unscale_in = in[i].tc_unscale;
          lp.w_rs = h->d_tcblob + rs[i].tc_off, lp.bias_rs = h->d_blob + rs[i].tc_b_off, lp.unscale_rs = rs[i].tc_unscale;
          lp.N_rs = rs[i].tc_N, lp.Cout_rs = rs[i].Cout;
          if (in[i].tc_N != w.N_in) err = cudaErrorInvalidValue;
          const bool first = i == 0, last = i == n - 1;
          flops += 2.0 * B * (double)T * ((double)2 * H * H * k + (double)rs[i].Cout * H);
          // x image in; fp32 x read + written and its new image (not on the last layer); out read (not on the first) + written; weights
          bytes += 4.0 * E * (img + (last ? 0.0 : 2.0 + img) + (first ? 1.0 : 2.0)) + 4.0 * ((double)2 * H * H * k + (double)rs[i].Cout * H);
          dup += last ? 0.0 : 4.0 * E * img;
        }
        if (err != cudaSuccess) break;
        bool open = false;
        if (h->profiling && h->prof_records.size() < h->prof_cap) {
          svk_launch_record r;
          memset(&r, 0, sizeof(r));
          r.layer = SVK_LAYER_WN_LAYER, r.cin = H, r.cout = per > 1 ? 2 * H * per : rs[i0].Cout, r.k = k, r.dilation = 1, r.batch = B, r.length = T, r.engine = 1;
          r.flops = flops, r.bytes = bytes, r.dup_bytes = dup;
          h->prof_records.push_back(r);
          cudaEventRecord(h->prof_events[2 * (h->prof_records.size() - 1)], stream);
          open = true;
        }
        // developer aid: $SVK_WN_TRACE=<file> dumps the clock64() stamps of the first multi-layer launch of the process
        static const char* trace_path = getenv("SVK_WN_TRACE");
        static bool traced = false;
        long long* d_trace = nullptr;
        const size_t trace_n = (size_t)wn_tiles * per * 32;
        if (trace_path && !traced && per > 1 && cudaMalloc(&d_trace, trace_n * sizeof(long long)) == cudaSuccess) {
          cudaMemsetAsync(d_trace, 0, trace_n * sizeof(long long), stream);
          w.trace = d_trace;
        }
        err = launch_wn_layers(w, stream);
        if (per > 1 && (err == cudaErrorCooperativeLaunchTooLarge || err == cudaErrorLaunchOutOfResources)) {
          // the cooperative grid does not fit THIS context right now (SMs reserved elsewhere: MPS limits, green contexts):
          // nothing has run; forget the whole-stack form for this handle and do the stack layer by layer
          cudaGetLastError();
          if (open) h->prof_records.pop_back();
          if (d_trace) cudaFree(d_trace);
          h->wn_stack = 0, err = cudaSuccess;
          wn(in, rs, x, acts, out, mask, T, x_img, acts_img);
          return;
        }
        if (d_trace) {
          traced = true;
          std::vector<long long> host(trace_n);
          cudaStreamSynchronize(stream);
          cudaMemcpy(host.data(), d_trace, trace_n * sizeof(long long), cudaMemcpyDeviceToHost);
          cudaFree(d_trace);
          if (FILE* f = fopen(trace_path, "w")) {
            fprintf(f, "tile,layer");
            for (int e = 0; e < 32; ++e) fprintf(f, ",e%d", e);
            fprintf(f, "\n");
            for (int t = 0; t < wn_tiles; ++t)
              for (int l = 0; l < per; ++l) {
                fprintf(f, "%d,%d", t, l);
                for (int e = 0; e < 32; ++e) fprintf(f, ",%lld", host[((size_t)t * per + l) * 32 + e]);
                fprintf(f, "\n");
              }
            fclose(f);
          }
        }
        prof_close(open);
        h->launches++;
      }
      return;
    }
    for (int i = 0; i < n; ++i) {
      ConvArgs a = base(in[i], x, H, 0, T, T, 1, (k - 1) / 2, T, T);
      a.mode = MODE_GATE;
      a.e[0].C = H;
      if (images) a.e[0].y = nullptr, a.e[0].split = acts_img, a.e[0].split_slope = 1.0f;
      else a.e[0].y = acts;
      run(a, SVK_LAYER_WN_IN, images ? x_img : nullptr);
      ConvArgs r = base(rs[i], acts, H, 0, T, T, 1, 0, T, T);
      r.out_mask = mask, r.mask_stride = T;
      if (i < n - 1) {
        r.split = H;
        r.e[0].res = x, r.e[0].y = x, r.e[0].C = H, r.e[0].use_mask = 1;       // x = (x + res) * mask
        if (images) r.e[0].split = x_img, r.e[0].split_slope = 1.0f;            // ... and its image for layer i+1
        r.e[1].acc_in = i ? out : nullptr, r.e[1].y = out, r.e[1].C = H;      // out += skip
      } else {
        r.e[0].acc_in = i ? out : nullptr, r.e[0].y = out, r.e[0].C = H, r.e[0].use_mask = 1;  // (out + rs) * mask
      }
      run(r, SVK_LAYER_WN_RES_SKIP, images ? acts_img : nullptr);
    }
  }
  bool wn_uses_images(const std::vector<PackedConv>& in, const std::vector<PackedConv>& rs) const {
    if (!h->tensor_engine() || h->cfg.hidden_channels % 32) return false;
    for (size_t i = 0; i < in.size(); ++i)
      if (!in[i].tc || !rs[i].tc) return false;
    return true;
  }

  // True when every conv of the block runs on the tcgen05 engine with operand-image I/O.
  bool resblock_uses_images(const ResBlock& rb) const {
    if (!h->tensor_engine() || rb.C % 16) return false;
    for (int l = 0; l < rb.n; ++l)
      if (!rb.c1[l].tc || (rb.type == 1 && !rb.c2[l].tc)) return false;
    return true;
  }

  // ResBlock1.forward on the tcgen05 engine.  Activations that feed a conv live in HBM as fp16 hi/lo
  // operand images of leaky_relu(x, 0.1) (written by the producing epilogue, loaded by TMA); the fp32
  // tensors are kept only where the residual adds need them (modules.py:220).  x_img = image of x.
  void resblock_images(const ResBlock& rb, const float* x, const uint16_t* x_img, uint16_t* xt_img, float* cur,
                       uint16_t* cur_img, float* dst, const float* acc_in, float post_div, int L,
                       uint16_t* dst_img = nullptr) {
    const int C = rb.C;
    if (rb.type == 2) {
      // ResBlock2 (modules.py:243-252): x = conv_l(leaky_relu(x), dilation d_l) + x, twice.  Each conv reads the operand
      // image of leaky_relu(x) and writes the next one (another buffer: its neighbours' halos are still being read);
      // the residual is the fp32 x, or -- on the wide, tensor-bound stages -- rebuilt from the input image itself.
      const bool img_stream = h->planes() == 2 && C >= 128;
      for (int l = 0; l < rb.n; ++l) {
        const bool last = l == rb.n - 1;
        const float* src = l == 0 ? x : cur;
        const uint16_t* src_img = l == 0 ? x_img : ((l & 1) ? cur_img : xt_img);
        const int d = rb.dil[l];
        ConvArgs a = base(rb.c1[l], src, C, 0, L, L, d, (rb.k * d - d) / 2, L, L);
        a.pre_slope = 0.1f;
        a.e[0].C = C;
        if (img_stream && (l > 0 || x_img)) a.e[0].res_img = src_img, a.e[0].res_slope = 0.1f;
        else a.e[0].res = src;
        if (!last) {
          a.e[0].y = img_stream ? nullptr : cur;
          a.e[0].split = (l & 1) ? xt_img : cur_img, a.e[0].split_slope = 0.1f;
        } else {
          a.e[0].y = dst, a.e[0].acc_in = acc_in, a.post_div = post_div;
          a.e[0].split = dst_img, a.e[0].split_slope = 0.1f;
        }
        run(a, SVK_LAYER_RESBLOCK_CONV1, src_img);
      }
      return;
    }
    bool fuse = true;
    for (int l = 0; l < SVK_RESBLOCK_PAIRS; ++l) fuse = fuse && pair_fusable(rb, l);
    if (fuse) {
      // narrow stages: one launch per pair, xt never leaves the SM (conv_tc_pair.cu).  The pair kernel reads its
      // input image with a halo, so images ping-pong: x_img -> cur_img -> xt_img (free here) -> dst_img.
      for (int l = 0; l < SVK_RESBLOCK_PAIRS; ++l) {
        ConvPairArgs p;
        memset(&p, 0, sizeof(p));
        p.L = L;
        p.x_img = l == 0 ? x_img : (l == 1 ? cur_img : xt_img);
        if (h->img_stream >= 2) {  // residual stream as images only: the pair's own input image is its residual
          p.res_img = p.x_img, p.res_slope = 0.1f;
        } else {
          p.res = l == 0 ? x : cur;  // fp32 residual stream: same element read then written by one thread
        }
        p.post_div = 1.0f;
        if (l < SVK_RESBLOCK_PAIRS - 1) {
          p.y = h->img_stream >= 2 ? nullptr : cur;
          p.y_img = l == 0 ? cur_img : xt_img, p.y_slope = 0.1f;
        } else {
          p.y = dst, p.acc_in = acc_in, p.post_div = post_div;
          p.y_img = dst_img, p.y_slope = 0.1f;
        }
        run_pair(rb, l, p);
      }
      return;
    }
    for (int l = 0; l < SVK_RESBLOCK_PAIRS; ++l) {
      const float* src = l == 0 ? x : cur;
      const uint16_t* src_img = l == 0 ? x_img : cur_img;
      const int d = rb.dil[l];
      ConvArgs a = base(rb.c1[l], src, C, 0, L, L, d, (rb.k * d - d) / 2, L, L);
      a.pre_slope = 0.1f;
      a.e[0].y = nullptr, a.e[0].C = C;
      a.e[0].split = xt_img, a.e[0].split_slope = 0.1f;  // xt is only ever read through leaky_relu by c2
      run(a, SVK_LAYER_RESBLOCK_CONV1, src_img);
      ConvArgs b = base(rb.c2[l], nullptr, C, 0, L, L, 1, (rb.k - 1) / 2, L, L);
      b.pre_slope = 0.1f;
      b.e[0].C = C;
      // x + xt (modules.py:220).  The residual stream between (and, since the end of round 2, into) the pairs of a block lives
      // in HBM as its operand image only: the epilogue rebuilds x = hi + lo * 2^-11 (leaky_relu inverted, ~2^-22 relative)
      // instead of reading a second, fp32 copy.  Round 1 kept the fp32 copy on the narrow layers (their conv2 epilogues are
      // issue-bound and the image form costs instructions: per launch the k = 11 blocks of C <= 64 are 4-10 % slower with
      // it, profiles/r2_img_stream_ab.txt) -- but the step is power-capped, and the variant that moves the fewest bytes
      // wins on the step as a whole (profiles/r2_ab_img_stream.log: 34.8 -> 34.4 ms).  A block's output is always fp32.
      const int ims = h->img_stream;
      const bool img_stream = h->planes() == 2 && (ims == 1 || ims == 2 || ((ims == 0 || ims == 3) && (C >= 128 || rb.k == 7)) ||
                                                     (ims < 0 && (C >= 256 || (C >= 128 && rb.k >= 7))));
      if ((l > 0 || ims == 2) && img_stream) b.e[0].res_img = src_img, b.e[0].res_slope = 0.1f;
      else b.e[0].res = src;
      if (l < SVK_RESBLOCK_PAIRS - 1) {
        b.e[0].y = img_stream ? nullptr : cur;
        b.e[0].split = cur_img, b.e[0].split_slope = 0.1f;
      } else {
        b.e[0].y = dst, b.e[0].acc_in = acc_in, b.post_div = post_div;
        b.e[0].split = dst_img, b.e[0].split_slope = 0.1f;  // next upsampler reads leaky_relu(x, 0.1) (models.py:147)
      }
      run(b, SVK_LAYER_RESBLOCK_CONV2, xt_img);
    }
  }

  // ResBlock1.forward (modules.py:210-223).  `dst` receives x_out (+ acc_in) / post_div.
  void resblock(const ResBlock& rb, const float* x, float* xt, float* cur, float* dst, const float* acc_in,
                float post_div, int L) {
    const int C = rb.C;
    if (rb.type == 2) {  // ResBlock2: x = conv_l(leaky_relu(x)) + x; ping-pong cur / xt so no conv runs in place
      for (int l = 0; l < rb.n; ++l) {
        const bool last = l == rb.n - 1;
        const float* src = l == 0 ? x : ((l & 1) ? cur : xt);
        const int d = rb.dil[l];
        ConvArgs a = base(rb.c1[l], src, C, 0, L, L, d, (rb.k * d - d) / 2, L, L);
        a.pre_slope = 0.1f;
        a.e[0].res = src, a.e[0].C = C;
        if (!last) {
          a.e[0].y = (l & 1) ? xt : cur;
        } else {
          a.e[0].y = dst, a.e[0].acc_in = acc_in, a.post_div = post_div;
        }
        run(a, SVK_LAYER_RESBLOCK_CONV1);
      }
      return;
    }
    for (int l = 0; l < SVK_RESBLOCK_PAIRS; ++l) {
      const float* src = l == 0 ? x : cur;
      const int d = rb.dil[l];
      ConvArgs a = base(rb.c1[l], src, C, 0, L, L, d, (rb.k * d - d) / 2, L, L);
      a.pre_slope = 0.1f;
      a.e[0].y = xt, a.e[0].C = C;
      run(a, SVK_LAYER_RESBLOCK_CONV1);
      ConvArgs b = base(rb.c2[l], xt, C, 0, L, L, 1, (rb.k - 1) / 2, L, L);
      b.pre_slope = 0.1f;
      b.e[0].res = src, b.e[0].C = C;
      if (l < SVK_RESBLOCK_PAIRS - 1) {
        b.e[0].y = cur;
      } else {
        b.e[0].y = dst, b.e[0].acc_in = acc_in, b.post_div = post_div;
      }
      run(b, SVK_LAYER_RESBLOCK_CONV2);
    }
  }
};

struct DecoderPlan {
  size_t buf_floats;  // each of the rotating stage buffers (an operand image has the same byte size)
  int n_bufs;         // 4 fp32 tensors + 3 operand images on the tcgen05 engine
};

DecoderPlan plan_decoder(const svk_handle* h, int B, int L) {
  size_t mx = (size_t)h->cfg.upsample_initial_channel * L;
  int64_t len = L;
  for (int i = 0; i < h->cfg.n_upsamples; ++i) {
    len *= h->cfg.upsample_rates[i];
    const size_t v = (size_t)h->stage_channels(i) * len;
    if (v > mx) mx = v;
  }
  return DecoderPlan{align_up(mx * (size_t)B, 64), h->tensor_engine() ? 7 : 4};
}

// Generator.forward (models.py:141-160).  z rows have stride z_stride; in_mask (optional) fuses
// the `(z * x_mask)[:, :, :max_len]` of models.py:338.
void run_decoder(Runner& R, const float* z, int z_stride, const float* in_mask, int L, float* o, float* ws) {
  svk_handle* h = R.h;
  const svk_config& c = h->cfg;
  const DecoderPlan plan = plan_decoder(h, R.B, L);
  float* buf[4] = {ws, ws + plan.buf_floats, ws + 2 * plan.buf_floats, ws + 3 * plan.buf_floats};
  uint16_t* img[3] = {nullptr, nullptr, nullptr};
  if (plan.n_bufs == 7)
    for (int i = 0; i < 3; ++i) img[i] = reinterpret_cast<uint16_t*>(ws + (4 + i) * plan.buf_floats);
  const int U = c.upsample_initial_channel;
  int pi = -1;  // index of the image buffer holding leaky_relu(prev, 0.1) when the producer wrote one, else -1
  {
    ConvArgs a = R.base(h->conv_pre, z, c.inter_channels, 0, z_stride, L, 1, 3, L, L);
    a.in_mask = in_mask, a.mask_stride = z_stride;
    a.e[0].y = buf[1], a.e[0].C = U;
    if (img[0] && h->conv_pre.tc && h->ups[0].tc) a.e[0].split = img[0], a.e[0].split_slope = 0.1f, pi = 0;
    R.run(a, SVK_LAYER_CONV_PRE);
  }
  const float* prev = buf[1];
  int prevC = U, len = L;
  for (int i = 0; i < c.n_upsamples; ++i) {
    const PackedConv& up = h->ups[i];
    const int C = h->stage_channels(i);
    const int Lout = (len - 1) * up.s - 2 * up.p + up.k;
    float* X = buf[0];
    float* XT = buf[1];
    float* CUR = buf[2];
    float* XS = buf[3];
    if (prev == XS) {  // previous stage summed into buf[3]; rotate so it is not overwritten early
      XS = buf[1], XT = buf[3];
    }
    const int nk = c.n_resblock_kernels;
    bool images = img[0] != nullptr;
    for (int j = 0; j < nk; ++j) images = images && R.resblock_uses_images(h->resblocks[i * nk + j]);
    // Image buffer roles of this stage.  A stride-2 upsampler writes the stage's x image from its own epilogue
    // (it may not overwrite the image it is reading: next buffer); otherwise split_image_kernel makes it from X
    // after the upsampler has finished with its input image (same buffer is fine).
    const bool up_writes_img = images && up.tc && ((up.s == 2 && (up.Cout % 16) == 0) || up.rmajor) && C % 32 == 0;
    const int xi = up_writes_img ? (pi < 0 ? 0 : (pi + 1) % 3) : (pi < 0 ? 0 : pi);
    uint16_t* x_img = img[xi];
    uint16_t* xt_img = img[(xi + 1) % 3];
    uint16_t* cur_img = img[(xi + 2) % 3];
    // ups[i](leaky_relu(x, 0.1)) (models.py:147-149)
    {
      const int nq = (Lout - 1 + up.p) / up.s + 1;
      ConvArgs a = R.base(up, prev, prevC, 0, len, len, 1, up.pad_virtual, nq, Lout);
      a.pre_slope = 0.1f;
      a.mode = MODE_SHUFFLE;
      a.shuf_s = up.s, a.shuf_p = up.p, a.shuf_Lout = Lout;
      a.shuf_rmajor = (up.tc && up.rmajor) ? 1 : 0;  // the channel order of the tcgen05 weight image
      a.e[0].y = X, a.e[0].C = C;
      if (up_writes_img) a.e[0].split = x_img, a.e[0].split_slope = 0.1f;
      // With the residual stream as images only (img_stream 2) no ResBlock1 reads the fp32 x of the stage: a stride-2
      // upsampler that writes the image from its own epilogue (TMA-fed: the lean polyphase epilogue) then writes nothing else
      bool x_fp32_dead = up_writes_img && up.tc && h->img_stream == 2 && h->planes() == 2 &&
                         (up.rmajor || (pi >= 0 && up.s == 2 && up.p == 1 && (Lout & 1) == 0));
      for (int j = 0; j < nk; ++j) x_fp32_dead = x_fp32_dead && h->resblocks[i * nk + j].type == 1;
      if (x_fp32_dead) a.e[0].y = nullptr;
      R.run(a, SVK_LAYER_UPSAMPLE, (up.tc && pi >= 0) ? img[pi] : nullptr);
    }
    // xs = sum_j resblock_j(x); x = xs / num_kernels (models.py:150-155)
    if (images && !up_writes_img) R.split_image(X, C, Lout, 0.1f, x_img);  // shared by the nk blocks
    const bool next_wants_img = images && i + 1 < c.n_upsamples && h->ups[i + 1].tc && C % 32 == 0;
    for (int j = 0; j < nk; ++j) {
      const ResBlock& rb = h->resblocks[i * nk + j];
      const float* acc = j ? XS : nullptr;
      const float div = j == nk - 1 ? (float)nk : 1.0f;
      // the last block's final epilogue also writes the next upsampler's operand image (over x_img, dead by then)
      uint16_t* next_img = (next_wants_img && j == nk - 1) ? x_img : nullptr;
      // ... and when the next upsampler reads that image (TMA-fed) nobody reads the stage's fp32 output: image only
      float* const out_fp32 = (next_img && h->img_stream == 2 && h->planes() == 2 && rb.type == 1) ? nullptr : XS;
      if (images) R.resblock_images(rb, X, x_img, xt_img, CUR, cur_img, out_fp32, acc, div, Lout, next_img);
      else R.resblock(rb, X, XT, CUR, XS, acc, div, Lout);
    }
    pi = next_wants_img ? xi : -1;
    prev = XS, prevC = C, len = Lout;
  }
  // tanh(conv_post(leaky_relu(x)))  -- default slope 0.01 (models.py:156-158)
  {
    ConvArgs a = R.base(h->conv_post, prev, prevC, 0, len, len, 1, 3, len, len);
    a.pre_slope = 0.01f;
    a.act_tanh = 1;
    a.range_flag = h->d_range_flag;
    a.e[0].y = o, a.e[0].C = 1;
    R.run(a, SVK_LAYER_CONV_POST);
  }
}

struct InferPlan {
  size_t mask, hbuf, acts, out, ximg, aimg, lat[4], dec, total;  // float offsets (operand images: same bytes as fp32)
};

InferPlan plan_infer(const svk_handle* h, int B, int T, int max_len) {
  const svk_config& c = h->cfg;
  const int Tp = max_len > 0 && max_len < T ? max_len : T;
  InferPlan p;
  size_t off = 0;
  auto take = [&](size_t n) {
    size_t o = off;
    off = align_up(off + n, 64);
    return o;
  };
  p.mask = take((size_t)B * T);
  p.hbuf = take((size_t)B * c.hidden_channels * T);
  p.acts = take((size_t)B * c.hidden_channels * T);
  p.out = take((size_t)B * c.hidden_channels * T);
  p.ximg = take((size_t)B * c.hidden_channels * T);
  p.aimg = take((size_t)B * c.hidden_channels * T);
  for (int i = 0; i < 4; ++i) p.lat[i] = take((size_t)B * c.inter_channels * T);
  {
    const DecoderPlan dp = plan_decoder(h, B, Tp);
    p.dec = take(dp.n_bufs * dp.buf_floats);
  }
  p.total = off;
  return p;
}

int check_ready(const svk_handle* h, const char* who) {
  if (!h) return fail(SVK_ERR_INVALID, "%s: null handle", who);
  if (!h->finalized) return fail(SVK_ERR_STATE, "%s: weights not finalized (svk_finalize_weights)", who);
  return SVK_OK;
}

void run_mel_encoder(Runner& R, const float* mel, const float* mask, int T, float* hbuf, float* acts, float* out,
                     float* m, float* logs, uint16_t* x_img = nullptr, uint16_t* acts_img = nullptr) {
  svk_handle* h = R.h;
  const svk_config& c = h->cfg;
  const int H = c.hidden_channels, C = c.inter_channels;
  // x = pre_enc(mel); encoder(x * x_mask, x_mask) (models.py:38-42)
  ConvArgs a = R.base(h->pre_enc, mel, c.n_mel, 0, T, T, 1, 0, T, T);
  a.out_mask = mask, a.mask_stride = T;
  a.e[0].y = hbuf, a.e[0].C = H, a.e[0].use_mask = 1;
  R.run(a, SVK_LAYER_PRE_ENC);
  const bool images = x_img && acts_img && R.wn_uses_images(h->enc_in, h->enc_rs);
  if (images) R.split_image(hbuf, H, T, 1.0f, x_img);  // pre_enc runs on the FFMA kernel
  R.wn(h->enc_in, h->enc_rs, hbuf, acts, out, mask, T, images ? x_img : nullptr, images ? acts_img : nullptr);
  // stats = proj(x) * x_mask; m, logs = split(stats) (models.py:44-46)
  ConvArgs p = R.base(h->proj, out, H, 0, T, T, 1, 0, T, T);
  p.out_mask = mask, p.mask_stride = T;
  p.split = C;
  p.e[0].y = m, p.e[0].C = C, p.e[0].use_mask = 1;
  p.e[1].y = logs, p.e[1].C = C, p.e[1].use_mask = 1;
  R.run(p, SVK_LAYER_PROJ);
}

// ResidualCouplingBlock.forward(reverse=True) in place on z (models.py:77-79, modules.py:324-343).
void run_flow_reverse(Runner& R, float* z, const float* mask, int T, float* hbuf, float* acts, float* out,
                      uint16_t* x_img = nullptr, uint16_t* acts_img = nullptr) {
  svk_handle* h = R.h;
  const svk_config& c = h->cfg;
  const int H = c.hidden_channels, C = c.inter_channels, half = C / 2;
  for (int fl = c.n_flows - 1; fl >= 0; --fl) {
    const FlowLayers& L = h->flows[fl];
    // h = pre(x0) * mask : x0 is the first half of the (possibly reversed) view
    ConvArgs a = R.base(L.pre, z, C, L.orient ? half : 0, T, T, 1, 0, T, T);
    a.out_mask = mask, a.mask_stride = T;
    a.e[0].y = hbuf, a.e[0].C = H, a.e[0].use_mask = 1;
    const bool images = x_img && acts_img && R.wn_uses_images(L.in, L.rs);
    if (images && L.pre.tc) a.e[0].split = x_img, a.e[0].split_slope = 1.0f;
    R.run(a, SVK_LAYER_FLOW_PRE);
    if (images && !L.pre.tc) R.split_image(hbuf, H, T, 1.0f, x_img);
    R.wn(L.in, L.rs, hbuf, acts, out, mask, T, images ? x_img : nullptr, images ? acts_img : nullptr);
    // x1 = (x1 - post(h) * mask) * mask, written over x1's storage channels
    ConvArgs p = R.base(L.post, out, H, 0, T, T, 1, 0, T, T);
    p.out_mask = mask, p.mask_stride = T;
    p.e[0].res = z, p.e[0].y = z, p.e[0].C = C, p.e[0].use_mask = 1;
    if (L.orient) {
      p.e[0].ch_off = half - 1, p.e[0].ch_sign = -1;
    } else {
      p.e[0].ch_off = half, p.e[0].ch_sign = 1;
    }
    R.run(p, SVK_LAYER_FLOW_POST);
  }
}

// PosteriorEncoder.forward, g=None (models.py:103-110): z = (m + eps * exp(logs)) * mask.
void run_posterior_encoder(Runner& R, const float* spec, const float* eps, const float* mask, int T, float* hbuf,
                           float* acts, float* out, float* z, float* m, float* logs, uint16_t* x_img, uint16_t* acts_img) {
  svk_handle* h = R.h;
  const svk_config& c = h->cfg;
  const int H = c.hidden_channels, C = c.inter_channels;
  ConvArgs a = R.base(h->encq_pre, spec, c.spec_channels, 0, T, T, 1, 0, T, T);  // x = pre(x) * x_mask
  a.out_mask = mask, a.mask_stride = T;
  a.e[0].y = hbuf, a.e[0].C = H, a.e[0].use_mask = 1;
  R.run(a, SVK_LAYER_PRE_ENC);
  const bool images = x_img && acts_img && R.wn_uses_images(h->encq_in, h->encq_rs);
  if (images) R.split_image(hbuf, H, T, 1.0f, x_img);
  R.wn(h->encq_in, h->encq_rs, hbuf, acts, out, mask, T, images ? x_img : nullptr, images ? acts_img : nullptr);
  ConvArgs p = R.base(h->encq_proj, out, H, 0, T, T, 1, 0, T, T);  // stats = proj(x) * x_mask
  p.out_mask = mask, p.mask_stride = T;
  p.split = C;
  p.e[0].y = m, p.e[0].C = C, p.e[0].use_mask = 1;
  p.e[1].y = logs, p.e[1].C = C, p.e[1].use_mask = 1;
  R.run(p, SVK_LAYER_PROJ);
  R.note(launch_posterior_sample(m, logs, eps, mask, R.B, C, T, z, R.stream));
}

// ResidualCouplingBlock.forward(reverse=False) in place on z (models.py:73-76, modules.py:324-339): RCL0, Flip, RCL1, ...
// Coupling fl has seen fl Flips; with an even number of flows that is the orientation the reverse pass packed
// (n_flows - fl), so the same folded pre weights serve both directions and storage ends un-flipped.
void run_flow_forward(Runner& R, float* z, const float* mask, int T, float* hbuf, float* acts, float* out,
                      uint16_t* x_img = nullptr, uint16_t* acts_img = nullptr) {
  svk_handle* h = R.h;
  const svk_config& c = h->cfg;
  const int H = c.hidden_channels, C = c.inter_channels, half = C / 2;
  for (int fl = 0; fl < c.n_flows; ++fl) {
    const FlowLayers& L = h->flows[fl];
    ConvArgs a = R.base(L.pre, z, C, L.orient ? half : 0, T, T, 1, 0, T, T);
    a.out_mask = mask, a.mask_stride = T;
    a.e[0].y = hbuf, a.e[0].C = H, a.e[0].use_mask = 1;
    const bool images = x_img && acts_img && R.wn_uses_images(L.in, L.rs);
    if (images && L.pre.tc) a.e[0].split = x_img, a.e[0].split_slope = 1.0f;
    R.run(a, SVK_LAYER_FLOW_PRE);
    if (images && !L.pre.tc) R.split_image(hbuf, H, T, 1.0f, x_img);
    R.wn(L.in, L.rs, hbuf, acts, out, mask, T, images ? x_img : nullptr, images ? acts_img : nullptr);
    // x1 = m + x1 * exp(0) * mask = (post(h) + x1) * mask, written over x1's storage channels
    ConvArgs p = R.base(L.post_fwd, out, H, 0, T, T, 1, 0, T, T);
    p.out_mask = mask, p.mask_stride = T;
    p.e[0].res = z, p.e[0].y = z, p.e[0].C = C, p.e[0].use_mask = 1;
    if (L.orient) {
      p.e[0].ch_off = half - 1, p.e[0].ch_sign = -1;
    } else {
      p.e[0].ch_off = half, p.e[0].ch_sign = 1;
    }
    R.run(p, SVK_LAYER_FLOW_POST);
  }
}

}  // namespace

// --------------------------------------------------------------------------------- the hot path
extern "C" size_t svk_workspace_bytes(const svk_handle* h, int B, int T, int max_len) {
  if (!h || B <= 0 || T <= 0) return 0;
  return plan_infer(h, B, T, max_len).total * sizeof(float);
}

extern "C" int64_t svk_last_launch_count(const svk_handle* h) { return h ? h->launches : 0; }

extern "C" int svk_profile_begin(svk_handle* h, int max_records) {
  if (!h || max_records <= 0) return fail(SVK_ERR_INVALID, "svk_profile_begin: bad argument");
  CUDA_TRY(cudaSetDevice(h->device));
  while (h->prof_events.size() < 2 * (size_t)max_records) {
    cudaEvent_t e;
    CUDA_TRY(cudaEventCreate(&e));
    h->prof_events.push_back(e);
  }
  h->prof_records.clear();
  h->prof_cap = (size_t)max_records;
  h->profiling = true;
  return SVK_OK;
}

extern "C" int svk_profile_end(svk_handle* h, svk_launch_record* out, int max_records, int* n_records) {
  if (!h || !n_records) return fail(SVK_ERR_INVALID, "svk_profile_end: bad argument");
  h->profiling = false;
  CUDA_TRY(cudaSetDevice(h->device));
  const size_t n = h->prof_records.size();
  for (size_t i = 0; i < n; ++i) {
    CUDA_TRY(cudaEventSynchronize(h->prof_events[2 * i + 1]));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, h->prof_events[2 * i], h->prof_events[2 * i + 1]));
    h->prof_records[i].ms = ms;
    float gap = 0.f;
    if (i > 0) CUDA_TRY(cudaEventElapsedTime(&gap, h->prof_events[2 * i - 1], h->prof_events[2 * i]));
    h->prof_records[i].gap_ms = gap;
  }
  int m = 0;
  for (size_t i = 0; i < n && out && m < max_records; ++i) out[m++] = h->prof_records[i];
  *n_records = (int)n;
  return SVK_OK;
}

extern "C" int svk_infer(svk_handle* h, const float* mel, const int64_t* lengths, const float* eps,
                         float noise_scale, int B, int T, int max_len, float* o, float* x_mask, float* z,
                         float* z_p, float* m_p, float* logs_p, void* workspace, size_t workspace_bytes,
                         void* stream) {
  SVK_TRY(check_ready(h, "svk_infer"));
  if (!mel || !lengths || !eps || !o) return fail(SVK_ERR_INVALID, "svk_infer: null mel/lengths/eps/o");
  if (B <= 0 || T <= 0) return fail(SVK_ERR_INVALID, "svk_infer: B and T must be positive");
  const InferPlan p = plan_infer(h, B, T, max_len);
  if (!workspace || workspace_bytes < p.total * sizeof(float))
    return fail(SVK_ERR_WORKSPACE, "svk_infer: workspace %zu < %zu bytes", workspace_bytes, p.total * sizeof(float));
  CUDA_TRY(cudaSetDevice(h->device));
  const svk_config& c = h->cfg;
  float* ws = (float*)workspace;
  float* mask = x_mask ? x_mask : ws + p.mask;
  float* lat_z = z ? z : ws + p.lat[0];
  float* lat_zp = z_p ? z_p : nullptr;
  float* lat_m = m_p ? m_p : ws + p.lat[2];
  float* lat_logs = logs_p ? logs_p : ws + p.lat[3];
  const int Tp = max_len > 0 && max_len < T ? max_len : T;

  h->launches = 0;
  Runner R{h, (cudaStream_t)stream, B};
  R.note(launch_sequence_mask(lengths, B, T, mask, R.stream));
  uint16_t* ximg = reinterpret_cast<uint16_t*>(ws + p.ximg);
  uint16_t* aimg = reinterpret_cast<uint16_t*>(ws + p.aimg);
  run_mel_encoder(R, mel, mask, T, ws + p.hbuf, ws + p.acts, ws + p.out, lat_m, lat_logs, ximg, aimg);
  // z_p = m_p + eps * exp(logs_p) * noise_scale (models.py:336)
  R.note(launch_sample(lat_m, lat_logs, eps, noise_scale, lat_zp, lat_z, (int64_t)B * c.inter_channels * T, R.stream));
  run_flow_reverse(R, lat_z, mask, T, ws + p.hbuf, ws + p.acts, ws + p.out, ximg, aimg);
  if (c.n_flows & 1) {  // odd number of Flips leaves storage reversed: materialise the last Flip
    float* tmp = ws + p.lat[1];
    R.note(launch_flip(lat_z, B, c.inter_channels, T, tmp, R.stream));
    if (R.err == cudaSuccess)
      R.err = cudaMemcpyAsync(lat_z, tmp, sizeof(float) * (size_t)B * c.inter_channels * T, cudaMemcpyDeviceToDevice, R.stream);
  }
  // o = dec((z * x_mask)[:, :, :max_len]) (models.py:338)
  run_decoder(R, lat_z, T, mask, Tp, o, ws + p.dec);
  if (R.err != cudaSuccess) return fail(SVK_ERR_CUDA, "svk_infer: launch failed: %s", cudaGetErrorString(R.err));
  return SVK_OK;
}

static const char* kRangeMessage =
    "non-finite waveform sample: an activation left the fp16 operand range of the tensor-core engine (|x| > 65504) or the "
    "inputs / weights are not finite; rescale the checkpoint or use engine=\"fp32\"";

extern "C" int svk_check_range(svk_handle* h, void* stream) {
  if (!h) return fail(SVK_ERR_INVALID, "svk_check_range: null handle");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  int flag = 0;
  CUDA_TRY(cudaMemcpyAsync(&flag, h->d_range_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  if (flag) {
    CUDA_TRY(cudaMemsetAsync(h->d_range_flag, 0, sizeof(int), s));
    return fail(SVK_ERR_RANGE, "%s", kRangeMessage);
  }
  return SVK_OK;
}

extern "C" int svk_infer_host(svk_handle* h, const float* mel, const int64_t* lengths, const float* eps,
                              float noise_scale, int B, int T, int max_len, float* o, float* x_mask, float* z,
                              float* z_p, float* m_p, float* logs_p) {
  SVK_TRY(check_ready(h, "svk_infer_host"));
  if (!mel || !lengths || !eps || !o) return fail(SVK_ERR_INVALID, "svk_infer_host: null mel/lengths/eps/o");
  if (B <= 0 || T <= 0) return fail(SVK_ERR_INVALID, "svk_infer_host: B and T must be positive");
  CUDA_TRY(cudaSetDevice(h->device));
  if (!h->host_stream) CUDA_TRY(cudaStreamCreateWithFlags(&h->host_stream, cudaStreamNonBlocking));
  const svk_config& c = h->cfg;
  const int Tp = max_len > 0 && max_len < T ? max_len : T;
  const size_t n_mel = (size_t)B * c.n_mel * T, n_lat = (size_t)B * c.inter_channels * T, n_o = (size_t)B * h->hop() * Tp;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o2 = off;
    off = align_up(off + bytes, 256);
    return o2;
  };
  const size_t o_mel = take(n_mel * 4), o_len = take((size_t)B * 8), o_eps = take(n_lat * 4), o_o = take(n_o * 4),
               o_mask = take((size_t)B * T * 4);
  size_t o_lat[4];
  for (int i = 0; i < 4; ++i) o_lat[i] = take(n_lat * 4);
  const size_t ws_bytes = svk_workspace_bytes(h, B, T, max_len);
  const size_t o_ws = take(ws_bytes);
  if (off > h->host_dev_bytes) {
    if (h->host_dev) CUDA_TRY(cudaFree(h->host_dev));
    h->host_dev = nullptr, h->host_dev_bytes = 0;
    CUDA_TRY(cudaMalloc(&h->host_dev, off));
    h->host_dev_bytes = off;
  }
  char* d = (char*)h->host_dev;
  cudaStream_t s = h->host_stream;
  CUDA_TRY(cudaMemcpyAsync(d + o_mel, mel, n_mel * 4, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(d + o_len, lengths, (size_t)B * 8, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(d + o_eps, eps, n_lat * 4, cudaMemcpyHostToDevice, s));
  float* dl[4];
  float* hl[4] = {z, z_p, m_p, logs_p};
  for (int i = 0; i < 4; ++i) dl[i] = hl[i] ? (float*)(d + o_lat[i]) : nullptr;
  SVK_TRY(svk_infer(h, (const float*)(d + o_mel), (const int64_t*)(d + o_len), (const float*)(d + o_eps), noise_scale,
                    B, T, max_len, (float*)(d + o_o), (float*)(d + o_mask), dl[0], dl[1], dl[2], dl[3], d + o_ws,
                    ws_bytes, s));
  CUDA_TRY(cudaMemcpyAsync(o, d + o_o, n_o * 4, cudaMemcpyDeviceToHost, s));
  if (x_mask) CUDA_TRY(cudaMemcpyAsync(x_mask, d + o_mask, (size_t)B * T * 4, cudaMemcpyDeviceToHost, s));
  for (int i = 0; i < 4; ++i)
    if (hl[i]) CUDA_TRY(cudaMemcpyAsync(hl[i], dl[i], n_lat * 4, cudaMemcpyDeviceToHost, s));
  return svk_check_range(h, s);  // synchronises
}

// ------------------------------------------------------------- windowed / chunked synthesis (8(f) rank 3)
namespace {

// Frames of context per side after which a window's output equals the whole-utterance result:
// WN stacks (k-1)/2 per layer (modules.py:133), decoder receptive field propagated back through
// conv_post, the ResBlock1 dilations (modules.py:191-206) and the transposed convs.
int halo_frames(const svk_handle* h) {
  const svk_config& c = h->cfg;
  const int wn = (c.wn_kernel - 1) / 2;
  int halo = 3;  // conv_post k=7 at the output rate
  for (int i = c.n_upsamples - 1; i >= 0; --i) {
    int rb = 0;
    for (int j = 0; j < c.n_resblock_kernels; ++j) {
      int r = 0;
      const int hk = (c.resblock_kernel_sizes[j] - 1) / 2;
      if (c.resblock_type == 2) {
        for (int l = 0; l < 2; ++l) r += hk * c.resblock_dilations[j][l];
      } else {
        for (int l = 0; l < SVK_RESBLOCK_PAIRS; ++l) r += hk * (c.resblock_dilations[j][l] + 1);
      }
      rb = r > rb ? r : rb;
    }
    halo += rb;
    const int s = c.upsample_rates[i];
    halo = (halo + s - 1) / s + 1;  // ConvTranspose1d k = 2s: an output reads inputs floor((t+p-k+1)/s) .. floor((t+p)/s)
  }
  halo += 3;  // conv_pre k=7
  return halo + c.enc_layers * wn + c.n_flows * c.flow_layers * wn;
}

struct WindowPlan {
  int64_t W;  // widest window: frames + 2 * halo
  size_t mel, eps, len, o, lat[4], ws, ws_bytes, total;  // byte offsets
};

WindowPlan plan_window(const svk_handle* h, int B, int frames) {
  const svk_config& c = h->cfg;
  WindowPlan p;
  p.W = (int64_t)frames + 2 * halo_frames(h);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  p.mel = take((size_t)B * c.n_mel * p.W * 4);
  p.eps = take((size_t)B * c.inter_channels * p.W * 4);
  p.len = take((size_t)B * 8);
  p.o = take((size_t)B * h->hop() * p.W * 4);
  for (int i = 0; i < 4; ++i) p.lat[i] = take((size_t)B * c.inter_channels * p.W * 4);
  p.ws_bytes = svk_workspace_bytes(h, B, (int)p.W, 0);
  p.ws = take(p.ws_bytes);
  p.total = off;
  return p;
}

}  // namespace

extern "C" int svk_halo_frames(const svk_handle* h) { return h ? halo_frames(h) : 0; }

extern "C" size_t svk_window_workspace_bytes(const svk_handle* h, int B, int frames) {
  if (!h || B <= 0 || frames <= 0) return 0;
  return plan_window(h, B, frames).total;
}

extern "C" int svk_infer_window(svk_handle* h, const float* mel, const int64_t* lengths, const float* eps,
                                float noise_scale, int B, int T, int max_len, int t0, int t1, float* o, int64_t o_row_stride,
                                float* z, float* z_p, float* m_p, float* logs_p, int64_t lat_row_stride,
                                void* workspace, size_t workspace_bytes, void* stream) {
  SVK_TRY(check_ready(h, "svk_infer_window"));
  if (!mel || !lengths || !eps || !o) return fail(SVK_ERR_INVALID, "svk_infer_window: null mel/lengths/eps/o");
  const int Tp = max_len > 0 && max_len < T ? max_len : T;  // decoder input is (z * mask)[:, :, :max_len] (models.py:338)
  if (B <= 0 || T <= 0 || t0 < 0 || t1 <= t0 || t1 > Tp) return fail(SVK_ERR_INVALID, "svk_infer_window: need 0 <= t0 < t1 <= min(T, max_len)");
  const svk_config& c = h->cfg;
  const int hop = h->hop(), halo = halo_frames(h);
  const WindowPlan p = plan_window(h, B, t1 - t0);
  if (!workspace || workspace_bytes < p.total)
    return fail(SVK_ERR_WORKSPACE, "svk_infer_window: workspace %zu < %zu bytes (svk_window_workspace_bytes)", workspace_bytes, p.total);
  if (o_row_stride < (int64_t)hop * (t1 - t0)) return fail(SVK_ERR_INVALID, "svk_infer_window: o_row_stride shorter than the window");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  // window with context, clipped at the true sequence ends (where the reference zero-pads too)
  const int a = t0 - halo > 0 ? t0 - halo : 0, b = t1 + halo < T ? t1 + halo : T, w = b - a;
  char* d = (char*)workspace;
  float* mel_w = (float*)(d + p.mel);
  float* eps_w = (float*)(d + p.eps);
  int64_t* len_w = (int64_t*)(d + p.len);
  float* o_w = (float*)(d + p.o);
  CUDA_TRY(cudaMemcpy2DAsync(mel_w, (size_t)w * 4, mel + a, (size_t)T * 4, (size_t)w * 4, (size_t)B * c.n_mel, cudaMemcpyDeviceToDevice, s));
  CUDA_TRY(cudaMemcpy2DAsync(eps_w, (size_t)w * 4, eps + a, (size_t)T * 4, (size_t)w * 4, (size_t)B * c.inter_channels, cudaMemcpyDeviceToDevice, s));
  CUDA_TRY(launch_window_lengths(lengths, B, a, w, len_w, s));
  float* user[4] = {z, z_p, m_p, logs_p};
  float* lat[4];
  for (int i = 0; i < 4; ++i) lat[i] = user[i] ? (float*)(d + p.lat[i]) : nullptr;
  // the encoder and the flow see frames up to b; the decoder stops at the true end Tp when the window crosses it
  const int wd = b > Tp ? Tp - a : w;
  SVK_TRY(svk_infer(h, mel_w, len_w, eps_w, noise_scale, B, w, wd < w ? wd : 0, o_w, nullptr, lat[0], lat[1], lat[2], lat[3],
                    d + p.ws, p.ws_bytes, stream));
  const int64_t launches = h->launches;
  const size_t n = (size_t)(t1 - t0);
  CUDA_TRY(cudaMemcpy2DAsync(o, (size_t)o_row_stride * 4, o_w + (size_t)(t0 - a) * hop, (size_t)wd * hop * 4, n * hop * 4, (size_t)B,
                             cudaMemcpyDeviceToDevice, s));
  for (int i = 0; i < 4; ++i)
    if (user[i]) {
      if (lat_row_stride < (int64_t)n) return fail(SVK_ERR_INVALID, "svk_infer_window: lat_row_stride shorter than the window");
      CUDA_TRY(cudaMemcpy2DAsync(user[i], (size_t)lat_row_stride * 4, lat[i] + (t0 - a), (size_t)w * 4, n * 4,
                                 (size_t)B * c.inter_channels, cudaMemcpyDeviceToDevice, s));
    }
  h->launches = launches;
  return SVK_OK;
}

extern "C" int svk_infer_chunked(svk_handle* h, const float* mel, const int64_t* lengths, const float* eps,
                                 float noise_scale, int B, int T, int max_len, int chunk_frames, float* o, float* x_mask, float* z,
                                 float* z_p, float* m_p, float* logs_p, void* workspace, size_t workspace_bytes,
                                 void* stream) {
  SVK_TRY(check_ready(h, "svk_infer_chunked"));
  if (chunk_frames <= 0) return fail(SVK_ERR_INVALID, "svk_infer_chunked: chunk_frames must be positive");
  if (B <= 0 || T <= 0 || !o) return fail(SVK_ERR_INVALID, "svk_infer_chunked: bad argument");
  const int hop = h->hop();
  const int Tp = max_len > 0 && max_len < T ? max_len : T;
  if (Tp < T && (z || z_p || m_p || logs_p))
    return fail(SVK_ERR_INVALID, "svk_infer_chunked: latents with max_len < T are not supported (use svk_infer)");
  int64_t total = 0;
  for (int t0 = 0; t0 < Tp; t0 += chunk_frames) {
    const int t1 = t0 + chunk_frames < Tp ? t0 + chunk_frames : Tp;
    SVK_TRY(svk_infer_window(h, mel, lengths, eps, noise_scale, B, T, max_len, t0, t1, o + (size_t)t0 * hop, (int64_t)hop * Tp,
                             z ? z + t0 : nullptr, z_p ? z_p + t0 : nullptr, m_p ? m_p + t0 : nullptr,
                             logs_p ? logs_p + t0 : nullptr, T, workspace, workspace_bytes, stream));
    total += h->launches;
  }
  if (x_mask) {
    const cudaError_t e = launch_sequence_mask(lengths, B, T, x_mask, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(SVK_ERR_CUDA, "svk_infer_chunked: %s", cudaGetErrorString(e));
    total++;
  }
  h->launches = total;
  return SVK_OK;
}

// ----------------------------------------------------------------------------- module-level ops
extern "C" int svk_mel_encoder(svk_handle* h, const float* mel, const int64_t* lengths, int B, int T, float* x_out,
                               float* m, float* logs, float* mask, void* workspace, size_t workspace_bytes,
                               void* stream) {
  SVK_TRY(check_ready(h, "svk_mel_encoder"));
  if (!mel || !lengths || !x_out || !m || !logs || !mask || B <= 0 || T <= 0)
    return fail(SVK_ERR_INVALID, "svk_mel_encoder: bad argument");
  const size_t n = align_up((size_t)B * h->cfg.hidden_channels * T, 64);
  if (!workspace || workspace_bytes < 4 * n * sizeof(float)) return fail(SVK_ERR_WORKSPACE, "svk_mel_encoder: workspace too small");
  CUDA_TRY(cudaSetDevice(h->device));
  float* ws = (float*)workspace;
  h->launches = 0;
  Runner R{h, (cudaStream_t)stream, B};
  R.note(launch_sequence_mask(lengths, B, T, mask, R.stream));
  run_mel_encoder(R, mel, mask, T, ws, ws + n, x_out, m, logs, reinterpret_cast<uint16_t*>(ws + 2 * n),
                  reinterpret_cast<uint16_t*>(ws + 3 * n));
  if (R.err != cudaSuccess) return fail(SVK_ERR_CUDA, "svk_mel_encoder: %s", cudaGetErrorString(R.err));
  return SVK_OK;
}

extern "C" int svk_flow_reverse(svk_handle* h, float* z, const float* mask, int B, int T, void* workspace,
                                size_t workspace_bytes, void* stream) {
  SVK_TRY(check_ready(h, "svk_flow_reverse"));
  if (!z || !mask || B <= 0 || T <= 0) return fail(SVK_ERR_INVALID, "svk_flow_reverse: bad argument");
  const svk_config& c = h->cfg;
  const size_t n = align_up((size_t)B * c.hidden_channels * T, 64), nl = align_up((size_t)B * c.inter_channels * T, 64);
  if (!workspace || workspace_bytes < (5 * n + nl) * sizeof(float)) return fail(SVK_ERR_WORKSPACE, "svk_flow_reverse: workspace too small");
  CUDA_TRY(cudaSetDevice(h->device));
  float* ws = (float*)workspace;
  h->launches = 0;
  Runner R{h, (cudaStream_t)stream, B};
  run_flow_reverse(R, z, mask, T, ws, ws + n, ws + 2 * n, reinterpret_cast<uint16_t*>(ws + 3 * n + nl),
                   reinterpret_cast<uint16_t*>(ws + 4 * n + nl));
  if (c.n_flows & 1) {
    float* tmp = ws + 3 * n;
    R.note(launch_flip(z, B, c.inter_channels, T, tmp, R.stream));
    if (R.err == cudaSuccess)
      R.err = cudaMemcpyAsync(z, tmp, sizeof(float) * (size_t)B * c.inter_channels * T, cudaMemcpyDeviceToDevice, R.stream);
  }
  if (R.err != cudaSuccess) return fail(SVK_ERR_CUDA, "svk_flow_reverse: %s", cudaGetErrorString(R.err));
  return SVK_OK;
}

extern "C" int svk_flow_forward(svk_handle* h, float* z, const float* mask, int B, int T, void* workspace,
                                size_t workspace_bytes, void* stream) {
  SVK_TRY(check_ready(h, "svk_flow_forward"));
  if (!z || !mask || B <= 0 || T <= 0) return fail(SVK_ERR_INVALID, "svk_flow_forward: bad argument");
  const svk_config& c = h->cfg;
  if (c.n_flows & 1) return fail(SVK_ERR_INVALID, "svk_flow_forward: an odd number of flows is not supported");
  const size_t n = align_up((size_t)B * c.hidden_channels * T, 64);
  if (!workspace || workspace_bytes < 5 * n * sizeof(float)) return fail(SVK_ERR_WORKSPACE, "svk_flow_forward: workspace too small");
  CUDA_TRY(cudaSetDevice(h->device));
  float* ws = (float*)workspace;
  h->launches = 0;
  Runner R{h, (cudaStream_t)stream, B};
  run_flow_forward(R, z, mask, T, ws, ws + n, ws + 2 * n, reinterpret_cast<uint16_t*>(ws + 3 * n),
                   reinterpret_cast<uint16_t*>(ws + 4 * n));
  if (R.err != cudaSuccess) return fail(SVK_ERR_CUDA, "svk_flow_forward: %s", cudaGetErrorString(R.err));
  return SVK_OK;
}

extern "C" int svk_posterior_encoder(svk_handle* h, const float* spec, const int64_t* lengths, const float* eps, int B,
                                     int T, float* z, float* m, float* logs, float* mask, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  SVK_TRY(check_ready(h, "svk_posterior_encoder"));
  if (!h->has_posterior) return fail(SVK_ERR_STATE, "svk_posterior_encoder: enc_q.* weights were not loaded");
  if (!spec || !lengths || !eps || !z || !m || !logs || !mask || B <= 0 || T <= 0)
    return fail(SVK_ERR_INVALID, "svk_posterior_encoder: bad argument");
  const size_t n = align_up((size_t)B * h->cfg.hidden_channels * T, 64);
  if (!workspace || workspace_bytes < 5 * n * sizeof(float)) return fail(SVK_ERR_WORKSPACE, "svk_posterior_encoder: workspace too small");
  CUDA_TRY(cudaSetDevice(h->device));
  float* ws = (float*)workspace;
  h->launches = 0;
  Runner R{h, (cudaStream_t)stream, B};
  R.note(launch_sequence_mask(lengths, B, T, mask, R.stream));
  run_posterior_encoder(R, spec, eps, mask, T, ws, ws + n, ws + 2 * n, z, m, logs, reinterpret_cast<uint16_t*>(ws + 3 * n),
                        reinterpret_cast<uint16_t*>(ws + 4 * n));
  if (R.err != cudaSuccess) return fail(SVK_ERR_CUDA, "svk_posterior_encoder: %s", cudaGetErrorString(R.err));
  return SVK_OK;
}

extern "C" int svk_generator(svk_handle* h, const float* z, int B, int L, float* o, void* workspace,
                             size_t workspace_bytes, void* stream) {
  SVK_TRY(check_ready(h, "svk_generator"));
  if (!z || !o || B <= 0 || L <= 0) return fail(SVK_ERR_INVALID, "svk_generator: bad argument");
  const DecoderPlan dplan = plan_decoder(h, B, L);
  const size_t need = dplan.n_bufs * dplan.buf_floats * sizeof(float);
  if (!workspace || workspace_bytes < need) return fail(SVK_ERR_WORKSPACE, "svk_generator: workspace %zu < %zu", workspace_bytes, need);
  CUDA_TRY(cudaSetDevice(h->device));
  h->launches = 0;
  Runner R{h, (cudaStream_t)stream, B};
  run_decoder(R, z, L, nullptr, L, o, (float*)workspace);
  if (R.err != cudaSuccess) return fail(SVK_ERR_CUDA, "svk_generator: %s", cudaGetErrorString(R.err));
  return SVK_OK;
}

extern "C" size_t svk_resblock1_workspace_bytes(const svk_handle* h, int index, int B, int L) {
  if (!h || index < 0 || index >= (int)h->resblocks.size() || B <= 0 || L <= 0) return 0;
  return 4 * align_up((size_t)B * h->resblocks[index].C * L, 64) * sizeof(float);
}

extern "C" int svk_resblock1(svk_handle* h, int index, const float* x, int B, int L, float* y, void* workspace,
                             size_t workspace_bytes, void* stream) {
  SVK_TRY(check_ready(h, "svk_resblock1"));
  if (index < 0 || index >= (int)h->resblocks.size()) return fail(SVK_ERR_INVALID, "svk_resblock1: index out of range");
  if (!x || !y || B <= 0 || L <= 0) return fail(SVK_ERR_INVALID, "svk_resblock1: bad argument");
  const ResBlock& rb = h->resblocks[index];
  const size_t n = align_up((size_t)B * rb.C * L, 64);
  if (!workspace || workspace_bytes < 4 * n * sizeof(float))
    return fail(SVK_ERR_WORKSPACE, "svk_resblock1: workspace %zu < %zu (svk_resblock1_workspace_bytes)", workspace_bytes,
                4 * n * sizeof(float));
  CUDA_TRY(cudaSetDevice(h->device));
  float* ws = (float*)workspace;
  h->launches = 0;
  Runner R{h, (cudaStream_t)stream, B};
  if (R.resblock_uses_images(rb)) {
    uint16_t* x_img = reinterpret_cast<uint16_t*>(ws + n);
    R.split_image(x, rb.C, L, 0.1f, x_img);
    R.resblock_images(rb, x, x_img, reinterpret_cast<uint16_t*>(ws + 2 * n), ws, reinterpret_cast<uint16_t*>(ws + 3 * n), y,
                      nullptr, 1.0f, L);
  } else {
    R.resblock(rb, x, ws, ws + n, y, nullptr, 1.0f, L);
  }
  if (R.err != cudaSuccess) return fail(SVK_ERR_CUDA, "svk_resblock1: %s", cudaGetErrorString(R.err));
  return SVK_OK;
}
