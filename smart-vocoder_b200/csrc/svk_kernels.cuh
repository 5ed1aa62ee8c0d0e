// Internal kernel interfaces of libsvk (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace svk {

// One side of a (possibly split) epilogue: where an output-channel range is written and what is
// folded into it.  Used e.g. by WN's res_skip conv (modules.py:169-175): channels [0,H) update
// x = (x + r) * mask in place, channels [H,2H) accumulate out += r.
struct EpiDesc {
  const float* res;     // added element-wise (same [B, C, L] geometry as y), may be null
  const float* acc_in;  // added element-wise after res (running sum), may be null
  float* y;             // destination tensor [B, C, y_stride]
  int C;                // channels of the destination tensor
  int ch_off;           // destination channel = ch_off + ch_sign * (o - first o of this side)
  int ch_sign;          // +1, or -1 to write channels in reversed order (folded Flip)
  int use_mask;         // multiply by out_mask[b, t]
  // Optional second output for the tcgen05 engine: leaky_relu(y, split_slope) as the fp16 hi/lo operand
  // image [hi|lo][B][C/32][L = y_stride][32] that the consuming conv loads by TMA (conv_tc.cu).  `y` may
  // then be null when nothing reads the fp32 tensor.  Needs ch_sign == +1, ch_off % 16 == 0, C % 32 == 0.
  uint16_t* split;
  float split_slope;
  // Optional residual taken from an operand image instead of `res` (hi/lo engine only): the image holds
  // leaky_relu(r, res_slope) with the destination tensor's geometry; the epilogue inverts the leaky_relu
  // (r = v >= 0 ? v : v / res_slope).  Lets a residual stream live in HBM as images only.
  const uint16_t* res_img;
  float res_slope;
};

enum ConvMode : int {
  MODE_STORE = 0,    // plain conv epilogue
  MODE_GATE = 1,     // tanh(a[c]) * sigmoid(a[c+H]) (commons.py:100-107); packed channel pairs
  MODE_SHUFFLE = 2,  // ConvTranspose1d polyphase store
};

// Generic stride-1 Conv1d as implicit GEMM on the fp32 FFMA pipe.
//   y[b,o,t] = bias[o] + sum_c sum_j wp[c][j][o] * f(x[b, x_ch_off + c, t - pad + j*dil])
// with f = leaky_relu(pre_slope) then * in_mask.  Weights are pre-packed [Cin][K][CoutPad].
struct ConvArgs {
  const float* x;
  int x_C;          // channels of the tensor x lives in
  int x_ch_off;     // first input channel used
  int x_stride;     // row stride in floats (>= Lin)
  int Lin;          // logical input length (zero padding outside [0, Lin))
  const float* in_mask;  // [B, mask_stride] or null
  int mask_stride;
  float pre_slope;  // 1.0f = none
  const float* wp;
  const float* bias;  // [CoutPad] (zeros when the layer has no bias)
  int Cin;
  int Cout;     // logical output channels (virtual channels for GATE / SHUFFLE)
  int CoutPad;  // packed row length (multiple of the CTA's channel tile)
  int K;
  int dil;
  int pad;
  int Lout;      // outputs computed per row (virtual positions q for SHUFFLE)
  int y_stride;  // row stride of destination tensors
  int mode;
  int split;     // o < split -> e[0], else e[1]
  EpiDesc e[2];
  float post_div;  // 1.0f = none
  int act_tanh;
  const float* out_mask;  // [B, mask_stride] or null
  int shuf_s, shuf_p, shuf_Lout;  // MODE_SHUFFLE: t = s*q + r - p, valid in [0, shuf_Lout)
  int shuf_rmajor;  // MODE_SHUFFLE, tcgen05 engine only: virtual channel o' = r * (Cout / s) + co instead of co * s + r, so a
                    // 16-column job holds 16 real channels of ONE output step = one 32 B sector of an operand-image row
  int B;
  int* range_flag;  // act_tanh launches: set to 1 when an output is outside [-1, 1] (NaN), may be null (svk_check_range)
};

// Channel tile (in output channels) the FFMA kernel will use for a layer with `cout` outputs.
int conv_ffma_channel_tile(int cout);
// true when a kernel instance exists for this tap count
bool conv_ffma_supports_k(int k);
cudaError_t launch_conv_ffma(const ConvArgs& a, cudaStream_t stream);

// ---- tcgen05 path (conv_tc.cu): same ConvArgs semantics, weights pre-split into fp16 hi/lo images.
constexpr int TC_KC = 32;  // input channels per staged chunk (Cin must be a multiple)
// Division by a launch-invariant divisor as multiply-high + shifts (Granlund-Montgomery, exact for all uint32).
struct FastDiv {
  uint32_t d, m, sh1, sh2;
};

struct ConvTcArgs {
  ConvArgs c;           // c.wp unused; c.bias = fp32 bias [Cout] (unscaled)
  const uint16_t* wtc;  // packed image, see conv_tc_pack
  float unscale;        // 1 / weight scale (power of two)
  int N;                // output channels per tile (multiple of 16, <= 128)
  int planes;           // 2 (default when 0): fp16 hi/lo three-product scheme; 1: single bf16 pass, fp32 accumulate
  const uint16_t* x_split;  // non-null: input comes from this operand image (geometry [B, c.x_C, c.Lin]) by TMA;
                            // leaky_relu / mask were applied when it was written, c.x / pre_slope / in_mask unused
  // filled by launch_conv_tc:
  int rows, tmem_cols, na, nw, resident, items, ntiles_t, bias_bytes, bias_count, nacc, epi_groups, a_off;
  int epi_fast;  // lean STORE epilogue (conv_tc.cu: the decoder ResBlock convs take it)
  int dual_issue;  // two MMA issuer warps taking alternate tiles (resident weights, short MMAs)
  int lazy_res;    // lean epilogue: residual image words decoded where they are used, not where they are loaded
  FastDiv div_t, div_b;  // by ntiles_t and by B (work-item decoding)
};
int conv_tc_rows(int K, int dil);
size_t conv_tc_packed_halves(int Cin, int Cout, int K, int N, int planes = 2);
float conv_tc_weight_scale(const float* w, size_t n);
void conv_tc_pack(const float* w_ock, int Cout, int Cin, int K, int N, float scale, uint16_t* out, int planes = 2);
void conv_tc_plan(int Cin, int Cout, int K, int dil, int N, bool tma, int planes, int* na, int* nw, int* resident,
                  size_t* smem_bytes);
cudaError_t launch_conv_tc(const ConvTcArgs& a, cudaStream_t stream);
// fp32 [B, C, L] -> operand image of leaky_relu(x, slope) (C % 32 == 0); bytes = 2 * planes * B * C * L
cudaError_t launch_split_image(const float* x, int B, int C, int L, float slope, uint16_t* img, int planes,
                               cudaStream_t stream);
inline size_t split_image_halves(int B, int C, int L, int planes = 2) { return (size_t)planes * B * C * L; }

// ---- fused ResBlock1 conv pair on the narrow stages (conv_tc_pair.cu): xt = conv1(x_img) stays on the SM,
// y = conv2(leaky_relu(xt)) + residual.  Weights are conv_tc_pack images with N = C, planes = 2.
struct ConvPairArgs {
  const uint16_t* x_img;  // operand image of leaky_relu(x) [B, C, L]
  int B, C, L, K, dil1;   // both convs have K taps; conv1 dilation dil1, conv2 dilation 1
  const uint16_t* w1;
  const uint16_t* w2;
  const float* bias1;     // [C]
  const float* bias2;
  float unscale1, unscale2;
  float xt_slope;         // leaky_relu between the convs
  const float* res;       // fp32 residual [B, C, L], or null
  const uint16_t* res_img;  // ... or its operand image (leaky_relu(res, res_slope))
  float res_slope;
  const float* acc_in;    // optional running sum added after the residual
  float post_div;         // 1.0f = none
  float* y;               // fp32 output or null
  uint16_t* y_img;        // operand image of leaky_relu(y, y_slope) or null
  float y_slope;
  // filled by launch_conv_tc_pair:
  int rows1, rows2, TO, h, ntiles_t, items, na, nw, resident, a_off, a2_off, w_off;
  int dual_issue;  // conv1 and conv2 issued by two warps (resident weights)
  int ns;          // stages of the accumulator sets and xt tiles = items in flight (2 or 4)
  FastDiv div_t;
};
bool conv_tc_pair_supported(int C, int K, int dil1);
cudaError_t launch_conv_tc_pair(const ConvPairArgs& a, cudaStream_t stream);

// ---- fused WN layer (wn_layer.cu): in_layer conv + gate + res_skip 1x1 + residual / skip update in ONE launch
// (modules.py:156-175).  Weights are the conv_tc_pack images of the two convs (in_layer gate-packed, planes as given).
constexpr int WN_MAX_LAYERS = 16;
struct WnLayerParams {       // what differs between the layers of one WN stack
  const uint16_t* w_in;      // in_layer image (gate-packed), N_in channels per tile
  const uint16_t* w_rs;      // res_skip image
  const float* bias_in;      // [nt_in * N_in], virtual (gate-packed) order
  const float* bias_rs;      // [nt_rs * N_rs]
  float unscale_in, unscale_rs;
  int N_rs, Cout_rs;         // res_skip: channels per tile, outputs (2H; H on the last layer of a stack)
  int nt_rs, bias_count_rs;  // filled by launch_wn_layers
};
// n_layers consecutive layers l0 .. l0 + n_layers - 1 of a WN stack of n_total layers in ONE launch (wn_layer.cu).
// n_layers == 1: one launch per layer.  n_layers > 1: every CTA keeps its 128-frame tile through all layers and adjacent
// tiles synchronise through `flags` (needs items <= SM count: all CTAs co-resident).
struct WnLayerArgs {
  int B, T, H, K, planes;
  uint16_t* img[2];          // operand images of x [B, H, T]: layer l0 + j reads img[j & 1] (with a (K-1)/2 halo, by TMA) and
                             // writes the updated x into img[(j + 1) & 1] (neighbours still read the old one)
  float* x;                  // fp32 x, updated in place (unused on the last layer of the stack)
  float* out;                // fp32 skip accumulator [B, H, T]
  const float* mask;         // [B, T]
  int l0, n_layers, n_total; // global layer 0: out = skip (nothing read); global layer n_total - 1: res_skip has H outputs,
                             // out = (out + rs) * mask, no x / image written
  int N_in;
  int* flags;                // [items] zeroed before the launch (n_layers > 1): layers completed by each tile
  long long* trace;          // developer aid ($SVK_WN_TRACE): [items][n_layers][32] clock64() stamps of the role loops, or null
  WnLayerParams layer[WN_MAX_LAYERS];
  // filled by launch_wn_layers:
  int rows, ntiles_t, items, nt_in, na, nw, a_off, acts_off, w_off, w_slot, bias_count_in, bias_max_rs, acc_stride, tmem_cols;
  FastDiv div_t;
};
bool wn_layer_supported(int H, int K, int N_in, int N_rs, int Cout_rs, int planes);
int wn_stack_max_items();  // tiles a multi-layer launch can take (= SM count of the current device)
cudaError_t launch_wn_layers(const WnLayerArgs& a, cudaStream_t stream);

// elementwise / small kernels
cudaError_t launch_sequence_mask(const int64_t* lengths, int B, int T, float* mask, cudaStream_t s);
cudaError_t launch_window_lengths(const int64_t* lengths, int B, int64_t a, int64_t w, int64_t* out, cudaStream_t s);
cudaError_t launch_pcm_to_int16(const float* x, int64_t n, float scale, int16_t* y, cudaStream_t s);
cudaError_t launch_flip(const float* x, int B, int C, int T, float* y, cudaStream_t s);
cudaError_t launch_sample(const float* m, const float* logs, const float* eps, float noise_scale,
                          float* z_p, float* z, int64_t n, cudaStream_t s);
cudaError_t launch_posterior_sample(const float* m, const float* logs, const float* eps, const float* mask, int B, int C,
                                    int T, float* z, cudaStream_t s);
cudaError_t launch_weight_norm(const float* v, const float* g, int64_t dim0, int64_t inner, float* w,
                               cudaStream_t s);
cudaError_t launch_rq_spline(const float* x, const float* uw, const float* uh, const float* ud,
                             int64_t n, int nb, int inverse, float tail_bound, float min_bw,
                             float min_bh, float min_d, float* y, float* lad, int32_t* bins,
                             cudaStream_t s);

}  // namespace svk
