/*
 * svk.h -- C ABI of the B200-native SMART-Vocoder mel->waveform path (libsvk.so).
 *
 * The reference (SMART-TTS/SMART-Vocoder) is pure Python and has no FFI of its own; its de-facto
 * boundary is the nn.Module surface used by inference.ipynb:64-71,118 and train.py:82-86,273
 * (SURVEY 8b).  Every entry point below names the reference interface it replaces (file:line in
 * /root/reference).  smart-vocoder_b200/models.py binds these with ctypes and re-exposes the
 * reference's `SynthesizerTrn` signature on top.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary;
 *   - every function returns an int status (SVK_OK == 0); svk_last_error() gives a thread-local
 *     message; nothing throws across the ABI;
 *   - tensors are contiguous fp32 [batch, channels, time] ("NCT"), like the reference;
 *   - "dev" pointers are CUDA device pointers on the handle's device; `stream` is a cudaStream_t
 *     passed as void* (0 = legacy default stream).  Device entry points are stream-ordered: no
 *     host synchronisation and no allocation happen inside them;
 *   - a handle is bound to one device and is not re-entrant (one call at a time per handle).
 *   - There is NO CPU fallback: every compute entry point fails with SVK_ERR_CUDA when no
 *     sm_100 device is usable.
 */
#ifndef SVK_H_
#define SVK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVK_ABI_VERSION 2

#define SVK_OK 0
#define SVK_IGNORED 1            /* svk_load_tensor: key is dead at inference, accepted and dropped */
#define SVK_ERR_INVALID (-1)     /* bad argument / shape / configuration                            */
#define SVK_ERR_CUDA (-2)        /* CUDA runtime error, or no usable device                         */
#define SVK_ERR_UNKNOWN_KEY (-3) /* svk_load_tensor: key is not part of the checkpoint surface      */
#define SVK_ERR_STATE (-4)       /* call order (e.g. infer before svk_finalize_weights)             */
#define SVK_ERR_WORKSPACE (-5)   /* workspace too small                                             */
#define SVK_ERR_RANGE (-6)       /* a non-finite waveform sample was produced: an activation left the
                                    fp16 operand range of the tensor-core engine (|x| > 65504), or the
                                    inputs / weights were not finite (svk_check_range)              */

#define SVK_MAX_UPSAMPLES 8
#define SVK_MAX_RESBLOCK_KERNELS 8
#define SVK_RESBLOCK_PAIRS 3

/* Arithmetic of the convolution stacks. */
#define SVK_PRECISION_FP32 0   /* fp32 FFMA everywhere                                    */
#define SVK_PRECISION_TC 1     /* fp32-class results on the tensor cores: every conv whose widths
                                  allow it runs as a 3-product fp16 split (hi*hi + hi*lo + lo*hi,
                                  fp32 accumulate in TMEM) on tcgen05; the rest stays on FFMA.
                                  Meets the same 1e-4 bar as SVK_PRECISION_FP32.            */
#define SVK_PRECISION_BF16 2   /* BASELINE configs[3]: bf16 operands (activations stored in HBM as bf16
                                  operand images, weights bf16), ONE tcgen05 pass per conv, fp32
                                  accumulate in TMEM, fp32 residual stream.  No reference counterpart
                                  (the reference has no reduced-precision inference): not bound by the
                                  1e-4 bar; tests state its tolerance (SNR vs the fp64 oracle).   */

/*
 * Effective hyper-parameters of SynthesizerTrn.__init__ (reference models.py:266-314).
 * Values the reference hard-codes (enc_layers 16, flow_layers 8, n_flows 4, wn_kernel 5,
 * n_mel 80) are explicit here so the library carries no hidden constants.
 */
typedef struct svk_config {
  int32_t n_mel;
  int32_t spec_channels;   /* only sizes the dead enc_q.* keys */
  int32_t inter_channels;
  int32_t hidden_channels;
  int32_t enc_layers;
  int32_t flow_layers;
  int32_t n_flows;
  int32_t wn_kernel;
  int32_t gin_channels;    /* only sizes the dead cond keys */
  int32_t upsample_initial_channel;
  int32_t n_upsamples;
  int32_t upsample_rates[SVK_MAX_UPSAMPLES];
  int32_t upsample_kernel_sizes[SVK_MAX_UPSAMPLES];
  int32_t n_resblock_kernels;
  int32_t resblock_kernel_sizes[SVK_MAX_RESBLOCK_KERNELS];
  int32_t resblock_dilations[SVK_MAX_RESBLOCK_KERNELS][SVK_RESBLOCK_PAIRS];
  int32_t precision;       /* SVK_PRECISION_* */
  int32_t resblock_type;   /* 0 or 1: modules.ResBlock1 (3 conv pairs, modules.py:187-229); 2: modules.ResBlock2 (two convs with
                              dilations resblock_dilations[j][0..1], each x = conv(lrelu(x)) + x, modules.py:232-256) --
                              Generator picks by the config's "resblock" key (models.py:121) */
} svk_config;

typedef struct svk_handle svk_handle;

int svk_abi_version(void);
const char *svk_last_error(void);

/* ---- lifetime ------------------------------------------------------------------------------
 * Replaces SynthesizerTrn(...).cuda().eval()  (models.py:266-314, inference.ipynb:64-70). */
int svk_create(const svk_config *cfg, int device, svk_handle **out);
void svk_destroy(svk_handle *h);

/* ---- weight ingress ------------------------------------------------------------------------
 * Replaces utils.load_checkpoint -> model.load_state_dict  (utils.py:18-43).
 * Called once per state_dict entry with the reference key name ("dec.ups.0.weight_v", ...) and a
 * HOST fp32 tensor.  Keys that no built path reads (*.cond_layer.*, dec.cond.*) are
 * accepted and return SVK_IGNORED.  enc_q.* keys are optional: svk_infer never reads them, loading
 * all of them enables svk_posterior_encoder.  Shapes are checked. */
int svk_load_tensor(svk_handle *h, const char *key, const float *host_data, const int64_t *shape,
                    int ndim);
/* Folds weight_norm (w = g * v / ||v||, norm over dims != 0; models.py:125, modules.py:128-145,
 * 191-206 -- the reference recomputes this 172 times per infer), folds the channel
 * Flip (modules.py:270-277) of odd couplings into the 1x1 pre/post weights, repacks for the
 * kernels and uploads.  Fails if a live key was never loaded. */
int svk_finalize_weights(svk_handle *h);
/* Number of live (inference-relevant) keys and how many of them have been loaded so far. */
int svk_weight_status(const svk_handle *h, int *n_live, int *n_loaded);

/* ---- the hot path --------------------------------------------------------------------------
 * Replaces SynthesizerTrn.infer(x, x_lengths, sid, noise_scale, length_scale, noise_scale_w,
 * max_len)  (models.py:331-339).  sid / length_scale / noise_scale_w are ignored by the
 * reference and have no counterpart here.
 *
 *   mel      [B, n_mel, T]                 lengths [B] int64
 *   eps      [B, inter, T]  the N(0,1) draw of models.py:336 (caller draws it: SURVEY F11)
 *   max_len  <= 0 means None;  T' = max_len>0 ? min(T,max_len) : T
 *   o        [B, 1, hop*T']                x_mask  [B, 1, T]
 *   z, z_p, m_p, logs_p  [B, inter, T]     (each may be NULL when the caller does not want it)
 *
 * Calls on ONE handle must be ordered with respect to each other (same stream, or event-ordered streams): besides the
 * caller's workspace they use small device scratch the handle owns (the range flag of svk_check_range, the tile
 * progress counters of the multi-layer WN launches).  Use one handle per concurrent stream.
 */
size_t svk_workspace_bytes(const svk_handle *h, int B, int T, int max_len);
int svk_infer(svk_handle *h, const float *mel_dev, const int64_t *lengths_dev, const float *eps_dev,
              float noise_scale, int B, int T, int max_len, float *o_dev, float *x_mask_dev,
              float *z_dev, float *z_p_dev, float *m_p_dev, float *logs_p_dev, void *workspace_dev,
              size_t workspace_bytes, void *stream);
/* Same call with HOST buffers (pinned or pageable): H2D of mel/lengths/eps, infer, D2H of o and
 * x_mask (and latents when non-NULL), then synchronises.  The library owns and caches the device
 * buffers.  This is what inference.ipynb:114-118 does around infer (`.cuda()` ... `.cpu()`). */
int svk_infer_host(svk_handle *h, const float *mel, const int64_t *lengths, const float *eps,
                   float noise_scale, int B, int T, int max_len, float *o, float *x_mask, float *z,
                   float *z_p, float *m_p, float *logs_p);
/* Range guard.  The default engine stores activations as fp16 hi/lo pairs: an activation above 65504
 * becomes inf and reaches the waveform as NaN.  The final tanh epilogue (models.py:158) raises a device
 * flag for every output sample outside [-1, 1] (only a NaN can be).  svk_check_range synchronises
 * `stream`, reads and clears the flag: SVK_OK, or SVK_ERR_RANGE if any svk_infer / svk_generator call
 * on this handle since the last check produced one.  svk_infer_host checks before it returns.
 * No reference counterpart: fp32 PyTorch would print the same NaN audio silently. */
int svk_check_range(svk_handle *h, void *stream);
/* The configuration the handle was created with. */
int svk_get_config(const svk_handle *h, svk_config *out);

/* ---- CUDA-graph replay -----------------------------------------------------------------------
 * svk_infer issues ~135 kernels; at small batch (inference.ipynb:114-118 runs ONE utterance) the host's launch calls
 * are a visible part of the latency.  svk_graph_create runs svk_infer once on an internal stream, captures the same
 * launch sequence (programmatic dependent launch edges kept when the driver accepts them) into a CUDA graph bound
 * to buffers the graph owns, and svk_graph_launch replays it with one call.  The caller writes mel / lengths / eps
 * into the buffers svk_graph_buffers reports (stream-ordered before the launch) and reads o / x_mask / latents from
 * them afterwards.  Shapes, max_len and noise_scale are fixed per graph.  No reference counterpart. */
typedef struct svk_graph svk_graph;
typedef struct svk_graph_io {
  float *mel;        /* [B, n_mel, T]  in  */
  int64_t *lengths;  /* [B]            in  */
  float *eps;        /* [B, inter, T]  in  */
  float *o;          /* [B, 1, hop*T_out] out */
  float *x_mask;     /* [B, 1, T] */
  float *z, *z_p, *m_p, *logs_p; /* [B, inter, T] */
  int32_t B, T, T_out;
  int32_t programmatic_edges; /* 1: captured with programmatic dependent launch */
  int64_t kernel_nodes;
} svk_graph_io;
int svk_graph_create(svk_handle *h, int B, int T, int max_len, float noise_scale, svk_graph **out);
int svk_graph_buffers(const svk_graph *g, svk_graph_io *io);
int svk_graph_launch(svk_graph *g, void *stream);
void svk_graph_destroy(svk_graph *g);

/* ---- pipelined host entry (pinned D2H overlapped with compute) ---------------
 * svk_infer_host is serial: H2D, kernels, D2H, sync.  A pipeline keeps `depth` slots of device I/O buffers and
 * three streams: svk_pipeline_submit enqueues the H2D of this call (overlapping the kernels of the previous one),
 * its kernels, and its D2H (overlapping the kernels of the next one), and returns a ticket without waiting;
 * svk_pipeline_wait blocks until that ticket's PCM (and x_mask) are in host memory and returns SVK_ERR_RANGE if it
 * contains a non-finite sample.  Host buffers must stay valid (and should be pinned) until the wait.  Tickets are
 * waited in submission order; a ticket not waited for before `depth` further submits is waited for implicitly.
 * eps_host may be NULL: the N(0,1) draw of models.py:336 is then made on the device by svk_randn from
 * (seed, a per-pipeline counter).  Shapes and max_len are fixed per pipeline. */
typedef struct svk_pipeline svk_pipeline;
int svk_pipeline_create(svk_handle *h, int B, int T, int max_len, int depth, svk_pipeline **out);
int svk_pipeline_submit(svk_pipeline *p, const float *mel_host, const int64_t *lengths_host, const float *eps_host,
                        uint64_t seed, float noise_scale, float *o_host, float *x_mask_host, int64_t *ticket);
int svk_pipeline_wait(svk_pipeline *p, int64_t ticket);
int svk_pipeline_drain(svk_pipeline *p);
void svk_pipeline_destroy(svk_pipeline *p);
/* Counter-based N(0,1) generator (Philox4x32-10 + Box-Muller): element i = f(seed, offset + i/4).  Stands in for
 * torch.randn_like (models.py:336) where the caller does not bring eps; not torch's element order. */
int svk_randn(svk_handle *h, uint64_t seed, uint64_t offset, int64_t n, float *out_dev, void *stream);

/* ---- windowed / chunked synthesis (streaming and T >> 1024) ----------------
 * svk_infer_window computes the PCM of frames [t0, t1) of the utterance batch EXACTLY as the whole-
 * utterance svk_infer would: the window is widened by svk_halo_frames() frames of context per side
 * (WN stacks + flow + decoder receptive field; 110 for iitp_base.json), clipped at the true sequence
 * ends, and only the interior is stored.  mel / eps / lengths are the FULL tensors ([B,n_mel,T],
 * [B,inter,T], [B]); o receives [B, hop*(t1-t0)] with row stride o_row_stride floats (so a chunk can
 * be written in place into a full [B,1,hop*T] buffer, or into a buffer of its own for streaming);
 * the optional latents receive [B, inter, t1-t0] with row stride lat_row_stride.  Workspace:
 * svk_window_workspace_bytes(B, t1-t0) -- bounded by the chunk, not by T.
 * svk_infer_chunked walks the windows of `chunk_frames` frames over [0, T) and fills o / x_mask /
 * latents of the svk_infer layout.  No reference counterpart (the reference synthesises whole
 * utterances, models.py:331-339); parity = equality with svk_infer on the same inputs.  max_len as in
 * svk_infer (<= 0: None): windows must lie inside [0, min(T, max_len)); svk_infer_chunked returns latents
 * only when max_len does not shorten the output. */
int svk_halo_frames(const svk_handle *h);
size_t svk_window_workspace_bytes(const svk_handle *h, int B, int frames);
int svk_infer_window(svk_handle *h, const float *mel_dev, const int64_t *lengths_dev, const float *eps_dev,
                     float noise_scale, int B, int T, int max_len, int t0, int t1, float *o_dev, int64_t o_row_stride,
                     float *z_dev, float *z_p_dev, float *m_p_dev, float *logs_p_dev, int64_t lat_row_stride,
                     void *workspace_dev, size_t workspace_bytes, void *stream);
int svk_infer_chunked(svk_handle *h, const float *mel_dev, const int64_t *lengths_dev, const float *eps_dev,
                      float noise_scale, int B, int T, int max_len, int chunk_frames, float *o_dev, float *x_mask_dev,
                      float *z_dev, float *z_p_dev, float *m_p_dev, float *logs_p_dev, void *workspace_dev,
                      size_t workspace_bytes, void *stream);
/* Kernels launched by the most recent svk_infer / module call on this handle. */
int64_t svk_last_launch_count(const svk_handle *h);

/* ---- per-launch timing (bench.py's roofline leg) ---------------------------------------------
 * Between svk_profile_begin and svk_profile_end every kernel the handle launches is bracketed by
 * CUDA events on the launching stream.  svk_profile_end synchronises those events and returns one
 * record per launch with its ALGORITHMIC work (convolution MACs*2 as SURVEY 8(d) counts them;
 * compulsory bytes = inputs + outputs + weights of that launch).  No reference counterpart
 * (the reference has no profiler). */
#define SVK_LAYER_OTHER 0      /* mask / sample / flip */
#define SVK_LAYER_PRE_ENC 1
#define SVK_LAYER_WN_IN 2      /* k=5 conv + gate */
#define SVK_LAYER_WN_RES_SKIP 3
#define SVK_LAYER_PROJ 4
#define SVK_LAYER_FLOW_PRE 5
#define SVK_LAYER_FLOW_POST 6
#define SVK_LAYER_CONV_PRE 7
#define SVK_LAYER_UPSAMPLE 8
#define SVK_LAYER_RESBLOCK_CONV1 9   /* dilated */
#define SVK_LAYER_RESBLOCK_CONV2 10
#define SVK_LAYER_CONV_POST 11
#define SVK_LAYER_RESBLOCK_PAIR 12 /* fused convs1[l] + convs2[l] + residual of a ResBlock1 (narrow stages) */
#define SVK_LAYER_SPLIT_IMAGE 13   /* fp32 tensor -> operand image copy (pure duplicate traffic) */
#define SVK_LAYER_WN_LAYER 14      /* fused WN layer: in_layer k5 + gate + res_skip 1x1 + residual / skip update */
typedef struct svk_launch_record {
  int32_t layer;      /* SVK_LAYER_* */
  int32_t cin, cout;  /* logical channels */
  int32_t k, dilation;
  int32_t batch;
  int64_t length;     /* output positions per row */
  double flops;       /* algorithmic */
  double bytes;       /* algorithmic (compulsory) */
  float ms;           /* device time between the bracketing events */
  int32_t engine;     /* 0 = fp32 FFMA kernel, 1 = tcgen05 kernel */
  float gap_ms;       /* device time between the previous record's end event and this one's start */
  int32_t reserved;
  double dup_bytes;   /* part of `bytes` that exists only because a tensor is kept twice (fp32 + fp16 hi/lo operand
                         image of the same values): bytes - dup_bytes is what an fp32-only layer would move */
} svk_launch_record;
int svk_profile_begin(svk_handle *h, int max_records);
int svk_profile_end(svk_handle *h, svk_launch_record *out, int max_records, int *n_records);

/* ---- module-level entry points (use the handle's folded weights; device pointers) -----------
 * MelEncoder.forward (models.py:35-47): x_out/m/logs [B,hidden|inter,T], mask [B,1,T]. */
int svk_mel_encoder(svk_handle *h, const float *mel_dev, const int64_t *lengths_dev, int B, int T,
                    float *x_out_dev, float *m_dev, float *logs_dev, float *mask_dev,
                    void *workspace_dev, size_t workspace_bytes, void *stream);
/* ResidualCouplingBlock.forward(reverse=True) (models.py:73-80), in place on z [B,inter,T]. */
int svk_flow_reverse(svk_handle *h, float *z_dev, const float *mask_dev, int B, int T,
                     void *workspace_dev, size_t workspace_bytes, void *stream);
/* ---- analysis direction -------------------------------------------------
 * ResidualCouplingBlock.forward(reverse=False) (models.py:73-76), in place on z [B,inter,T]
 * (log-determinants are identically zero for mean_only couplings and are not returned, as in the
 * reference's block).  Workspace >= 5 * B * hidden * T floats. */
int svk_flow_forward(svk_handle *h, float *z_dev, const float *mask_dev, int B, int T,
                     void *workspace_dev, size_t workspace_bytes, void *stream);
/* PosteriorEncoder.forward(x, x_lengths, g=None) (models.py:103-110): spec [B,spec_channels,T] linear
 * spectrogram, eps [B,inter,T] = the randn_like draw of models.py:109 -> z = (m + eps*exp(logs))*mask,
 * m, logs [B,inter,T], mask [B,1,T].  Needs the enc_q.* keys (optional for svk_infer; SVK_ERR_STATE
 * if they were never loaded).  Workspace >= 5 * B * hidden * T floats. */
int svk_posterior_encoder(svk_handle *h, const float *spec_dev, const int64_t *lengths_dev,
                          const float *eps_dev, int B, int T, float *z_dev, float *m_dev, float *logs_dev,
                          float *mask_dev, void *workspace_dev, size_t workspace_bytes, void *stream);
/* Generator.forward(x, g=None) (models.py:141-160): z [B,inter,L] -> o [B,1,hop*L]. */
int svk_generator(svk_handle *h, const float *z_dev, int B, int L, float *o_dev, void *workspace_dev,
                  size_t workspace_bytes, void *stream);
/* ResBlock1.forward(x) (modules.py:210-223) -- or ResBlock2.forward(x) (modules.py:243-252) when the handle was created
 * with resblock_type 2 -- of dec.resblocks[index]: [B,C,L] -> [B,C,L]. */
size_t svk_resblock1_workspace_bytes(const svk_handle *h, int index, int B, int L);
int svk_resblock1(svk_handle *h, int index, const float *x_dev, int B, int L, float *y_dev,
                  void *workspace_dev, size_t workspace_bytes, void *stream);

/* ---- stateless operators (weights passed explicitly, reference layouts; device pointers) ----
 * nn.Conv1d stride 1: w [Cout,Cin,k], bias may be NULL,
 * pre_slope != 1 applies leaky_relu(x, pre_slope) to the input first (modules.py:212-217). */
int svk_conv1d(const float *x_dev, int B, int Cin, int L, const float *w_dev, const float *bias_dev,
               int Cout, int k, int dilation, int padding, float pre_slope, float *y_dev,
               void *stream);
/* nn.ConvTranspose1d (models.py:123-127): w [Cin,Cout,k]. */
int svk_conv_transpose1d(const float *x_dev, int B, int Cin, int L, const float *w_dev,
                         const float *bias_dev, int Cout, int k, int stride, int padding,
                         float pre_slope, float *y_dev, void *stream);
/* The same two operators on the tcgen05 engine (SVK_PRECISION_TC arithmetic: fp16 hi/lo split,
 * fp32 accumulate in TMEM).  Cin must be a multiple of 32.  Weight images are rebuilt per call
 * (these are parity-test entry points; the hot path packs once in svk_finalize_weights). */
int svk_conv1d_tc(const float *x_dev, int B, int Cin, int L, const float *w_dev, const float *bias_dev,
                  int Cout, int k, int dilation, int padding, float pre_slope, float *y_dev,
                  void *stream);
int svk_conv_transpose1d_tc(const float *x_dev, int B, int Cin, int L, const float *w_dev,
                            const float *bias_dev, int Cout, int k, int stride, int padding,
                            float pre_slope, float *y_dev, void *stream);
/* commons.sequence_mask(...).to(float) (commons.py:121-125, models.py:40): mask [B,T]. */
int svk_sequence_mask(const int64_t *lengths_dev, int B, int T, float *mask_dev, void *stream);
/* PCM egress: float waveform -> int16 samples, y = saturate(round_to_nearest_even(x * max_wav_value))
 * with max_wav_value = 32768 (configs/iitp_base.json:23), the inverse of `audio / 32768.0` (inference.ipynb cell 4); halves
 * the D2H bytes of a serving path that ships int16.  Both buffers 16 B-aligned. */
int svk_pcm_to_int16(const float *pcm_dev, int64_t n, float max_wav_value, int16_t *out_dev, void *stream);
/* modules.Flip (modules.py:270-277): y[b,c,t] = x[b,C-1-c,t]. Bit-exact copy. */
int svk_flip(const float *x_dev, int B, int C, int T, float *y_dev, void *stream);
/* torch.nn.utils.weight_norm fold: v [dim0, inner], g [dim0] -> w. */
int svk_weight_norm(const float *v_dev, const float *g_dev, int64_t dim0, int64_t inner,
                    float *w_dev, void *stream);
/* transforms.piecewise_rational_quadratic_transform(tails='linear') (transforms.py:12-193):
 * n elements, parameters contiguous per element ([n,nb], [n,nb], [n,nb-1]); bin_dev (may be
 * NULL) receives the searchsorted bin index, -1 outside [-tail_bound, tail_bound]. */
int svk_rq_spline(const float *x_dev, const float *uw_dev, const float *uh_dev, const float *ud_dev,
                  int64_t n, int num_bins, int inverse, float tail_bound, float min_bin_width,
                  float min_bin_height, float min_derivative, float *y_dev, float *logabsdet_dev,
                  int32_t *bin_dev, void *stream);

/* modules.ConvFlow.forward(x, x_mask, g=None, reverse) (modules.py:346-390) with its DDSConv (modules.py:70-108) and
 * LayerNorm (modules.py:20-32): the spline coupling flow.  The reference never instantiates it, so this is a
 * stateless operator with explicit weights in the reference's layouts, the per-layer tensors stacked along a leading
 * n_layers axis (device pointers):
 *   pre    Conv1d(C/2 -> F, 1)            convs_sep[i] Conv1d(F, F, k, groups=F, dilation k^i)    convs_1x1[i] Conv1d(F -> F, 1)
 *   norms_1[i], norms_2[i]  LayerNorm(F), eps 1e-5      proj   Conv1d(F -> C/2 * (3*num_bins - 1), 1)
 *   x [B, C, T], mask [B, T] -> y [B, C, T] = cat(x0, spline(x1)) * mask; logdet [B] = sum(logabsdet * mask) (may be NULL;
 *   the reference returns it for reverse = 0 only); bins [B, C/2, T] (may be NULL) = the searchsorted bin of every
 *   element (-1 in the linear tails) -- integer work, compared bit for bit by the tests. */
typedef struct svk_convflow_weights {
  const float *pre_w, *pre_b;       /* [F, C/2, 1], [F] */
  const float *sep_w, *sep_b;       /* [n_layers, F, 1, k], [n_layers, F] */
  const float *pw_w, *pw_b;         /* [n_layers, F, F, 1], [n_layers, F] */
  const float *norm1_g, *norm1_b;   /* [n_layers, F] gamma / beta of norms_1 */
  const float *norm2_g, *norm2_b;   /* [n_layers, F] */
  const float *proj_w, *proj_b;     /* [C/2 * (3*num_bins - 1), F, 1], [C/2 * (3*num_bins - 1)] */
} svk_convflow_weights;
size_t svk_convflow_workspace_bytes(int B, int C, int T, int filter_channels, int num_bins);
int svk_convflow(const float *x_dev, const float *mask_dev, int B, int C, int T, int filter_channels, int kernel_size,
                 int n_layers, int num_bins, float tail_bound, const svk_convflow_weights *w, int reverse, float *y_dev,
                 float *logdet_dev, int32_t *bins_dev, void *workspace_dev, size_t workspace_bytes, void *stream);

/* ---- mel front-end (the step before infer in both callers) -----------------
 * Replaces mel_processing.spectrogram_torch (mel_processing.py:51-69), spec_to_mel_torch (:72-81) and
 * mel_spectrogram_torch (:84-112) as inference.ipynb:100-111 / train.py:266-272 use them (center=False).
 *   y    [B, n_samples] fp32 waveform in [-1, 1]      T = svk_frontend_frames(n_samples)
 *   spec [B, n_fft/2+1, T] = sqrt(re^2 + im^2 + 1e-6) of the Hann(win_size) STFT, hop `hop_size`, over
 *        the signal reflect-padded by (n_fft - hop_size)/2 per side
 *   mel  [B, n_mels, T]    = log(clamp(mel_basis @ spec, 1e-5))
 * fmax <= 0 means None (sampling_rate / 2), as in configs/iitp_base.json.  The handle owns the window,
 * twiddle and mel-basis tables on its device; calls are stream-ordered, no allocation, no sync. */
typedef struct svk_frontend svk_frontend;
int svk_frontend_create(int n_fft, int hop_size, int win_size, int sampling_rate, int n_mels, float fmin,
                        float fmax, int device, svk_frontend **out);
void svk_frontend_destroy(svk_frontend *f);
int64_t svk_frontend_frames(const svk_frontend *f, int64_t n_samples);
int svk_spectrogram(svk_frontend *f, const float *y_dev, int B, int64_t n_samples, float *spec_dev,
                    void *stream);
int svk_spec_to_mel(svk_frontend *f, const float *spec_dev, int B, int64_t T, float *mel_dev, void *stream);
/* Fused: the linear spectrogram stays in shared memory (spec_dev may be NULL; non-NULL also stores it). */
int svk_mel_spectrogram(svk_frontend *f, const float *y_dev, int B, int64_t n_samples, float *mel_dev,
                        float *spec_dev, void *stream);
/* Host-side table builders (no device needed): librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) with its
 * defaults htk=False, norm="slaney" (mel_processing.py:76,95; third-party, restated) -> [n_mels, n_fft/2+1];
 * torch.hann_window(win_size) in fp32 centred in n_fft taps (mel_processing.py:59-60) -> [n_fft]. */
int svk_mel_basis(int sampling_rate, int n_fft, int n_mels, float fmin, float fmax, float *basis_host);
int svk_hann_window(int win_size, int n_fft, float *window_host);

#ifdef __cplusplus
}
#endif
#endif /* SVK_H_ */
