// One launch = one layer of modules.WN.forward (reference modules.py:156-175, g = None):
//
//     x_in = in_layer(x)                      Conv1d(H -> 2H, k, padding (k-1)/2)
//     acts = tanh(x_in[:H]) * sigmoid(x_in[H:])        commons.fused_add_tanh_sigmoid_multiply, g_l = 0
//     rs   = res_skip(acts)                   Conv1d(H -> 2H, 1)   (last layer: H -> H)
//     x    = (x + rs[:H]) * mask ; out += rs[H:]        (last layer: out = (out + rs) * mask)
//
// `acts` never leaves the SM: the gate epilogue writes it as the A operand of the 1x1 conv straight into shared
// memory (fp16 hi/lo, 64 B-swizzled rows, fence.proxy.async), and the second GEMM runs on that tile.  Unfused this was
// two launches per layer (wn_in + wn_res_skip, 49 + 35 us at 16 x 1024 frames) with `acts` written to and read back
// from HBM as an operand image; fused, HBM sees x (image in, fp32 + image out), out and the weights.
//
// Same machinery as conv_tc.cu (three-product fp16 split, taps as row shifts of one 64 B-swizzled tile, TMA-fed
// operand images, weight stages by cp.async.bulk, accumulators in TMEM, warp roles, PDL).  One CTA owns one tile of
// 128 frames through BOTH GEMMs:
//     in_layer   : N-tiles nt = 0..2 of 128 virtual channels (64 tanh + the matching 64 sigmoid channels, packed that
//                  way by pack_tc(gate_half)), K = H x taps; two accumulator stages (main + cross, 2 x 256 columns =
//                  all of TMEM), so the gate epilogue of N-tile nt runs under the MMAs of nt + 1
//     res_skip   : N-tiles mt of 128 (96 on the last layer), K = H, A = the acts tile; its first chunks are issued as
//                  soon as the gate epilogues that produce them have finished (one mbarrier per N-tile of in_layer)
// The x image is ping-ponged between two buffers: neighbouring CTAs read halo rows of the OLD x while this CTA writes
// the new one.
//
// A launch covers n_layers consecutive layers.  n_layers == 1 is the one-launch-per-layer form.  With n_layers > 1 (the
// whole stack: `items` <= SM count, one tile per CTA, all CTAs co-resident) a CTA keeps its tile through every layer:
// barriers, TMEM, the weight ring (which simply keeps streaming: the next layer's weights arrive under this layer's
// tail) and the role loops carry on, and the only thing a layer waits for is the 2-frame halo of its two neighbour
// tiles -- tile i starts layer l + 1 once tiles i - 1, i, i + 1 have published `flags[] > l` (epilogue stores ->
// __threadfence -> named barrier -> st.release.gpu; the TMA thread polls with ld.acquire.gpu, then fence.proxy.async).
// That removes the per-layer launch, prologue (barrier init, TMEM allocation, bias staging) and the grid-wide join
// at every layer boundary: a fast tile is at most one layer ahead of its neighbours instead of waiting for the
// slowest tile of the grid.
#include <string.h>

#include "svk_kernels.cuh"
#include "tc_common.cuh"

namespace svk {

namespace {

constexpr int WN_NA_MAX = 4, WN_NW_MAX = 8, WN_MAX_NT = 8;
constexpr int WN_EPI_WARPS = 8;
constexpr int WN_THREADS = 128 + 32 * WN_EPI_WARPS;
constexpr int WN_EPI_THREADS = 32 * WN_EPI_WARPS;

struct __align__(8) WnHeader {
  uint64_t a_full[WN_NA_MAX], a_empty[WN_NA_MAX];
  uint64_t w_full[WN_NW_MAX], w_empty[WN_NW_MAX];
  uint64_t acc_full[2], acc_empty[2];
  uint64_t acts_full[WN_MAX_NT];
  uint64_t x_ready;  // multi-layer launches: this CTA's epilogue has stored the layer's x image
  uint32_t tmem_base;
  uint32_t pad;
};
constexpr int WN_HEADER_BYTES = 512;
static_assert(sizeof(WnHeader) <= WN_HEADER_BYTES, "header");

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// $SVK_WN_TRACE: event `ev` of layer `l` of this CTA's tile (one thread per role writes)
__device__ __forceinline__ void wn_stamp(const WnLayerArgs& wa, int l, int ev) {
  if (wa.trace) wa.trace[((size_t)blockIdx.x * wa.n_layers + l) * 32 + ev] = clock64();
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(WN_EPI_THREADS) : "memory"); }

__global__ void __launch_bounds__(WN_THREADS, 1)
    wn_layer_kernel(const __grid_constant__ WnLayerArgs wa, const __grid_constant__ CUtensorMap tmap0, const __grid_constant__ CUtensorMap tmap1) {
  extern __shared__ __align__(128) uint8_t smem[];
  WnHeader* hdr = reinterpret_cast<WnHeader*>(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = wa.H, K = wa.K, planes = wa.planes;
  const int nchunks = H / KC;
  const int N1 = wa.N_in, nt_in = wa.nt_in;
  const int nl = wa.n_layers;
  const int na = wa.na, nw = wa.nw;
  const uint32_t a_plane = (uint32_t)wa.rows * 64u, a_stage = a_plane * planes;
  const uint32_t acts_plane = 128u * 64u, acts_chunk = acts_plane * planes;
  const uint32_t w1_plane2 = (uint32_t)N1 * 16u * planes, w1_stage = w1_plane2 * KG;
  const uint32_t w_slot = wa.w_slot;
  const int acc_stride = wa.acc_stride;  // TMEM columns between the two accumulator stages
  float* bias_s = reinterpret_cast<float*>(smem + WN_HEADER_BYTES);  // in_layer bias (virtual order), then res_skip bias
  float* bias2_s = bias_s + wa.bias_count_in;
  uint8_t* a_smem = smem + wa.a_off;
  uint8_t* acts_smem = smem + wa.acts_off;
  uint8_t* w_smem = smem + wa.w_off;
  const int items = wa.items;
  const int n_my = (items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // 1 when nl > 1

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < na; ++i) mbar_init(&hdr->a_full[i], 1), mbar_init(&hdr->a_empty[i], 1);
    for (int i = 0; i < WN_NW_MAX; ++i) mbar_init(&hdr->w_full[i], 1), mbar_init(&hdr->w_empty[i], 1);
    for (int i = 0; i < 2; ++i) mbar_init(&hdr->acc_full[i], 1), mbar_init(&hdr->acc_empty[i], WN_EPI_THREADS);
    for (int i = 0; i < WN_MAX_NT; ++i) mbar_init(&hdr->acts_full[i], WN_EPI_THREADS);
    mbar_init(&hdr->x_ready, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&hdr->tmem_base, (uint32_t)wa.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = hdr->tmem_base;
  griddep_launch_dependents();  // PDL: only weights / biases are touched before griddep_wait()

  if (warp == 0) {
    // ------------------------------------------------ weight producer: one bulk copy per (N-tile, chunk, tap)
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      for (int i = 0; i < n_my; ++i)
        for (int l = 0; l < nl; ++l) {
          const WnLayerParams& lp = wa.layer[l];
          const uint8_t* s1 = reinterpret_cast<const uint8_t*>(lp.w_in);
          const uint8_t* s2 = reinterpret_cast<const uint8_t*>(lp.w_rs);
          const uint32_t w2_stage = (uint32_t)lp.N_rs * 16u * planes * KG;
          const int n1 = nt_in * nchunks * K, n2 = lp.nt_rs * nchunks;
          for (int it = 0; it < n1 + n2; ++it) {
            const bool first = it < n1;
            const uint32_t bytes = first ? w1_stage : w2_stage;
            const uint8_t* src = first ? s1 + (size_t)it * w1_stage : s2 + (size_t)(it - n1) * w2_stage;
            mbar_wait(&hdr->w_empty[st], ph ^ 1);
            mbar_arrive_expect_tx(&hdr->w_full[st], bytes);
            bulk_g2s(w_smem + (size_t)st * w_slot, src, bytes, &hdr->w_full[st]);
            if (++st == nw) st = 0, ph ^= 1;
          }
        }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (whole warp walks the loop, one elected lane issues)
    const uint32_t fmt = planes == 1 ? ((1u << 7) | (1u << 10)) : 0u;  // bf16 x bf16 for the single-plane engine
    auto idesc = [&](int n) { return (1u << 4) | fmt | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24); };
    const uint32_t id1_main = idesc(N1), id1_wide = idesc(2 * N1);
    const uint32_t b_hi = (128u >> 4) | (1u << 14);               // B: no swizzle, SBO = 128 B
    const uint32_t a_hi = (512u >> 4) | (1u << 14) | (4u << 29);  // A: SWIZZLE_64B, SBO = 8 rows x 64 B
    const uint32_t a_lo0 = ((smem_u32(a_smem) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t acts_lo0 = ((smem_u32(acts_smem) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t w_addr16 = (smem_u32(w_smem) & 0x3FFFFu) >> 4;
    const uint32_t a_stage16 = a_stage >> 4, acts_chunk16 = acts_chunk >> 4, w_slot16 = w_slot >> 4;
    const uint32_t lo1_16 = a_plane >> 4, lo2_16 = acts_plane >> 4;
    const uint32_t bar_a_full = smem_u32(&hdr->a_full[0]), bar_a_empty = smem_u32(&hdr->a_empty[0]);
    const uint32_t bar_w_full = smem_u32(&hdr->w_full[0]), bar_w_empty = smem_u32(&hdr->w_empty[0]);
    const uint32_t bar_acc_full = smem_u32(&hdr->acc_full[0]), bar_acc_empty = smem_u32(&hdr->acc_empty[0]);
    const uint32_t bar_acts_full = smem_u32(&hdr->acts_full[0]);
    const bool leader = elect_one();
    int ast = 0, wst = 0;
    uint32_t aph = 0, wph = 0;
    uint32_t q = 0;  // accumulation counter: stage q & 1, phase (q >> 1) & 1

    // one (chunk, tap): two K = 16 steps of [main | cross] (+)= xh . [wh | wl] ; cross += xl . wh
    auto tap = [&](uint32_t dmain, int N, uint32_t ah, uint32_t lo16, uint32_t bw, uint32_t ks_b16, uint32_t id_main,
                   uint32_t id_wide, uint32_t acc) {
      if (planes == 2) {
        umma_f16_lo(dmain, ah, bw, a_hi, b_hi, id_wide, acc);
        umma_f16_lo(dmain + (uint32_t)N, ah + lo16, bw, a_hi, b_hi, id_main, 1u);
        umma_f16_lo(dmain, ah + 2u, bw + ks_b16, a_hi, b_hi, id_wide, 1u);
        umma_f16_lo(dmain + (uint32_t)N, ah + 2u + lo16, bw + ks_b16, a_hi, b_hi, id_main, 1u);
      } else {
        umma_f16_lo(dmain, ah, bw, a_hi, b_hi, id_main, acc);
        umma_f16_lo(dmain, ah + 2u, bw + ks_b16, a_hi, b_hi, id_main, 1u);
      }
    };
    uint32_t lq = 0;  // (item, layer) counter: parity of the acts_full barriers
    for (int i = 0; i < n_my; ++i)
      for (int l = 0; l < nl; ++l, ++lq) {
        const int N2 = wa.layer[l].N_rs, nt_rs = wa.layer[l].nt_rs;
        const uint32_t w2_plane2 = (uint32_t)N2 * 16u * planes;
        const uint32_t id2_main = idesc(N2), id2_wide = idesc(2 * N2);
        // ---- in_layer: N-tiles of N1 virtual channels, K = H x taps
        for (int nt = 0; nt < nt_in; ++nt, ++q) {
          const uint32_t s = q & 1u;
          mbar_wait_u32(bar_acc_empty + 8u * s, ((q >> 1) & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t dmain = tmem + s * (uint32_t)acc_stride;
          uint32_t acc = 0;
          for (int ch = 0; ch < nchunks; ++ch) {
            mbar_wait_u32(bar_a_full + 8u * ast, aph);
            tc_fence_after();
            if (leader && nt == 0 && ch == 0) wn_stamp(wa, l, 4);
            uint32_t ah = a_lo0 + (uint32_t)ast * a_stage16;
            for (int j = 0; j < K; ++j) {
              mbar_wait_u32(bar_w_full + 8u * wst, wph);
              tc_fence_after();
              if (leader) {
                const uint32_t bw = (w_addr16 + (uint32_t)wst * w_slot16) | ((w1_plane2 >> 4) << 16);
                tap(dmain, N1, ah, lo1_16, bw, (2 * w1_plane2) >> 4, id1_main, id1_wide, acc);
                umma_commit_u32(bar_w_empty + 8u * wst);
              }
              acc = 1u;
              ah += 4u;  // dilation 1: next tap = next 64 B row
              if (++wst == nw) wst = 0, wph ^= 1;
            }
            if (leader) umma_commit_u32(bar_a_empty + 8u * ast);
            if (++ast == na) ast = 0, aph ^= 1;
          }
          if (leader) umma_commit_u32(bar_acc_full + 8u * s);
        }
        if (leader) wn_stamp(wa, l, 5);
        // ---- res_skip: N-tiles of N2 channels, K = H, A = the acts tile the gate epilogues wrote
        for (int mt = 0; mt < nt_rs; ++mt, ++q) {
          const uint32_t s = q & 1u;
          mbar_wait_u32(bar_acc_empty + 8u * s, ((q >> 1) & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t dmain = tmem + s * (uint32_t)acc_stride;
          uint32_t acc = 0;
          for (int ch = 0; ch < nchunks; ++ch) {
            // acts chunk ch comes from in_layer N-tile (ch * 32) / (N1 / 2); completes once per (item, layer)
            mbar_wait_u32(bar_acts_full + 8u * (uint32_t)((ch * KC) / (N1 >> 1)), lq & 1u);
            mbar_wait_u32(bar_w_full + 8u * wst, wph);
            tc_fence_after();
            if (leader && mt == 0 && ch == nchunks - 1) wn_stamp(wa, l, 6);
            if (leader) {
              const uint32_t ah = acts_lo0 + (uint32_t)ch * acts_chunk16;
              const uint32_t bw = (w_addr16 + (uint32_t)wst * w_slot16) | ((w2_plane2 >> 4) << 16);
              tap(dmain, N2, ah, lo2_16, bw, (2 * w2_plane2) >> 4, id2_main, id2_wide, acc);
              umma_commit_u32(bar_w_empty + 8u * wst);
            }
            acc = 1u;
            if (++wst == nw) wst = 0, wph ^= 1;
          }
          if (leader) umma_commit_u32(bar_acc_full + 8u * s);
        }
        if (leader) wn_stamp(wa, l, 7);
      }
    __syncwarp();
  } else if (warp == 2) {
    // ------------------------------------------------ x-image loader: one TMA box per 32-channel chunk and N-tile
    if (lane == 0) {
      griddep_wait();
      int as = 0;
      uint32_t ph = 0;
      const int cgs = H >> 5, pad = (K - 1) / 2;
      for (int i = 0; i < n_my; ++i) {
        const int item = (int)blockIdx.x + i * (int)gridDim.x;
        const int b = (int)fast_div((uint32_t)item, wa.div_t), tt = item - b * wa.ntiles_t;
        for (int l = 0; l < nl; ++l) {
          wn_stamp(wa, l, 0);
          if (l > 0) {
            // layer l reads the image layer l - 1 wrote: this tile's rows (x_ready: our own epilogue has stored them) and
            // the halo rows of the two neighbour tiles of the same utterance (their flags)
            mbar_wait(&hdr->x_ready, (uint32_t)(l - 1) & 1u);
            wn_stamp(wa, l, 1);
            for (int nb = -1; nb <= 1; nb += 2) {
              if (tt + nb < 0 || tt + nb >= wa.ntiles_t) continue;
              const int* f = wa.flags + item + nb;
              long long t0 = 0;
              for (uint32_t spin = 0; ld_acquire_gpu(f) < l; ++spin) {
                if ((spin & 0xFFF) == 0xFFF) {
                  const long long now = clock64();
                  if (t0 == 0) t0 = now;
                  else if (now - t0 > 4000000000LL) __trap();  // a neighbour that never arrives: fail, do not hang
                }
              }
            }
            fence_proxy_async_all();  // the generic-proxy stores we have just acquired -> visible to the TMA (async proxy) reads
          }
          wn_stamp(wa, l, 2);
          const CUtensorMap* map = (l & 1) ? &tmap1 : &tmap0;
          for (int nt = 0; nt < nt_in; ++nt)
            for (int ch = 0; ch < nchunks; ++ch) {
              mbar_wait(&hdr->a_empty[as], ph ^ 1);
              mbar_arrive_expect_tx(&hdr->a_full[as], a_stage);
              tma_load_4d(a_smem + (size_t)as * a_stage, map, 0, tt * 128 - pad, b * cgs + ch, 0, &hdr->a_full[as]);
              if (++as == na) as = 0, ph ^= 1;
            }
          wn_stamp(wa, l, 3);
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------------------ epilogue warps: gate -> acts tile ; res_skip -> x, out
    griddep_wait();
    const int q4 = warp & 3, part = (warp - 4) >> 2;  // TMEM lane quarter; which half of a tile's columns
    const int row = q4 * 32 + lane;
    const int T = wa.T;
    const size_t img_plane = (size_t)wa.B * H * T;  // halves between the hi and lo planes of an x image
    const int hN = N1 >> 1;                          // gate channels per in_layer N-tile
    uint32_t q = 0;
    const int etid = tid - 128;  // 0 .. WN_EPI_THREADS - 1
    for (int i = 0; i < n_my; ++i) {
      const int item = (int)blockIdx.x + i * (int)gridDim.x;
      const int b = (int)fast_div((uint32_t)item, wa.div_t), tt = item - b * wa.ntiles_t;
      const int t = tt * 128 + row;
      const bool tin = t < T;
      const int tl = tin ? t : T - 1;
      const float mv = __ldg(wa.mask + (size_t)b * T + tl);
     for (int l = 0; l < nl; ++l) {
      const WnLayerParams& lp = wa.layer[l];
      const int N2 = lp.N_rs, nt_rs = lp.nt_rs, Cout_rs = lp.Cout_rs;
      const bool first = wa.l0 + l == 0, last = wa.l0 + l == wa.n_total - 1;
      uint16_t* const x_img_out = last ? nullptr : wa.img[(l + 1) & 1];
      // ---- this layer's biases -> shared memory (the epilogue warps only: nobody else reads them).  The barrier in
      // front closes the previous layer: nobody reads its biases any more.
      if (i + l > 0) epi_bar_sync();
      for (int k = etid; k < wa.bias_count_in; k += WN_EPI_THREADS) bias_s[k] = __ldg(lp.bias_in + k);
      for (int k = etid; k < lp.bias_count_rs; k += WN_EPI_THREADS) bias2_s[k] = __ldg(lp.bias_rs + k);
      epi_bar_sync();
      // ---- gate epilogues: acts[c] = tanh(a[c]) * sigmoid(a[c + H]) -> fp16 hi/lo rows of the acts tile
      for (int nt = 0; nt < nt_in; ++nt, ++q) {
        const uint32_t s = q & 1u;
        mbar_wait(&hdr->acc_full[s], (q >> 1) & 1u);
        tc_fence_after();
        if (etid == 0) wn_stamp(wa, l, nt == 0 ? 8 : (nt == nt_in - 1 ? 13 : 14));
        const uint32_t tsub = tmem + ((uint32_t)(q4 * 32) << 16) + s * (uint32_t)acc_stride;
        const float* bptr = bias_s + nt * N1;
        const int npair = hN >> 4, hp = (npair + 1) >> 1;  // 16-channel jobs of this tile, split between the two parts
        for (int jb = part * hp; jb < min(npair, (part + 1) * hp); ++jb) {
          const int n0 = jb * 16;
          uint32_t m[16], c[16];
          float g[16];
          tmem_ld16(tsub + (uint32_t)n0, m);
          if (planes == 2) {
            tmem_ld16(tsub + (uint32_t)(N1 + n0), c);
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) c[e] = 0u;
          }
          tmem_wait_ld();
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            const float4 bq = reinterpret_cast<const float4*>(bptr + n0)[e4];
            const float bb[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
              g[4 * e4 + e] = fmaf(fmaf(__uint_as_float(c[4 * e4 + e]), LO_INV, __uint_as_float(m[4 * e4 + e])), lp.unscale_in, bb[e]);  // tanh side, pre-activation
          }
          tmem_ld16(tsub + (uint32_t)(hN + n0), m);
          if (planes == 2) tmem_ld16(tsub + (uint32_t)(N1 + hN + n0), c);
          tmem_wait_ld();
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            const float4 bq = reinterpret_cast<const float4*>(bptr + hN + n0)[e4];
            const float bb[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
              g[4 * e4 + e] = gate_tanh_sigmoid(g[4 * e4 + e], fmaf(fmaf(__uint_as_float(c[4 * e4 + e]), LO_INV, __uint_as_float(m[4 * e4 + e])), lp.unscale_in, bb[e]));
          }
          // 16 channels = two 16 B pieces of this row of chunk (channel / 32), per plane
          const int ch0 = nt * hN + n0;
          uint8_t* tile = acts_smem + (size_t)(ch0 >> 5) * acts_chunk;
          uint4* hi = reinterpret_cast<uint4*>(tile) + row * KG;
          uint4* lo = reinterpret_cast<uint4*>(tile + acts_plane) + row * KG;
          const int swz = (int)((smem_u32(hi) >> 7) & 3u);  // 64 B swizzle on absolute address bits 7-8
          const int kg0 = (ch0 & 31) >> 3;
#pragma unroll
          for (int g8 = 0; g8 < 2; ++g8) {
            if (planes == 2) {
              uint4 hq, lq;
              split2(g[8 * g8 + 0], g[8 * g8 + 1], hq.x, lq.x);
              split2(g[8 * g8 + 2], g[8 * g8 + 3], hq.y, lq.y);
              split2(g[8 * g8 + 4], g[8 * g8 + 5], hq.z, lq.z);
              split2(g[8 * g8 + 6], g[8 * g8 + 7], hq.w, lq.w);
              hi[(kg0 + g8) ^ swz] = hq;
              lo[(kg0 + g8) ^ swz] = lq;
            } else {
              hi[(kg0 + g8) ^ swz] = pack_bf16x8(&g[8 * g8]);
            }
          }
        }
        fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
        tc_fence_before();
        mbar_arrive(&hdr->acts_full[nt]);
        mbar_arrive(&hdr->acc_empty[s]);
        if (etid == 0 && nt == nt_in - 1) wn_stamp(wa, l, 9);
      }
      // ---- res_skip epilogues: x = (x + rs[:H]) * mask (+ its operand image), out += rs[H:]
      const bool publish = nl > 1 && l < nl - 1;  // another layer of this launch follows (never the stack's last layer)
      bool x_done = false;
      // The two warps of a lane quarter take alternate 16-column jobs (job = part + 2 k): the x half of rs, which the next
      // layer waits for, is then shared evenly between them (contiguous halves gave one warp 8 of its 12 jobs).
      // The residual / running-sum operands of a job do not depend on the accumulator: those of the first N-tile are
      // requested before it is waited for, and each job, once its operands are consumed, requests the operands of the
      // same job of the NEXT N-tile into the same registers -- a tile's operands are in flight for a whole tile.
      const int nch = N2 >> 4;
      constexpr int MAXJ = 4;  // N2 <= 128: at most 4 jobs of 16 columns per part
      float r[MAXJ][16];
      auto load_ops = [&](int mt_, int jb, float (&o)[16]) {
        const int o0 = mt_ * N2 + jb * 16;
        const bool res_side = !last && o0 < H;
        const float* src = nullptr;
        if (o0 < Cout_rs) {
          if (res_side) src = wa.x + ((size_t)b * H + o0) * T + tl;
          else if (!first) src = wa.out + ((size_t)b * H + (last ? o0 : o0 - H)) * T + tl;
        }
        if (src) {
#pragma unroll
          for (int e = 0; e < 16; ++e) o[e] = src[(size_t)e * T];
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) o[e] = 0.f;
        }
      };
#pragma unroll
      for (int k = 0; k < MAXJ; ++k)
        if (part + 2 * k < nch) load_ops(0, part + 2 * k, r[k]);
      for (int mt = 0; mt < nt_rs; ++mt, ++q) {
        const uint32_t s = q & 1u;
        mbar_wait(&hdr->acc_full[s], (q >> 1) & 1u);
        tc_fence_after();
        if (etid == 0 && mt == 0) wn_stamp(wa, l, 10);
        const uint32_t tsub = tmem + ((uint32_t)(q4 * 32) << 16) + s * (uint32_t)acc_stride;
#pragma unroll
        for (int k = 0; k < MAXJ; ++k) {
          const int jb = part + 2 * k;
          if (jb >= nch) break;
          const int n0 = jb * 16, o0 = mt * N2 + n0;
          uint32_t m[16], c[16];
          tmem_ld16(tsub + (uint32_t)n0, m);
          if (planes == 2) {
            tmem_ld16(tsub + (uint32_t)(N2 + n0), c);
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) c[e] = 0u;
          }
          tmem_wait_ld();
          if (etid == 0 && mt < 2) wn_stamp(wa, l, 16 + mt * 8 + 2 * k);      // job k of N-tile mt: accumulators in registers
          if (o0 < Cout_rs) {
            float v[16];
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
              const float4 bq = reinterpret_cast<const float4*>(bias2_s + o0)[e4];
              const float bb[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
              for (int e = 0; e < 4; ++e)
                v[4 * e4 + e] = fmaf(fmaf(__uint_as_float(c[4 * e4 + e]), LO_INV, __uint_as_float(m[4 * e4 + e])), lp.unscale_rs, bb[e]) + r[k][4 * e4 + e];
            }
            if (mt + 1 < nt_rs) load_ops(mt + 1, jb, r[k]);  // this job's operands of the next N-tile
            const bool res_side = !last && o0 < H;
            if (res_side || last) {
#pragma unroll
              for (int e = 0; e < 16; ++e) v[e] *= mv;
            }
            if (tin) {
              float* dst = res_side ? wa.x + ((size_t)b * H + o0) * T + t : wa.out + ((size_t)b * H + (last ? o0 : o0 - H)) * T + t;
#pragma unroll
              for (int e = 0; e < 16; ++e) dst[(size_t)e * T] = v[e];
              if (res_side && x_img_out) {
                uint16_t* sp = x_img_out + (((size_t)b * (H >> 5) + (o0 >> 5)) * T + t) * 32 + (o0 & 31);
                uint4 h2[2], l2[2];
#pragma unroll
                for (int g8 = 0; g8 < 2; ++g8) {
                  if (planes == 2) {
                    split2(v[8 * g8 + 0], v[8 * g8 + 1], h2[g8].x, l2[g8].x);
                    split2(v[8 * g8 + 2], v[8 * g8 + 3], h2[g8].y, l2[g8].y);
                    split2(v[8 * g8 + 4], v[8 * g8 + 5], h2[g8].z, l2[g8].z);
                    split2(v[8 * g8 + 6], v[8 * g8 + 7], h2[g8].w, l2[g8].w);
                  } else {
                    h2[g8] = pack_bf16x8(&v[8 * g8]);
                  }
                }
                st_global_v8(sp, h2[0], h2[1]);  // 16 channels = one 32 B sector per plane
                if (planes == 2) st_global_v8(sp + img_plane, l2[0], l2[1]);
              }
            }
          }
          if (etid == 0 && mt < 2) wn_stamp(wa, l, 17 + mt * 8 + 2 * k);      // ... and stored
        }
        tc_fence_before();
        mbar_arrive(&hdr->acc_empty[s]);
        // Multi-layer launch: the next layer only needs the new x (channels < H of rs); the skip half (out += rs[H:]) can
        // finish under the next layer's first MMAs.  A warp whose remaining jobs of this layer are all on the skip side
        // fences its x / image stores (device-wide for the neighbour CTAs, and towards the async proxy for the TMA
        // reads) and arrives -- without blocking -- on named barrier 2; warp 4, itself on the x side until N-tile
        // (H - 1) / N2, waits there for all of them and its first thread publishes the tile.
        if (publish && !x_done && (mt + 1) * N2 + part * 16 >= H) {
          x_done = true;
          __threadfence();
          fence_proxy_async_all();
          if (warp == 4) {  // the publishing thread's warp waits (a barrier instruction counts whole, converged warps)
            asm volatile("bar.sync 2, %0;" ::"n"(WN_EPI_THREADS) : "memory");
            if (etid == 0) {
              st_release_gpu(wa.flags + item, l + 1);  // l + 1 layers of this tile's x are complete: neighbours may read its halo
              mbar_arrive(&hdr->x_ready);              // ... and so may our own TMA thread
              wn_stamp(wa, l, 12);
            }
            __syncwarp();
          } else {
            asm volatile("bar.arrive 2, %0;" ::"n"(WN_EPI_THREADS) : "memory");
          }
        }
      }
      if (etid == 0) wn_stamp(wa, l, 11);
     }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, (uint32_t)wa.tmem_cols);
}

}  // namespace

// Shared-memory plan: header | biases | A ring (x tile chunks) | acts tile | weight ring, sized for the widest layer of
// the launch.  false when a layer does not fit (the caller then runs the two-launch form).
static bool wn_plan(WnLayerArgs& wa, size_t* smem_bytes) {
  const int H = wa.H, planes = wa.planes, nchunks = H / KC;
  if (wa.n_layers < 1 || wa.n_layers > WN_MAX_LAYERS) return false;
  if (H % KC || wa.K < 1 || !(wa.K & 1) || wa.N_in % 32 || wa.N_in > 128 || wa.N_in < 32) return false;
  wa.nt_in = (2 * H + wa.N_in - 1) / wa.N_in;
  // every 32-channel chunk of the acts tile must come from ONE in_layer N-tile (one acts_full barrier per N-tile)
  if (wa.nt_in > WN_MAX_NT || (H % (wa.N_in / 2)) != 0 || (wa.N_in / 2) % KC != 0) return false;
  wa.rows = conv_tc_rows(wa.K, 1);
  wa.bias_count_in = wa.nt_in * wa.N_in;
  int nmax = wa.N_in, nrs_max = 0;
  wa.bias_max_rs = 0;
  for (int l = 0; l < wa.n_layers; ++l) {
    WnLayerParams& lp = wa.layer[l];
    if (lp.N_rs % 16 || lp.N_rs > 128 || lp.N_rs < 16 || lp.Cout_rs < 1) return false;
    lp.nt_rs = (lp.Cout_rs + lp.N_rs - 1) / lp.N_rs;
    lp.bias_count_rs = lp.nt_rs * lp.N_rs;
    if (lp.bias_count_rs > wa.bias_max_rs) wa.bias_max_rs = lp.bias_count_rs;
    if (lp.N_rs > nmax) nmax = lp.N_rs;
    if (lp.N_rs > nrs_max) nrs_max = lp.N_rs;
  }
  wa.acc_stride = planes * nmax;
  int cols = 32;
  while (cols < 2 * wa.acc_stride) cols <<= 1;
  if (cols > 512) return false;
  wa.tmem_cols = cols;
  const size_t fixed = (WN_HEADER_BYTES + (size_t)(wa.bias_count_in + wa.bias_max_rs) * 4 + 1023) & ~(size_t)1023;
  const size_t a_stage = (size_t)wa.rows * 64 * planes, acts = (size_t)nchunks * 128 * 64 * planes;
  const size_t w1 = (size_t)wa.N_in * 16 * planes * KG, w2 = (size_t)nrs_max * 16 * planes * KG;
  wa.w_slot = (int)(w1 > w2 ? w1 : w2);
  const size_t budget = 227 * 1024;
  wa.na = 2;
  if (fixed + wa.na * a_stage + acts + 3 * (size_t)wa.w_slot > budget) return false;
  // (a third x-tile stage at the price of a weight stage measured slower: 0.764 vs 0.733 ms for the 16-layer encoder)
  wa.nw = (int)((budget - fixed - wa.na * a_stage - acts) / wa.w_slot);
  if (wa.nw > WN_NW_MAX) wa.nw = WN_NW_MAX;
  // left-over room goes back to the A ring
  while (wa.na < WN_NA_MAX && fixed + (wa.na + 1) * a_stage + acts + (size_t)wa.nw * wa.w_slot <= budget) wa.na++;
  wa.a_off = (int)fixed;
  wa.acts_off = (int)(fixed + wa.na * a_stage);
  wa.w_off = (int)(wa.acts_off + acts);
  *smem_bytes = wa.w_off + (size_t)wa.nw * wa.w_slot;
  return *smem_bytes <= budget;
}

bool wn_layer_supported(int H, int K, int N_in, int N_rs, int Cout_rs, int planes) {
  WnLayerArgs wa;
  memset(&wa, 0, sizeof(wa));
  wa.H = H, wa.K = K, wa.N_in = N_in, wa.planes = planes, wa.n_layers = 1;
  wa.layer[0].N_rs = N_rs, wa.layer[0].Cout_rs = Cout_rs;
  size_t smem = 0;
  return (planes == 1 || planes == 2) && wn_plan(wa, &smem);
}

static int g_wn_sm_count[64] = {0};
static bool g_wn_configured[64] = {false};
static cudaError_t wn_configure(int* dev_out) {
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (!g_wn_configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(wn_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&g_wn_sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    g_wn_configured[dev] = true;
  }
  *dev_out = dev;
  return cudaSuccess;
}

int wn_stack_max_items() {
  int dev = 0;
  return wn_configure(&dev) == cudaSuccess ? g_wn_sm_count[dev] : 0;
}

cudaError_t launch_wn_layers(const WnLayerArgs& in, cudaStream_t stream) {
  WnLayerArgs wa = in;
  if (wa.planes != 1) wa.planes = 2;
  size_t smem = 0;
  if (!wn_plan(wa, &smem) || !wa.img[0] || !wa.out || !wa.mask) return cudaErrorInvalidValue;
  if (wa.l0 < 0 || wa.l0 + wa.n_layers > wa.n_total) return cudaErrorInvalidValue;
  for (int l = 0; l < wa.n_layers; ++l) {
    const WnLayerParams& lp = wa.layer[l];
    const bool last = wa.l0 + l == wa.n_total - 1;
    if (!lp.w_in || !lp.w_rs || !lp.bias_in || !lp.bias_rs) return cudaErrorInvalidValue;
    if (!last && (!wa.x || lp.Cout_rs != 2 * wa.H || !wa.img[1])) return cudaErrorInvalidValue;
    if (last && lp.Cout_rs != wa.H) return cudaErrorInvalidValue;
  }
  if (wa.B <= 0 || wa.T <= 0) return cudaSuccess;
  wa.ntiles_t = (wa.T + 127) / 128;
  const long long items = (long long)wa.B * wa.ntiles_t;
  if (items > 0x7FFFFFFFLL / 8) return cudaErrorInvalidValue;
  wa.items = (int)items;
  wa.div_t = make_fast_div((uint32_t)wa.ntiles_t);
  int dev = 0;
  cudaError_t e = wn_configure(&dev);
  if (e != cudaSuccess) return e;
  // several layers per launch: one tile per CTA, every CTA resident (they wait for their neighbours inside the kernel)
  if (wa.n_layers > 1 && (wa.items > g_wn_sm_count[dev] || !wa.flags || !wa.img[1])) return cudaErrorInvalidValue;
  CUtensorMap map0, map1;
  memset(&map0, 0, sizeof(map0));
  memset(&map1, 0, sizeof(map1));
  e = tc_make_image_map(wa.img[0], wa.B, wa.H, wa.T, wa.rows, wa.planes, &map0);
  if (e != cudaSuccess) return e;
  if (wa.img[1]) {
    e = tc_make_image_map(wa.img[1], wa.B, wa.H, wa.T, wa.rows, wa.planes, &map1);
    if (e != cudaSuccess) return e;
  }
  if (wa.n_layers > 1) {
    e = cudaMemsetAsync(wa.flags, 0, sizeof(int) * (size_t)wa.items, stream);
    if (e != cudaSuccess) return e;
  }
  const int grid = wa.items < g_wn_sm_count[dev] ? wa.items : g_wn_sm_count[dev];
  if (wa.n_layers > 1) {
    // The tiles of a multi-layer launch wait for each other inside the kernel, so ALL its CTAs must be resident at once.
    // Counting SMs is not enough: a kernel of another stream (an NCCL send / recv of the sharded serving path, a
    // caller's own kernel) can hold SMs for as long as it likes, and resident tiles would spin on neighbours that
    // cannot be scheduled (measured at N = 2: the NCCL-overlapped end-to-end legs went from 37 to 47-67 ms per step).
    // A cooperative launch is the guarantee: the grid starts only when every CTA can be placed.  (No programmatic
    // dependent launch for this one: griddepcontrol is a no-op without the attribute.)
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid), cfg.blockDim = dim3(WN_THREADS), cfg.dynamicSmemBytes = smem, cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, wn_layer_kernel, wa, map0, map1);
  }
  return launch_pdl(wn_layer_kernel, grid, WN_THREADS, smem, stream, wa, map0, map1);
}

}  // namespace svk
