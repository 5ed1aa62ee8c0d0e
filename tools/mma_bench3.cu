// Developer micro-benchmark #3 (round 2): the conv kernels' bare MMA issue loop (conv_tc.cu, resident weights) in isolation --
// is a narrow layer (N' = 2N <= 128) bound by the tensor pipe, by reading the row-shifted A tile, or by how fast one lane can
// ISSUE?  mma_bench.cu's "74 cycles per row-shifted MMA" had a runtime modulo in its issue path; this one replays the real
// loop (4 MMAs + 2 adds per tap, whole warp walks, elected lane issues) and variants of it, with no barriers in the way.
//   mma_bench3 [tiles]
#include "../smart-vocoder_b200/csrc/tc_common.cuh"

#include <stdio.h>

int svk_g_pdl_enabled = 0;

using namespace svk;

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_));   \
      exit(3);                                                                                  \
    }                                                                                           \
  } while (0)

// MODE 0: the library's loop (runtime K, `#pragma unroll 1`).  MODE 1: taps unrolled x KU (descriptor offsets become
// immediates of uniform adds).  MODE 2: like 0 but only the N' = 2N MMAs (xh products): what the xl pass costs.
// MODE 3: one MMA per k-group with N' = 2N only, A row pitch unchanged -- the bf16 engine's loop.
template <int MODE, int KU>
__global__ void __launch_bounds__(128, 1) issue_kernel(int tiles, int nchunks, int K, int N, int dil, long long* out_cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base;
  if (warp == 1) {
    uint8_t* a_smem = smem;                // A stages: [hi|lo][rows][64 B], 2 stages of 2 x 192 rows
    uint8_t* w_smem = smem + 64 * 1024;    // weight image, resident
    const uint32_t a_plane = 192 * 64, a_stage = 2 * a_plane;
    const uint32_t w_plane2 = (uint32_t)(2 * N) * 16, w_stage = w_plane2 * 4;  // [k-group 0..3][hi n | lo n][8]
    const uint32_t idesc1 = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | ((uint32_t)((2 * N) >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t b_hi = (128u >> 4) | (1u << 14);
    const uint32_t a_hi = (512u >> 4) | (1u << 14) | (4u << 29);
    const uint32_t a_lo0 = ((smem_u32(a_smem) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t w_lo0 = ((smem_u32(w_smem) & 0x3FFFFu) >> 4) | ((w_plane2 >> 4) << 16);
    const uint32_t a_stage16 = a_stage >> 4;
    const uint32_t w_stage16 = (size_t)K * nchunks * w_stage <= 128 * 1024 ? w_stage >> 4 : 0u;  // image too large: re-read tap 0
    const uint32_t lo_plane16 = a_plane >> 4;
    const uint32_t ks_a16 = 32u >> 4, ks_b16 = (2 * w_plane2) >> 4;
    const uint32_t dil16 = (uint32_t)dil * 4u;
    const uint32_t bar_addr = smem_u32(&bar);
    const bool leader = elect_one();
    const uint32_t dmain = tmem, dcross = tmem + (uint32_t)N;
    int ast = 0;
    __syncwarp();
    const long long t0 = clock64();
    for (int i = 0; i < tiles; ++i) {
      uint32_t acc = 0;
      for (int ch = 0; ch < nchunks; ++ch) {
        uint32_t ah = a_lo0 + (uint32_t)ast * a_stage16;
        if (leader) {
          uint32_t bw = w_lo0 + (uint32_t)(ch * K) * w_stage16;
          if constexpr (MODE == 0) {
#pragma unroll 1
            for (int j = 0; j < K; ++j) {
              umma_f16_lo(dmain, ah, bw, a_hi, b_hi, idesc2, acc);
              umma_f16_lo(dcross, ah + lo_plane16, bw, a_hi, b_hi, idesc1, 1u);
              umma_f16_lo(dmain, ah + ks_a16, bw + ks_b16, a_hi, b_hi, idesc2, 1u);
              umma_f16_lo(dcross, ah + ks_a16 + lo_plane16, bw + ks_b16, a_hi, b_hi, idesc1, 1u);
              acc = 1u;
              ah += dil16, bw += w_stage16;
            }
          } else if constexpr (MODE == 1) {
#pragma unroll 1
            for (int j = 0; j < K; j += KU) {
#pragma unroll
              for (int u = 0; u < KU; ++u) {
                if (j + u < K) {
                  umma_f16_lo(dmain, ah, bw, a_hi, b_hi, idesc2, acc);
                  umma_f16_lo(dcross, ah + lo_plane16, bw, a_hi, b_hi, idesc1, 1u);
                  umma_f16_lo(dmain, ah + ks_a16, bw + ks_b16, a_hi, b_hi, idesc2, 1u);
                  umma_f16_lo(dcross, ah + ks_a16 + lo_plane16, bw + ks_b16, a_hi, b_hi, idesc1, 1u);
                  acc = 1u;
                  ah += dil16, bw += w_stage16;
                }
              }
            }
          } else if constexpr (MODE == 2) {
#pragma unroll 1
            for (int j = 0; j < K; ++j) {
              umma_f16_lo(dmain, ah, bw, a_hi, b_hi, idesc2, acc);
              umma_f16_lo(dmain, ah + ks_a16, bw + ks_b16, a_hi, b_hi, idesc2, 1u);
              acc = 1u;
              ah += dil16, bw += w_stage16;
            }
          } else {
#pragma unroll 1
            for (int j = 0; j < K; ++j) {
              umma_f16_lo(dmain, ah, bw, a_hi, b_hi, idesc1, acc);
              umma_f16_lo(dmain, ah + ks_a16, bw + ks_b16, a_hi, b_hi, idesc1, 1u);
              acc = 1u;
              ah += dil16, bw += w_stage16;
            }
          }
          umma_commit_u32(bar_addr);
        }
        acc = 1u;
        ast ^= 1;
      }
    }
    if (leader) {
      umma_commit_u32(bar_addr);
    }
    __syncwarp();
    // drain: wait until the pipe is idle (the barrier has been arrived on many times; poll a fresh phase by time instead)
    long long t1 = clock64();
    if (leader) {
      // a final commit to a second barrier gives an exact end point
      __shared__ uint64_t bar2;
      mbar_init(&bar2, 1);
      fence_barrier_init();
      umma_commit(&bar2);
      mbar_wait(&bar2, 0);
      t1 = clock64();
      out_cycles[blockIdx.x] = t1 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int MODE, int KU>
static void run(const char* name, int tiles, int nchunks, int K, int N, int dil, int sms, long long* d, long long* h) {
  const int smem = 200 * 1024;
  CK(cudaFuncSetAttribute(issue_kernel<MODE, KU>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int rep = 0; rep < 2; ++rep) {
    issue_kernel<MODE, KU><<<sms, 128, smem>>>(tiles, nchunks, K, N, dil, d);
    CK(cudaDeviceSynchronize());
  }
  CK(cudaMemcpy(h, d, sms * sizeof(long long), cudaMemcpyDeviceToHost));
  double sum = 0;
  for (int i = 0; i < sms; ++i) sum += (double)h[i];
  const double mmas = (double)tiles * nchunks * K * (MODE >= 2 ? 2 : 4);
  printf("%-44s C=%3d K=%2d dil=%d  cycles/MMA %6.1f  cycles/tile %8.0f\n", name, N, K, dil, sum / sms / mmas,
         sum / sms / tiles);
}

int main(int argc, char** argv) {
  const int tiles = argc > 1 ? atoi(argv[1]) : 200;
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  long long* d;
  CK(cudaMalloc(&d, sms * sizeof(long long)));
  long long* h = (long long*)malloc(sms * sizeof(long long));
  for (int N : {32, 64, 128}) {
    const int nch = N / 32;
    for (int K : {3, 7, 11})
      for (int dil : {0, 1, 3, 5, 8}) {
        if (K != 11 && dil != 1 && dil != 0) continue;
        run<0, 1>("library loop (4 MMAs / tap, unroll 1)", tiles, nch, K, N, dil, sms, d, h);
      }
    for (int dil : {0, 1, 3}) {
      run<1, 4>("taps unrolled x4", tiles, nch, 11, N, dil, sms, d, h);
      run<1, 11>("taps unrolled x11", tiles, nch, 11, N, dil, sms, d, h);
      run<2, 1>("xh products only (N' = 2N, 2 MMAs / tap)", tiles, nch, 11, N, dil, sms, d, h);
      run<3, 1>("single plane (N, 2 MMAs / tap)", tiles, nch, 11, N, dil, sms, d, h);
    }
  }
  return 0;
}
