// Micro-benchmark behind DESIGN.md's "whole-sector stores": how fast HBM absorbs 1 GiB written as
//   v4x2   two 16 B stores per thread to the two halves of a 32 B sector, back to back (the old operand-image store)
//   v8     one 32 B store per thread and sector (st.global.v8.b32 / STG.256)
//   late   the two halves of every sector written by two passes over the tile, far apart in time (the old stride-2 upsampler)
//   lines  plain coalesced float4 stores (each warp writes 512 contiguous bytes)
// Thread t of a warp owns row t of a [rows][64 B] tile, as the epilogues do.  Build: tools/build_probe.sh (or nvcc directly).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ void st_v8(void* p, uint4 a, uint4 b) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
               "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}

// rows of 64 B; a CTA of 256 threads writes 256 rows x 64 B per iteration, grid-stride over the buffer
template <int MODE>
__global__ void __launch_bounds__(256) store_kernel(uint8_t* buf, size_t rows, uint32_t seed) {
  const uint4 a = make_uint4(seed, seed + 1, seed + 2, seed + 3), b = make_uint4(seed + 4, seed + 5, seed + 6, seed + 7);
  const size_t per_iter = (size_t)gridDim.x * 256;
  for (size_t r0 = (size_t)blockIdx.x * 256; r0 < rows; r0 += per_iter) {
    const size_t r = r0 + threadIdx.x;
    if (r >= rows) continue;
    uint8_t* row = buf + r * 64;
    if (MODE == 0) {  // v4x2: sector 0 then sector 1 of the row, each as two 16 B stores
      *reinterpret_cast<uint4*>(row) = a;
      *reinterpret_cast<uint4*>(row + 16) = b;
      *reinterpret_cast<uint4*>(row + 32) = b;
      *reinterpret_cast<uint4*>(row + 48) = a;
    } else if (MODE == 1) {  // v8
      st_v8(row, a, b);
      st_v8(row + 32, b, a);
    } else if (MODE == 3) {  // lines: thread t writes 16 B at t * 16 of a 4 KB block, four times
      uint8_t* blk = buf + r0 * 64;
#pragma unroll
      for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(blk + k * 4096 + threadIdx.x * 16) = k & 1 ? b : a;
    }
  }
  if (MODE == 2) {  // late: first pass writes the low half of every sector, second pass the high halves
    for (int pass = 0; pass < 2; ++pass)
      for (size_t r0 = (size_t)blockIdx.x * 256; r0 < rows; r0 += per_iter) {
        const size_t r = r0 + threadIdx.x;
        if (r >= rows) continue;
        uint8_t* row = buf + r * 64 + pass * 16;
        *reinterpret_cast<uint4*>(row) = a;
        *reinterpret_cast<uint4*>(row + 32) = b;
      }
  }
}

template <int MODE>
static float run(uint8_t* buf, size_t rows, int grid) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  float best = 1e30f;
  for (int it = 0; it < 6; ++it) {
    cudaEventRecord(e0);
    store_kernel<MODE><<<grid, 256>>>(buf, rows, (uint32_t)it);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (it > 0 && ms < best) best = ms;
  }
  return best;
}

int main() {
  const size_t bytes = 1ull << 30, rows = bytes / 64;
  uint8_t* buf;
  if (cudaMalloc(&buf, bytes) != cudaSuccess) return 1;
  const char* names[4] = {"v4x2 (two 16 B halves, back to back)", "v8   (one 32 B store per sector)", "late (halves of a sector in two passes)",
                          "lines (coalesced float4)"};
  for (int grid : {148 * 2, 148 * 8}) {
    const float t[4] = {run<0>(buf, rows, grid), run<1>(buf, rows, grid), run<2>(buf, rows, grid), run<3>(buf, rows, grid)};
    for (int m = 0; m < 4; ++m) printf("grid %4d  %-42s %7.3f ms  %7.1f GB/s\n", grid, names[m], t[m], bytes / (t[m] * 1e-3) / 1e9);
  }
  cudaFree(buf);
  return 0;
}
