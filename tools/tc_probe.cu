// Developer probe for the tcgen05 conv kernel (conv_tc.cu): runs one shape, compares against a CPU
// fp64 reference (small shapes) or the FFMA kernel (large shapes), prints error statistics and time.
//   tc_probe B Cin Cout K dil L N [ref=cpu|ffma] [res=0|1] [reps] [flags: t = TMA input from an operand image,
//            s = also write + check the output operand image, n = with s: no fp32 output]
// Built by tools/build_probe.sh; not part of the product.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../smart-vocoder_b200/csrc/svk_kernels.cuh"

using namespace svk;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_)); \
      exit(3);                                                                         \
    }                                                                                  \
  } while (0)

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static float frand() {  // uniform (-1, 1)
  rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull;
  return (float)((double)(rng_state >> 11) / 9007199254740992.0 * 2.0 - 1.0);
}
static float gauss() {
  float s = 0;
  for (int i = 0; i < 6; ++i) s += frand();
  return s * 0.7071f;
}

int main(int argc, char** argv) {
  if (argc < 8) {
    printf("usage: tc_probe B Cin Cout K dil L N [cpu|ffma] [res] [reps]\n");
    return 2;
  }
  const int B = atoi(argv[1]), Cin = atoi(argv[2]), Cout = atoi(argv[3]), K = atoi(argv[4]), dil = atoi(argv[5]),
            L = atoi(argv[6]), N = atoi(argv[7]);
  const bool cpu_ref = argc > 8 ? !strcmp(argv[8], "cpu") : true;
  const int use_res = argc > 9 ? atoi(argv[9]) : 0;
  const int reps = argc > 10 ? atoi(argv[10]) : 3;
  const char* flags = argc > 11 ? argv[11] : "";
  const bool f_tma = strchr(flags, 't') != nullptr, f_split = strchr(flags, 's') != nullptr, f_noraw = strchr(flags, 'n') != nullptr;
  const float out_slope = 0.25f;
  const int pad = (K - 1) * dil / 2;
  const float slope = 0.1f;

  const size_t nx = (size_t)B * Cin * L, ny = (size_t)B * Cout * L, nwts = (size_t)Cout * Cin * K;
  std::vector<float> hx(nx), hw(nwts), hb(Cout), hres(use_res ? ny : 0);
  for (auto& v : hx) v = gauss() * 1.5f;
  const float wsd = 1.0f / sqrtf((float)Cin * K);
  for (auto& v : hw) v = gauss() * wsd;
  for (auto& v : hb) v = gauss() * 0.1f;
  for (auto& v : hres) v = gauss();

  float *dx, *dy, *dyref, *db, *dres = nullptr;
  CK(cudaMalloc(&dx, nx * 4));
  CK(cudaMalloc(&dy, ny * 4));
  CK(cudaMalloc(&dyref, ny * 4));
  CK(cudaMalloc(&db, Cout * 4));
  CK(cudaMemcpy(dx, hx.data(), nx * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, hb.data(), Cout * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dy, 0xFF, ny * 4));
  if (use_res) {
    CK(cudaMalloc(&dres, ny * 4));
    CK(cudaMemcpy(dres, hres.data(), ny * 4, cudaMemcpyHostToDevice));
  }

  // ---- tensor-core weights
  const float scale = conv_tc_weight_scale(hw.data(), nwts);
  std::vector<uint16_t> himg(conv_tc_packed_halves(Cin, Cout, K, N));
  conv_tc_pack(hw.data(), Cout, Cin, K, N, scale, himg.data());
  uint16_t* dimg;
  CK(cudaMalloc(&dimg, himg.size() * 2));
  CK(cudaMemcpy(dimg, himg.data(), himg.size() * 2, cudaMemcpyHostToDevice));

  ConvTcArgs ta;
  memset(&ta, 0, sizeof(ta));
  ConvArgs& a = ta.c;
  a.x = dx, a.x_C = Cin, a.x_stride = L, a.Lin = L, a.pre_slope = slope;
  a.bias = db, a.Cin = Cin, a.Cout = Cout, a.CoutPad = Cout, a.K = K, a.dil = dil, a.pad = pad;
  a.Lout = L, a.y_stride = L, a.mode = MODE_STORE, a.split = 1 << 30, a.post_div = 1.0f, a.B = B;
  a.e[0].y = dy, a.e[0].C = Cout, a.e[0].ch_sign = 1, a.e[1].ch_sign = 1;
  a.e[0].res = dres;
  ta.wtc = dimg, ta.unscale = 1.0f / scale, ta.N = N;
  uint16_t *dximg = nullptr, *dyimg = nullptr;
  if (f_tma) {
    CK(cudaMalloc(&dximg, split_image_halves(B, Cin, L) * 2));
    CK(launch_split_image(dx, B, Cin, L, slope, dximg, 2, 0));
    ta.x_split = dximg;
  }
  if (f_split) {
    CK(cudaMalloc(&dyimg, split_image_halves(B, Cout, L) * 2));
    CK(cudaMemset(dyimg, 0xFF, split_image_halves(B, Cout, L) * 2));
    a.e[0].split = dyimg, a.e[0].split_slope = out_slope;
    if (f_noraw) a.e[0].y = nullptr;
  }

  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(launch_conv_tc(ta, 0));
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    CK(launch_conv_tc(ta, 0));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  std::vector<float> hy(ny);
  CK(cudaMemcpy(hy.data(), dy, ny * 4, cudaMemcpyDeviceToHost));
  std::vector<uint16_t> himgy;
  if (f_split) {
    himgy.resize(split_image_halves(B, Cout, L));
    CK(cudaMemcpy(himgy.data(), dyimg, himgy.size() * 2, cudaMemcpyDeviceToHost));
    if (f_noraw) {  // reconstruct y from the image for the comparison below (slope inverted)
      for (int b = 0; b < B; ++b)
        for (int o = 0; o < Cout; ++o)
          for (int t = 0; t < L; ++t) {
            const size_t cell = (((size_t)b * (Cout / 32) + o / 32) * L + t) * 32 + o % 32;
            const float v = __half2float(*reinterpret_cast<const __half*>(&himgy[cell])) +
                            __half2float(*reinterpret_cast<const __half*>(&himgy[(size_t)B * Cout * L + cell]));
            hy[((size_t)b * Cout + o) * L + t] = v >= 0.f ? v : v / out_slope;
          }
    }
  }

  // ---- reference
  std::vector<double> ref;
  std::vector<float> hyref;
  float ffma_ms = 0;
  if (cpu_ref) {
    ref.assign(ny, 0.0);
    for (int b = 0; b < B; ++b)
      for (int o = 0; o < Cout; ++o) {
        double* yr = &ref[((size_t)b * Cout + o) * L];
        for (int t = 0; t < L; ++t) yr[t] = hb[o] + (use_res ? (double)hres[((size_t)b * Cout + o) * L + t] : 0.0);
        for (int c = 0; c < Cin; ++c) {
          const float* xr = &hx[((size_t)b * Cin + c) * L];
          for (int j = 0; j < K; ++j) {
            const double w = hw[((size_t)o * Cin + c) * K + j];
            const int off = j * dil - pad;
            const int lo = off < 0 ? -off : 0, hi = off > 0 ? L - off : L;
            for (int t = lo; t < hi; ++t) {
              const float xv = xr[t + off];
              yr[t] += w * (double)(xv > 0.f ? xv : xv * slope);
            }
          }
        }
      }
  } else {
    const int ot = conv_ffma_channel_tile(Cout);
    const int CoutPad = (Cout + ot - 1) / ot * ot;
    std::vector<float> wp((size_t)Cin * K * CoutPad, 0.f), bp(CoutPad, 0.f);
    for (int c = 0; c < Cin; ++c)
      for (int j = 0; j < K; ++j)
        for (int o = 0; o < Cout; ++o) wp[((size_t)c * K + j) * CoutPad + o] = hw[((size_t)o * Cin + c) * K + j];
    for (int o = 0; o < Cout; ++o) bp[o] = hb[o];
    float *dwp, *dbp;
    CK(cudaMalloc(&dwp, wp.size() * 4));
    CK(cudaMalloc(&dbp, bp.size() * 4));
    CK(cudaMemcpy(dwp, wp.data(), wp.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dbp, bp.data(), bp.size() * 4, cudaMemcpyHostToDevice));
    ConvArgs f = a;
    f.wp = dwp, f.bias = dbp, f.CoutPad = CoutPad;
    f.e[0].y = dyref;
    CK(launch_conv_ffma(f, 0));
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    CK(launch_conv_ffma(f, 0));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ffma_ms, e0, e1));
    hyref.resize(ny);
    CK(cudaMemcpy(hyref.data(), dyref, ny * 4, cudaMemcpyDeviceToHost));
  }

  double maxabs = 0, sumsq = 0, sumref2 = 0, bias_num = 0, bias_den = 0;
  size_t nbad = 0, first_bad = (size_t)-1, nnan = 0;
  for (size_t i = 0; i < ny; ++i) {
    const double r = cpu_ref ? ref[i] : (double)hyref[i];
    const double d = (double)hy[i] - r;
    if (!(hy[i] == hy[i])) {
      nnan++;
      if (first_bad == (size_t)-1) first_bad = i;
      continue;
    }
    if (fabs(d) > maxabs) maxabs = fabs(d);
    sumsq += d * d, sumref2 += r * r;
    bias_num += d * (r >= 0 ? 1.0 : -1.0), bias_den += fabs(r);  // >0: magnitude inflated, <0: shrunk (RZ)
    if (fabs(d) > 1e-3 * (1.0 + fabs(r))) {
      nbad++;
      if (first_bad == (size_t)-1) first_bad = i;
    }
  }
  if (f_split && !f_noraw) {  // image must equal leaky_relu(y) to fp16x2 precision
    double worst = 0;
    for (int b = 0; b < B; ++b)
      for (int o = 0; o < Cout; ++o)
        for (int t = 0; t < L; ++t) {
          const size_t cell = (((size_t)b * (Cout / 32) + o / 32) * L + t) * 32 + o % 32;
          const float v = __half2float(*reinterpret_cast<const __half*>(&himgy[cell])) +
                          __half2float(*reinterpret_cast<const __half*>(&himgy[(size_t)B * Cout * L + cell]));
          const float y = hy[((size_t)b * Cout + o) * L + t];
          const float want = y > 0.f ? y : y * out_slope;
          const double d = fabs((double)v - want) / (4e-8 + 1e-6 * fabs(want));  // fp16 subnormal floor + 2^-20 relative
          if (!(d <= worst)) worst = d;
        }
    printf("  [split image vs leaky_relu(y): worst err / tolerance %.3f]\n", worst);
    if (!(worst < 1.0)) nbad++;
  }
  const double flops = 2.0 * B * (double)L * Cout * Cin * K;
  int pna = 0, pnw = 0, pres = 0;
  size_t psmem = 0;
  conv_tc_plan(Cin, Cout, K, dil, N, f_tma, 2, &pna, &pnw, &pres, &psmem);
  printf("B=%d Cin=%d Cout=%d K=%d dil=%d L=%d N=%d na=%d nw=%d resident=%d smem=%zu res=%d ref=%s | maxabs %.3e rms %.3e (ref rms %.3e) "
         "signed-rel-bias %.3e nan %zu bad %zu flags=%s",
         B, Cin, Cout, K, dil, L, N, pna, pnw, pres, psmem, use_res, cpu_ref ? "cpu64" : "ffma", maxabs, sqrt(sumsq / ny),
         sqrt(sumref2 / ny), bias_num / (bias_den + 1e-30), nnan, nbad, flags);
  printf(" | %.3f ms %.1f TFLOP/s(useful)", best, flops / best / 1e9);
  if (!cpu_ref) printf(" ffma %.3f ms", ffma_ms);
  printf("\n");
  if (nbad || nnan) {
    const size_t i = first_bad;
    const int t = (int)(i % L), o = (int)((i / L) % Cout), b = (int)(i / ((size_t)L * Cout));
    printf("  first bad at b=%d o=%d t=%d: got %.6f want %.6f\n", b, o, t, hy[i], cpu_ref ? ref[i] : (double)hyref[i]);
    // error map over (o mod 32, t mod 32) helps spot layout mistakes
    int shown = 0;
    for (size_t k = 0; k < ny && shown < 12; ++k) {
      const double r = cpu_ref ? ref[k] : (double)hyref[k];
      if (!(fabs((double)hy[k] - r) <= 1e-3 * (1.0 + fabs(r)))) {
        printf("    o=%d t=%d got %.5f want %.5f\n", (int)((k / L) % Cout), (int)(k % L), hy[k], r);
        shown++;
      }
    }
  }
  return (nbad || nnan) ? 1 : 0;
}
