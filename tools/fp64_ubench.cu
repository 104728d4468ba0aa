// Micro-benchmarks behind the direct-sum kernel design: FP64 pipe, MUFU.RSQ64H and mixed-loop
// throughput on the current GPU.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_ubench fp64_ubench.cu
#include <cmath>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double rsq64h(double a) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  return y;
}
__device__ __forceinline__ double halve(double y) {
  return __hiloint2double(__double2hiint(y) - 0x00100000, __double2loint(y));
}
// VAR 0: cubic Goldschmidt (5 FP64 ops, ~1 ulp)   VAR 1: Newton (3 FP64 ops, ~2^-42)
// VAR 2: FP32 seed (F2F + MUFU.RSQ + F2F) + cubic  VAR 3: no seed at all: 5 dependent DFMA (pipe ceiling of the loop)
// VAR 4: IEEE sqrt()                                VAR 5: FP32 seed + Newton (3 ops)
template <int VAR>
__device__ __forceinline__ double sqrt_var(double a) {
  if (VAR == 0) {
    const double y0 = rsq64h(a), h = halve(y0), r = a * y0, e = fma(-r, h, 0.5), c = fma(1.5, e, 1.0);
    return fma(r * e, c, r);
  } else if (VAR == 1) {
    const double y0 = rsq64h(a), h = halve(y0), r = a * y0, d = fma(-r, r, a);
    return fma(d, h, r);
  } else if (VAR == 2) {
    const double y0 = (double)rsqrtf((float)a), h = 0.5 * y0, r = a * y0, e = fma(-r, h, 0.5), c = fma(1.5, e, 1.0);
    return fma(r * e, c, r);
  } else if (VAR == 3) {
    double r = a;
    r = fma(r, 0.999, 1e-3); r = fma(r, 0.999, 1e-3); r = fma(r, 0.999, 1e-3); r = fma(r, 0.999, 1e-3);
    return fma(r, 0.999, 1e-3);
  } else if (VAR == 4) {
    return sqrt(a);
  } else if (VAR == 5) {
    const float yf = rsqrtf((float)a);
    const double y0 = (double)yf, h = (double)(0.5f * yf), r = a * y0, d = fma(-r, r, a);
    return fma(d, h, r);
  } else if (VAR == 6) {  // cost probe: integer "magic" seed instead of MUFU (accuracy is not the point)
    const double y0 = __hiloint2double(0x5fe6eb50 - (__double2hiint(a) >> 1), 0);
    const double r = a * y0, e = fma(-r, y0, 1.0), c = fma(e, 0.375, 0.5);
    return fma(r * e, c, r);
  } else if (VAR == 7) {  // cubic step without the halved seed (the sequence the library uses)
    const double y0 = rsq64h(a), r = a * y0, e = fma(-r, y0, 1.0), c = fma(e, 0.375, 0.5);
    return fma(r * e, c, r);
  } else if (VAR == 8) {  // quadratic step, 3 FP64 ops:  r + r (1/2 - r y0/2)
    const double y0 = rsq64h(a), h = halve(y0), r = a * y0, e = fma(-r, h, 0.5);
    return fma(r, e, r);
  } else {  // VAR 9: 2 sqrt(a) = r (3 - r y0): the factor 1/2 is folded into the weights, no halved seed
    const double y0 = rsq64h(a), r = a * y0, g = fma(-r, y0, 3.0);
    return r * g;
  }
}

// accuracy probe: seed, quadratic and cubic results for log-uniform inputs
__global__ void k_acc(const double *x, double *seed, double *quad, double *cub, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  seed[i] = rsq64h(x[i]);
  quad[i] = sqrt_var<8>(x[i]);
  cub[i] = sqrt_var<7>(x[i]);
}

// the M2P / P2L inner loop: r2 = axy + dz2[i2]; v = sqrt(r2); acc -= v * w  (7 FP64 ops per pair with VAR 0)
template <int VAR, int ILP>
__global__ void __launch_bounds__(256) k_far(const double *in, double *out, int iters) {
  __shared__ double tab[64][33], wt[512];
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) tab[i / 32][i % 32] = 1.0 + in[i & 255];
  for (int i = threadIdx.x; i < 512; i += blockDim.x) wt[i] = in[i & 255];
  __syncthreads();
  double dzr[7];
  for (int i = 0; i < 7; ++i) dzr[i] = 0.5 + in[i] + lane * 1e-3;
  double acc = 0.0;
  for (int it = 0; it < iters; ++it) {
    double axy[ILP];
#pragma unroll
    for (int u = 0; u < ILP; ++u) axy[u] = tab[(it * ILP + u) & 63][lane];
    const double *w = wt + ((it * 8) & 511 & ~7);
#pragma unroll
    for (int i2 = 0; i2 < 7; ++i2) {
      double v[ILP];
#pragma unroll
      for (int u = 0; u < ILP; ++u) v[u] = sqrt_var<VAR>(axy[u] + dzr[i2]);
#pragma unroll
      for (int u = 0; u < ILP; ++u) acc -= v[u] * w[i2];
    }
  }
  if (acc == 1.2345) out[0] = acc;
}

// the P2P inner loop: 3 sub, mul + 2 fma, sqrt, guarded select, fma; sources broadcast from shared memory
template <int VAR, int UNR>
__global__ void __launch_bounds__(256) k_p2p(const double *in, double *out, int iters) {
  __shared__ double sx[128], sy[128], sz[128], sw[128];
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 128; i += blockDim.x) {
    sx[i] = in[i];
    sy[i] = in[i + 64];
    sz[i] = in[(i + 128) & 255];
    sw[i] = in[(i + 32) & 255];
  }
  __syncthreads();
  const double xt = in[lane] + 0.3, yt = in[lane + 32] + 0.1, zt = in[lane + 64] - 0.2;
  double acc = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll UNR
    for (int j = 0; j < 128; ++j) {
      const double dx = xt - sx[j], dy = yt - sy[j], dz = zt - sz[j];
      double r2 = dx * dx;
      r2 += dy * dy;
      r2 += dz * dz;
      const double r = sqrt_var<VAR>(r2);
      const double m = __double2hiint(r2) >= 0x00100000 ? r : 0.0;
      acc -= m * sw[j];
    }
  }
  if (acc == 1.2345) out[0] = acc;
}

__global__ void __launch_bounds__(256) k_mufu(const double *in, double *out, int iters) {
  double x[8];
  for (int i = 0; i < 8; ++i) x[i] = 1.0 + in[i] + threadIdx.x * 1e-3;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = rsq64h(x[i]);
  double s = 0;
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 1.2345) out[0] = s;
}
__global__ void __launch_bounds__(256) k_mufu32(const double *in, double *out, int iters) {
  float x[8];
  for (int i = 0; i < 8; ++i) x[i] = 1.0f + (float)in[i] + threadIdx.x * 1e-3f;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = rsqrtf(x[i]);
  float s = 0;
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 1.2345f) out[0] = s;
}
__global__ void __launch_bounds__(256) k_dfma(const double *in, double *out, int iters) {
  double x[8];
  for (int i = 0; i < 8; ++i) x[i] = in[i] + threadIdx.x * 1e-3;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fma(x[i], 0.999999, 1e-9);
  double s = 0;
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 1.2345) out[0] = s;
}
// one dependent chain per thread, one warp per SM sub-partition: DFMA / MUFU latency in cycles
__global__ void k_lat(const double *in, double *out, int iters, int what, long long *cyc) {
  double x = 1.0 + in[threadIdx.x & 7];
  const long long t0 = clock64();
  if (what == 0)
    for (int it = 0; it < iters; ++it) x = fma(x, 0.999999, 1e-9);
  else
    for (int it = 0; it < iters; ++it) x = rsq64h(x);
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  if (x == 1.2345) out[0] = x;
}

template <class F>
double time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  int sms, khz;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double clk = khz * 1e3;
  double h_in[256];
  for (int i = 0; i < 256; ++i) h_in[i] = 0.001 * (i + 1);
  double *in, *out;
  long long *cyc;
  cudaMalloc(&in, sizeof(h_in));
  cudaMalloc(&out, 8);
  cudaMalloc(&cyc, 8);
  cudaMemcpy(in, h_in, sizeof(h_in), cudaMemcpyHostToDevice);
  printf("SMs %d, nominal clock %.0f MHz (rates below are per SM per nominal clock)\n", sms, clk / 1e6);
  const int iters = 2048;
  auto rate = [&](double n_warp_instr, double ms) { return n_warp_instr / (ms * 1e-3) / clk / sms; };
  {
    const int blocks = sms * 8, threads = 256;
    double ms = time_ms([&] { k_dfma<<<blocks, threads>>>(in, out, iters); });
    printf("DFMA         : %.3f warp-instr/clk/SM  (%.2f TFLOP/s)\n", rate(8.0 * iters * blocks * threads / 32, ms),
           2.0 * 8 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12);
    ms = time_ms([&] { k_mufu<<<blocks, threads>>>(in, out, iters); });
    printf("MUFU.RSQ64H  : %.3f warp-instr/clk/SM\n", rate(8.0 * iters * blocks * threads / 32, ms));
    ms = time_ms([&] { k_mufu32<<<blocks, threads>>>(in, out, iters); });
    printf("MUFU.RSQ f32 : %.3f warp-instr/clk/SM\n", rate(8.0 * iters * blocks * threads / 32, ms));
  }
  for (int what = 0; what < 2; ++what) {
    k_lat<<<1, 32>>>(in, out, 4096, what, cyc);
    long long h;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%s dependent latency: %.1f cycles\n", what == 0 ? "DFMA" : "MUFU.RSQ64H", (double)h / 4096);
  }
#define FAR(VAR, ILP, BPS, TH)                                                                                    \
  {                                                                                                               \
    const int blocks = sms * BPS;                                                                                 \
    double ms = time_ms([&] { k_far<VAR, ILP><<<blocks, TH>>>(in, out, iters); });                                \
    const double pairs = 7.0 * ILP * iters * (double)blocks * TH;                                                 \
    printf("far  var %d ilp %d  %2d warps/SM: %7.2f Gpair/s  (%.3f pair-warps/clk/SM)\n", VAR, ILP, BPS * TH / 32, \
           pairs / (ms * 1e-3) / 1e9, rate(pairs / 32, ms));                                                      \
  }
  FAR(0, 1, 8, 256) FAR(0, 2, 8, 256) FAR(0, 4, 8, 256) FAR(0, 4, 4, 256) FAR(0, 4, 2, 256) FAR(0, 8, 4, 256)
  FAR(1, 4, 8, 256) FAR(1, 4, 4, 256) FAR(2, 4, 8, 256) FAR(3, 4, 8, 256) FAR(4, 4, 8, 256) FAR(5, 4, 8, 256)
#define P2P(VAR, UNR, BPS, TH)                                                                                     \
  {                                                                                                                \
    const int blocks = sms * BPS, it2 = 64;                                                                        \
    double ms = time_ms([&] { k_p2p<VAR, UNR><<<blocks, TH>>>(in, out, it2); });                                   \
    const double pairs = 128.0 * it2 * (double)blocks * TH;                                                        \
    printf("p2p  var %d unr %d  %2d warps/SM: %7.2f Gpair/s  (%.3f pair-warps/clk/SM)\n", VAR, UNR, BPS * TH / 32, \
           pairs / (ms * 1e-3) / 1e9, rate(pairs / 32, ms));                                                       \
  }
  FAR(6, 4, 8, 256) FAR(6, 4, 2, 256) FAR(7, 4, 8, 256) FAR(7, 4, 2, 256) FAR(8, 4, 8, 256) FAR(8, 4, 2, 256)
  FAR(9, 4, 8, 256) FAR(9, 4, 2, 256) FAR(9, 4, 3, 256) FAR(7, 4, 1, 256) FAR(7, 4, 3, 256) FAR(7, 2, 3, 256) FAR(7, 2, 4, 256) FAR(7, 8, 2, 256) FAR(8, 4, 3, 256)
  {
    const int n = 1 << 20;
    double *hx = new double[n], *hs = new double[n], *hq = new double[n], *hc = new double[n];
    unsigned long long st = 88172645463325252ull;
    for (int i = 0; i < n; ++i) {
      st ^= st << 13; st ^= st >> 7; st ^= st << 17;
      const double u = (st >> 11) * (1.0 / 9007199254740992.0);
      hx[i] = exp2(-40.0 + 80.0 * u);
    }
    double *dx, *ds, *dq, *dc;
    cudaMalloc(&dx, n * 8); cudaMalloc(&ds, n * 8); cudaMalloc(&dq, n * 8); cudaMalloc(&dc, n * 8);
    cudaMemcpy(dx, hx, n * 8, cudaMemcpyHostToDevice);
    k_acc<<<n / 256, 256>>>(dx, ds, dq, dc, n);
    cudaMemcpy(hs, ds, n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(hq, dq, n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(hc, dc, n * 8, cudaMemcpyDeviceToHost);
    long double es = 0, eq = 0, ec = 0;
    for (int i = 0; i < n; ++i) {
      const long double t = sqrtl((long double)hx[i]);
      const long double a = fabsl((long double)hs[i] * t - 1.0L), b = fabsl((long double)hq[i] / t - 1.0L),
                        c = fabsl((long double)hc[i] / t - 1.0L);
      if (a > es) es = a;
      if (b > eq) eq = b;
      if (c > ec) ec = c;
    }
    printf("max relative error over 2^20 log-uniform inputs: MUFU.RSQ64H seed %.3Le (2^%.1Lf), 3-op sqrt %.3Le, 5-op sqrt %.3Le\n",
           es, log2l(es), eq, ec);
  }
  P2P(0, 1, 8, 256) P2P(0, 2, 8, 256) P2P(0, 4, 8, 256) P2P(0, 8, 8, 256) P2P(0, 4, 4, 256) P2P(0, 4, 2, 256)
  P2P(1, 4, 8, 256) P2P(2, 4, 8, 256) P2P(3, 4, 8, 256) P2P(4, 4, 8, 256) P2P(5, 4, 8, 256) P2P(6, 4, 8, 256) P2P(7, 4, 8, 256) P2P(7, 4, 2, 256) P2P(8, 4, 8, 256) P2P(8, 4, 2, 256) P2P(9, 4, 8, 256) P2P(9, 4, 2, 256)
  printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
