// Micro-benchmark: FP64 tensor (DMMA, mma.sync f64) vs FP64 FMA-pipe throughput on the current GPU.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_bench dmma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double *out, int iters) {
  double x[16];
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = fma(x[i], 0.999999, 1e-9);
  double s = 0;
  for (int i = 0; i < 16; ++i) s += x[i];
  if (s == 1.2345) out[0] = s;
}

// m8n8k4: A 1 reg, B 1 reg, C 2 regs per lane
__global__ void k_dmma884(double *out, int iters) {
  double c[8][2];
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  double s = 0;
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  if (s == 1.2345) out[0] = s;
}

#if __CUDA_ARCH__ >= 900 || !defined(__CUDA_ARCH__)
// m16n8k16: A 8 regs, B 4 regs, C 4 regs per lane
__global__ void k_dmma16816(double *out, int iters) {
  double c[4][4];
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.0;
  double a[8], b[4];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int i = 0; i < 4; ++i) b[i] = threadIdx.x * 2e-3 + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                     "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  double s = 0;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  if (s == 1.2345) out[0] = s;
}
#endif

// Do DMMA and DFMA overlap?  Per loop trip: 4 independent m8n8k4 products and NF independent DFMAs per lane (the P2P
// tensor-core kernel's mix is 4 : 32).  If the pipes are separate the time is the max of the two, else the sum.
template <int NF, int ND>
__global__ void k_mix(double *out, int iters) {
  double c[4][2], x[32];
  for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = 0.0;
  for (int i = 0; i < 32; ++i) x[i] = threadIdx.x * 1e-3 + i;
  double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ND; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
#pragma unroll
    for (int i = 0; i < NF; ++i) x[i] = fma(x[i], 0.999999, 1e-9);
  }
  double s = 0;
  for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
  for (int i = 0; i < 32; ++i) s += x[i];
  if (s == 1.2345) out[0] = s;
}

template <class F>
double time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double *out; cudaMalloc(&out, 8);
  const int iters = 4096, blocks = sms * 8, threads = 256;
  double ms = time_ms([&] { k_dfma<<<blocks, threads>>>(out, iters); });
  printf("DFMA       : %.2f TFLOP/s\n", 2.0 * 16 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12);
  ms = time_ms([&] { k_dmma884<<<blocks, threads>>>(out, iters); });
  printf("DMMA m8n8k4 : %.2f TFLOP/s\n", 2.0 * 256 * 8 * iters * (double)blocks * (threads / 32) / (ms * 1e-3) / 1e12);
  ms = time_ms([&] { k_dmma16816<<<blocks, threads>>>(out, iters); });
  printf("DMMA m16n8k16: %.2f TFLOP/s\n", 2.0 * 2048 * 4 * iters * (double)blocks * (threads / 32) / (ms * 1e-3) / 1e12);
  {
    const int it2 = 2048;
    const double ms_d = time_ms([&] { k_mix<0, 4><<<blocks, threads>>>(out, it2); });
    const double ms_f = time_ms([&] { k_mix<32, 0><<<blocks, threads>>>(out, it2); });
    const double ms_m = time_ms([&] { k_mix<32, 4><<<blocks, threads>>>(out, it2); });
    const double ms_h = time_ms([&] { k_mix<16, 4><<<blocks, threads>>>(out, it2); });
    printf("mix (4 DMMA : 32 DFMA per trip): DMMA only %.3f ms, DFMA only %.3f ms, both %.3f ms (sum %.3f, max %.3f); 4 : 16 -> %.3f ms\n",
           ms_d, ms_f, ms_m, ms_d + ms_f, ms_d > ms_f ? ms_d : ms_f, ms_h);
  }
  cudaError_t e = cudaGetLastError();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
