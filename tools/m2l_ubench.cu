// Micro-benchmarks behind the round-2 M2L design (tools/, not product code):
//   1. coalesced RED.E.ADD.F64 of P-long columns into a [cells][Ps] array (the scatter of Z = U Y into the locals)
//   2. the same columns added by the TMA engine: cp.reduce.async.bulk .add.f64 (SASS UBLKRED.G.S.ADD.F64)
//   3. cp.async.bulk loads of P-long multipole columns into a shared-memory ring (SASS UBLKCP.S.G + SYNCS)
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/m2l_ubench tools/m2l_ubench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int Ps = 344;  // padded column length (doubles): 2752 B, a multiple of 16

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *b, int cnt) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(cnt));
}
__device__ __forceinline__ void mbar_expect(unsigned long long *b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *b, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_load(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_red_add(double *dst, const void *src, unsigned bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)),
               "r"(bytes)
               : "memory");
}

// 1. every warp adds whole columns with coalesced REDs
__global__ void k_red_lsu(double *loc, const int *tgt, int n_entries) {
  const int lane = threadIdx.x & 31;
  const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int e = wg; e < n_entries; e += nw) {
    double *dst = loc + (size_t)tgt[e] * Ps;
#pragma unroll
    for (int i = 0; i < 11; ++i)
      if (lane + 32 * i < 343) atomicAdd(dst + lane + 32 * i, 1.0);
  }
}

// 2. one thread per CTA hands columns staged in shared memory to the TMA engine
__global__ void k_red_bulk(double *loc, const int *tgt, int n_entries, int cols_per_op) {
  extern __shared__ __align__(128) double sm[];  // 16 columns
  for (int i = threadIdx.x; i < 16 * Ps; i += blockDim.x) sm[i] = 1.0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    int k = 0;
    for (int e = blockIdx.x; e < n_entries; e += gridDim.x, ++k) {
      bulk_red_add(loc + (size_t)tgt[e] * Ps, sm + (k & 15) * Ps, Ps * 8);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if ((k & 7) == 7) asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

// 3. producer thread fills a ring of STAGES x COLS columns, the other warps just read one value per column and release
template <int STAGES, int COLS>
__global__ void k_bulk_load(const double *mult, const int *src, int n_entries, double *sink) {
  extern __shared__ __align__(128) double sm[];
  __shared__ unsigned long long full[STAGES], empty[STAGES];
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], blockDim.x / 32 - 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int per_cta = (n_entries / COLS + gridDim.x - 1) / gridDim.x;  // stages of COLS columns
  const int st0 = blockIdx.x * per_cta, st1 = min(n_entries / COLS, st0 + per_cta);
  if (tid < 32) {
    if (tid == 0) {
      for (int st = st0, k = 0; st < st1; ++st, ++k) {
        const int s = k % STAGES;
        if (k >= STAGES) mbar_wait(&empty[s], ((k / STAGES) - 1) & 1);
        mbar_expect(&full[s], COLS * Ps * 8);
        for (int c = 0; c < COLS; ++c)
          bulk_load(sm + ((size_t)s * COLS + c) * Ps, mult + (size_t)src[st * COLS + c] * Ps, Ps * 8, &full[s]);
      }
    }
  } else {
    double acc = 0;
    for (int st = st0, k = 0; st < st1; ++st, ++k) {
      const int s = k % STAGES;
      mbar_wait(&full[s], (k / STAGES) & 1);
      acc += sm[((size_t)s * COLS + (tid % COLS)) * Ps + (tid >> 5)];
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(&empty[s]);
    }
    if (acc == 123.456) sink[0] = acc;
  }
}

int main() {
  const int ncell = 12000, n_entries = 926436;
  double *loc, *mult, *sink;
  int *tgt_seq, *tgt_rnd;
  CK(cudaMalloc(&loc, (size_t)ncell * Ps * 8));
  CK(cudaMalloc(&mult, (size_t)ncell * Ps * 8));
  CK(cudaMalloc(&sink, 8));
  CK(cudaMemset(loc, 0, (size_t)ncell * Ps * 8));
  CK(cudaMemset(mult, 0, (size_t)ncell * Ps * 8));
  std::vector<int> hs(n_entries), hr(n_entries);
  // "seq": 316 groups, each a sorted sweep over the cells (what a (level, vector) grouping produces)
  const int per_group = (n_entries + 315) / 316;
  for (int e = 0; e < n_entries; ++e) {
    hs[e] = (int)(((long long)(e % per_group) * ncell) / per_group);
    hr[e] = (int)((e * 2654435761u) % (unsigned)ncell);
  }
  CK(cudaMalloc(&tgt_seq, n_entries * 4));
  CK(cudaMalloc(&tgt_rnd, n_entries * 4));
  CK(cudaMemcpy(tgt_seq, hs.data(), n_entries * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(tgt_rnd, hr.data(), n_entries * 4, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  auto timeit = [&](const char *name, auto launch) {
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      CK(cudaEventRecord(e0));
      launch();
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      CK(cudaGetLastError());
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep > 0 && ms < best) best = ms;
    }
    const double bytes = (double)n_entries * 343 * 8;
    printf("%-44s %8.3f ms  %8.1f GB/s  (%.1f G doubles/s)\n", name, best, bytes / best / 1e6, bytes / 8 / best / 1e6);
  };
  for (int pat = 0; pat < 2; ++pat) {
    const int *tg = pat ? tgt_rnd : tgt_seq;
    const char *pn = pat ? "random" : "group-sorted";
    char nm[128];
    snprintf(nm, sizeof nm, "RED.F64 coalesced (LSU), %s targets", pn);
    timeit(nm, [&] { k_red_lsu<<<148 * 8, 256>>>(loc, tg, n_entries); });
    for (int ctas : {1, 2, 4, 8}) {
      snprintf(nm, sizeof nm, "UBLKRED add.f64, %d CTA/SM, %s", ctas, pn);
      CK(cudaFuncSetAttribute(k_red_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * Ps * 8));
      timeit(nm, [&] { k_red_bulk<<<148 * ctas, 64, 16 * Ps * 8>>>(loc, tg, n_entries, 1); });
    }
    snprintf(nm, sizeof nm, "UBLKCP ring 4x8 cols, 1 CTA/SM, %s", pn);
    CK(cudaFuncSetAttribute(k_bulk_load<4, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 8 * Ps * 8));
    timeit(nm, [&] { k_bulk_load<4, 8><<<148, 288, 4 * 8 * Ps * 8>>>(mult, tg, n_entries, sink); });
    snprintf(nm, sizeof nm, "UBLKCP ring 4x8 cols, 2 CTA/SM, %s", pn);
    timeit(nm, [&] { k_bulk_load<4, 8><<<296, 288, 4 * 8 * Ps * 8>>>(mult, tg, n_entries, sink); });
    snprintf(nm, sizeof nm, "UBLKCP ring 3x16 cols, 1 CTA/SM, %s", pn);
    CK(cudaFuncSetAttribute(k_bulk_load<3, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 16 * Ps * 8));
    timeit(nm, [&] { k_bulk_load<3, 16><<<148, 288, 3 * 16 * Ps * 8>>>(mult, tg, n_entries, sink); });
  }
  // sanity: every add landed (sum of loc == total adds)
  std::vector<double> h((size_t)ncell * Ps);
  CK(cudaMemcpy(h.data(), loc, h.size() * 8, cudaMemcpyDeviceToHost));
  double s = 0;
  for (double v : h) s += v;
  printf("sum(loc) = %.0f\n", s);
  return 0;
}
