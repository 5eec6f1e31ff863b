// Micro-benchmark: issue rate of tcgen05.mma.cta_group::1.kind::f16 (M=128, K=16) for several N,
// operands in shared memory (SWIZZLE_128B K-major), aligned vs row-shifted A start addresses.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "../../buctd_b200/csrc/common.cuh"
namespace buctd { void set_error(const char*, ...) {} EncodeTiledFn get_encode_tiled() { return nullptr; } }
using namespace buctd;

template <int N>
__global__ void __launch_bounds__(128) k(long long* out, int iters, int shift_rows, int nacc, int commit_every) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, bar_mid;
  __shared__ uint32_t slot;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar_mid, 1 << 20); fence_mbar_init(); }
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0;
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x < 32) {
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_f16(0, 128, N);
      const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 32 * 1024);
      long long t0 = clock64();
      uint32_t par = 0;
      for (int it = 0; it < iters; ++it) {
        const uint32_t sh = (uint32_t)((it % 9) * shift_rows) * 128;
        const uint64_t ad = umma_desc_sw128(a0 + sh), bd = umma_desc_sw128(b0);
        const uint32_t d = tm + (it % nacc) * N;
        tc_mma_f16(d, ad, bd, idesc, 1u);
        tc_mma_f16(d, ad + 2, bd + 2, idesc, 1u);
        tc_mma_f16(d, ad + 4, bd + 4, idesc, 1u);
        if (commit_every && (it % commit_every) == commit_every - 1) { tc_commit(&bar_mid); }
      }
      long long t1 = clock64();
      tc_commit(&bar);
      // drain: wait for all commits (phase flips per commit; just spin on the last one via try_wait loop)
      mbar_wait(&bar, par);
      long long t2 = clock64();
      out[blockIdx.x * 2] = t1 - t0;
      out[blockIdx.x * 2 + 1] = t2 - t0;
    }
    __syncwarp();
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

template <int N> void run(int shift, int nacc, int commit_every) {
  long long* d; cudaMalloc(&d, 148 * 2 * sizeof(long long));
  cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 600;
  k<N><<<148, 128, 100 * 1024>>>(d, iters, shift, nacc, commit_every);
  k<N><<<148, 128, 100 * 1024>>>(d, iters, shift, nacc, commit_every);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148 * 2]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double issue = 0, total = 0; for (int i = 0; i < 148; ++i) { issue += h[2 * i]; total += h[2 * i + 1]; }
  printf("N=%3d shift_rows=%2d nacc=%d commit_every=%d: issue %.1f cyc/MMA, issue+drain %.1f cyc/MMA  (tensor floor %d) %s\n", N, shift, nacc,
         commit_every, issue / 148 / (iters * 3), total / 148 / (iters * 3), N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int ce : {0, 1, 3, 9, 27, 81}) run<48>(1, 3, ce);
  for (int ce : {0, 1, 3, 9, 27, 81}) run<96>(1, 3, ce);
  for (int ce : {0, 1, 9}) run<128>(1, 2, ce);
  for (int ce : {0, 1, 9}) run<192>(1, 1, ce);
  return 0;
}
