// Instruction-throughput microbenchmark for the soft-max inner loop on sm_100a: MUFU.EX2, F2FP pack, FFMA, FMNMX
// per SM sub-partition, as a function of resident warps.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sfu_rate sfu_rate.cu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = threadIdx.x * 0.001f + i * 0.1f;
  unsigned acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (MODE == 1) v[i] = fmaf(v[i], 1.0001f, 0.5f);
      if (MODE == 2) { __half2 h = __floats2half2_rn(v[i], v[(i + 1) & 15]); acc ^= *reinterpret_cast<unsigned*>(&h); v[i] += 1.f; }
      if (MODE == 3) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i])); v[i] = fmaf(v[i], 0.5f, -1.f); }
    }
    if (MODE == 4) {     // the real mix: 16 x (ffma, ex2), 8 packs, 8 max3
#pragma unroll
      for (int i = 0; i < 16; ++i) { float t = fmaf(v[i], 0.999f, -0.3f); asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(v[i]) : "f"(t)); }
#pragma unroll
      for (int i = 0; i < 16; i += 2) { __half2 h = __floats2half2_rn(v[i], v[i + 1]); acc ^= *reinterpret_cast<unsigned*>(&h); }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  const char* names[5] = {"MUFU.EX2", "FFMA", "F2FP pack (+FADD)", "EX2+FFMA dependent pairs", "softmax mix (16 ffma+ex2, 8 pack)"};
  for (int mode = 0; mode < 5; ++mode)
    for (int warps = 4; warps <= 32; warps *= 2) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, warps * 32>>>(out, iters, cyc);
        if (mode == 1) k<1><<<148, warps * 32>>>(out, iters, cyc);
        if (mode == 2) k<2><<<148, warps * 32>>>(out, iters, cyc);
        if (mode == 3) k<3><<<148, warps * 32>>>(out, iters, cyc);
        if (mode == 4) k<4><<<148, warps * 32>>>(out, iters, cyc);
      }
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      const double per_warp_instr = (double)c / (iters * 16.0);
      printf("%-36s warps/SM %2d (per SMSP %d): %.2f clk per 16-op group per warp -> %.1f ops/clk/SM\n", names[mode], warps, warps / 4,
             per_warp_instr * 16, warps * 32 * 16.0 * iters / c);
    }
  return 0;
}
