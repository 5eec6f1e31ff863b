// HBM-bound kernels of the training step (SURVEY.md section 8 row a-T; lib/core/function.py:113-138): train-mode
// BatchNorm forward / backward, bias gradients, zero-stuffing for stride-2 data / weight gradients, the backward of the
// exchange/fuse sum, layout changes, the soft-max backward of both attentions (with the hash dropout mask), the stem
// weight gradient, split-partial reduction into torch-layout float32 gradients and the per-step weight re-pack.
// Reductions go through per-block float64 partials summed in index order: bit-reproducible, no atomics.
#include "common.cuh"
#include "internal.h"

namespace buctd {

namespace {

constexpr int kRedBlocks = BUCTD_BN_BLOCKS;

template <typename T> __device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  float2 t;
  t = T16<T>::unpack2(u.x); f[0] = t.x; f[1] = t.y;
  t = T16<T>::unpack2(u.y); f[2] = t.x; f[3] = t.y;
  t = T16<T>::unpack2(u.z); f[4] = t.x; f[5] = t.y;
  t = T16<T>::unpack2(u.w); f[6] = t.x; f[7] = t.y;
}
template <typename T> __device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  u.x = T16<T>::pack2(f[0], f[1]); u.y = T16<T>::pack2(f[2], f[3]);
  u.z = T16<T>::pack2(f[4], f[5]); u.w = T16<T>::pack2(f[6], f[7]);
  return u;
}
template <typename T> __device__ __forceinline__ uint4 ld8(const void* base, long long row, int cs, int c0) {
  return *reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(base) + row * cs + c0);
}
template <typename T> __device__ __forceinline__ void st8(void* base, long long row, int cs, int c0, const uint4& u) {
  *reinterpret_cast<uint4*>(reinterpret_cast<T*>(base) + row * cs + c0) = u;
}

// Column reduction skeleton: NV values per channel accumulated over the rows of a [rows][cs] matrix.
// Block = 256 threads = (row lanes) x (8-channel groups); per-block partial[blk][v][c] in float64.
template <int NV, typename F>
__device__ __forceinline__ void col_reduce(long long rows, int c, double* partial, F&& load) {
  extern __shared__ double s_red[];                 // [nrl][NV][gc*8]
  const int groups = (c + 7) / 8;
  const int nrl = blockDim.x / groups;
  const int g = threadIdx.x % groups, rl = threadIdx.x / groups;
  float acc[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[v][j] = 0.f;
  double dacc[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int j = 0; j < 8; ++j) dacc[v][j] = 0.0;
  if (rl < nrl) {
    // A thread walks its rows in runs of <= 64 (then the float accumulators are flushed into float64), four rows per
    // trip: the four rows' loads are independent of the accumulators, so they are in flight together -- with one 16-byte
    // load per thread and trip the wide layers (256 channels: 8 row lanes per block, ~100 trips) ran at DRAM latency,
    // 0.2 TB/s.  Same additions in the same order as the one-row loop.
    const long long step = (long long)gridDim.x * nrl;
    long long r = (long long)blockIdx.x * nrl + rl;
    while (r < rows) {
      const long long left = (rows - r + step - 1) / step;
      const int n = left < 64 ? (int)left : 64;
      int i = 0;
      for (; i + 4 <= n; i += 4, r += 4 * step) {
        load(r, g * 8, acc);
        load(r + step, g * 8, acc);
        load(r + 2 * step, g * 8, acc);
        load(r + 3 * step, g * 8, acc);
      }
      for (; i < n; ++i, r += step) load(r, g * 8, acc);
#pragma unroll
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int j = 0; j < 8; ++j) { dacc[v][j] += (double)acc[v][j]; acc[v][j] = 0.f; }
    }
    const int cw = groups * 8;
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int j = 0; j < 8; ++j) s_red[((size_t)rl * NV + v) * cw + g * 8 + j] = dacc[v][j];
  }
  __syncthreads();
  const int cw = groups * 8;
  for (int i = threadIdx.x; i < NV * cw; i += blockDim.x) {
    const int v = i / cw, ch = i - v * cw;
    if (ch >= c) continue;
    double t = 0.0;
    for (int l = 0; l < nrl; ++l) t += s_red[((size_t)l * NV + v) * cw + ch];
    partial[((size_t)blockIdx.x * NV + v) * c + ch] = t;
  }
}
size_t col_reduce_smem(int c, int nv) {
  const int groups = (c + 7) / 8;
  const int nrl = 256 / groups;
  return sizeof(double) * (size_t)nrl * nv * groups * 8;
}

// row indices of one tensor fit 32 bits (<= 7.2 M rows at batch 256): 32-bit divisions, not 64-bit ones
__device__ __forceinline__ bool is_pad_row(unsigned r, int H, int W) {
  const unsigned ry = r / (unsigned)(W + 1);
  const unsigned x = r - ry * (unsigned)(W + 1);
  const unsigned y = ry % (unsigned)(H + 1);
  return x == (unsigned)W || y == (unsigned)H;
}

}  // namespace

// =================================================================================================
// BatchNorm statistics
// =================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) bn_stats_kernel(buctd_bn_stats_args a) {
  pdl_launch_dependents();
  pdl_wait();
  const long long rows = (long long)a.batch * (a.h + 1) * (a.w + 1);
  col_reduce<2>(rows, a.c, a.scratch, [&](long long r, int c0, float (&acc)[2][8]) {
    float f[8];
    unpack8<T>(ld8<T>(a.z, r, a.z_cs, c0), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[0][j] += f[j]; acc[1][j] = fmaf(f[j], f[j], acc[1][j]); }
  });
}
// one warp per channel: lane l adds the partials of blocks l, l + 32, ... in order, then a fixed shuffle tree (bit-reproducible)
__device__ __forceinline__ void warp_sum2(const double* partial, int nblocks, int nv, int v0, int v1, int c, int cstride, double* o0, double* o1) {
  const int lane = threadIdx.x & 31;
  double s = 0.0, ss = 0.0;
  for (int b = lane; b < nblocks; b += 32) {
    s += partial[((size_t)b * nv + v0) * cstride + c];
    if (v1 >= 0) ss += partial[((size_t)b * nv + v1) * cstride + c];
  }
  for (int o = 16; o; o >>= 1) {
    s += __shfl_xor_sync(0xffffffff, s, o);
    ss += __shfl_xor_sync(0xffffffff, ss, o);
  }
  *o0 = s; *o1 = ss;
}
__global__ void __launch_bounds__(256) bn_stats_finalize_kernel(buctd_bn_stats_args a, int nblocks) {
  pdl_launch_dependents();
  pdl_wait();
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= a.c) return;
  double s, ss;
  warp_sum2(a.scratch, nblocks, 2, 0, 1, c, a.c, &s, &ss);
  if ((threadIdx.x & 31) != 0) return;
  const double n = (double)a.batch * a.h * a.w;
  const double mean = s / n;
  double var = ss / n - mean * mean;
  if (var < 0.0) var = 0.0;
  a.mean_rstd[c] = (float)mean;
  a.mean_rstd[a.c + c] = (float)(1.0 / sqrt(var + (double)a.eps));
  if (a.running_mean != nullptr && a.running_var != nullptr) {
    const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
    a.running_mean[c] = (float)((1.0 - a.momentum) * (double)a.running_mean[c] + a.momentum * mean);
    a.running_var[c] = (float)((1.0 - a.momentum) * (double)a.running_var[c] + a.momentum * unbiased);
  }
}
int launch_bn_stats(const buctd_bn_stats_args& a, cudaStream_t st) {
  BUCTD_CHECK(a.z && a.mean_rstd && a.scratch, BUCTD_ERR_INVALID, "bn_stats: null pointer");
  BUCTD_CHECK(a.c >= 1 && a.c <= 2048 && a.z_cs % 8 == 0 && a.c <= a.z_cs && (a.c + 7) / 8 <= 256, BUCTD_ERR_INVALID, "bn_stats: bad channels");
  BUCTD_CHECK(a.batch >= 1, BUCTD_ERR_INVALID, "bn_stats: empty batch");
  const size_t smem = col_reduce_smem(a.c, 2);
  if (a.dtype == BUCTD_F16) {
    BUCTD_CUDA(cudaFuncSetAttribute(bn_stats_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    launch_pdl(bn_stats_kernel<__half>, dim3(kRedBlocks), dim3(256), smem, st, a);
  } else {
    BUCTD_CUDA(cudaFuncSetAttribute(bn_stats_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    launch_pdl(bn_stats_kernel<__nv_bfloat16>, dim3(kRedBlocks), dim3(256), smem, st, a);
  }
  launch_pdl(bn_stats_finalize_kernel, dim3(ceil_div(a.c, 8)), dim3(256), 0, st, a, kRedBlocks);
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}

// =================================================================================================
// BatchNorm apply (+ residual, ReLU)
// =================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) bn_apply_kernel(buctd_bn_apply_args a) {
  pdl_launch_dependents();
  pdl_wait();
  const unsigned rows = (unsigned)a.batch * (a.h + 1) * (a.w + 1);
  const unsigned groups = a.out_cs / 8;
  const unsigned total = rows * groups;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned r = i / groups;
    const int c0 = (int)(i - r * groups) * 8;
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = 0.f;
    if (c0 < a.c && !is_pad_row(r, a.h, a.w)) {
      float z[8], rs[8];
      unpack8<T>(ld8<T>(a.z, r, a.z_cs, c0), z);
      if (a.residual) unpack8<T>(ld8<T>(a.residual, r, a.res_cs, c0), rs);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = c0 + j;
        if (c < a.c) {
          float v = (z[j] - a.mean_rstd[c]) * a.mean_rstd[a.c + c] * a.gamma[c] + a.beta[c];
          if (a.residual) v += rs[j];
          o[j] = a.relu ? fmaxf(v, 0.f) : v;
        }
      }
    }
    st8<T>(a.out, r, a.out_cs, c0, pack8<T>(o));
  }
}
int launch_bn_apply(const buctd_bn_apply_args& a, cudaStream_t st) {
  BUCTD_CHECK(a.z && a.mean_rstd && a.gamma && a.beta && a.out, BUCTD_ERR_INVALID, "bn_apply: null pointer");
  BUCTD_CHECK(a.z_cs % 8 == 0 && a.out_cs % 8 == 0 && (a.residual == nullptr || a.res_cs % 8 == 0) && a.c <= a.z_cs && a.c <= a.out_cs &&
                  a.out_cs <= a.z_cs + 7, BUCTD_ERR_INVALID, "bn_apply: bad channel strides");
  const long long total = (long long)a.batch * (a.h + 1) * (a.w + 1) * (a.out_cs / 8);
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  if (blocks == 0) return BUCTD_OK;
  if (a.dtype == BUCTD_F16) launch_pdl(bn_apply_kernel<__half>, dim3(blocks), dim3(256), 0, st, a);
  else launch_pdl(bn_apply_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, st, a);
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}

// =================================================================================================
// BatchNorm backward
// =================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(buctd_bn_bwd_args a) {
  pdl_launch_dependents();
  pdl_wait();
  const long long rows = (long long)a.batch * (a.h + 1) * (a.w + 1);
  col_reduce<2>(rows, a.c, a.scratch, [&](long long r, int c0, float (&acc)[2][8]) {
    float g[8], z[8], y[8];
    unpack8<T>(ld8<T>(a.dy, r, a.dy_cs, c0), g);
    unpack8<T>(ld8<T>(a.z, r, a.z_cs, c0), z);
    if (a.relu) unpack8<T>(ld8<T>(a.y, r, a.y_cs, c0), y);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      if (c < a.c) {
        const float gj = (a.relu && !(y[j] > 0.f)) ? 0.f : g[j];
        const float zh = (z[j] - a.mean_rstd[c]) * a.mean_rstd[a.c + c];
        acc[0][j] += gj;
        acc[1][j] = fmaf(gj, zh, acc[1][j]);
      }
    }
  });
}
__global__ void __launch_bounds__(256) bn_bwd_finalize_kernel(buctd_bn_bwd_args a, int nblocks) {
  pdl_launch_dependents();
  pdl_wait();
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= a.c) return;
  double s1, s2;
  warp_sum2(a.scratch, nblocks, 2, 0, 1, c, a.c, &s1, &s2);
  if ((threadIdx.x & 31) != 0) return;
  // note: pad rows have z = 0 -> zhat != 0, but dy = 0 there (gradient tensors keep zero pads), so they add nothing
  if (a.dbeta) a.dbeta[c] = (float)(s1 * (double)a.inv_scale);
  if (a.dgamma) a.dgamma[c] = (float)(s2 * (double)a.inv_scale);
  const double n = (double)a.batch * a.h * a.w;
  double* m = a.scratch + (size_t)nblocks * 2 * a.c;
  m[c] = s1 / n;
  m[a.c + c] = s2 / n;
}
template <typename T>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(buctd_bn_bwd_args a, int nblocks) {
  pdl_launch_dependents();
  pdl_wait();
  const unsigned rows = (unsigned)a.batch * (a.h + 1) * (a.w + 1);
  const unsigned groups = a.dz_cs / 8;
  const unsigned total = rows * groups;
  const double* m = a.scratch + (size_t)nblocks * 2 * a.c;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned r = i / groups;
    const int c0 = (int)(i - r * groups) * 8;
    float o[8], gr[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { o[j] = 0.f; gr[j] = 0.f; }
    const bool live = c0 < a.c && !is_pad_row(r, a.h, a.w);
    if (live) {
      float g[8], z[8], y[8];
      unpack8<T>(ld8<T>(a.dy, r, a.dy_cs, c0), g);
      unpack8<T>(ld8<T>(a.z, r, a.z_cs, c0), z);
      if (a.relu) unpack8<T>(ld8<T>(a.y, r, a.y_cs, c0), y);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = c0 + j;
        if (c < a.c) {
          const float gj = (a.relu && !(y[j] > 0.f)) ? 0.f : g[j];
          const float rstd = a.mean_rstd[a.c + c];
          const float zh = (z[j] - a.mean_rstd[c]) * rstd;
          o[j] = a.gamma[c] * rstd * (gj - (float)m[c] - zh * (float)m[a.c + c]);
          gr[j] = gj;
        }
      }
    }
    st8<T>(a.dz, r, a.dz_cs, c0, pack8<T>(o));
    if (a.dres != nullptr && c0 < a.dres_cs) {
      if (a.res_accumulate && live) {
        float old[8];
        unpack8<T>(ld8<T>(a.dres, r, a.dres_cs, c0), old);
#pragma unroll
        for (int j = 0; j < 8; ++j) gr[j] += old[j];
      }
      if (!a.res_accumulate || live) st8<T>(a.dres, r, a.dres_cs, c0, pack8<T>(gr));
    }
  }
}
int launch_bn_bwd(const buctd_bn_bwd_args& a, cudaStream_t st) {
  BUCTD_CHECK(a.dy && a.z && a.mean_rstd && a.gamma && a.scratch && a.dz && (!a.relu || a.y), BUCTD_ERR_INVALID, "bn_bwd: null pointer");
  BUCTD_CHECK(a.dy_cs % 8 == 0 && a.z_cs % 8 == 0 && a.dz_cs % 8 == 0 && (!a.relu || a.y_cs % 8 == 0) && (!a.dres || a.dres_cs % 8 == 0) &&
                  a.c <= a.dy_cs && a.c <= a.z_cs && a.c <= a.dz_cs && a.dz_cs <= a.dy_cs + 7 && a.dz_cs <= a.z_cs + 7 && (a.c + 7) / 8 <= 256,
              BUCTD_ERR_INVALID, "bn_bwd: bad channel strides");
  const size_t smem = col_reduce_smem(a.c, 2);
  const long long total = (long long)a.batch * (a.h + 1) * (a.w + 1) * (a.dz_cs / 8);
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  if (a.dtype == BUCTD_F16) {
    BUCTD_CUDA(cudaFuncSetAttribute(bn_bwd_reduce_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    launch_pdl(bn_bwd_reduce_kernel<__half>, dim3(kRedBlocks), dim3(256), smem, st, a);
    launch_pdl(bn_bwd_finalize_kernel, dim3(ceil_div(a.c, 8)), dim3(256), 0, st, a, kRedBlocks);
    launch_pdl(bn_bwd_apply_kernel<__half>, dim3(blocks), dim3(256), 0, st, a, kRedBlocks);
  } else {
    BUCTD_CUDA(cudaFuncSetAttribute(bn_bwd_reduce_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    launch_pdl(bn_bwd_reduce_kernel<__nv_bfloat16>, dim3(kRedBlocks), dim3(256), smem, st, a);
    launch_pdl(bn_bwd_finalize_kernel, dim3(ceil_div(a.c, 8)), dim3(256), 0, st, a, kRedBlocks);
    launch_pdl(bn_bwd_apply_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, st, a, kRedBlocks);
  }
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}

// =================================================================================================
// Bias gradient: column sums of a 16-bit [rows][cs] matrix
// =================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(buctd_colsum_args a) {
  pdl_launch_dependents();
  pdl_wait();
  col_reduce<1>(a.rows, a.c, a.scratch, [&](long long r, int c0, float (&acc)[1][8]) {
    float f[8];
    unpack8<T>(ld8<T>(a.dy, r, a.cs, c0), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[0][j] += f[j];
  });
}
__global__ void __launch_bounds__(256) colsum_finalize_kernel(buctd_colsum_args a, int nblocks) {
  pdl_launch_dependents();
  pdl_wait();
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= a.c) return;
  double s, unused;
  warp_sum2(a.scratch, nblocks, 1, 0, -1, c, a.c, &s, &unused);
  if ((threadIdx.x & 31) != 0) return;
  const float v = (float)(s * (double)a.inv_scale);
  a.dbias[c] = a.accumulate ? a.dbias[c] + v : v;
}
int launch_colsum(const buctd_colsum_args& a, cudaStream_t st) {
  BUCTD_CHECK(a.dy && a.dbias && a.scratch, BUCTD_ERR_INVALID, "colsum: null pointer");
  BUCTD_CHECK(a.cs % 8 == 0 && a.c >= 1 && (a.c + 7) / 8 * 8 <= a.cs + 7 && (a.c + 7) / 8 <= 256 * 4, BUCTD_ERR_INVALID, "colsum: bad channels %d / %d", a.c, a.cs);
  // wide matrices (the HW x HW projection: 6912 columns) go in column slices of 2048
  for (int c0 = 0; c0 < a.c; c0 += 2048) {
    buctd_colsum_args b = a;
    b.c = a.c - c0 < 2048 ? a.c - c0 : 2048;
    b.dy = reinterpret_cast<const char*>(a.dy) + (size_t)c0 * 2;
    b.dbias = a.dbias + c0;
    BUCTD_CHECK((b.c + 7) / 8 * 8 <= a.cs - c0 + 7, BUCTD_ERR_INVALID, "colsum: slice beyond the row stride");
    const size_t smem = col_reduce_smem(b.c, 1);
    if (a.dtype == BUCTD_F16) {
      BUCTD_CUDA(cudaFuncSetAttribute(colsum_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      launch_pdl(colsum_kernel<__half>, dim3(kRedBlocks), dim3(256), smem, st, b);
    } else {
      BUCTD_CUDA(cudaFuncSetAttribute(colsum_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      launch_pdl(colsum_kernel<__nv_bfloat16>, dim3(kRedBlocks), dim3(256), smem, st, b);
    }
    launch_pdl(colsum_finalize_kernel, dim3(ceil_div(b.c, 8)), dim3(256), 0, st, b, kRedBlocks);
  }
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}

// =================================================================================================
// Zero-stuffing (stride-2 gradients)
// =================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) zero_stuff_kernel(buctd_zero_stuff_args a) {
  pdl_launch_dependents();
  pdl_wait();
  const int OH = 2 * a.h, OW = 2 * a.w;
  const unsigned groups = a.cs / 8;
  const unsigned total = (unsigned)a.batch * (OH + 1) * (OW + 1) * groups;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned r = i / groups;
    const int c0 = (int)(i - r * groups) * 8;
    const unsigned ry = r / (unsigned)(OW + 1);
    const int X = (int)(r - ry * (unsigned)(OW + 1));
    const int b = (int)(ry / (unsigned)(OH + 1));
    const int Y = (int)(ry - (unsigned)b * (unsigned)(OH + 1));
    uint4 u = make_uint4(0, 0, 0, 0);
    if (!(X & 1) && !(Y & 1) && X < OW && Y < OH) u = ld8<T>(a.in, ((long long)b * (a.h + 1) + (Y >> 1)) * (a.w + 1) + (X >> 1), a.cs, c0);
    st8<T>(a.out, r, a.cs, c0, u);
  }
}
int launch_zero_stuff(const buctd_zero_stuff_args& a, cudaStream_t st) {
  BUCTD_CHECK(a.in && a.out && a.cs % 8 == 0, BUCTD_ERR_INVALID, "zero_stuff: bad arguments");
  const long long total = (long long)a.batch * (2 * a.h + 1) * (2 * a.w + 1) * (a.cs / 8);
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  if (blocks == 0) return BUCTD_OK;
  if (a.dtype == BUCTD_F16) launch_pdl(zero_stuff_kernel<__half>, dim3(blocks), dim3(256), 0, st, a);
  else launch_pdl(zero_stuff_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, st, a);
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}

// =================================================================================================
// Backward of the exchange/fuse sum
// =================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) fuse_bwd_kernel(buctd_fuse_bwd_args a, int j) {
  pdl_launch_dependents();
  pdl_wait();
  const int sh = a.term_shift[j];
  const int th = a.h >> sh, tw = a.w >> sh;
  const int tcs = a.term_cs[j];
  const unsigned groups = tcs / 8;
  const unsigned total = (unsigned)a.batch * (th + 1) * (tw + 1) * groups;
  const int n = 1 << sh;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned r = i / groups;
    const int c0 = (int)(i - r * groups) * 8;
    const unsigned ry = r / (unsigned)(tw + 1);
    const int X = (int)(r - ry * (unsigned)(tw + 1));
    const int b = (int)(ry / (unsigned)(th + 1));
    const int Y = (int)(ry - (unsigned)b * (unsigned)(th + 1));
    float acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.f;
    const bool live = X < tw && Y < th && c0 < a.c;
    if (live) {
      for (int dy = 0; dy < n; ++dy)
        for (int dx = 0; dx < n; ++dx) {
          const long long fr = ((long long)b * (a.h + 1) + (Y * n + dy)) * (a.w + 1) + (X * n + dx);
          float g[8], o[8];
          unpack8<T>(ld8<T>(a.dout, fr, a.cs, c0), g);
          if (a.relu) unpack8<T>(ld8<T>(a.out, fr, a.cs, c0), o);
#pragma unroll
          for (int q = 0; q < 8; ++q) acc[q] += (a.relu && !(o[q] > 0.f)) ? 0.f : g[q];
        }
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (c0 + q >= a.c) acc[q] = 0.f;
      if (a.term_accumulate[j]) {
        float old[8];
        unpack8<T>(ld8<T>(a.dterm[j], r, tcs, c0), old);
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] += old[q];
      }
    }
    if (!a.term_accumulate[j] || live) st8<T>(a.dterm[j], r, tcs, c0, pack8<T>(acc));
  }
}
int launch_fuse_bwd(const buctd_fuse_bwd_args& a, cudaStream_t st) {
  BUCTD_CHECK(a.dout && (!a.relu || a.out) && a.nterms >= 1 && a.nterms <= 4 && a.cs % 8 == 0, BUCTD_ERR_INVALID, "fuse_bwd: bad arguments");
  for (int j = 0; j < a.nterms; ++j) {
    BUCTD_CHECK(a.dterm[j] && a.term_cs[j] % 8 == 0 && a.term_cs[j] <= a.cs + 7 && a.term_shift[j] >= 0 && a.term_shift[j] <= 3 &&
                    a.h % (1 << a.term_shift[j]) == 0 && a.w % (1 << a.term_shift[j]) == 0,
                BUCTD_ERR_INVALID, "fuse_bwd: bad term %d", j);
    const long long total = (long long)a.batch * ((a.h >> a.term_shift[j]) + 1) * ((a.w >> a.term_shift[j]) + 1) * (a.term_cs[j] / 8);
    const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    if (blocks == 0) continue;
    if (a.dtype == BUCTD_F16) launch_pdl(fuse_bwd_kernel<__half>, dim3(blocks), dim3(256), 0, st, a, j);
    else launch_pdl(fuse_bwd_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, st, a, j);
  }
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}

// =================================================================================================
// out (+)= a
// =================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) add16_kernel(buctd_add16_args a) {
  pdl_launch_dependents();
  pdl_wait();
  const int groups = (a.c + 7) / 8;
  const long long total = a.rows * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / groups;
    const int c0 = (int)(i - r * groups) * 8;
    float x[8];
    unpack8<T>(ld8<T>(a.a, r, a.cs_a, c0), x);
    if (a.accumulate) {
      float o[8];
      unpack8<T>(ld8<T>(a.out, r, a.cs_out, c0), o);
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] += o[j];
    }
    st8<T>(a.out, r, a.cs_out, c0, pack8<T>(x));
  }
}
int launch_add16(const buctd_add16_args& a, cudaStream_t st) {
  BUCTD_CHECK(a.a && a.out && a.cs_a % 8 == 0 && a.cs_out % 8 == 0 && (a.c + 7) / 8 * 8 <= a.cs_a && (a.c + 7) / 8 * 8 <= a.cs_out, BUCTD_ERR_INVALID,
              "add16: bad arguments");
  const long long total = a.rows * ((a.c + 7) / 8);
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  if (blocks == 0) return BUCTD_OK;
  if (a.dtype == BUCTD_F16) launch_pdl(add16_kernel<__half>, dim3(blocks), dim3(256), 0, st, a);
  else launch_pdl(add16_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, st, a);
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}

// =================================================================================================
// planar [B][C][ps] <-> act16 [B][H+1][W+1][cs]: 32 pixels x 32 channels through shared memory
// =================================================================================================
template <typename T, bool TO_ACT>
__global__ void __launch_bounds__(256) relayout_kernel(buctd_relayout_args a) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_t[32][33];
  const int b = blockIdx.z;
  const int hw = a.h * a.w;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  const long long rows = (long long)(a.h + 1) * (a.w + 1);
  if (TO_ACT) {
    const T* src = reinterpret_cast<const T*>(a.in) + (long long)b * a.c * a.p_stride;
    for (int cc = ty; cc < 32; cc += 8) {
      const int c = c0 + cc, p = p0 + tx;
      s_t[cc][tx] = (c < a.c && p < hw) ? T16<T>::to_f(src[(long long)c * a.p_stride + p]) : 0.f;
    }
    __syncthreads();
    T* dst = reinterpret_cast<T*>(a.out) + (long long)b * rows * a.cs;
    for (int pp = ty; pp < 32; pp += 8) {
      const int p = p0 + pp, c = c0 + tx;
      if (p >= hw || c >= a.cs) continue;
      const long long r = (long long)(p / a.w) * (a.w + 1) + (p % a.w);
      float v = c < a.c ? s_t[tx][pp] : 0.f;
      if (a.accumulate && c < a.c) v += T16<T>::to_f(dst[r * a.cs + c]);
      dst[r * a.cs + c] = T16<T>::from_f(v);
    }
  } else {
    const T* src = reinterpret_cast<const T*>(a.in) + (long long)b * rows * a.cs;
    for (int pp = ty; pp < 32; pp += 8) {
      const int p = p0 + pp, c = c0 + tx;
      float v = 0.f;
      if (p < hw && c < a.c) v = T16<T>::to_f(src[((long long)(p / a.w) * (a.w + 1) + (p % a.w)) * a.cs + c]);
      s_t[pp][tx] = v;
    }
    __syncthreads();
    T* dst = reinterpret_cast<T*>(a.out) + (long long)b * a.c * a.p_stride;
    for (int cc = ty; cc < 32; cc += 8) {
      const int c = c0 + cc, p = p0 + tx;
      if (c >= a.c || p >= hw) continue;
      float v = s_t[tx][cc];
      if (a.accumulate) v += T16<T>::to_f(dst[(long long)c * a.p_stride + p]);
      dst[(long long)c * a.p_stride + p] = T16<T>::from_f(v);
    }
  }
}
template <bool TO_ACT>
static int launch_relayout(const buctd_relayout_args& a, cudaStream_t st) {
  BUCTD_CHECK(a.in && a.out && a.cs % 8 == 0 && a.c <= a.cs && a.p_stride >= a.h * a.w, BUCTD_ERR_INVALID, "relayout: bad arguments");
  if (a.batch == 0) return BUCTD_OK;
  if (TO_ACT && !a.accumulate) {
    // pads and spare channels of the act16 tensor must be zero: clear it, then scatter the valid pixels
    BUCTD_CUDA(cudaMemsetAsync(a.out, 0, (size_t)a.batch * (a.h + 1) * (a.w + 1) * a.cs * 2, st));
  }
  dim3 grid(ceil_div(a.h * a.w, 32), ceil_div(TO_ACT ? a.cs : a.c, 32), a.batch);
  if (a.dtype == BUCTD_F16) launch_pdl(relayout_kernel<__half, TO_ACT>, dim3(grid), dim3(256), 0, st, a);
  else launch_pdl(relayout_kernel<__nv_bfloat16, TO_ACT>, dim3(grid), dim3(256), 0, st, a);
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}
int launch_planar_to_act16(const buctd_relayout_args& a, cudaStream_t st) { return launch_relayout<true>(a, st); }
int launch_act16_to_planar(const buctd_relayout_args& a, cudaStream_t st) { return launch_relayout<false>(a, st); }

// =================================================================================================
// Channel-attention soft-max backward: one warp per (crop, c1) row of the [C x C] matrices
// =================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) chan_softmax_bwd_kernel(buctd_chan_softmax_bwd_args a) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x & 31;
  if (row >= a.batch * a.c) return;
  const int b = row / a.c, c1 = row - b * a.c;
  const float* pr = a.probs + (long long)b * a.prob_stride + (long long)c1 * a.c;
  const float* da = a.da_d + ((long long)b * a.c + c1) * a.c;
  const uint32_t thr = drop_threshold(a.drop_p);
  const float keep_scale = a.drop_p > 0.f ? 1.f / (1.f - a.drop_p) : 1.f;
  const uint32_t seed = a.seed ^ (a.seed_xor ? *a.seed_xor : 0u);
  float dot = 0.f;
  for (int c2 = lane; c2 < a.c; c2 += 32) {
    const bool keep = a.drop_p <= 0.f || drop_hash(seed, (uint32_t)b, (uint32_t)c1, (uint32_t)c2) >= thr;
    dot += (keep ? da[c2] * keep_scale : 0.f) * pr[c2];
  }
  for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(0xffffffff, dot, o);
  T* ds = reinterpret_cast<T*>(a.ds) + ((long long)b * a.c + c1) * a.ld;
  T* pd = reinterpret_cast<T*>(a.probs_d) + ((long long)b * a.c + c1) * a.ld;
  for (int c2 = lane; c2 < a.ld; c2 += 32) {
    float dsv = 0.f, pdv = 0.f;
    if (c2 < a.c) {
      const bool keep = a.drop_p <= 0.f || drop_hash(seed, (uint32_t)b, (uint32_t)c1, (uint32_t)c2) >= thr;
      const float p = pr[c2];
      dsv = p * ((keep ? da[c2] * keep_scale : 0.f) - dot) * a.scale;
      pdv = keep ? p * keep_scale : 0.f;
    }
    ds[c2] = T16<T>::from_f(dsv);
    pd[c2] = T16<T>::from_f(pdv);
  }
}
int launch_chan_softmax_bwd(const buctd_chan_softmax_bwd_args& a, cudaStream_t st) {
  BUCTD_CHECK(a.probs && a.da_d && a.ds && a.probs_d && a.ld >= a.c && a.ld % 8 == 0, BUCTD_ERR_INVALID, "chan_softmax_bwd: bad arguments");
  if (a.batch == 0) return BUCTD_OK;
  const int blocks = ceil_div(a.batch * a.c, 8);
  if (a.dtype == BUCTD_F16) launch_pdl(chan_softmax_bwd_kernel<__half>, dim3(blocks), dim3(256), 0, st, a);
  else launch_pdl(chan_softmax_bwd_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, st, a);
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}

// =================================================================================================
// Position-attention soft-max backward (element-wise over [rows x ld] per crop)
// =================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) pos_softmax_bwd_kernel(buctd_pos_softmax_bwd_args a) {
  pdl_launch_dependents();
  pdl_wait();
  const int rows = (a.h + 1) * (a.w + 1);
  const unsigned groups = a.ld / 8;
  const unsigned total = (unsigned)a.batch * rows * groups;            // < 2^32 (checked by the launcher)
  const uint32_t thr = drop_threshold(a.drop_p);
  const float keep_scale = a.drop_p > 0.f ? 1.f / (1.f - a.drop_p) : 1.f;
  const uint32_t seed = a.seed ^ (a.seed_xor ? *a.seed_xor : 0u);
  const unsigned wp = a.w + 1;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned qr = i / groups;                  // global query row
    const int k0 = (int)(i - qr * groups) * 8;
    const int b = (int)(qr / (unsigned)rows), q = (int)(qr - (unsigned)b * (unsigned)rows);
    const bool qv = ((unsigned)q % wp) != (unsigned)a.w && ((unsigned)q / wp) != (unsigned)a.h;
    float pd[8], ds[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { pd[j] = 0.f; ds[j] = 0.f; }
    if (qv) {
      const float lse = a.lse2[qr], dl = a.delta[qr];
      const float4* sp = reinterpret_cast<const float4*>(a.s + (long long)qr * a.ld + k0);
      const float4* dp = reinterpret_cast<const float4*>(a.dpd + (long long)qr * a.ld + k0);
      const float4 s0 = sp[0], s1 = sp[1], d0 = dp[0], d1 = dp[1];
      const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
      const float dv[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const unsigned k = (unsigned)(k0 + j);
        if (k < (unsigned)rows && (k % wp) != (unsigned)a.w && (k / wp) != (unsigned)a.h) {
          const float p = exp2f(sv[j] - lse);
          const bool keep = a.drop_p <= 0.f || drop_hash(seed, (uint32_t)b, (uint32_t)q, (uint32_t)k) >= thr;
          pd[j] = keep ? p * keep_scale : 0.f;
          ds[j] = p * ((keep ? dv[j] * keep_scale : 0.f) - dl) * a.inv_sqrt_c;
        }
      }
    }
    st8<T>(a.pd, qr, a.ld, k0, pack8<T>(pd));
    st8<T>(a.ds, qr, a.ld, k0, pack8<T>(ds));
  }
}
int launch_pos_softmax_bwd(const buctd_pos_softmax_bwd_args& a, cudaStream_t st) {
  BUCTD_CHECK(a.s && a.dpd && a.lse2 && a.delta && a.pd && a.ds, BUCTD_ERR_INVALID, "pos_softmax_bwd: null pointer");
  BUCTD_CHECK(a.ld % 8 == 0 && a.ld >= (a.h + 1) * (a.w + 1), BUCTD_ERR_INVALID, "pos_softmax_bwd: bad row stride %d", a.ld);
  const long long total = (long long)a.batch * (a.h + 1) * (a.w + 1) * (a.ld / 8);
  BUCTD_CHECK(total < (1ll << 32), BUCTD_ERR_INVALID, "pos_softmax_bwd: %lld element groups exceed 32-bit indexing (split the batch)", total);
  const int blocks = (int)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
  if (blocks == 0) return BUCTD_OK;
  if (a.dtype == BUCTD_F16) launch_pdl(pos_softmax_bwd_kernel<__half>, dim3(blocks), dim3(256), 0, st, a);
  else launch_pdl(pos_softmax_bwd_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, st, a);
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}

// =================================================================================================
// delta[r] = sum_c a[r][c] b[r][c]
// =================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) rowdot_kernel(buctd_rowdot_args a) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int groups = (a.c + 7) / 8;
  for (long long r = (long long)blockIdx.x * 8 + threadIdx.x / 32; r < a.rows; r += (long long)gridDim.x * 8) {
    float acc = 0.f;
    for (int g = lane; g < groups; g += 32) {
      float x[8], y[8];
      unpack8<T>(ld8<T>(a.a, r, a.a_cs, g * 8), x);
      unpack8<T>(ld8<T>(a.b, r, a.b_cs, g * 8), y);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (g * 8 + j < a.c) acc = fmaf(x[j], y[j], acc);
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffff, acc, o);
    if (lane == 0) a.out[r] = acc;
  }
}
int launch_rowdot(const buctd_rowdot_args& a, cudaStream_t st) {
  BUCTD_CHECK(a.a && a.b && a.out && a.a_cs % 8 == 0 && a.b_cs % 8 == 0, BUCTD_ERR_INVALID, "rowdot: bad arguments");
  const int blocks = (int)((a.rows + 7) / 8 < 148 * 16 ? (a.rows + 7) / 8 : 148 * 16);
  if (blocks == 0) return BUCTD_OK;
  if (a.dtype == BUCTD_F16) launch_pdl(rowdot_kernel<__half>, dim3(blocks), dim3(256), 0, st, a);
  else launch_pdl(rowdot_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, st, a);
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}

// =================================================================================================
// Stem conv1 weight gradient: dW[co][ci][ky][kx] = sum dz[b][oy][ox][co] x[b][ci][2 oy + ky - 1][2 ox + kx - 1]
// block = 256 threads = 64 output channels x 4 pixel lanes; blocks stride over the output pixels
// =================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) stem_wgrad_kernel(buctd_stem_wgrad_args a) {
  pdl_launch_dependents();
  pdl_wait();
  // one output row (b, oy) at a time: its three input rows of the three colour planes go through shared memory (coalesced
  // loads, broadcast reads), thread (co, lane) walks every fourth output column
  extern __shared__ float s_dyn[];
  float* s_x = s_dyn;                                   // [3 ci][3 ky][w + 2]  (column -1 and w are zero)
  float* s_p = s_dyn + 9 * (a.w + 2);                   // [4][64][27]
  const int oh = a.h / 2, ow = a.w / 2;
  const int nrows = a.batch * oh;
  const int co = threadIdx.x & 63, pl = threadIdx.x >> 6;
  const int wp = a.w + 2;
  float acc[27];
#pragma unroll
  for (int t = 0; t < 27; ++t) acc[t] = 0.f;
  for (int row = blockIdx.x; row < nrows; row += gridDim.x) {
    const int b = row / oh, oy = row - b * oh;
    __syncthreads();
    for (int i = threadIdx.x; i < 9 * wp; i += 256) {
      const int plane = i / wp, xx = i - plane * wp - 1;
      const int ci = plane / 3, ky = plane - ci * 3;
      const int iy = 2 * oy + ky - 1;
      float v = 0.f;
      if (iy >= 0 && iy < a.h && xx >= 0 && xx < a.w) {
        const long long o = (((long long)b * a.x_channels + ci) * a.h + iy) * a.w + xx;
        v = a.x_half ? __half2float(reinterpret_cast<const __half*>(a.x)[o]) : reinterpret_cast<const float*>(a.x)[o];
      }
      s_x[i] = v;
    }
    __syncthreads();
    const T* dz = reinterpret_cast<const T*>(a.dz) + (((long long)b * (oh + 1) + oy) * (ow + 1)) * a.dz_cs + co;
    for (int ox = pl; ox < ow; ox += 4) {
      const float g = T16<T>::to_f(dz[(long long)ox * a.dz_cs]);
#pragma unroll
      for (int p9 = 0; p9 < 9; ++p9) {                  // (ci, ky)
        const float* xr = s_x + p9 * wp + 2 * ox;        // columns 2 ox - 1 .. 2 ox + 1 (+1 for the zero column)
        acc[p9 * 3 + 0] = fmaf(g, xr[0], acc[p9 * 3 + 0]);
        acc[p9 * 3 + 1] = fmaf(g, xr[1], acc[p9 * 3 + 1]);
        acc[p9 * 3 + 2] = fmaf(g, xr[2], acc[p9 * 3 + 2]);
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int t = 0; t < 27; ++t) s_p[(pl * 64 + co) * 27 + t] = acc[t];
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * 27; i += 256)
    a.partial[(long long)blockIdx.x * 64 * 27 + i] = (s_p[i] + s_p[64 * 27 + i]) + (s_p[2 * 64 * 27 + i] + s_p[3 * 64 * 27 + i]);
}
int launch_stem_wgrad(const buctd_stem_wgrad_args& a, cudaStream_t st) {
  BUCTD_CHECK(a.x && a.dz && a.partial && a.h % 2 == 0 && a.w % 2 == 0 && a.dz_cs >= 64, BUCTD_ERR_INVALID, "stem_wgrad: bad arguments");
  const size_t smem = sizeof(float) * (9 * (size_t)(a.w + 2) + 4 * 64 * 27);
  BUCTD_CHECK(smem <= 48 * 1024, BUCTD_ERR_INVALID, "stem_wgrad: width %d too large", a.w);
  if (a.dtype == BUCTD_F16) launch_pdl(stem_wgrad_kernel<__half>, dim3(kRedBlocks), dim3(256), smem, st, a);
  else launch_pdl(stem_wgrad_kernel<__nv_bfloat16>, dim3(kRedBlocks), dim3(256), smem, st, a);
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}

// =================================================================================================
// Split partials -> torch-layout float32 gradient
// =================================================================================================
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(buctd_wgrad_reduce_args a) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = (long long)a.n * a.k * a.ntaps;
  const long long per = (long long)a.ntaps * a.n * a.k;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    // i walks the PARTIAL layout [t][n][k] (coalesced reads); the write goes to [n][k][t]
    const int k = (int)(i % a.k);
    const long long t1 = i / a.k;
    const int n = (int)(t1 % a.n), t = (int)(t1 / a.n);
    float s = 0.f;
    for (int sp = 0; sp < a.nsplit; ++sp) s += a.partial[(long long)sp * per + i];
    s *= a.inv_scale;
    float* g = a.grad + ((long long)n * a.k + k) * a.ntaps + t;
    *g = a.accumulate ? *g + s : s;
  }
}
int launch_wgrad_reduce(const buctd_wgrad_reduce_args& a, cudaStream_t st) {
  BUCTD_CHECK(a.partial && a.grad && a.n >= 1 && a.k >= 1 && a.ntaps >= 1 && a.nsplit >= 1, BUCTD_ERR_INVALID, "wgrad_reduce: bad arguments");
  const long long total = (long long)a.n * a.k * a.ntaps;
  const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  launch_pdl(wgrad_reduce_kernel, dim3(blocks), dim3(256), 0, st, a);
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}

// =================================================================================================
// Weight re-pack (float32 masters -> 16-bit tap-GEMM layouts), table driven.  One block = one 32 (cout) x 32 (cin) tile of one
// entry, all taps: the float32 rows are read in contiguous 32*taps-float segments and written as 64-byte runs along the
// destination's contiguous dimension in both layouts (shared-memory transpose).  Padding rows / columns of the packed
// buffers are never touched (they are zero from allocation).  units: int32 [n_units][4] = {entry, cout tile, cin tile, 0}.
// =================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) repack_kernel(buctd_repack_args a) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_w[32][32 * 9 + 1];
  const int* u = a.units + (long long)blockIdx.x * 4;
  const int64_t* e = a.table + (long long)u[0] * 8;
  const long long src_off = e[0], dst_off = e[1];
  const int cout = (int)e[2], cin = (int)e[3], taps = (int)e[4], rows_alloc = (int)e[5], cs = (int)e[6], tr = (int)e[7];
  const int co0 = u[1] * 32, ci0 = u[2] * 32;
  const int nco = min(32, cout - co0), nci = min(32, cin - ci0);
  const float* src = a.param + src_off;
  T* dst = reinterpret_cast<T*>(a.w16) + dst_off;
  const int seg = nci * taps;                                      // contiguous floats per cout row
  for (int i = threadIdx.x; i < nco * seg; i += 256) {
    const int r = i / seg, j = i - r * seg;
    s_w[r][j] = src[((long long)(co0 + r) * cin + ci0) * taps + j];
  }
  __syncthreads();
  if (!tr) {                      // dst[tap][co][ci]: ci contiguous
    for (int i = threadIdx.x; i < taps * nco * nci; i += 256) {
      const int ci = i % nci;
      const int t1 = i / nci;
      const int co = t1 % nco, tap = t1 / nco;
      dst[((long long)tap * rows_alloc + co0 + co) * cs + ci0 + ci] = T16<T>::from_f(s_w[co][ci * taps + tap]);
    }
  } else {                        // dst[taps-1-tap][ci][co]: co contiguous
    for (int i = threadIdx.x; i < taps * nco * nci; i += 256) {
      const int co = i % nco;
      const int t1 = i / nco;
      const int ci = t1 % nci, tap = t1 / nci;
      dst[((long long)(taps - 1 - tap) * rows_alloc + ci0 + ci) * cs + co0 + co] = T16<T>::from_f(s_w[co][ci * taps + tap]);
    }
  }
}
int launch_repack(const buctd_repack_args& a, cudaStream_t st) {
  BUCTD_CHECK(a.param && a.w16 && a.table && a.units && a.n_entries >= 0 && a.n_units >= 0, BUCTD_ERR_INVALID, "repack: bad arguments");
  if (a.n_units == 0) return BUCTD_OK;
  if (a.dtype == BUCTD_F16) launch_pdl(repack_kernel<__half>, dim3(a.n_units), dim3(256), 0, st, a);
  else launch_pdl(repack_kernel<__nv_bfloat16>, dim3(a.n_units), dim3(256), 0, st, a);
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}

}  // namespace buctd
