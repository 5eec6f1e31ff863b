// gemm16: D[M][N] = alpha * sum_k A[m][k] B[n][k] on tcgen05 with either operand stored K-major ([MN rows][K contiguous])
// or MN-major ([K rows][MN contiguous]) -- the contractions of the TRAINING step that the tap-GEMM (K-major activations x
// K-major packed weights) cannot express:
//   * weight gradients: dW_t[n][k] = sum_r dY[r][n] X[r + off_t][k], a contraction over the PIXEL rows of two act16
//     tensors (both MN-major; the tap is a row offset of the flat padded matrix, out-of-range rows are TMA zero fill),
//     split over `nsplit` CTAs along r with float32 partials (reduced in fixed order by wgrad_reduce: deterministic);
//   * the matrix products of the attention backward: S = Q K^T, dP = dO V^T (K-major x K-major), dV = P^T dO,
//     dK = dS^T Q (MN-major A), dQ = dS K (MN-major B), batched over the crops;
//   * nn.Linear weight gradients (dW = dY^T X).
// One 128 x BN output tile per CTA, K in stages of 64 through a TMA ring (SWIZZLE_128B, 64-element boxes), one
// accumulator in TMEM, four epilogue warps (thread = output row) writing float32 or 16-bit rows.  Tensor maps are 3-D
// {inner, rows of one batch, batches}: a K chunk that runs past the rows of ITS batch reads zeros, not the next crop.
// Replaces (with the tap-GEMM for the data gradients) the autograd backward of nn.Conv2d / nn.Linear / torch.matmul,
// lib/core/function.py:130-138.
#include "common.cuh"
#include "internal.h"

namespace buctd {

constexpr int kGmThreads = 192;
constexpr int kGmStageK = 64;

struct Gemm16Params {
  int M, N, K;
  int a_mn, b_mn;
  int ntaps, nsplit, nbatch;
  int k_per_split;
  int tap_off[9];
  long long d_t_stride, d_s_stride, d_b_stride;
  int ldd, out_f32, accumulate;
  int group;             // 3: one CTA computes the three dx taps of a kernel row from ONE A tile and a 72-row B stage
  float alpha;
  void* d;
  int bn, bn_pad, stages;
};

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// MN-major SWIZZLE_128B operand: [K rows][64 MN elements = 128 B] blocks of 64 K rows (8 KB); 8-row groups are 1024 B apart
// (SBO), consecutive 64-element MN blocks 8192 B apart (LBO)
__device__ __forceinline__ uint64_t gm_desc_mn(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint64_t gm_desc_mn_lbo(uint32_t smem_addr, uint32_t lbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void gm_tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
      "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

template <typename T>
__global__ void __launch_bounds__(kGmThreads)
gemm16_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const Gemm16Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int a_bytes = 128 * 128;
  const int b_blk = p.group == 3 ? 72 * 128 : 64 * 128;          // one 64-channel block of an MN-major B stage (+ 8 halo rows)
  const int b_bytes = p.group == 3 ? (p.bn_pad / 64) * b_blk : p.bn_pad * 128;
  const int stage_bytes = a_bytes + b_bytes;
  const int S = p.stages;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_bytes);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* acc_bar = empty_bar + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128, n0 = blockIdx.y * p.bn;
  int z = blockIdx.z;
  const int b = z % p.nbatch; z /= p.nbatch;
  const int s = z % p.nsplit;
  const int t = (z / p.nsplit) * (p.group == 3 ? 3 : 1);         // first tap of this CTA
  const int k_begin = s * p.k_per_split;
  const int k_end = min(p.K, k_begin + p.k_per_split);
  const int nstage = k_end > k_begin ? (k_end - k_begin + kGmStageK - 1) / kGmStageK : 0;
  const int nacc = p.group == 3 ? 3 : 1;
  const int acc_stride = (p.bn + 31) / 32 * 32;
  const int need_cols = nacc * acc_stride;
  const uint32_t tmem_cols = need_cols <= 32 ? 32 : need_cols <= 64 ? 64 : need_cols <= 128 ? 128 : need_cols <= 256 ? 256 : 512;
  // an MN-major A whose M fits one 64-element block: the second block of every stage is zeroed once and never loaded
  const bool a_half = p.a_mn && (p.M - m0) <= 64;
  if (a_half) {
    for (int i = threadIdx.x; i < S * 512; i += blockDim.x)
      *reinterpret_cast<uint4*>(smem + (size_t)(i / 512) * stage_bytes + 8192 + (i % 512) * 16) = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int i = 0; i < S; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(acc_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      const uint32_t tx = (uint32_t)((a_half ? a_bytes / 2 : a_bytes) + (p.b_mn ? b_bytes : p.bn * 128));
      uint32_t slot = 0, ph = 1;
      for (int i = 0; i < nstage; ++i) {
        const int k0 = k_begin + i * kGmStageK;
        uint8_t* sa = smem + (size_t)slot * stage_bytes;
        uint8_t* sb = sa + a_bytes;
        mbar_wait(&empty_bar[slot], ph);
        mbar_expect_tx(&full_bar[slot], tx);
        if (p.a_mn) {
          tma_load_3d(sa, &map_a, &full_bar[slot], m0, k0, b);
          if (!a_half) tma_load_3d(sa + 8192, &map_a, &full_bar[slot], m0 + 64, k0, b);
        } else {
          tma_load_3d(sa, &map_a, &full_bar[slot], k0, m0, b);
        }
        if (p.group == 3) {
          // rows k0 + off(dx = -1) .. + 72 of the input: the three dx taps are this stage shifted by 0 / 1 / 2 rows
          for (int j = 0; j * 64 < p.bn_pad; ++j) tma_load_3d(sb + j * b_blk, &map_b, &full_bar[slot], n0 + 64 * j, k0 + p.tap_off[t], b);
        } else if (p.b_mn) {
          for (int j = 0; j * 64 < p.bn_pad; ++j) tma_load_3d(sb + j * 8192, &map_b, &full_bar[slot], n0 + 64 * j, k0 + p.tap_off[t], b);
        } else {
          tma_load_3d(sb, &map_b, &full_bar[slot], k0, n0, b);
        }
        if (++slot == (uint32_t)S) { slot = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_f16(T16<T>::kFmt, 128, p.bn) | ((uint32_t)(p.a_mn != 0) << 15) | ((uint32_t)(p.b_mn != 0) << 16);
      uint32_t slot = 0, ph = 0;
      for (int i = 0; i < nstage; ++i) {
        mbar_wait(&full_bar[slot], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)slot * stage_bytes);
        const uint32_t sb = sa + a_bytes;
        if (p.group == 3) {
#pragma unroll
          for (int ks = 0; ks < kGmStageK / 16; ++ks) {
            const uint64_t ad = gm_desc_mn(sa + ks * 2048);
#pragma unroll
            for (int j = 0; j < 3; ++j)        // a row-shifted view of a SWIZZLE_128B tile is still a valid operand (address-based swizzle)
              tc_mma_f16(tmem_base + j * acc_stride, ad, gm_desc_mn_lbo(sb + ks * 2048 + j * 128, 72 * 128), idesc, (i | ks) != 0);
          }
        } else {
#pragma unroll
          for (int ks = 0; ks < kGmStageK / 16; ++ks) {
            const uint64_t ad = p.a_mn ? gm_desc_mn(sa + ks * 2048) : umma_desc_sw128(sa + ks * 32);
            const uint64_t bd = p.b_mn ? gm_desc_mn(sb + ks * 2048) : umma_desc_sw128(sb + ks * 32);
            tc_mma_f16(tmem_base, ad, bd, idesc, (i | ks) != 0);
          }
        }
        tc_commit(&empty_bar[slot]);
        if (++slot == (uint32_t)S) { slot = 0; ph ^= 1; }
      }
      tc_commit(acc_bar);
    }
    __syncwarp();
  } else {
    // epilogue: thread = output row (TMEM lane); warp w of the CTA may only touch lanes 32 (w % 4) ..
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int m = m0 + row;
    if (nstage > 0) {
      mbar_wait(acc_bar, 0);
      tc_fence_after();
    }
    const int ncols = min(p.bn, p.N - n0);
    for (int ja = 0; ja < nacc; ++ja) {
    const long long off = (long long)(t + ja) * p.d_t_stride + (long long)s * p.d_s_stride + (long long)b * p.d_b_stride + (long long)m * p.ldd + n0;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ja * acc_stride);
    for (int c0 = 0; c0 < ncols; c0 += 32) {
      uint32_t v[32];
      if (nstage > 0) {
        gm_tmem_ld32(taddr + c0, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0u;               // an empty K range contributes zeros
      }
      if (m >= p.M) continue;
      const int nc = min(32, ncols - c0);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * p.alpha);
      if (p.accumulate) {                          // D += ...: the same thread owns the element, no hazard
        if (p.out_f32) {
          const float* src = reinterpret_cast<const float*>(p.d) + off + c0;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < nc) v[i] = __float_as_uint(__uint_as_float(v[i]) + src[i]);
        } else {
          const T* src = reinterpret_cast<const T*>(p.d) + off + c0;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < nc) v[i] = __float_as_uint(__uint_as_float(v[i]) + T16<T>::to_f(src[i]));
        }
      }
      if (p.out_f32) {
        float* dst = reinterpret_cast<float*>(p.d) + off + c0;
        if (nc == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<float4*>(dst + i) = make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < nc) dst[i] = __uint_as_float(v[i]);
        }
      } else {
        T* dst = reinterpret_cast<T*>(p.d) + off + c0;
        if (nc == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 u;
            u.x = T16<T>::pack2(__uint_as_float(v[i]), __uint_as_float(v[i + 1]));
            u.y = T16<T>::pack2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
            u.z = T16<T>::pack2(__uint_as_float(v[i + 4]), __uint_as_float(v[i + 5]));
            u.w = T16<T>::pack2(__uint_as_float(v[i + 6]), __uint_as_float(v[i + 7]));
            *reinterpret_cast<uint4*>(dst + i) = u;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < nc) dst[i] = T16<T>::from_f(__uint_as_float(v[i]));
        }
      }
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

static int gm_encode(CUtensorMap* m, int dtype, const void* base, int inner, int rows, int nbatch, int ld, long long batch_stride,
                     int box_inner, int box_rows, const char* what) {
  EncodeTiledFn enc = get_encode_tiled();
  BUCTD_CHECK(enc != nullptr, BUCTD_ERR_NO_DEVICE, "cuTensorMapEncodeTiled entry point not available");
  const CUtensorMapDataType dt = dtype == BUCTD_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)nbatch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(nbatch > 1 ? batch_stride : (long long)rows * ld) * 2};
  cuuint32_t box[3] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult cr = enc(m, dt, 3, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BUCTD_CHECK(cr == CUDA_SUCCESS, BUCTD_ERR_CUDA, "gemm16: cuTensorMapEncodeTiled(%s) failed: %d (inner %d rows %d batches %d ld %d)", what,
              (int)cr, inner, rows, nbatch, ld);
  return BUCTD_OK;
}

int gemm16_init_device() {
  BUCTD_CUDA(cudaFuncSetAttribute(gemm16_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  BUCTD_CUDA(cudaFuncSetAttribute(gemm16_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  return BUCTD_OK;
}

int launch_gemm16(const buctd_gemm16_args& a, cudaStream_t st) {
  BUCTD_CHECK(a.a && a.b && a.d, BUCTD_ERR_INVALID, "gemm16: null pointer");
  BUCTD_CHECK(a.m >= 1 && a.n >= 1 && a.k >= 0, BUCTD_ERR_INVALID, "gemm16: bad sizes %d %d %d", a.m, a.n, a.k);
  BUCTD_CHECK(a.nbatch >= 1 && a.nsplit >= 1 && a.ntaps >= 1 && a.ntaps <= 9, BUCTD_ERR_INVALID, "gemm16: bad batch / split / taps");
  BUCTD_CHECK(a.lda % 8 == 0 && a.ldb % 8 == 0 && (((uintptr_t)a.a | (uintptr_t)a.b) & 15) == 0, BUCTD_ERR_INVALID,
              "gemm16: operands need 16-byte aligned rows (lda %d ldb %d)", a.lda, a.ldb);
  BUCTD_CHECK(a.ntaps == 1 || a.b_mn_major, BUCTD_ERR_INVALID, "gemm16: taps shift the K rows of an MN-major B");
  const bool group3 = a.tap_group == 3;
  if (group3) {
    BUCTD_CHECK(a.ntaps % 3 == 0 && a.a_mn_major && a.b_mn_major && a.nbatch == 1, BUCTD_ERR_INVALID, "gemm16: tap_group 3 needs 3k taps, MN-major operands, one batch");
    for (int g = 0; g < a.ntaps; g += 3)
      BUCTD_CHECK(a.tap_off[g + 1] == a.tap_off[g] + 1 && a.tap_off[g + 2] == a.tap_off[g] + 2, BUCTD_ERR_INVALID,
                  "gemm16: tap_group 3 needs consecutive row offsets within a group");
  }
  Gemm16Params p;
  memset(&p, 0, sizeof(p));
  p.M = a.m; p.N = a.n; p.K = a.k;
  p.a_mn = a.a_mn_major != 0; p.b_mn = a.b_mn_major != 0;
  p.ntaps = a.ntaps; p.nsplit = a.nsplit; p.nbatch = a.nbatch;
  const int per = (a.k + a.nsplit - 1) / a.nsplit;
  p.k_per_split = (per + kGmStageK - 1) / kGmStageK * kGmStageK;
  if (p.k_per_split == 0) p.k_per_split = kGmStageK;
  for (int i = 0; i < 9; ++i) p.tap_off[i] = a.tap_off[i];
  p.d_t_stride = a.d_t_stride; p.d_s_stride = a.d_s_stride; p.d_b_stride = a.d_b_stride;
  p.ldd = a.ldd; p.out_f32 = a.out_f32; p.accumulate = a.accumulate; p.alpha = a.alpha; p.d = a.d;
  // N tile: the whole N when it fits one instruction (multiple of 16, <= 256), else 256 / 128
  int bn = a.n <= 256 ? (a.n + 15) / 16 * 16 : 256;
  if (a.n > 256 && a.n % 256 != 0 && a.n % 128 == 0) bn = 128;
  if (group3 && bn > 128) bn = a.n % 128 == 0 ? 128 : (a.n % 96 == 0 ? 96 : 128);     // three accumulators in 512 TMEM columns
  p.group = group3 ? 3 : 1;
  p.bn = bn;
  p.bn_pad = (bn + 63) / 64 * 64;
  const int stage_bytes = 128 * 128 + (group3 ? (p.bn_pad / 64) * 72 * 128 : p.bn_pad * 128);
  int stages = (227 * 1024 - 2048) / stage_bytes;
  p.stages = stages > 8 ? 8 : stages;
  // a short K loop does not need the whole shared memory: fewer stages let several CTAs share an SM (the attention-sized
  // products are thousands of one- to three-stage CTAs)
  const int kstages = (p.k_per_split + kGmStageK - 1) / kGmStageK;
  if (p.stages > kstages + 1) p.stages = kstages + 1 < 2 ? 2 : kstages + 1;
  // K-major: inner = K, rows = MN; MN-major: inner = MN, rows = K (+ tap offsets: signed coordinates, zero fill outside)
  CUtensorMap ma, mb;
  int rc;
  if (p.a_mn) rc = gm_encode(&ma, a.dtype, a.a, a.m, a.k, a.nbatch, a.lda, a.a_batch_stride, 64, 64, "A");
  else rc = gm_encode(&ma, a.dtype, a.a, a.k, a.m, a.nbatch, a.lda, a.a_batch_stride, 64, 128, "A");
  if (rc) return rc;
  if (p.b_mn) rc = gm_encode(&mb, a.dtype, a.b, a.n, a.b_rows > 0 ? a.b_rows : a.k, a.nbatch, a.ldb, a.b_batch_stride, 64, group3 ? 72 : 64, "B");
  else rc = gm_encode(&mb, a.dtype, a.b, a.k, a.n, a.nbatch, a.ldb, a.b_batch_stride, 64, bn, "B");
  if (rc) return rc;
  dim3 grid((a.m + 127) / 128, (a.n + bn - 1) / bn, (group3 ? a.ntaps / 3 : a.ntaps) * a.nsplit * a.nbatch);
  const size_t smem = 1024 + (size_t)p.stages * stage_bytes + 256;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(kGmThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaError_t e;
  if (a.dtype == BUCTD_F16) e = cudaLaunchKernelEx(&cfg, gemm16_kernel<__half>, ma, mb, p);
  else e = cudaLaunchKernelEx(&cfg, gemm16_kernel<__nv_bfloat16>, ma, mb, p);
  BUCTD_CHECK(e == cudaSuccess, BUCTD_ERR_CUDA, "gemm16 launch failed: %s", cudaGetErrorString(e));
  return BUCTD_OK;
}

}  // namespace buctd
