// Position-attention backward, score stage: for every (query tile, key tile) of every crop
//     S  = Q K^T            (log2-domain scale folded in)         both on tcgen05, two accumulators in TMEM
//     dP = dO V^T
// and, in the epilogue (thread = query row), the soft-max backward with the dropout mask
//     p = exp2(S - lse2[q]),  pd = mask p / (1 - drop),  ds = p (mask dP / (1 - drop) - delta[q]) / sqrt(C)
// written as 16-bit [rows][ld] matrices -- the operands of dV = pd^T dO, dK = ds^T Q, dQ = ds K (buctd_gemm16).
// The float32 [rows x rows] score / dP matrices (2 x 6.4 GB at batch 32, 96x72) never exist: this kernel replaces two
// buctd_gemm16 launches + buctd_pos_softmax_bwd (kept as the unfused reference path, tests/test_gpu_train_ops.py).
// Autograd of torch.softmax / nn.Dropout / torch.matmul in ScaledDotProductAttention.forward, lib/models/self_attention.py:78-86.
#include "common.cuh"
#include "internal.h"

namespace buctd {

constexpr int kPbThreads = 64 + 256;     // TMA warp, MMA warp, 8 epilogue warps (two per TMEM lane quarter, splitting the columns)
constexpr int kPbBN = 256;

struct PosBwdParams {
  int rows, W, H, c, ld, nkc;
  float sc, inv_sqrt_c, drop_p;
  uint32_t seed;
  const uint32_t* seed_xor;
  const float* lse2;
  const float* delta;
  void* pd;
  void* ds;
};

__device__ __forceinline__ void pb_tma_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void pb_tmem_ld32(uint32_t taddr, uint32_t* r) {
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
__global__ void __launch_bounds__(kPbThreads, 1)
pos_att_bwd_scores_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                          const __grid_constant__ CUtensorMap map_do, const __grid_constant__ CUtensorMap map_v, const PosBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kA = 128 * 128, kB = kPbBN * 128;
  constexpr int kStage = 2 * (kA + kB);                 // Q | K | dO | V blocks of one 64-channel chunk
  uint8_t* s_valid = smem + 2 * kStage;                  // [256] key validity
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_valid + 256);
  uint64_t* empty_bar = full_bar + 2;
  uint64_t* acc_bar = empty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128, n0 = blockIdx.y * kPbBN, b = blockIdx.z;
  const unsigned wp = p.W + 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_do); tma_prefetch_desc(&map_v);
    for (int i = 0; i < 2; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(acc_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  if (threadIdx.x >= 64) {
    const unsigned k = (unsigned)(n0 + (int)threadIdx.x - 64);
    s_valid[threadIdx.x - 64] = (k < (unsigned)p.rows && (k % wp) != (unsigned)p.W && (k / wp) != (unsigned)p.H) ? 1 : 0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      uint32_t slot = 0, ph = 1;
      for (int kc = 0; kc < p.nkc; ++kc) {
        uint8_t* st = smem + (size_t)slot * kStage;
        mbar_wait(&empty_bar[slot], ph);
        mbar_expect_tx(&full_bar[slot], (uint32_t)kStage);
        pb_tma_3d(st, &map_q, &full_bar[slot], kc * 64, m0, b);
        pb_tma_3d(st + kA, &map_k, &full_bar[slot], kc * 64, n0, b);
        pb_tma_3d(st + kA + kB, &map_do, &full_bar[slot], kc * 64, m0, b);
        pb_tma_3d(st + 2 * kA + kB, &map_v, &full_bar[slot], kc * 64, n0, b);
        if (++slot == 2) { slot = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_f16(T16<T>::kFmt, 128, kPbBN);
      uint32_t slot = 0, ph = 0;
      for (int kc = 0; kc < p.nkc; ++kc) {
        mbar_wait(&full_bar[slot], ph);
        tc_fence_after();
        const uint32_t st = smem_u32(smem + (size_t)slot * kStage);
        const int ks_n = (min(64, p.c - kc * 64) + 15) >> 4;
        for (int ks = 0; ks < ks_n; ++ks) {
          tc_mma_f16(tmem_base, umma_desc_sw128(st + ks * 32), umma_desc_sw128(st + kA + ks * 32), idesc, (kc | ks) != 0);
          tc_mma_f16(tmem_base + kPbBN, umma_desc_sw128(st + kA + kB + ks * 32), umma_desc_sw128(st + 2 * kA + kB + ks * 32), idesc, (kc | ks) != 0);
        }
        tc_commit(&empty_bar[slot]);
        if (++slot == 2) { slot = 0; ph ^= 1; }
      }
      tc_commit(acc_bar);
    }
    __syncwarp();
  } else {
    const int q = warp & 3;                              // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;                    // which 128 columns of the tile
    const int row = q * 32 + lane;
    const int m = m0 + row;
    const bool qv = m < p.rows && ((unsigned)m % wp) != (unsigned)p.W && ((unsigned)m / wp) != (unsigned)p.H;
    const long long qr = (long long)b * p.rows + m;
    const float lse = qv ? p.lse2[qr] : 0.f;
    const float dl = qv ? p.delta[qr] : 0.f;
    const uint32_t thr = drop_threshold(p.drop_p);
    const float ks = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
    const uint32_t seed = p.seed ^ ((p.drop_p > 0.f && p.seed_xor) ? *p.seed_xor : 0u);
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    T* pd = reinterpret_cast<T*>(p.pd) + qr * p.ld;
    T* ds = reinterpret_cast<T*>(p.ds) + qr * p.ld;
    const int ncols = min(kPbBN, p.ld - n0);             // ld is a multiple of 8 >= rows: columns in [rows, ld) are written as zero
    for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 32) {
      if (c0 >= ncols) break;                            // warp-uniform
      uint32_t sv[32], dv[32];
      pb_tmem_ld32(taddr + c0, sv);
      pb_tmem_ld32(taddr + kPbBN + c0, dv);
      tmem_ld_wait();
      if (m >= p.rows) continue;
      uint32_t opd[16], ods[16];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        float pv[2], dsv[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int cc = c0 + i + e;
          float pe = 0.f, de = 0.f;
          if (qv && s_valid[cc]) {
            const float pr = exp2f(__uint_as_float(sv[i + e]) * p.sc - lse);
            const bool keep = p.drop_p <= 0.f || drop_hash(seed, (uint32_t)b, (uint32_t)m, (uint32_t)(n0 + cc)) >= thr;
            pe = keep ? pr * ks : 0.f;
            de = pr * ((keep ? __uint_as_float(dv[i + e]) * ks : 0.f) - dl) * p.inv_sqrt_c;
          }
          pv[e] = pe; dsv[e] = de;
        }
        opd[i >> 1] = T16<T>::pack2(pv[0], pv[1]);
        ods[i >> 1] = T16<T>::pack2(dsv[0], dsv[1]);
      }
      const int nc = min(32, ncols - c0);                // multiple of 8
#pragma unroll
      for (int g = 0; g < 4; ++g)
        if (g * 8 < nc) {
          *reinterpret_cast<uint4*>(pd + n0 + c0 + g * 8) = make_uint4(opd[4 * g], opd[4 * g + 1], opd[4 * g + 2], opd[4 * g + 3]);
          *reinterpret_cast<uint4*>(ds + n0 + c0 + g * 8) = make_uint4(ods[4 * g], ods[4 * g + 1], ods[4 * g + 2], ods[4 * g + 3]);
        }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

static size_t pb_smem() { return 1024 + 2 * (size_t)(2 * (128 * 128 + kPbBN * 128)) + 256 + 128; }

int pos_att_bwd_init_device() {
  BUCTD_CUDA(cudaFuncSetAttribute(pos_att_bwd_scores_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pb_smem()));
  BUCTD_CUDA(cudaFuncSetAttribute(pos_att_bwd_scores_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pb_smem()));
  return BUCTD_OK;
}

int launch_pos_bwd_scores(const buctd_pos_bwd_scores_args& a, cudaStream_t st) {
  BUCTD_CHECK(a.q && a.k && a.v && a.d_o && a.lse2 && a.delta && a.pd && a.ds, BUCTD_ERR_INVALID, "pos_bwd_scores: null pointer");
  const int rows = (a.h + 1) * (a.w + 1);
  BUCTD_CHECK(a.c % 16 == 0 && a.c >= 16 && a.c <= 384, BUCTD_ERR_INVALID, "pos_bwd_scores: C=%d must be a multiple of 16 in [16, 384]", a.c);
  BUCTD_CHECK(a.ld % 8 == 0 && a.ld >= rows, BUCTD_ERR_INVALID, "pos_bwd_scores: bad row stride %d", a.ld);
  BUCTD_CHECK(a.q_cs % 8 == 0 && a.k_cs % 8 == 0 && a.v_cs % 8 == 0 && a.do_cs % 8 == 0, BUCTD_ERR_INVALID, "pos_bwd_scores: channel strides");
  if (a.batch == 0) return BUCTD_OK;
  EncodeTiledFn enc = get_encode_tiled();
  BUCTD_CHECK(enc != nullptr, BUCTD_ERR_NO_DEVICE, "cuTensorMapEncodeTiled entry point not available");
  const CUtensorMapDataType dt = a.dtype == BUCTD_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap maps[4];
  const void* ptrs[4] = {a.q, a.k, a.d_o, a.v};
  const int cs[4] = {a.q_cs, a.k_cs, a.do_cs, a.v_cs};
  const int box_rows[4] = {128, kPbBN, 128, kPbBN};
  for (int i = 0; i < 4; ++i) {
    cuuint64_t dims[3] = {(cuuint64_t)a.c, (cuuint64_t)rows, (cuuint64_t)a.batch};
    cuuint64_t strides[2] = {(cuuint64_t)cs[i] * 2, (cuuint64_t)rows * cs[i] * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows[i], 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult cr = enc(&maps[i], dt, 3, const_cast<void*>(ptrs[i]), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    BUCTD_CHECK(cr == CUDA_SUCCESS, BUCTD_ERR_CUDA, "pos_bwd_scores: cuTensorMapEncodeTiled(%d) failed: %d", i, (int)cr);
  }
  PosBwdParams p;
  p.rows = rows; p.W = a.w; p.H = a.h; p.c = a.c; p.ld = a.ld; p.nkc = (a.c + 63) / 64;
  p.sc = 1.4426950408889634f / sqrtf((float)a.c);
  p.inv_sqrt_c = 1.f / sqrtf((float)a.c);
  p.drop_p = a.drop_p; p.seed = a.seed; p.seed_xor = a.seed_xor;
  p.lse2 = a.lse2; p.delta = a.delta; p.pd = a.pd; p.ds = a.ds;
  dim3 grid((rows + 127) / 128, (a.ld + kPbBN - 1) / kPbBN, a.batch);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(kPbThreads); cfg.dynamicSmemBytes = pb_smem(); cfg.stream = st;
  cudaError_t e;
  if (a.dtype == BUCTD_F16) e = cudaLaunchKernelEx(&cfg, pos_att_bwd_scores_kernel<__half>, maps[0], maps[1], maps[2], maps[3], p);
  else e = cudaLaunchKernelEx(&cfg, pos_att_bwd_scores_kernel<__nv_bfloat16>, maps[0], maps[1], maps[2], maps[3], p);
  BUCTD_CHECK(e == cudaSuccess, BUCTD_ERR_CUDA, "pos_bwd_scores launch failed: %s", cudaGetErrorString(e));
  return BUCTD_OK;
}

}  // namespace buctd
