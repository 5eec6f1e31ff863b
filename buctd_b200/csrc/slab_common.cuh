// Pieces shared by the slab tap-GEMM kernel (tapgemm.cu) and its multi-layer chain variant (tapgemm_chain.cu).
#pragma once
#include "common.cuh"
#include "internal.h"

namespace buctd {

constexpr int kBlockM = 128;
constexpr int kSlabThreads = 352;      // producer + MMA + 8 epilogue warps (2 per TMEM lane quarter) + DMA warp
constexpr int kSmemBudget = 227 * 1024;

constexpr int kMaxAcc = 8;   // accumulator ring slots (barrier array size)
// TMEM column stride of one accumulator slot: BN rounded up to a power of two (BUCTD_ACC_ALIGN experiment: an N = 48
// accumulator that starts on a multiple of 64 columns)
#ifndef BUCTD_ACC_PACKED
__host__ __device__ constexpr int acc_stride_of(int bn) { return bn <= 32 ? 32 : bn <= 64 ? 64 : bn <= 128 ? 128 : 256; }
#else
__host__ __device__ constexpr int acc_stride_of(int bn) { return bn; }
#endif
template <int BN> __host__ __device__ constexpr int acc_stride() { return acc_stride_of(BN); }
template <int BN> __host__ __device__ constexpr int slab_nacc() { return (512 / acc_stride<BN>()) < kMaxAcc ? (512 / acc_stride<BN>()) : kMaxAcc; }

// Walks the 128-row sub-tiles of one CTA in processing order: work item k = blockIdx.x + k*gridDim.x is
// (super-tile s, N tile nt) in mixed radix; every item has MSUB sub-tiles.  No division after init().
struct SubIter {
  int s, nt, j, ds, dn, num_super;
  __device__ __forceinline__ void init(int num_super_) {
    num_super = num_super_;
    s = (int)blockIdx.x % num_super; nt = (int)blockIdx.x / num_super;
    ds = (int)gridDim.x % num_super; dn = (int)gridDim.x / num_super;
    j = 0;
  }
  // the same walk for a virtual CTA `vb` of `vg` over `num_units` units per N tile (CTA pairs: a unit = two super-tiles)
  __device__ __forceinline__ void init2(int num_units, int vb, int vg) {
    num_super = num_units;
    s = vb % num_units; nt = vb / num_units;
    ds = vg % num_units; dn = vg / num_units;
    j = 0;
  }
  __device__ __forceinline__ void next_item() {
    s += ds; nt += dn;
    if (s >= num_super) { s -= num_super; ++nt; }
  }
  template <int MSUB> __device__ __forceinline__ void next() {
    if (++j == MSUB) { j = 0; next_item(); }
  }
  template <int MSUB> __device__ __forceinline__ long long m0() const {
    return (long long)s * (MSUB * kBlockM) + (long long)j * kBlockM;
  }
};

// all k-steps of one (sub-tile, tap, chunk): fully unrolled, every descriptor is base + immediate
template <int BN, int KS, bool PAIR = false>
__device__ __forceinline__ void mma_ksteps(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t first) {
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    if constexpr (PAIR) tc_mma_f16_pair(d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, k == 0 ? first : 1u);
    else tc_mma_f16(d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, k == 0 ? first : 1u);
  }
}
template <int BN, bool PAIR = false>
__device__ __forceinline__ void mma_ksteps_rt(int ksteps, uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t first) {
  switch (ksteps) {
    case 1: mma_ksteps<BN, 1, PAIR>(d, adesc, bdesc, idesc, first); break;
    case 2: mma_ksteps<BN, 2, PAIR>(d, adesc, bdesc, idesc, first); break;
    case 3: mma_ksteps<BN, 3, PAIR>(d, adesc, bdesc, idesc, first); break;
    default: mma_ksteps<BN, 4, PAIR>(d, adesc, bdesc, idesc, first); break;
  }
}


}  // namespace buctd
