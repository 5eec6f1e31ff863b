// Shared epilogue of the tap-GEMM kernels: TMEM accumulator tile (128 rows x BN fp32 columns)
//   -> + bias[n] + residual[m,n] -> ReLU -> pad mask -> 16-bit -> global.
//
// The 16-bit tile is staged in shared memory as ceil(BN/64) column blocks of [128 rows][128 B] in the
// SWIZZLE_128B layout, which is (a) what a TMA tensor load produces -- the residual tile is fetched
// by TMA ahead of time, so its HBM latency is off the critical path -- (b) bank-conflict free for the
// one-thread-one-row access pattern that tcgen05.ld imposes, and (c) what a TMA tensor store consumes:
// the output leaves with cp.async.bulk.tensor stores, which also clip the M and N tails.
// Runs on the 4 epilogue warps (128 threads); warp w may only touch TMEM lanes [32*(w%4), +32).
#pragma once
#include "common.cuh"
#include "internal.h"

namespace buctd {

constexpr int kEpiBarrier = 1;     // named barrier id used by the epilogue warps
constexpr int kEpiBlockBytes = 128 * 128;

// NT = number of epilogue threads (128: one warp per TMEM lane quarter; 256: two warps per quarter,
// which split the columns of a tile between them)
template <int NT> __device__ __forceinline__ void epi_sync() {
  asm volatile("bar.sync %0, %1;" ::"n"(kEpiBarrier), "n"(NT) : "memory");
}

// group-private barrier (ids 1 and 2) of the 4 warps of one epilogue group
__device__ __forceinline__ void epi_sync_grp(int grp) {
  asm volatile("bar.sync %0, 128;" ::"r"(kEpiBarrier + grp) : "memory");
}
// n / d for n < 2^31 with a host-computed (mul, shift) pair (internal.h, fastdiv_make): no integer division on
// the epilogue's critical path
__device__ __forceinline__ unsigned fastdiv(unsigned n, unsigned mul, unsigned shift) {
  return (__umulhi(n, mul) + n) >> shift;
}

template <int BN> __host__ __device__ constexpr int epi_blocks() { return (BN + 63) / 64; }
template <int BN> __host__ __device__ constexpr int epi_stage_bytes() { return epi_blocks<BN>() * kEpiBlockBytes; }

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(m), "r"(smem_u32(src)),
               "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(m),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

struct EpiTile {          // where the 128 accumulator rows go
  long long m0;           // FLAT: first row
  int tb, ty0, tx0;       // PATCH: image and patch origin
};

struct EpiRow {
  bool valid, pad;
  int b, y, x;
};

__device__ __forceinline__ EpiRow epi_row(const TapGemmKParams& p, const EpiTile& t, int row) {
  EpiRow o;
  o.pad = false; o.b = o.y = o.x = 0;
  const int Hp = p.H + 1, Wp = p.W + 1;
  if (p.mode == BUCTD_TG_FLAT) {
    const long long r = t.m0 + row;
    o.valid = r < p.m_rows;
    if (o.valid && p.out_mode != BUCTD_OUT_ROWS16) {
      // 32-bit multiply-high division (m_rows < 2^31 is checked on the host)
      const unsigned ru = (unsigned)r;
      const unsigned ry = fastdiv(ru, p.wp_mul, p.wp_sh);
      o.x = (int)(ru - ry * (unsigned)Wp);
      o.b = (int)fastdiv(ry, p.hp_mul, p.hp_sh);
      o.y = (int)(ry - (unsigned)o.b * (unsigned)Hp);
      o.pad = (o.x == p.W) || (o.y == p.H);
    }
  } else {
    const int ty = row / p.tw, tx = row - ty * p.tw;
    o.y = t.ty0 + ty; o.x = t.tx0 + tx; o.b = t.tb;
    o.valid = (row < p.tw * p.th) && o.y < p.H && o.x < p.W;
  }
  return o;
}

// issue the TMA load of the residual tile into `stage` (one elected thread); completes on `bar`
template <int BN>
__device__ __forceinline__ void epi_load_residual(const TapGemmKParams& p, const CUtensorMap* map_res, const EpiTile& t, int n0,
                                                  uint8_t* stage, uint64_t* bar) {
  constexpr int NB = epi_blocks<BN>();
  int nblk = 0;
  for (int blk = 0; blk < NB; ++blk)
    if (n0 + blk * 64 < p.res_cs) ++nblk;
  const uint32_t rows_bytes = (p.mode == BUCTD_TG_FLAT ? 128u : (uint32_t)(p.tw * p.th)) * 128u;
  mbar_expect_tx(bar, nblk * rows_bytes);
  for (int blk = 0; blk < nblk; ++blk) {
    if (p.mode == BUCTD_TG_FLAT) tma_load_2d(stage + blk * kEpiBlockBytes, map_res, bar, n0 + blk * 64, (int)t.m0);
    else tma_load_4d(stage + blk * kEpiBlockBytes, map_res, bar, n0 + blk * 64, t.tx0, t.ty0, t.tb);
  }
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}

template <typename T> __device__ __forceinline__ uint32_t relu2(uint32_t u);
template <> __device__ __forceinline__ uint32_t relu2<__half>(uint32_t u) {
  __half2 h = __hmax2(*reinterpret_cast<__half2*>(&u), __float2half2_rn(0.f));
  return *reinterpret_cast<uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t relu2<__nv_bfloat16>(uint32_t u) {
  __nv_bfloat162 h = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&u), __float2bfloat162_rn(0.f));
  return *reinterpret_cast<uint32_t*>(&h);
}

// ---- fast path: a full tile (all BN columns real), one group per tile -------------------------------
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// round two fp32 values to the 16-bit type, clamp to [lo, +max finite] on the packed pair (lo = 0 for ReLU)
template <typename T> struct Finish;
template <> struct Finish<__half> {
  __device__ static __forceinline__ uint32_t lo(bool relu) { return relu ? 0u : 0xFBFFFBFFu; }     // -65504 | -65504
  __device__ static __forceinline__ uint32_t run(float a, float b, uint32_t lo2) {
    __half2 h = __floats2half2_rn(a, b);
    const uint32_t hi2 = 0x7BFF7BFFu;                                                               // 65504 | 65504
    h = __hmin2(__hmax2(h, *reinterpret_cast<const __half2*>(&lo2)), *reinterpret_cast<const __half2*>(&hi2));
    return *reinterpret_cast<uint32_t*>(&h);
  }
};
template <> struct Finish<__nv_bfloat16> {
  __device__ static __forceinline__ uint32_t lo(bool relu) { return relu ? 0u : 0xFF7FFF7Fu; }     // lowest finite bf16
  __device__ static __forceinline__ uint32_t run(float a, float b, uint32_t lo2) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    h = __hmax2(h, *reinterpret_cast<const __nv_bfloat162*>(&lo2));
    return *reinterpret_cast<uint32_t*>(&h);
  }
};

// Straight-line code: 64 columns per TMEM round trip, shared-space 128-bit accesses (no generic loads), packed
// clamps, no divergent branches.  Same arithmetic as the general path below (fp32 bias + residual add, one rounding).
template <typename T, int BN, bool RES>
__device__ __forceinline__ void epi_compute_fast(const TapGemmKParams& p, const EpiTile& tile, uint32_t tmem_acc, uint8_t* stage,
                                                 const float* s_bias, int q, int lane) {
  const int row = q * 32 + lane;
  const EpiRow me = epi_row(p, tile, row);
  const bool zero = me.pad || !me.valid;
  const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16);
  const uint32_t srow = smem_u32(stage) + (uint32_t)row * 128u;
  const uint32_t sw = (uint32_t)row & 7u;
  const uint32_t bias_addr = smem_u32(s_bias);
  const uint32_t lo2 = Finish<T>::lo(p.relu != 0);
#pragma unroll
  for (int c0 = 0; c0 < BN; c0 += 64) {
    constexpr int kMaxW = BN < 64 ? BN : 64;
    const int cw = (BN - c0) < 64 ? (BN - c0) : 64;          // compile-time after unrolling
    uint32_t acc[kMaxW];
#pragma unroll
    for (int i = 0; i < kMaxW / 16; ++i)
      if (i * 16 < cw) tmem_ld16(taddr + c0 + i * 16, &acc[i * 16]);
    tmem_ld_wait();
    const uint32_t blk = srow + (uint32_t)(c0 >> 6) * kEpiBlockBytes;
#pragma unroll
    for (int u = 0; u < kMaxW / 8; ++u) {
      if (u * 8 < cw) {
        const uint32_t a = blk + ((((uint32_t)u) ^ sw) << 4);
        const uint4 b0 = lds128(bias_addr + (uint32_t)(c0 + u * 8) * 4u), b1 = lds128(bias_addr + (uint32_t)(c0 + u * 8 + 4) * 4u);
        float v[8];
        v[0] = __uint_as_float(acc[u * 8 + 0]) + __uint_as_float(b0.x); v[1] = __uint_as_float(acc[u * 8 + 1]) + __uint_as_float(b0.y);
        v[2] = __uint_as_float(acc[u * 8 + 2]) + __uint_as_float(b0.z); v[3] = __uint_as_float(acc[u * 8 + 3]) + __uint_as_float(b0.w);
        v[4] = __uint_as_float(acc[u * 8 + 4]) + __uint_as_float(b1.x); v[5] = __uint_as_float(acc[u * 8 + 5]) + __uint_as_float(b1.y);
        v[6] = __uint_as_float(acc[u * 8 + 6]) + __uint_as_float(b1.z); v[7] = __uint_as_float(acc[u * 8 + 7]) + __uint_as_float(b1.w);
        if (RES) {
          const uint4 r = lds128(a);
          float2 f;
          f = T16<T>::unpack2(r.x); v[0] += f.x; v[1] += f.y;
          f = T16<T>::unpack2(r.y); v[2] += f.x; v[3] += f.y;
          f = T16<T>::unpack2(r.z); v[4] += f.x; v[5] += f.y;
          f = T16<T>::unpack2(r.w); v[6] += f.x; v[7] += f.y;
        }
        uint4 o;
        o.x = Finish<T>::run(v[0], v[1], lo2); o.y = Finish<T>::run(v[2], v[3], lo2);
        o.z = Finish<T>::run(v[4], v[5], lo2); o.w = Finish<T>::run(v[6], v[7], lo2);
        if (zero) o = make_uint4(0, 0, 0, 0);
        sts128(a, o);
      }
    }
  }
}

// ---- direct path: no shared-memory staging at all ---------------------------------------------------
// The SS-mode tcgen05.mma of these layers is bound by shared-memory operand bandwidth, so every byte
// the epilogue moves through shared memory (stage write, TMA-store read, residual TMA write + read) is taken from the
// tensor core.  Here a thread owns one output row and writes it with 16-byte st.global (every 32-byte sector of the
// row is completely written by two of them); only the residual tile still comes through a TMA-filled stage.
// one full 32-byte sector per thread and instruction (sm_100 256-bit store)
__device__ __forceinline__ void stg256(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x),
               "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}
__device__ __forceinline__ void stg128(void* p, const uint4& v) {
  asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <typename T, int BN, bool RES>
__device__ __forceinline__ void epi_compute_direct(const TapGemmKParams& p, const EpiTile& tile, int n0, uint32_t tmem_acc,
                                                   const uint8_t* stage, const float* s_bias, int q, int lane) {
  const int row = q * 32 + lane;
  const EpiRow me = epi_row(p, tile, row);
  const bool zero = me.pad;
  const long long r = tile.m0 + row;
  const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16);
  const uint32_t bias_addr = smem_u32(s_bias);
  const uint32_t lo2 = Finish<T>::lo(p.relu != 0);
  T* orow = reinterpret_cast<T*>(p.out) + r * p.out_cs + n0;
  // the residual tile was prefetched by TMA into the SWIZZLE_128B stage (its latency is off the critical path; a
  // per-thread ld.global of the row measured 20-80 % slower kernels)
  const uint32_t srow = smem_u32(stage) + (uint32_t)row * 128u;
  const uint32_t sw = (uint32_t)row & 7u;
  constexpr int CW = 64;                                    // columns per TMEM round trip
#pragma unroll
  for (int c0 = 0; c0 < BN; c0 += CW) {
    constexpr int kMaxW = BN < CW ? BN : CW;
    const int cw = (BN - c0) < CW ? (BN - c0) : CW;          // compile-time after unrolling
    uint32_t acc[kMaxW];
#pragma unroll
    for (int i = 0; i < kMaxW / 16; ++i)
      if (i * 16 < cw) tmem_ld16(taddr + c0 + i * 16, &acc[i * 16]);
    tmem_ld_wait();
    const uint32_t blk = srow + (uint32_t)(c0 >> 6) * kEpiBlockBytes;
    uint4 prev = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int u = 0; u < kMaxW / 8; ++u) {
      if (u * 8 < cw) {
        const uint4 b0 = lds128(bias_addr + (uint32_t)(c0 + u * 8) * 4u), b1 = lds128(bias_addr + (uint32_t)(c0 + u * 8 + 4) * 4u);
        float v[8];
        v[0] = __uint_as_float(acc[u * 8 + 0]) + __uint_as_float(b0.x); v[1] = __uint_as_float(acc[u * 8 + 1]) + __uint_as_float(b0.y);
        v[2] = __uint_as_float(acc[u * 8 + 2]) + __uint_as_float(b0.z); v[3] = __uint_as_float(acc[u * 8 + 3]) + __uint_as_float(b0.w);
        v[4] = __uint_as_float(acc[u * 8 + 4]) + __uint_as_float(b1.x); v[5] = __uint_as_float(acc[u * 8 + 5]) + __uint_as_float(b1.y);
        v[6] = __uint_as_float(acc[u * 8 + 6]) + __uint_as_float(b1.z); v[7] = __uint_as_float(acc[u * 8 + 7]) + __uint_as_float(b1.w);
        if (RES) {
          const uint4 rs = lds128(blk + ((((uint32_t)u) ^ sw) << 4));
          float2 f;
          f = T16<T>::unpack2(rs.x); v[0] += f.x; v[1] += f.y;
          f = T16<T>::unpack2(rs.y); v[2] += f.x; v[3] += f.y;
          f = T16<T>::unpack2(rs.z); v[4] += f.x; v[5] += f.y;
          f = T16<T>::unpack2(rs.w); v[6] += f.x; v[7] += f.y;
        }
        uint4 o;
        o.x = Finish<T>::run(v[0], v[1], lo2); o.y = Finish<T>::run(v[2], v[3], lo2);
        o.z = Finish<T>::run(v[4], v[5], lo2); o.w = Finish<T>::run(v[6], v[7], lo2);
        if (zero) o = make_uint4(0, 0, 0, 0);
        // BN % 16 == 0: units pair up into 32-byte sectors (rows and N tiles are 32-byte aligned, checked on the host)
        if (u & 1) { if (me.valid) stg256(orow + c0 + (u - 1) * 8, prev, o); }
        else prev = o;
      }
    }
  }
}

// TMEM -> (+bias, +residual from stage, ReLU, mask) -> 16-bit -> stage (in place).  Work unit = 8
// columns (one 16-byte chunk); warp group `grp` of `ngrp` takes every ngrp-th unit of its rows.  The
// next unit's tcgen05.ld is issued before the current one is processed.
template <typename T, int BN>
__device__ __forceinline__ void epi_compute(const TapGemmKParams& p, const EpiTile& tile, int n0, uint32_t tmem_acc,
                                            uint8_t* stage, const float* s_bias, bool has_res, int q, int lane, int grp,
                                            int ngrp) {
  // fast path: every column of the tile has weight rows behind it (rows >= n are zero by the packing contract, their
  // bias is 0, so those columns come out as exact zeros and the TMA store clips whatever lies beyond out_cs)
  if (ngrp == 1 && n0 + BN <= p.n_alloc) {
    if (has_res) epi_compute_fast<T, BN, true>(p, tile, tmem_acc, stage, s_bias, q, lane);
    else epi_compute_fast<T, BN, false>(p, tile, tmem_acc, stage, s_bias, q, lane);
    return;
  }
  const int row = q * 32 + lane;
  const EpiRow me = epi_row(p, tile, row);
  const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16);
  const bool zero = me.pad || !me.valid;
  const int limit = p.n > p.out_cs ? p.n : p.out_cs;
  int units = (min(limit - n0, BN) + 7) >> 3;              // 8-column units this tile really has
  if (units <= grp) return;
  uint8_t* srow = stage + row * 128;
  const int sw = row & 7;
  auto process = [&](int u, const uint32_t (&ac)[8]) {
    const int c0 = u * 8;
    uint4* a = reinterpret_cast<uint4*>(srow + (c0 >> 6) * kEpiBlockBytes + ((((c0 & 63) >> 3) ^ sw) << 4));
    float v[8];
    const float4 b0 = *reinterpret_cast<const float4*>(s_bias + c0), b1 = *reinterpret_cast<const float4*>(s_bias + c0 + 4);
    v[0] = __uint_as_float(ac[0]) + b0.x; v[1] = __uint_as_float(ac[1]) + b0.y;
    v[2] = __uint_as_float(ac[2]) + b0.z; v[3] = __uint_as_float(ac[3]) + b0.w;
    v[4] = __uint_as_float(ac[4]) + b1.x; v[5] = __uint_as_float(ac[5]) + b1.y;
    v[6] = __uint_as_float(ac[6]) + b1.z; v[7] = __uint_as_float(ac[7]) + b1.w;
    if (has_res) {
      const uint4 r = *a;
      const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = T16<T>::unpack2(w[j]);
        v[2 * j] += f.x;
        v[2 * j + 1] += f.y;
      }
    }
    if (n0 + c0 + 8 > p.n) {                                // N tail (and the zero pad channels of act16)
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (n0 + c0 + j >= p.n) v[j] = 0.f;
    }
    uint4 o;
    o.x = T16<T>::pack2(v[0], v[1]); o.y = T16<T>::pack2(v[2], v[3]);
    o.z = T16<T>::pack2(v[4], v[5]); o.w = T16<T>::pack2(v[6], v[7]);
    if (p.relu) { o.x = relu2<T>(o.x); o.y = relu2<T>(o.y); o.z = relu2<T>(o.z); o.w = relu2<T>(o.w); }
    if (zero) o = make_uint4(0, 0, 0, 0);
    *a = o;
  };
  // up to 6 units in flight per tcgen05.wait::ld (the TMEM read latency, not its bandwidth, is what costs)
  constexpr int kFly = 6;
  uint32_t acc[kFly][8];
#pragma unroll 1
  for (int u0 = grp; u0 < units; u0 += kFly * ngrp) {
#pragma unroll
    for (int i = 0; i < kFly; ++i)
      if (u0 + i * ngrp < units) tmem_ld8(taddr + (u0 + i * ngrp) * 8, acc[i]);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < kFly; ++i)
      if (u0 + i * ngrp < units) process(u0 + i * ngrp, acc[i]);
  }
}

// one elected thread: TMA-store the staged tile (one bulk group)
template <int BN>
__device__ __forceinline__ void epi_store(const TapGemmKParams& p, const CUtensorMap* map_out, const EpiTile& t, int n0,
                                          const uint8_t* stage) {
  constexpr int NB = epi_blocks<BN>();
  for (int blk = 0; blk < NB; ++blk) {
    if (n0 + blk * 64 >= p.out_cs) break;
    if (p.mode == BUCTD_TG_FLAT) tma_store_2d(map_out, stage + blk * kEpiBlockBytes, n0 + blk * 64, (int)t.m0);
    else tma_store_4d(map_out, stage + blk * kEpiBlockBytes, n0 + blk * 64, t.tx0, t.ty0, t.tb);
  }
  tma_store_commit();
}

// small-N head: fp32 planes [B][N][H][W], written straight from registers (coalesced across the
// warp's consecutive pixels).  No staging.
template <int BN>
__device__ __forceinline__ void epi_head_nchw32(const TapGemmKParams& p, const EpiTile& tile, int n0, uint32_t tmem_acc, int q,
                                                int lane) {
  const int row = q * 32 + lane;
  const EpiRow me = epi_row(p, tile, row);
  const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
  for (int c0 = 0; c0 < BN; c0 += 16) {
    const int nb = n0 + c0;
    if (nb >= p.n) break;
    uint32_t acc[16];
    tmem_ld16(taddr + c0, acc);
    tmem_ld_wait();
    if (!me.valid || me.pad) continue;
    float* o = reinterpret_cast<float*>(p.out);
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (nb + j < p.n) {
        float v = __uint_as_float(acc[j]) + (p.bias ? __ldg(p.bias + nb + j) : 0.f);
        if (p.relu) v = fmaxf(v, 0.f);
        o[(((long long)me.b * p.n + nb + j) * p.H + me.y) * p.W + me.x] = v;
      }
  }
}

// bias tile -> shared memory (NT epilogue threads, no sync inside)
template <int BN, int NT>
__device__ __forceinline__ void epi_load_bias(const TapGemmKParams& p, int n0, float* s_bias, int etid) {
  for (int i = etid; i < BN; i += NT) s_bias[i] = (p.bias != nullptr && n0 + i < p.n) ? __ldg(p.bias + n0 + i) : 0.f;
}

}  // namespace buctd
