// Shared epilogue of the tap-GEMM kernels: TMEM accumulator tile (128 rows x BN fp32 columns)
//   -> + bias[n] + residual[m,n] -> ReLU -> pad mask -> 16-bit -> global,
// staged through shared memory so that both the residual read and the output write are fully
// coalesced 16-byte accesses (a thread owns one accumulator ROW; writing rows straight from
// registers would scatter 16-byte pieces over 32 different lines per instruction).
// Runs on the 4 epilogue warps (128 threads); warp w may only touch TMEM lanes [32*(w%4), +32).
#pragma once
#include "common.cuh"
#include "internal.h"

namespace buctd {

constexpr int kEpiThreads = 128;
constexpr int kEpiBarrier = 1;     // named barrier id used by the epilogue warps

__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync %0, %1;" ::"n"(kEpiBarrier), "n"(kEpiThreads) : "memory"); }

template <int BN> __host__ __device__ constexpr int epi_pitch() { return BN * 2 + 16; }   // odd number of 16-B chunks
template <int BN> __host__ __device__ constexpr int epi_stage_bytes() { return 128 * epi_pitch<BN>(); }

struct EpiTile {          // where the 128 accumulator rows go
  long long m0;           // FLAT: first row
  int tb, ty0, tx0;       // PATCH: image and patch origin
};

// row of the tile -> global row index, validity, pad flag and (b,y,x) for NCHW output
struct EpiRow {
  long long r;
  bool valid, pad;
  int b, y, x;
};

__device__ __forceinline__ EpiRow epi_row(const TapGemmKParams& p, const EpiTile& t, int row, bool want_yx) {
  EpiRow o;
  o.pad = false; o.b = o.y = o.x = 0;
  const int Hp = p.H + 1, Wp = p.W + 1;
  if (p.mode == BUCTD_TG_FLAT) {
    o.r = t.m0 + row;
    o.valid = o.r < p.m_rows;
    if (want_yx && o.valid && p.out_mode != BUCTD_OUT_ROWS16) {
      o.x = (int)(o.r % Wp);
      const long long ry = o.r / Wp;
      o.y = (int)(ry % Hp);
      o.b = (int)(ry / Hp);
      o.pad = (o.x == p.W) || (o.y == p.H);
    }
  } else {
    const int ty = row / p.tw, tx = row - ty * p.tw;
    o.y = t.ty0 + ty; o.x = t.tx0 + tx; o.b = t.tb;
    o.valid = (row < p.tw * p.th) && o.y < p.H && o.x < p.W;
    o.r = ((long long)t.tb * Hp + o.y) * Wp + o.x;
  }
  return o;
}

// etid: 0..127 index among the epilogue threads; q = TMEM lane quarter of this warp; tmem_acc: TMEM
// address (lane 0, first column) of this tile's accumulator.
template <typename T, int BN>
__device__ void epilogue_tile(const TapGemmKParams& p, const EpiTile& tile, int n0, uint32_t tmem_acc, uint8_t* stage,
                              int etid, int q, int lane) {
  constexpr int PITCH = epi_pitch<BN>();
  constexpr int CH = BN / 8;                       // 16-byte chunks per row
  const T* res = reinterpret_cast<const T*>(p.residual);
  const int row = q * 32 + lane;
  const EpiRow me = epi_row(p, tile, row, true);

  if (p.out_mode == BUCTD_OUT_NCHW32) {            // small-N head: fp32 planes, coalesced across the warp's pixels
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
    return;
  }

  // ---- 1. residual tile -> stage (coalesced) ------------------------------------------------------
  if (res != nullptr) {
#pragma unroll 1
    for (int g = etid; g < 128 * CH; g += kEpiThreads) {
      const int rr = g / CH, ch = g - rr * CH;
      const EpiRow er = epi_row(p, tile, rr, false);
      uint4 v = make_uint4(0, 0, 0, 0);
      const int col = n0 + ch * 8;
      if (er.valid && col + 8 <= p.res_cs && col < p.n) v = *reinterpret_cast<const uint4*>(res + er.r * p.res_cs + col);
      *reinterpret_cast<uint4*>(stage + rr * PITCH + ch * 16) = v;
    }
    epi_sync();
  }
  // ---- 2. own row: TMEM -> bias/residual/ReLU/mask -> 16 bit -> stage -----------------------------
  {
    const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16);
    uint8_t* srow = stage + row * PITCH;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      const int nb = n0 + c0;
      if (nb >= p.n && nb >= p.out_cs) break;       // warp-uniform
      uint32_t acc[16];
      tmem_ld16(taddr + c0, acc);
      tmem_ld_wait();
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        v[j] = __uint_as_float(acc[j]);
        if (p.bias != nullptr && nb + j < p.n) v[j] += __ldg(p.bias + nb + j);
      }
      if (res != nullptr) {
        const uint4 u0 = *reinterpret_cast<const uint4*>(srow + c0 * 2);
        const uint4 u1 = *reinterpret_cast<const uint4*>(srow + c0 * 2 + 16);
        const uint32_t w[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 f = T16<T>::unpack2(w[j]);
          v[2 * j] += f.x;
          v[2 * j + 1] += f.y;
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (p.relu) v[j] = fmaxf(v[j], 0.f);
        if (me.pad || nb + j >= p.n) v[j] = 0.f;
      }
      uint4 u0, u1;
      u0.x = T16<T>::pack2(v[0], v[1]);   u0.y = T16<T>::pack2(v[2], v[3]);
      u0.z = T16<T>::pack2(v[4], v[5]);   u0.w = T16<T>::pack2(v[6], v[7]);
      u1.x = T16<T>::pack2(v[8], v[9]);   u1.y = T16<T>::pack2(v[10], v[11]);
      u1.z = T16<T>::pack2(v[12], v[13]); u1.w = T16<T>::pack2(v[14], v[15]);
      *reinterpret_cast<uint4*>(srow + c0 * 2) = u0;
      *reinterpret_cast<uint4*>(srow + c0 * 2 + 16) = u1;
    }
  }
  epi_sync();
  // ---- 3. stage -> global (coalesced 16-byte stores, rows are contiguous in global memory) ---------
  {
    T* out = reinterpret_cast<T*>(p.out);
#pragma unroll 1
    for (int g = etid; g < 128 * CH; g += kEpiThreads) {
      const int rr = g / CH, ch = g - rr * CH;
      const int col = n0 + ch * 8;
      if (col + 8 > p.out_cs) continue;
      const EpiRow er = epi_row(p, tile, rr, false);
      if (!er.valid) continue;
      *reinterpret_cast<uint4*>(out + er.r * p.out_cs + col) = *reinterpret_cast<const uint4*>(stage + rr * PITCH + ch * 16);
    }
  }
  epi_sync();      // stage may be reused by the next tile
}

}  // namespace buctd
