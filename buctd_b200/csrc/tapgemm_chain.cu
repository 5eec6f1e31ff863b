// Chain tap-GEMM: up to BUCTD_MAX_CHAIN FLAT tap-GEMMs of IDENTICAL geometry (the eight 3x3 convolutions of the four
// BasicBlocks of one HRNet branch) in ONE persistent launch.
//
// Why: a slab tap-GEMM launch costs ~8 us around its ~29 us of tensor work at B = 64 (inter-kernel gap ~4 us even with
// programmatic dependent launch -- a CTA of the next kernel needs the whole SM -- plus prologue, pipeline fill and
// drain; tools/conv_sweep.py).  Here the rings (slabs, weights, TMEM accumulators, stage buffers) keep
// running across layers, and there is NO grid-wide barrier between them: the dependency of a 3x3 convolution is local,
// super-tile s of layer l+1 reads rows of the super-tiles s-1, s, s+1 of layer l (halo W+2 < 128 rows) and, as residual,
// super-tile s of an earlier layer.  Every layer has the same geometry, hence the same (super-tile -> CTA) assignment.
// flag[l][s] is bumped (+1 per N tile) once the TMA stores of that work item are complete -- the DMA thread hands the
// index to the publisher warp, which owns the gpu-scope release -- and the producer warp polls the flags of the previous
// layer (all lanes, ten items ahead) before lane 0 loads a slab.  CTAs drift apart freely: a slow CTA only delays its
// neighbours.  Writes are safe by the same argument: whoever overwrites rows of super-tile s at layer
// l+1 has waited for every layer-l reader of those rows (buffers rotate as in a BasicBlock: in -> tmp -> out).
// All CTAs are co-resident by construction (grid <= SM count, one CTA per SM) and dependencies only point to lower
// layers, so the spins cannot deadlock; they trap after ~4 s instead of hanging the device.
//
// Same roles and pipeline as tapgemm_slab_kernel (tapgemm.cu) plus the publisher warp (12 warps); staged (TMA store)
// epilogue only.  Layers with C <= 48 run the dx-fused plan (DXF, below): one N = 3C MMA per kernel row.
// Replaces the conv1/bn1/relu/conv2/bn2/+residual/relu sequences of HighResolutionModule._make_one_branch,
// lib/models/pose_hrnet_coam.py:31-60, 142-165, 254-255.
#include "common.cuh"
#include "epilogue.cuh"
#include "internal.h"
#include "slab_common.cuh"

namespace buctd {

struct alignas(64) ChainLayerMaps {
  CUtensorMap a, b, out, res;
};
struct ChainParams {
  ChainLayerMaps m[BUCTD_MAX_CHAIN];
  const float* bias[BUCTD_MAX_CHAIN];
  int has_res[BUCTD_MAX_CHAIN];
  int relu[BUCTD_MAX_CHAIN];
  unsigned int* flags;       // [n_layers][num_super]: N tiles of that super-tile whose output is in global memory
  int n_layers;
  int n_tiles;               // N tiles per super-tile
  int lag;                   // items whose stores stay in flight before a flag is published (0, 1, 2 or 4)
  int debug;                 // BUCTD_CHAIN_DEBUG timing experiments (results are WRONG): 1 no dataflow flags, 2 dx-fused
                             // epilogue without the +-1 row exchange, 4 no MMAs
};

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned int ld_relaxed(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Publishing a flag needs a gpu-scope release.  On the DMA thread (which has TMA stores in flight) the fences of a
// release cost ~1 us EACH time and stall the stage-buffer ring with it (measured: 49 -> 33 us per dx-fused 48->48
// convolution without them).  So the DMA thread, once cp.async.bulk.wait_group says the stores of an item are complete,
// only pushes the flag index into a small shared-memory FIFO; the publisher warp (one thread, nothing in flight) pops
// it, fences once per batch and bumps the flags.  Order: stores complete (wait_group) -> st.release.cta (FIFO head) ->
// ld.acquire.cta -> fence.acq_rel.gpu -> red.relaxed.gpu (flag) -> consumer's ld.relaxed + fence + fence.proxy.async.
constexpr int kChainThreads = kSlabThreads + 32;       // + the publisher warp
constexpr int kPubRing = 16;
struct PubFifo {
  unsigned int head;             // entries pushed so far; bit 31: the DMA thread is done
  unsigned int tail;             // entries published so far
  unsigned int ring[kPubRing];   // flag indices
};
__device__ __forceinline__ unsigned int lds_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void sts_release(unsigned int* p, unsigned int v) {
  asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void pub_push(PubFifo* f, unsigned int& pushed, unsigned int idx) {
  while (pushed - lds_acquire(&f->tail) >= (unsigned)kPubRing) __nanosleep(32);
  f->ring[pushed % kPubRing] = idx;
  sts_release(&f->head, ++pushed);
}
template <int N> __device__ __forceinline__ void tma_store_wait_complete() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ---- dx-fused variant (DXF) ---------------------------------------------------------------------
// The SS-mode tcgen05.mma of a C <= 64 layer is bound by the shared-memory fetch of its A tile (128 rows x 128 B per
// (tap, sub-tile) whatever N is).  The three dx taps of a kernel row read the SAME activation rows
// shifted by one: with B = [W(dy,0); W(dy,1); W(dy,2)] (N = 3 BN <= 256) one MMA per dy fetches A once instead of three
// times and leaves three partial sums per tile row; out[g] = D0[g-1] + D1[g] + D2[g+1].  Tiles overlap by one row on
// each side (126 output rows per 128-lane accumulator), the +-1 row shift is a warp shuffle plus a 2 x 16-float
// exchange between neighbouring warps per 16 columns.
constexpr int kDxfRows = 126;
constexpr int kDxfPass = 48;                     // most columns one epilogue pass handles (register budget)
constexpr int kDxfXchFloats = 2 /*groups*/ * 2 /*parity*/ * 4 /*warps*/ * 2 /*L,R*/ * kDxfPass;
template <int BN> __host__ __device__ constexpr int dxf_acc_stride() { return 3 * BN <= 128 ? 128 : 3 * BN <= 160 ? 160 : 256; }
template <int BN> __host__ __device__ constexpr int dxf_nacc() { return 512 / dxf_acc_stride<BN>(); }

template <int BN>
__device__ __forceinline__ void dxf_load_residual(const TapGemmKParams& p, const CUtensorMap* map_res, long long m0, uint8_t* stage,
                                                  uint64_t* bar) {
  constexpr int NB = epi_blocks<BN>();
  int nblk = 0;
  for (int blk = 0; blk < NB; ++blk)
    if (blk * 64 < p.res_cs) ++nblk;
  mbar_expect_tx(bar, (uint32_t)nblk * kDxfRows * 128u);
  for (int blk = 0; blk < nblk; ++blk) tma_load_2d(stage + blk * kEpiBlockBytes, map_res, bar, blk * 64, (int)m0);
}
template <int BN>
__device__ __forceinline__ void dxf_store(const TapGemmKParams& p, const CUtensorMap* map_out, long long m0, const uint8_t* stage) {
  constexpr int NB = epi_blocks<BN>();
  for (int blk = 0; blk < NB; ++blk) {
    if (blk * 64 >= p.out_cs) break;
    tma_store_2d(map_out, stage + blk * kEpiBlockBytes, blk * 64, (int)m0);
  }
  tma_store_commit();
}

// One accumulator tile: lane rho <-> activation row g = m0 - 1 + rho; rows rho = 1..126 are written to stage row rho-1
// (the residual tile lies there too).  Per pass of CP columns: the left / right partials are loaded from TMEM, the
// warp-boundary rows go to the exchange buffer, the centre partial's load is issued, the +-1 row shift happens in
// registers (warp shuffles), ONE group barrier publishes the boundary rows, and lanes 0 / 31 patch theirs in.  The
// epilogue of these layers is latency bound (every TMEM round trip and barrier is serial per tile), hence one pass for
// BN <= 48.  fp32 tap sum: centre + (left + right), then bias, residual, one rounding.
template <typename T, int BN, bool RES>
__device__ __forceinline__ void dxf_compute(const TapGemmKParams& p, long long m0, uint32_t tmem_acc, uint8_t* stage,
                                            const float* s_bias, float* xch, uint32_t& xpar, int q, int lane, int grp) {
  static_assert(BN % 16 == 0 && BN <= 64, "dx-fused tiles: N = 3 BN <= 256 and one 64-channel stage block");
  constexpr int CP = BN <= kDxfPass ? BN : 32;                 // columns per pass
  const int rho = q * 32 + lane;
  const long long g = m0 - 1 + rho;
  const bool writes = rho >= 1 && rho <= kDxfRows;
  bool zero = true;
  if (writes && g < p.m_rows) {
    const unsigned ru = (unsigned)g;
    const unsigned ry = fastdiv(ru, p.wp_mul, p.wp_sh);
    const unsigned x = ru - ry * (unsigned)(p.W + 1);
    const unsigned b = fastdiv(ry, p.hp_mul, p.hp_sh);
    const unsigned y = ry - b * (unsigned)(p.H + 1);
    zero = (x == (unsigned)p.W) || (y == (unsigned)p.H);
  }
  const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16);
  const uint32_t srow = smem_u32(stage) + (uint32_t)(rho - 1) * 128u;
  const uint32_t sw = (uint32_t)(rho - 1) & 7u;
  const uint32_t bias_addr = smem_u32(s_bias);
  const uint32_t lo2 = Finish<T>::lo(p.relu != 0);
#pragma unroll
  for (int c0 = 0; c0 < BN; c0 += CP) {
    float sum[CP];
    uint32_t c[CP];
    {
      uint32_t l[CP], r[CP];
#pragma unroll
      for (int i = 0; i < CP; i += 16) {
        tmem_ld16(taddr + c0 + i, &l[i]);
        tmem_ld16(taddr + 2 * BN + c0 + i, &r[i]);
      }
      tmem_ld_wait();
      // boundary rows for the neighbouring warps: lane 31's left partial belongs to row 0 of warp q+1, lane 0's right
      // partial to row 31 of warp q-1.  Double buffered by a running parity: a writer of pass n+2 has passed barrier n+1,
      // which every reader of pass n reaches only after its reads.
      const uint32_t xb = smem_u32(xch) + xpar * (4 * 2 * kDxfPass * 4);
      xpar ^= 1u;
      if (lane == 31) {
#pragma unroll
        for (int i = 0; i < CP; i += 4) sts128(xb + (uint32_t)((q * 2 + 0) * kDxfPass + i) * 4u, make_uint4(l[i], l[i + 1], l[i + 2], l[i + 3]));
      }
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < CP; i += 4) sts128(xb + (uint32_t)((q * 2 + 1) * kDxfPass + i) * 4u, make_uint4(r[i], r[i + 1], r[i + 2], r[i + 3]));
      }
#pragma unroll
      for (int i = 0; i < CP; ++i) {
        const float lf = __shfl_up_sync(0xffffffffu, __uint_as_float(l[i]), 1);
        const float rf = __shfl_down_sync(0xffffffffu, __uint_as_float(r[i]), 1);
        sum[i] = (lane == 0 ? 0.f : lf) + (lane == 31 ? 0.f : rf);
      }
      // the centre partial's TMEM round trip overlaps the barrier (issued here: l / r are dead, registers are tight)
#pragma unroll
      for (int i = 0; i < CP; i += 16) tmem_ld16(taddr + BN + c0 + i, &c[i]);
      epi_sync_grp(grp);
      if ((lane == 0 && q > 0) || (lane == 31 && q < 3)) {
        const uint32_t src = xb + (uint32_t)((lane == 0 ? (q - 1) * 2 + 0 : (q + 1) * 2 + 1) * kDxfPass) * 4u;
#pragma unroll
        for (int i = 0; i < CP; i += 4) {
          const uint4 t = lds128(src + (uint32_t)i * 4u);
          sum[i] += __uint_as_float(t.x); sum[i + 1] += __uint_as_float(t.y);
          sum[i + 2] += __uint_as_float(t.z); sum[i + 3] += __uint_as_float(t.w);
        }
      }
    }
    tmem_ld_wait();
    if (writes) {
#pragma unroll
      for (int u = 0; u < CP / 8; ++u) {
        const uint32_t a = srow + ((((uint32_t)(c0 >> 3) + u) ^ sw) << 4);
        const uint4 b0 = lds128(bias_addr + (uint32_t)(c0 + u * 8) * 4u), b1 = lds128(bias_addr + (uint32_t)(c0 + u * 8 + 4) * 4u);
        float w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = __uint_as_float(c[u * 8 + j]) + sum[u * 8 + j];
        w[0] += __uint_as_float(b0.x); w[1] += __uint_as_float(b0.y); w[2] += __uint_as_float(b0.z); w[3] += __uint_as_float(b0.w);
        w[4] += __uint_as_float(b1.x); w[5] += __uint_as_float(b1.y); w[6] += __uint_as_float(b1.z); w[7] += __uint_as_float(b1.w);
        if (RES) {
          const uint4 rs = lds128(a);
          float2 f;
          f = T16<T>::unpack2(rs.x); w[0] += f.x; w[1] += f.y;
          f = T16<T>::unpack2(rs.y); w[2] += f.x; w[3] += f.y;
          f = T16<T>::unpack2(rs.z); w[4] += f.x; w[5] += f.y;
          f = T16<T>::unpack2(rs.w); w[6] += f.x; w[7] += f.y;
        }
        uint4 o;
        o.x = Finish<T>::run(w[0], w[1], lo2); o.y = Finish<T>::run(w[2], w[3], lo2);
        o.z = Finish<T>::run(w[4], w[5], lo2); o.w = Finish<T>::run(w[6], w[7], lo2);
        if (zero) o = make_uint4(0, 0, 0, 0);
        sts128(a, o);
      }
    }
  }
}

// PAIR: the grid is launched in clusters of two CTAs that work as ONE tcgen05 CTA pair (cta_group::2, M = 256): CTA r of
// pair v processes super-tile 2 su + r of unit su (both with the same N tile), each CTA loads its own slab but only HALF
// of every weight tile (BN / 2 rows), and the leader (rank 0) issues the MMAs of both.  The layers this is chosen for
// (C >= 96) stream 166 KB .. 1.3 MB of weights per work item against <= 64 KB of activations: at C >= 192 the one-CTA
// kernel is bound by that stream (TMA fill of one SM ~ 36 B/clk: 10 us per 128-row item against 6.4 us of MMAs).
// Everything behind the accumulator (epilogue, stage buffers, DMA, dataflow flags, publisher) stays per CTA.
template <typename T, int BN, int MSUB, bool DXF, bool PAIR = false>
__global__ void __launch_bounds__(kChainThreads, 1)
tapgemm_chain_kernel(const __grid_constant__ ChainParams cp, const TapGemmKParams p0, const SlabParams sp) {
  static_assert(!(PAIR && DXF), "the dx-fused plan is a one-CTA plan");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kWTileBytes = (PAIR ? BN / 2 : BN) * 128;            // this CTA's rows of one weight tile
  constexpr int TR = DXF ? kDxfRows : kBlockM;                       // output rows per accumulator tile
  constexpr int ASTRIDE = DXF ? dxf_acc_stride<BN>() : acc_stride<BN>();
  constexpr int NACC = DXF ? dxf_nacc<BN>() : slab_nacc<BN>();
  static_assert(DXF || MSUB <= NACC, "a super-tile needs MSUB accumulators at once");
  uint8_t* slabs = smem;
  uint8_t* wbuf = slabs + (size_t)sp.sa * sp.slab_bytes;
  uint8_t* stage = wbuf + (size_t)sp.w_tiles * kWTileBytes;
  float* s_bias = reinterpret_cast<float*>(stage + (size_t)sp.stage_bufs * epi_stage_bytes<BN>());   // [2][BN]
  float* s_xch = s_bias + 2 * BN;                                    // DXF: boundary rows between the warps of a group
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_xch + (DXF ? kDxfXchFloats : 0));
  uint64_t* slab_full = bars;
  uint64_t* slab_empty = slab_full + 4;
  uint64_t* w_full = slab_empty + 4;
  uint64_t* w_empty = w_full + 8;
  uint64_t* acc_full = w_empty + 8;
  uint64_t* acc_empty = acc_full + kMaxAcc;
  uint64_t* st_ready = acc_empty + kMaxAcc;
  uint64_t* st_done = st_ready + 4;
  uint64_t* w_free = st_done + 4;             // resident weights: the MMAs of the layer have read them
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_free + 1);
  PubFifo* pub = reinterpret_cast<PubFifo*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool split = sp.stage_bufs == 1;
  const int NL = cp.n_layers;
  // virtual CTA / grid and the units of one N tile: a CTA and a super-tile, or a CTA pair and two super-tiles
  const int rank = PAIR ? (int)cluster_ctarank() : 0;
  const int vb = PAIR ? (int)blockIdx.x >> 1 : (int)blockIdx.x;
  const int vg = PAIR ? (int)gridDim.x >> 1 : (int)gridDim.x;
  const int units = PAIR ? (sp.num_super + 1) >> 1 : sp.num_super;
  auto real_s = [&](int su) { return PAIR ? 2 * su + rank : su; };     // may be == num_super (odd count): a dummy tile

  if (warp == 0 && lane == 0) {
    for (int l = 0; l < NL; ++l) {
      tma_prefetch_desc(&cp.m[l].a);
      tma_prefetch_desc(&cp.m[l].b);
    }
    for (int s = 0; s < 4; ++s) { mbar_init(&slab_full[s], 1); mbar_init(&slab_empty[s], 1); }
    for (int s = 0; s < 8; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    // PAIR: the leader's acc_empty collects the epilogue warps of both CTAs
    for (int s = 0; s < kMaxAcc; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], (split ? 8 : 4) * (PAIR ? 2 : 1)); }
    for (int s = 0; s < 4; ++s) { mbar_init(&st_ready[s], 1); mbar_init(&st_done[s], split ? 8 : 4); }
    mbar_init(w_free, 1);
    pub->head = 0;
    pub->tail = 0;
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) tmem_alloc_pair(tmem_slot, sp.tmem_cols);
    else tmem_alloc(tmem_slot, sp.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();      // the barriers of both CTAs exist before either signals the other's
  tc_fence_after();
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;
  constexpr int R = MSUB * TR;
  const int num_items_v = units * cp.n_tiles;
  const int my_items = (num_items_v - vb + vg - 1) / vg;
  const int total_u = my_items * MSUB;

  if (warp == 0) {
    // Producer warp.  Lane 0 issues every TMA load; ALL lanes poll the dataflow flags: lane L looks at flag (L % per) of
    // the item (L / per) positions ahead (per = flags per item), so one L2 round trip usually clears the next ~10 items
    // (flags only grow).  With a single polling thread the serial latency per item (three dependent acquire loads) was
    // longer than a dx-fused tile takes to compute.
    uint32_t slot = 0, slot_ph = 1, ws = 0, ws_ph = 1;
    // full barriers: this CTA's, or (PAIR) the leader's -- which counts the bytes of both CTAs
    const uint32_t slab_full_a = PAIR ? mapa_u32(&slab_full[0], 0) : smem_u32(&slab_full[0]);
    const uint32_t w_full_a = PAIR ? mapa_u32(&w_full[0], 0) : smem_u32(&w_full[0]);
    const bool expects = !PAIR || rank == 0;
    constexpr uint32_t kTx = PAIR ? 2u : 1u;
    const int w_row0 = PAIR ? rank * (BN / 2) : 0;
    const int reach = (sp.halo + R - 1) / R;
    const int per = 2 * reach + 1, look = 32 / per;
    const unsigned full = (1u << per) - 1u;
    for (int l = 0; l < NL; ++l) {
      const CUtensorMap* map_a = &cp.m[l].a;
      const CUtensorMap* map_b = &cp.m[l].b;
      if (lane == 0 && sp.w_resident) {
        if (l > 0) mbar_wait(w_free, (uint32_t)(l - 1) & 1u);
        if (expects) mbar_expect_tx(&w_full[0], kTx * (uint32_t)(sp.w_tiles * kWTileBytes));
        for (int kc = 0; kc < p0.num_kc; ++kc)
          for (int t = 0; t < p0.ntaps; ++t)
            tma_load_2d_to<PAIR>(wbuf + (size_t)(kc * p0.ntaps + t) * kWTileBytes, map_b, w_full_a, kc * 64, t * p0.n_alloc + w_row0);
      }
      SubIter it;
      it.init2(units, vb, vg);
      int ready_upto = (l == 0 || (cp.debug & 1)) ? my_items : 0;       // items [k, ready_upto) have their inputs
      for (int k = 0; k < my_items; ++k, it.next_item()) {
        if (k >= ready_upto) {
          // the rows a slab covers: super-tiles s-reach .. s+reach of the previous layer
          const int o = lane / per, j = lane - o * per;
          const unsigned int* row = cp.flags + (size_t)(l - 1) * sp.num_super;
          int fidx = -1;
          if (o < look && k + o < my_items) {
            const long long item = (long long)vb + (long long)(k + o) * vg;
            const int f = real_s((int)(item % units)) - reach + j;
            if (f >= 0 && f < sp.num_super) fidx = f;
          }
          const long long t0 = clock64();
          for (;;) {
            const bool ok = fidx < 0 || ld_relaxed(row + fidx) >= (unsigned)cp.n_tiles;
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            int cnt = 0;
            while (cnt < look && ((m >> (cnt * per)) & full) == full) ++cnt;
            if (cnt > 0) { ready_upto = k + cnt; break; }
            __nanosleep(32);
            if (clock64() - t0 > (8ll << 30)) __trap();      // ~4 s: a CTA of this grid never became resident
          }
          __threadfence();                 // every polling lane: order what follows after the flags it has seen
          __syncwarp();
          if (lane == 0) fence_proxy_async_all();            // ... and lane 0's TMA (async proxy) reads after that
        }
        if (lane == 0) {
          const long long r0 = (long long)real_s(it.s) * R - sp.halo;      // a dummy tile lies past the last row: zero fill
          for (int kc = 0; kc < p0.num_kc; ++kc) {
            mbar_wait(&slab_empty[slot], slot_ph);
            if (expects) mbar_expect_tx(&slab_full[slot], kTx * (uint32_t)sp.slab_bytes);
            for (int i = 0; i < sp.nboxes; ++i)
              tma_load_2d_to<PAIR>(slabs + (size_t)slot * sp.slab_bytes + (size_t)i * sp.box_rows * 128, map_a, slab_full_a + slot * 8,
                                   kc * 64, (int)(r0 + (long long)i * sp.box_rows));
            if (++slot == (uint32_t)sp.sa) { slot = 0; slot_ph ^= 1; }
            if (!sp.w_resident) {
              for (int t = 0; t < p0.ntaps; ++t) {
                mbar_wait(&w_empty[ws], ws_ph);
                if (expects) mbar_expect_tx(&w_full[ws], kTx * kWTileBytes);
                tma_load_2d_to<PAIR>(wbuf + (size_t)ws * kWTileBytes, map_b, w_full_a + ws * 8, kc * 64,
                                     t * p0.n_alloc + it.nt * BN + w_row0);
                if (++ws == (uint32_t)sp.sw) { ws = 0; ws_ph ^= 1; }
              }
            }
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    if ((!PAIR || rank == 0) && elect_one()) {
      const uint32_t idesc = umma_idesc_f16(T16<T>::kFmt, PAIR ? 2 * kBlockM : kBlockM, DXF ? 3 * BN : BN);
      uint32_t slot = 0, slot_ph = 0, ws = 0, ws_ph = 0, aslot = 0, aph = 1;
      const uint32_t slab0_addr = smem_u32(slabs);
      const uint32_t wbuf_addr = smem_u32(wbuf);
      int tap_shift[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) tap_shift[t] = sp.halo + (p0.ntaps == 9 ? (t / 3 - 1) * p0.wp + (t % 3 - 1) : 0);
      for (int l = 0; l < NL; ++l) {
        if (sp.w_resident) { mbar_wait(&w_full[0], (uint32_t)l & 1u); }
        tc_fence_after();
        for (int k = 0; k < my_items; ++k) {
          if constexpr (DXF) {
            // dx-fused: one slab (one K chunk, resident weights), sub-tile-major.  A sub-tile takes its accumulator
            // when its turn comes and is handed to the epilogue after its own 3 x ksteps MMAs, so a super-tile may
            // have as many sub-tiles as fit the slab ring, not as many as there are TMEM slots (3 at N = 144).
            mbar_wait(&slab_full[slot], slot_ph);
            tc_fence_after();
            const uint32_t slab_addr = slab0_addr + slot * (uint32_t)sp.slab_bytes;
            const int ksteps = min(64, p0.k16) >> 4;
#pragma unroll
            for (int j = 0; j < MSUB; ++j) {
              mbar_wait(&acc_empty[aslot], aph);
              tc_fence_after();
              const uint32_t dj = tmem_base + aslot * ASTRIDE;
              // one MMA per kernel row dy: B = the three dx weight tiles of that row stacked along N (they are
              // contiguous in the resident weight buffer), A = the centre-column view of the slab, fetched ONCE
              // for the three taps.  Accumulator columns [dx*BN, +BN) hold sum_dy A[g+(dy-1)wp] W(dy,dx) for the
              // tile row g; the epilogue adds the dx = 0 block of row g-1 and the dx = 2 block of row g+1.
#pragma unroll
              for (int dy = 0; dy < 3; ++dy) {
                if (cp.debug & 4) break;
                const uint64_t bdesc = umma_desc_sw128(wbuf_addr + (uint32_t)(dy * 3) * kWTileBytes);
                const uint64_t adesc = umma_desc_sw128(slab_addr + (uint32_t)(dy * p0.wp + j * TR) * 128);
                mma_ksteps_rt<BN>(ksteps, dj, adesc, bdesc, idesc, dy == 0 ? 0u : 1u);
              }
              tc_commit(&acc_full[aslot]);
              if (++aslot == NACC) { aslot = 0; aph ^= 1; }
            }
            tc_commit(&slab_empty[slot]);
            if (++slot == (uint32_t)sp.sa) { slot = 0; slot_ph ^= 1; }
            continue;
          }
          uint32_t d[MSUB], dslot[MSUB];
#pragma unroll
          for (int j = 0; j < MSUB; ++j) {
            if constexpr (PAIR) mbar_wait_cluster(&acc_empty[aslot], aph);      // arrivals come from both CTAs
            else mbar_wait(&acc_empty[aslot], aph);
            dslot[j] = aslot;
            d[j] = tmem_base + aslot * ASTRIDE;
            if (++aslot == NACC) { aslot = 0; aph ^= 1; }
          }
          tc_fence_after();
          uint32_t w_res_addr = wbuf_addr;
          for (int kc = 0; kc < p0.num_kc; ++kc) {
            mbar_wait(&slab_full[slot], slot_ph);
            tc_fence_after();
            const uint32_t slab_addr = slab0_addr + slot * (uint32_t)sp.slab_bytes;
            const int ksteps = min(64, p0.k16 - kc * 64) >> 4;
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              if (t < p0.ntaps) {
                uint32_t w_addr;
                if (sp.w_resident) {
                  w_addr = w_res_addr;
                  w_res_addr += kWTileBytes;
                } else {
                  mbar_wait(&w_full[ws], ws_ph);
                  tc_fence_after();
                  w_addr = wbuf_addr + ws * kWTileBytes;
                }
                const uint64_t bdesc = umma_desc_sw128(w_addr);
                const uint64_t adesc0 = umma_desc_sw128(slab_addr + (uint32_t)tap_shift[t] * 128);
                const uint32_t first = (kc | t) == 0 ? 0u : 1u;
#pragma unroll
                for (int j = 0; j < MSUB; ++j)
                  mma_ksteps_rt<BN, PAIR>(ksteps, d[j], adesc0 + (uint64_t)(j * (kBlockM * 128 / 16)), bdesc, idesc, first);
                if (!sp.w_resident) {
                  tc_commit_x<PAIR>(&w_empty[ws]);
                  if (++ws == (uint32_t)sp.sw) { ws = 0; ws_ph ^= 1; }
                }
              }
            }
            tc_commit_x<PAIR>(&slab_empty[slot]);
            if (++slot == (uint32_t)sp.sa) { slot = 0; slot_ph ^= 1; }
          }
#pragma unroll
          for (int j = 0; j < MSUB; ++j) tc_commit_x<PAIR>(&acc_full[dslot[j]]);
        }
        if (sp.w_resident) tc_commit_x<PAIR>(w_free);
      }
    }
    __syncwarp();
  } else {
    // stage buffers: 1 (both groups share it), 2 (one per group) or 4 (two per group: a group computes into one while
    // the other drains through the TMA store and receives the next residual tile)
    const int SE = sp.stage_bufs;
    const int se_mask = SE - 1, se_sh = SE == 4 ? 2 : SE == 2 ? 1 : 0;
    // every role numbers the sub-tiles of the WHOLE chain: ug = l * total_u + u.  Stage buffer, its barrier phase and
    // the epilogue group of a sub-tile are functions of ug, so the rings simply keep turning across layers.
    if (warp == 11) {
      if (lane == 0) {
        unsigned int seen = 0;
        for (;;) {
          const unsigned int head = lds_acquire(&pub->head);
          const unsigned int n = head & 0x7fffffffu;
          if (seen != n) {
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
            for (; seen != n; ++seen)
              asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(cp.flags + pub->ring[seen % kPubRing]) : "memory");
            sts_release(&pub->tail, seen);
          } else if (head >> 31) {
            break;
          } else {
            __nanosleep(64);
          }
        }
      }
      __syncwarp();
    } else if (warp == 10) {
      if (lane == 0) {
        unsigned int pushed = 0;
        int ubase = 0;
        // super-tiles whose stores are committed but not yet published (FIFO, newest last): an item is published once
        // `lag` younger items have been committed, so that waiting for ITS stores does not wait for recent ones.  The
        // host sizes the lag to ~2.5 us of tiles: 4 items for the 0.5 us dx-fused tiles, 0 for the >= 6 us tiles of
        // C >= 192, where a CTA has one or two items per layer and a late flag stalls the neighbours for a whole item.
        constexpr int kLagMax = 4;
        const int lag = cp.lag;
        const bool drain = num_items_v / vg <= lag + 1;      // otherwise the FIFO runs across the layer boundary
        int pend[kLagMax + 1];
        int npend = 0;
        for (int l = 0; l < NL; ++l, ubase += total_u) {
          TapGemmKParams p = p0;
          p.residual = cp.has_res[l] ? reinterpret_cast<const void*>(1) : nullptr;      // epi_load_residual only reads geometry
          const bool has_res = cp.has_res[l] != 0;
          const CUtensorMap* map_out = &cp.m[l].out;
          const CUtensorMap* map_res = &cp.m[l].res;
          // the residual tile of super-tile s is an earlier layer's output of super-tile s: written, completed and
          // published by THIS thread (same super-tile -> CTA assignment in every layer)
          SubIter st_it, pr_it;
          st_it.init2(units, vb, vg);
          pr_it.init2(units, vb, vg);
          auto prepare = [&](int u) {
            const int ug = ubase + u;
            const int e = ug & se_mask;
            if (has_res) {
              EpiTile t;
              t.m0 = (long long)real_s(pr_it.s) * R + (long long)pr_it.j * TR;
              t.tb = t.ty0 = t.tx0 = 0;
              if constexpr (DXF) dxf_load_residual<BN>(p, map_res, t.m0, stage + (size_t)e * epi_stage_bytes<BN>(), &st_ready[e]);
              else epi_load_residual<BN>(p, map_res, t, pr_it.nt * BN, stage + (size_t)e * epi_stage_bytes<BN>(), &st_ready[e]);
            } else {
              mbar_arrive(&st_ready[e]);
            }
            pr_it.template next<MSUB>();
          };
          tma_store_wait_read<0>();               // (no drain at the previous layer's end) the stage buffers have been read
          for (int u = 0; u < SE && u < total_u; ++u) prepare(u);
          for (int u = 0; u < total_u; ++u) {
            const int ug = ubase + u;
            const int e = ug & se_mask;
            mbar_wait(&st_done[e], (uint32_t)(ug >> se_sh) & 1u);
            EpiTile t;
            t.m0 = (long long)real_s(st_it.s) * R + (long long)st_it.j * TR;
            t.tb = t.ty0 = t.tx0 = 0;
            if constexpr (DXF) dxf_store<BN>(p, map_out, t.m0, stage + (size_t)e * epi_stage_bytes<BN>());
            else epi_store<BN>(p, map_out, t, st_it.nt * BN, stage + (size_t)e * epi_stage_bytes<BN>());
            const int s_now = real_s(st_it.s);
            const bool item_end = st_it.j == MSUB - 1;
            st_it.template next<MSUB>();
            if (item_end && l + 1 < NL && !(cp.debug & 1)) {
#pragma unroll
              for (int i = 0; i <= kLagMax; ++i)
                if (i == npend) pend[i] = s_now < sp.num_super ? l * sp.num_super + s_now : -1;      // -1: a pair's dummy tile
              ++npend;
              // publish the oldest pending item: its MSUB store groups precede the lag * MSUB most recent ones, which
              // stay in flight (waiting for the newest stores would stall this thread, and with it the stage buffers)
              if (npend > lag) {
                switch (lag) {
                  case 0: tma_store_wait_complete<0>(); break;
                  case 1: tma_store_wait_complete<MSUB>(); break;
                  case 2: tma_store_wait_complete<2 * MSUB>(); break;
                  default: tma_store_wait_complete<kLagMax * MSUB>(); break;
                }
                if (pend[0] >= 0) pub_push(pub, pushed, (unsigned)pend[0]);
#pragma unroll
                for (int i = 0; i < kLagMax; ++i) pend[i] = pend[i + 1];
                --npend;
              }
            }
            if (u + SE < total_u) {
              tma_store_wait_read<0>();
              prepare(u + SE);
            }
          }
          if (drain || l + 2 >= NL) {
            // Few items per layer: publish everything now (a neighbour's next-layer item may need our LAST items while
            // we would wait for `lag` more commits that need ITS flags).  With more than lag + 1 items per CTA and layer
            // the youngest items are published during the next layer's first ones, which only need flags of items that
            // were published inside this layer: no cycle, and no drain bubble at the layer boundary.  The last layer
            // pushes nothing (nobody reads its flags), so the layer before it drains as well.
            tma_store_wait_all();                 // this CTA's tiles so far are in global memory
#pragma unroll
            for (int i = 0; i <= kLagMax; ++i)
              if (i < npend && pend[i] >= 0) pub_push(pub, pushed, (unsigned)pend[i]);
            npend = 0;
          }
        }
        sts_release(&pub->head, pushed | 0x80000000u);
      }
      __syncwarp();
    } else {
      const int q = warp & 3;
      const int grp = (warp - 2) >> 2;
      const int step = split ? 1 : 2;
      const int gtid = (threadIdx.x - 64) & (split ? 255 : 127);
      float* my_bias = s_bias + (split ? 0 : grp * BN);
      int ug = split ? 0 : grp;                 // next sub-tile (chain numbering) of this group
      uint32_t aslot = (uint32_t)ug, aph = 0;
      uint32_t xpar = 0;
      int ubase = 0;
      const uint32_t acc_empty_a = PAIR ? mapa_u32(&acc_empty[0], 0) : 0u;      // the leader's accumulator-free barriers
      for (int l = 0; l < NL; ++l, ubase += total_u) {
        TapGemmKParams p = p0;
        p.relu = cp.relu[l];
        p.bias = cp.bias[l];
        const bool has_res = cp.has_res[l] != 0;
        SubIter it;
        it.init2(units, vb, vg);
        int u = ug - ubase;                     // 0 or 1 at a layer start
        if (u == 1) it.template next<MSUB>();
        int cur_nt = -1;
        for (; u < total_u; u += step, ug += step) {
          const int nt = it.nt;
          EpiTile tile;
          tile.m0 = (long long)real_s(it.s) * R + (long long)it.j * TR;
          tile.tb = tile.ty0 = tile.tx0 = 0;
          if (nt != cur_nt) {
            if (split) { epi_sync<256>(); epi_load_bias<BN, 256>(p, nt * BN, my_bias, gtid); epi_sync<256>(); }
            else { epi_sync_grp(grp); epi_load_bias<BN, 128>(p, nt * BN, my_bias, gtid); epi_sync_grp(grp); }
            cur_nt = nt;
          }
          mbar_wait(&acc_full[aslot], aph);
          tc_fence_after();
          const uint32_t acc = tmem_base + aslot * ASTRIDE;
          const int e = ug & se_mask;               // == grp with one buffer per group
          uint8_t* st = stage + (size_t)e * epi_stage_bytes<BN>();
          mbar_wait(&st_ready[e], (uint32_t)(ug >> se_sh) & 1u);
          if constexpr (DXF) {
            float* xch = s_xch + grp * (kDxfXchFloats / 2);
            if (cp.debug & 2) {
              EpiTile t2 = tile;
              t2.m0 = tile.m0 - 1;
              epi_compute<T, BN>(p, t2, 0, acc + BN, st, my_bias, has_res, q, lane, 0, 1);
            } else if (has_res) dxf_compute<T, BN, true>(p, tile.m0, acc, st, my_bias, xch, xpar, q, lane, grp);
            else dxf_compute<T, BN, false>(p, tile.m0, acc, st, my_bias, xch, xpar, q, lane, grp);
          } else {
            epi_compute<T, BN>(p, tile, nt * BN, acc, st, my_bias, has_res, q, lane, split ? grp : 0, split ? 2 : 1);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&st_done[e]);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (PAIR) mbar_arrive_cluster(acc_empty_a + aslot * 8);
            else mbar_arrive(&acc_empty[aslot]);
          }
          it.template next<MSUB>();
          if (step == 2) it.template next<MSUB>();
          aslot += step;
          if (aslot >= NACC) { aslot -= NACC; aph ^= 1; }
        }
      }
    }
  }
  tc_fence_before();
  if constexpr (PAIR) {
    cluster_sync_all();                   // neither CTA leaves (or frees TMEM) while the other may still touch it
    if (warp == 1) tmem_dealloc_pair(tmem_base, sp.tmem_cols);
  } else {
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, sp.tmem_cols);
  }
}

// ---- host side --------------------------------------------------------------------------------
struct ChainLaunch {
  ChainParams cp;
  TapGemmKParams kp;
  SlabParams sp;
  dim3 grid;
  int bn, dtype;
  int dxf;
  int pair;        // CTA-pair plan (cta_group::2): grid = 2 x pairs, launched in clusters of two
};

template <typename T, int BN, int MSUB>
static int chain_set_attr() {
  if constexpr (MSUB <= slab_nacc<BN>()) {
    BUCTD_CUDA(cudaFuncSetAttribute(tapgemm_chain_kernel<T, BN, MSUB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
  }
  if constexpr (BN <= 64) {
    BUCTD_CUDA(cudaFuncSetAttribute(tapgemm_chain_kernel<T, BN, MSUB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
  }
  if constexpr (BN >= 96 && MSUB <= slab_nacc<BN>()) {
    BUCTD_CUDA(cudaFuncSetAttribute(tapgemm_chain_kernel<T, BN, MSUB, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
  }
  return BUCTD_OK;
}
template <typename T, int BN>
static int chain_set_attr_bn() {
  int rc;
  if ((rc = chain_set_attr<T, BN, 1>())) return rc;
  if ((rc = chain_set_attr<T, BN, 2>())) return rc;
  if ((rc = chain_set_attr<T, BN, 3>())) return rc;
  return chain_set_attr<T, BN, 4>();
}
template <typename T>
static int chain_set_attrs() {
  int rc;
  if ((rc = chain_set_attr_bn<T, 32>())) return rc;
  if ((rc = chain_set_attr_bn<T, 48>())) return rc;
  if ((rc = chain_set_attr_bn<T, 64>())) return rc;
  if ((rc = chain_set_attr_bn<T, 96>())) return rc;
  if ((rc = chain_set_attr_bn<T, 128>())) return rc;
  if ((rc = chain_set_attr_bn<T, 192>())) return rc;
  return chain_set_attr_bn<T, 256>();
}
int tapgemm_chain_init_device() {
  int rc = chain_set_attrs<__half>();
  return rc ? rc : chain_set_attrs<__nv_bfloat16>();
}

// Re-plan a chain whose layers qualify for the dx-fused kernel: 126-row tiles, slab = msub*126 + 2*halo rows, four
// stage buffers, resident weights.  Returns false when the shape does not qualify (the regular plan stays).
// BUCTD_DXF=0 disables it, BUCTD_DXF=2 also admits C = 64; BUCTD_DXF_MSUB pins the super-tile size (default 1).
static bool chain_plan_dxf(const buctd_tapgemm_args& f, const TapGemmLaunch& L, int sm_count, SlabParams* out, dim3* grid) {
  const char* e = getenv("BUCTD_DXF");
  if (e != nullptr && atoi(e) == 0) return false;
  const int bn = L.bn;
  // N = 192 (C = 64) measured slower than the regular plan (44.7 vs 42.3 us per convolution at 96x72, B = 64): the
  // epilogue with two passes and two accumulator slots no longer hides behind the MMAs; BUCTD_DXF=2 forces it (tests)
  const bool allow64 = e != nullptr && atoi(e) == 2;
  if (f.ntaps != 9 || L.kp.num_kc != 1 || !(bn == 32 || bn == 48 || (bn == 64 && allow64)) || f.n_alloc != bn || f.n != bn || !L.sp.w_resident)
    return false;
  // super-tile: the slab of two 126-row tiles carries 1.59x its rows through TMA instead of 2.17x (one tile) --
  // measured 32.5 / 33.1 us (C = 48 @96x72) and 22.4 / 23.8 us (C = 32 @96x72) -- as long as the four stage
  // buffers still fit beside it; three tiles only fit with two stage buffers (41.4 us).  BUCTD_DXF_MSUB pins it.
  const char* em = getenv("BUCTD_DXF_MSUB");
  const char* eb = getenv("BUCTD_DXF_STAGES");
  SlabParams c = L.sp;
  const int halo = f.out_w + 2;
  const long long total_w = 9ll * bn * 128;
  const long long fixed = 1024 + 2 * bn * 4 + kDxfXchFloats * 4 + 512;
  int msub = 0, sa = 0, se = 0, nboxes = 0, box_rows = 0;
  long long slab = 0, need = 0;
  auto plan = [&](int m, int want_se) {
    const int rows = m * kDxfRows + 2 * halo;
    const int nb = ceil_div(rows, 256);
    const int br = (ceil_div(rows, nb) + 7) / 8 * 8;
    const long long sl = (long long)nb * br * 128;
    for (int csa = 3; csa >= 2; --csa) {
      const long long nd = fixed + (long long)want_se * kEpiBlockBytes + csa * sl + total_w;
      if (nd <= kSmemBudget) {
        msub = m; sa = csa; se = want_se; nboxes = nb; box_rows = br; slab = sl; need = nd;
        return true;
      }
    }
    return false;
  };
  if (em != nullptr) {
    int m = atoi(em);
    m = m < 1 ? 1 : m > 4 ? 4 : m;
    if (!(eb && atoi(eb) == 2) && plan(m, 4)) {}
    else if (!plan(m, 2)) return false;
  } else if (eb && atoi(eb) == 2) {
    if (!plan(2, 2) && !plan(1, 2)) return false;
  } else if (!plan(2, 4) && !plan(1, 4) && !plan(1, 2)) {
    return false;
  }
  const long long supers = (f.m_rows + (long long)msub * kDxfRows - 1) / ((long long)msub * kDxfRows);
  c.msub = msub; c.halo = halo; c.nboxes = nboxes; c.box_rows = box_rows; c.slab_bytes = (int)slab;
  c.sa = sa; c.sw = 0; c.stage_bufs = se; c.w_resident = 1; c.w_tiles = 9;
  c.num_super = (int)supers; c.num_items = (int)supers; c.smem_bytes = (int)need;
  c.tmem_cols = 512; c.direct = 0; c.jmajor = 0;
  *out = c;
  *grid = dim3((unsigned)(c.num_items < sm_count ? c.num_items : sm_count), 1);
  return true;
}

size_t tapgemm_chain_flag_count(const ChainLaunch& C) { return (size_t)C.cp.n_layers * C.sp.num_super; }

int tapgemm_chain_prepare(const buctd_tapgemm_chain_args& a, ChainLaunch* C, int sm_count, unsigned int* counter) {
  BUCTD_CHECK(a.n_layers >= 1 && a.n_layers <= BUCTD_MAX_CHAIN, BUCTD_ERR_INVALID, "tapgemm_chain: n_layers %d out of range [1,%d]",
              a.n_layers, BUCTD_MAX_CHAIN);
  BUCTD_CHECK(counter != nullptr, BUCTD_ERR_INVALID, "tapgemm_chain: no barrier counter");
  const buctd_tapgemm_args& f = a.layer[0];
  memset(&C->cp, 0, sizeof(C->cp));
  SlabParams dxf_sp;
  dim3 dxf_grid;
  C->dxf = 0;
  // CTA-pair plan for the wide layers: the one-CTA kernel streams the whole weight set per work item through a ring
  // that is bound by the TMA round trip (4 tiles per ~0.9 us).  Measured per convolution at B = 64 (tools/chain_bench.py,
  // profiles/r3_chain_pair.txt): C = 192 @24x18 22.1 -> 17.1 us, C = 384 @12x9 22.4 -> 19.8 us; C = 96 @48x36 28.3 -> 27.9
  // (its floor is the activation traffic, not the weights), so pairs start at 192 channels.  BUCTD_CHAIN_PAIR=0 keeps
  // the one-CTA plan everywhere, =2 pairs from 96 channels on (tests).
  {
    const char* e = getenv("BUCTD_CHAIN_PAIR");
    const int mode = e != nullptr ? atoi(e) : 1;
    C->pair = mode != 0 && f.n_alloc >= (mode == 2 ? 96 : 192) && sm_count >= 2;
  }
  struct PairScope {       // tapgemm_prepare plans half weight tiles while this is alive
    explicit PairScope(int on) { tl_slab_pair = on; }
    ~PairScope() { tl_slab_pair = 0; }
  } pair_scope(C->pair);
  for (int l = 0; l < a.n_layers; ++l) {
    const buctd_tapgemm_args& g = a.layer[l];
    BUCTD_CHECK(g.mode == BUCTD_TG_FLAT && g.out_mode == BUCTD_OUT_ACT16, BUCTD_ERR_INVALID,
                "tapgemm_chain: layer %d must be a FLAT tap-GEMM with act16 output", l);
    BUCTD_CHECK(g.dtype == f.dtype && g.ntaps == f.ntaps && g.batch == f.batch && g.out_h == f.out_h && g.out_w == f.out_w &&
                    g.m_rows == f.m_rows && g.k == f.k && g.a_cs == f.a_cs && g.n == f.n && g.n_alloc == f.n_alloc &&
                    g.w_cs == f.w_cs && g.out_cs == f.out_cs && (g.residual == nullptr || g.res_cs == g.out_cs),
                BUCTD_ERR_INVALID, "tapgemm_chain: layer %d does not have the geometry of layer 0", l);
    TapGemmLaunch L;
    // every layer is planned as if it carried a residual when ANY layer does: the stage / ring carve-up must be common
    bool any_res = false;
    for (int j = 0; j < a.n_layers; ++j) any_res = any_res || a.layer[j].residual != nullptr;
    buctd_tapgemm_args planned = g;
    if (any_res && planned.residual == nullptr) { planned.residual = planned.out; planned.res_cs = planned.out_cs; }
    int rc = tapgemm_prepare(planned, &L, sm_count);
    if (rc) return rc;
    if (l == 0) {
      C->kp = L.kp; C->sp = L.sp; C->grid = L.grid; C->bn = L.bn; C->dtype = L.dtype;
      C->sp.direct = 0; C->sp.jmajor = 0;
      C->dxf = C->pair ? 0 : (chain_plan_dxf(g, L, sm_count, &dxf_sp, &dxf_grid) ? 1 : 0);
      if (C->pair) {
        BUCTD_CHECK(L.bn >= 96 && L.bn % 32 == 0, BUCTD_ERR_INVALID, "tapgemm_chain: pair plan with N tile %d", L.bn);
        int mc = 0;
        if ((rc = pair_max_clusters(&mc))) return rc;
        const int units = (L.sp.num_super + 1) / 2, items_v = units * (L.sp.num_items / L.sp.num_super);
        int pairs = sm_count / 2 < mc ? sm_count / 2 : mc;
        if (pairs > items_v) pairs = items_v;
        C->grid = dim3((unsigned)(2 * pairs), 1);
      }
    } else {
      BUCTD_CHECK(L.bn == C->bn && L.sp.msub == C->sp.msub && L.sp.sa == C->sp.sa && L.sp.sw == C->sp.sw &&
                      L.sp.stage_bufs == C->sp.stage_bufs && L.sp.smem_bytes == C->sp.smem_bytes,
                  BUCTD_ERR_INVALID, "tapgemm_chain: layer %d planned differently from layer 0", l);
    }
    BUCTD_CHECK(L.sp.direct == 0, BUCTD_ERR_INVALID, "tapgemm_chain: direct epilogue is not supported in chains (unset BUCTD_DIRECT_MAX_BN)");
    C->cp.m[l].a = L.map_a; C->cp.m[l].b = L.map_b; C->cp.m[l].out = L.map_out; C->cp.m[l].res = L.map_res;
    if (C->dxf) {
      // slab boxes of the dx-fused plan; 126-row output / residual tiles
      const int inner = g.a_cs < L.kp.k16 ? g.a_cs : L.kp.k16;
      if ((rc = tapgemm_encode_rows_map(&C->cp.m[l].a, g.dtype, g.a, inner, g.a_cs, g.m_rows, dxf_sp.box_rows))) return rc;
      if ((rc = tapgemm_encode_rows_map(&C->cp.m[l].out, g.dtype, g.out, g.out_cs, g.out_cs, g.m_rows, kDxfRows))) return rc;
      if (g.residual != nullptr &&
          (rc = tapgemm_encode_rows_map(&C->cp.m[l].res, g.dtype, g.residual, g.res_cs, g.res_cs, g.m_rows, kDxfRows)))
        return rc;
    }
    C->cp.bias[l] = g.bias;
    C->cp.has_res[l] = g.residual != nullptr;
    C->cp.relu[l] = g.relu;
    if (g.residual == nullptr) C->cp.m[l].res = C->cp.m[l].out;
  }
  if (C->dxf) { C->sp = dxf_sp; C->grid = dxf_grid; }
  C->kp.bias = nullptr; C->kp.residual = nullptr; C->kp.out = nullptr; C->kp.relu = 0;
  C->cp.flags = counter;
  C->cp.n_layers = a.n_layers;
  C->cp.n_tiles = C->sp.num_items / C->sp.num_super;
  {
    // publish lag ~ 2.5 us of tiles, from the measured MMA issue interval (50 + 0.35 N clk per 128 x N x 16 MMA)
    const int n_mma = (C->dxf ? 3 : f.ntaps) * (C->kp.k16 / 16);
    const double tile_us = n_mma * (50.0 + 0.35 * (C->dxf ? 3 * C->bn : C->bn)) / 1965.0;
    const int lag_tiles = (int)(2.5 / tile_us);
    int lag = (lag_tiles + C->sp.msub - 1) / C->sp.msub;
    if (getenv("BUCTD_CHAIN_LAG")) lag = atoi(getenv("BUCTD_CHAIN_LAG"));
    C->cp.lag = lag <= 0 ? 0 : lag == 1 ? 1 : lag == 2 ? 2 : 4;
  }
  C->cp.debug = getenv("BUCTD_CHAIN_DEBUG") ? atoi(getenv("BUCTD_CHAIN_DEBUG")) : 0;
  if (getenv("BUCTD_PLAN_DEBUG") != nullptr)
    fprintf(stderr, "[buctd] chain plan: %d layers bn %d dxf %d pair %d lag %d msub %d sa %d stage_bufs %d slab %d B supers %d grid %u smem %d\n", a.n_layers, C->bn,
            C->dxf, C->pair, C->cp.lag, C->sp.msub, C->sp.sa, C->sp.stage_bufs, C->sp.slab_bytes, C->sp.num_super, C->grid.x, C->sp.smem_bytes);
  return BUCTD_OK;
}

template <typename T, int BN, int MSUB>
static void chain_launch_one(const ChainLaunch& C, cudaStream_t st) {
  if (C.dxf) {
    if constexpr (BN <= 64) {
      launch_pdl(tapgemm_chain_kernel<T, BN, MSUB, true>, C.grid, dim3(kChainThreads), C.sp.smem_bytes, st, C.cp, C.kp, C.sp);
    }
    return;
  }
  if (C.pair) {
    if constexpr (BN >= 96 && MSUB <= slab_nacc<BN>()) {
      launch_pdl_cluster(tapgemm_chain_kernel<T, BN, MSUB, false, true>, C.grid, dim3(kChainThreads), C.sp.smem_bytes, st, 2u, C.cp, C.kp,
                         C.sp);
    }
    return;
  }
  if constexpr (MSUB <= slab_nacc<BN>()) {
    launch_pdl(tapgemm_chain_kernel<T, BN, MSUB, false>, C.grid, dim3(kChainThreads), C.sp.smem_bytes, st, C.cp, C.kp, C.sp);
  }
}
template <typename T, int BN>
static void chain_launch_bn(const ChainLaunch& C, cudaStream_t st) {
  switch (C.sp.msub) {
    case 1: chain_launch_one<T, BN, 1>(C, st); break;
    case 2: chain_launch_one<T, BN, 2>(C, st); break;
    case 3: chain_launch_one<T, BN, 3>(C, st); break;
    default: chain_launch_one<T, BN, 4>(C, st); break;
  }
}
template <typename T>
static int chain_dispatch(const ChainLaunch& C, cudaStream_t st) {
  switch (C.bn) {
    case 32: chain_launch_bn<T, 32>(C, st); break;
    case 48: chain_launch_bn<T, 48>(C, st); break;
    case 64: chain_launch_bn<T, 64>(C, st); break;
    case 96: chain_launch_bn<T, 96>(C, st); break;
    case 128: chain_launch_bn<T, 128>(C, st); break;
    case 192: chain_launch_bn<T, 192>(C, st); break;
    case 256: chain_launch_bn<T, 256>(C, st); break;
    default: set_error("tapgemm_chain: unsupported BN %d", C.bn); return BUCTD_ERR_INVALID;
  }
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}

int tapgemm_chain_launch(const ChainLaunch& C, cudaStream_t st) {
  BUCTD_CUDA(cudaMemsetAsync(C.cp.flags, 0, sizeof(unsigned int) * tapgemm_chain_flag_count(C), st));
  if (C.dtype == BUCTD_F16) return chain_dispatch<__half>(C, st);
  return chain_dispatch<__nv_bfloat16>(C, st);
}

ChainLaunch* tapgemm_chain_new() { return new ChainLaunch(); }
void tapgemm_chain_set_flags(ChainLaunch* C, unsigned int* flags) { C->cp.flags = flags; }
void tapgemm_chain_delete(ChainLaunch* c) { delete c; }

}  // namespace buctd
