// Tap-GEMM: implicit-GEMM convolutions and linear layers on the sm_100a tensor cores.
//
//   out[m, n] = act( sum_t sum_k A[m (+) tap_t, k] * W[t][n][k] + bias[n] + residual[m, n] )
//
// Two kernels share one epilogue (epilogue.cuh):
//
//  * slab kernel (FLAT mode: 3x3 stride-1 convolutions, 1x1 convolutions, nn.Linear) -- persistent,
//    one CTA per SM.  The convolution is im2col-free AND tap-free in memory traffic: for a
//    super-tile of MSUB x 128 consecutive rows of the padded NHWC matrix the producer TMA-loads ONE
//    slab of rows [r0 - (W+2), r0 + MSUB*128 + (W+2)) per 64-channel chunk; the nine taps are nine
//    tcgen05.mma passes whose A descriptor simply starts dy*(W+1)+dx rows further into the same
//    SWIZZLE_128B slab (the swizzle is a function of the shared-memory address bits, so a row-shifted
//    view of a TMA-written tile is still a valid K-major operand).  Weights stay resident in shared
//    memory when they fit (C <= 64), otherwise they stream through a ring.  Accumulators are double
//    buffered in TMEM so the epilogue of super-tile i overlaps the MMAs of super-tile i+1.
//  * patch kernel (PATCH mode: stride-2 3x3 convolutions) -- one 128-pixel output patch per CTA, A
//    tiles per (tap, chunk) through a 4-D TMA box traversed with element stride 2.
//
// Warp roles in both: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one lane),
// warps 2..5 = epilogue.
//
// Replaces: nn.Conv2d/BatchNorm2d/ReLU/residual of lib/models/pose_hrnet_coam.py:44-60, 81-101,
// 200-242, 405-432, 565, 635-641, 666-670 and nn.Linear of lib/models/self_attention.py:74-76,87,159.
#include "common.cuh"
#include "epilogue.cuh"
#include "internal.h"
#include "slab_common.cuh"

namespace buctd {

// Optional role timers (debug builds: -DBUCTD_ROLE_TIMERS): cycles each warp role spends waiting / working,
// written to a global buffer [grid][8] that tools/role_timers.py reads back.
#ifdef BUCTD_ROLE_TIMERS
__device__ unsigned long long g_role_cycles[160 * 16];
#define RT_DECL(name) long long name = 0
#define RT_BEGIN(t) const long long t = clock64()
#define RT_ADD(acc, t) acc += clock64() - t
#define RT_STORE(slot, v) g_role_cycles[blockIdx.x * 16 + (slot)] = (unsigned long long)(v)
#else
#define RT_DECL(name)
#define RT_BEGIN(t)
#define RT_ADD(acc, t)
#define RT_STORE(slot, v)
#endif

constexpr int kATileBytes = kBlockM * 128;
constexpr int kThreads = 192;          // patch kernel: producer + MMA + 4 epilogue warps

// Stages of the (tap, K chunk) ring.  One tile per CTA means prologue (barriers, TMEM allocation), the nine taps and the
// epilogue of a tile run back to back; with a 2-deep ring (BUCTD_PATCH_STAGES, default) two or three CTAs fit an SM and
// one's epilogue overlaps the other's loads -- measured against the 4-deep ring that fills the SM with one CTA.
#ifndef BUCTD_PATCH_STAGES
#define BUCTD_PATCH_STAGES 2
#endif
template <int BN> __host__ __device__ constexpr int patch_stages() { return BN <= 128 ? BUCTD_PATCH_STAGES : (BUCTD_PATCH_STAGES < 3 ? BUCTD_PATCH_STAGES : 3); }

// =================================================================================================
// PATCH kernel (stride-2 3x3): one tile per CTA
// =================================================================================================
template <typename T, int BN>
__global__ void __launch_bounds__(kThreads)
tapgemm_patch_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                     const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_res,
                     const TapGemmKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int S = patch_stages<BN>();
  constexpr int kBTileBytes = BN * 128;
  constexpr int kTmemCols = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + S * kATileBytes;                  // BN*128 is a multiple of 1024 for BN % 8 == 0
  uint8_t* stage = smem_b + S * kBTileBytes;
  float* s_bias = reinterpret_cast<float*>(stage + epi_stage_bytes<BN>());
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_bias + BN);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* tmem_full_bar = empty_bar + S;
  uint64_t* res_bar = tmem_full_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * BN;
  const bool has_res = p.residual != nullptr;

  EpiTile tile;
  tile.m0 = 0;
  {
    const int t = blockIdx.x;
    const int per_img = p.tiles_x * p.tiles_y;
    tile.tb = t / per_img;
    const int r = t - tile.tb * per_img;
    tile.ty0 = (r / p.tiles_x) * p.th;
    tile.tx0 = (r % p.tiles_x) * p.tw;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    tma_prefetch_desc(&map_out);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_init(res_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_launch_dependents();
  pdl_wait();                                   // inputs of this kernel are produced by the previous one
  const uint32_t tmem_base = *tmem_slot;
  const int iters = p.ntaps * p.num_kc;

  if (warp == 0) {
    if (lane == 0) {
      if (has_res) epi_load_residual<BN>(p, &map_res, tile, n0, stage, res_bar);
      for (int it = 0; it < iters; ++it) {
        const int s = it % S;
        const int t = it / p.num_kc;
        const int kc = it - t * p.num_kc;
        mbar_wait(&empty_bar[s], ((it / S) & 1) ^ 1);
        mbar_expect_tx(&full_bar[s], p.a_tile_bytes + kBTileBytes);
        int dy = 0, dx = 0;
        if (p.ntaps == 9) { dy = t / 3 - 1; dx = t % 3 - 1; }
        tma_load_4d(smem_a + s * kATileBytes, &map_a, &full_bar[s], kc * 64, tile.tx0 * p.stride + dx,
                    tile.ty0 * p.stride + dy, tile.tb);
        tma_load_2d(smem_b + s * kBTileBytes, &map_b, &full_bar[s], kc * 64, t * p.n_alloc + n0);
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_f16(T16<T>::kFmt, kBlockM, BN);
    const uint32_t a_addr = smem_u32(smem_a), b_addr = smem_u32(smem_b);
    for (int it = 0; it < iters; ++it) {
      const int s = it % S;
      const int kc = it % p.num_kc;
      mbar_wait(&full_bar[s], (it / S) & 1);
      tc_fence_after();
      const uint64_t adesc = umma_desc_sw128(a_addr + s * kATileBytes);
      const uint64_t bdesc = umma_desc_sw128(b_addr + s * kBTileBytes);
      const int ksteps = min(64, p.k16 - kc * 64) >> 4;
      if (elect_one()) {
        tc_mma_f16(tmem_base, adesc, bdesc, idesc, it != 0 ? 1u : 0u);
        if (ksteps > 1) tc_mma_f16(tmem_base, adesc + 2, bdesc + 2, idesc, 1u);
        if (ksteps > 2) tc_mma_f16(tmem_base, adesc + 4, bdesc + 4, idesc, 1u);
        if (ksteps > 3) tc_mma_f16(tmem_base, adesc + 6, bdesc + 6, idesc, 1u);
        tc_commit(&empty_bar[s]);
      }
      __syncwarp();
    }
    if (elect_one()) tc_commit(tmem_full_bar);
    __syncwarp();
  } else {
    const int etid = threadIdx.x - 64;
    epi_load_bias<BN, 128>(p, n0, s_bias, etid);
    epi_sync<128>();
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    if (has_res) mbar_wait(res_bar, 0);
    epi_compute<T, BN>(p, tile, n0, tmem_base, stage, s_bias, has_res, warp & 3, lane, 0, 1);
    fence_proxy_async_smem();
    epi_sync<128>();
    if (etid == 0) {
      epi_store<BN>(p, &map_out, tile, n0, stage);
      tma_store_wait_all();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// =================================================================================================
// SLAB kernel (FLAT mode): persistent, A-slab + row-shifted tap views, ring of TMEM accumulators
// =================================================================================================
template <typename T, int BN, int MSUB>
__global__ void __launch_bounds__(kSlabThreads, 1)
tapgemm_slab_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_res,
                    const TapGemmKParams p, const SlabParams sp) {
  RT_BEGIN(t_cta_start);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kWTileBytes = BN * 128;
  constexpr int NACC = slab_nacc<BN>();
  static_assert(MSUB <= NACC, "a super-tile needs MSUB accumulators at once");
  uint8_t* slabs = smem;                                         // [sa][slab_bytes]
  uint8_t* wbuf = slabs + (size_t)sp.sa * sp.slab_bytes;         // resident: [num_kc*ntaps] tiles; ring: [sw] tiles
  uint8_t* stage = wbuf + (size_t)sp.w_tiles * kWTileBytes;       // [stage_bufs][epi_stage_bytes]
  float* s_bias = reinterpret_cast<float*>(stage + (size_t)sp.stage_bufs * epi_stage_bytes<BN>());   // [2][BN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + 2 * BN);
  uint64_t* slab_full = bars;                 // [sa]
  uint64_t* slab_empty = slab_full + 4;       // [sa]
  uint64_t* w_full = slab_empty + 4;          // [sw] (resident: [0] only)
  uint64_t* w_empty = w_full + 8;             // [sw]
  uint64_t* acc_full = w_empty + 8;           // [NACC] accumulator slot holds a finished sub-tile
  uint64_t* acc_empty = acc_full + kMaxAcc;   // [NACC] accumulator slot has been read out
  uint64_t* st_ready = acc_empty + kMaxAcc;   // [2] stage buffer e may be written (its residual tile, if any, has landed)
  uint64_t* st_done = st_ready + 2;           // [2] stage buffer e holds a finished 16-bit tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(st_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // Epilogue organisation.  Two groups of 4 warps (one warp per TMEM lane quarter).  With two stage buffers
  // each group owns one buffer and takes every other sub-tile (whole tiles in parallel); with one stage
  // buffer (shared memory is tight) both groups work on the same tile and split its columns.
  const bool split = sp.stage_bufs == 1 && !sp.direct;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int s = 0; s < 4; ++s) { mbar_init(&slab_full[s], 1); mbar_init(&slab_empty[s], 1); }
    for (int s = 0; s < 8; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int s = 0; s < kMaxAcc; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], split ? 8 : 4); }
    for (int s = 0; s < 2; ++s) { mbar_init(&st_ready[s], 1); mbar_init(&st_done[s], split ? 8 : 4); }
    tma_prefetch_desc(&map_out);
    fence_mbar_init();
    if (sp.w_resident) {
      // Start the pipeline before the CTA-wide sync (the TMA engine only needs the barriers this thread has just
      // initialised): the weights are constants and go first, then -- once the previous kernel of the stream has
      // completed -- the first slabs.  Their latency overlaps the TMEM allocation and the prologue of the other warps.
      mbar_expect_tx(&w_full[0], (uint32_t)(sp.w_tiles * kWTileBytes));
      for (int kc = 0; kc < p.num_kc; ++kc)
        for (int t = 0; t < p.ntaps; ++t)
          tma_load_2d(wbuf + (size_t)(kc * p.ntaps + t) * kWTileBytes, &map_b, &w_full[0], kc * 64, t * p.n_alloc);
      pdl_wait();
      const long long r0 = (long long)((int)blockIdx.x % sp.num_super) * (MSUB * kBlockM) - sp.halo;
      const int npre = p.num_kc < sp.sa ? p.num_kc : sp.sa;
      for (int kc = 0; kc < npre; ++kc) {
        mbar_expect_tx(&slab_full[kc], (uint32_t)sp.slab_bytes);
        for (int i = 0; i < sp.nboxes; ++i)
          tma_load_2d(slabs + (size_t)kc * sp.slab_bytes + (size_t)i * sp.box_rows * 128, &map_a, &slab_full[kc], kc * 64,
                      (int)(r0 + (long long)i * sp.box_rows));
      }
    }
  }
  if (warp == 1) tmem_alloc(tmem_slot, sp.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_launch_dependents();
  RT_DECL(t_pro); RT_DECL(t_pdl);
  RT_ADD(t_pro, t_cta_start);
  pdl_wait();                                   // activations are produced by the previous kernel
  RT_ADD(t_pdl, t_cta_start);
  const uint32_t tmem_base = *tmem_slot;
  constexpr int R = MSUB * kBlockM;
  const int my_items = (sp.num_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t slot = 0, slot_ph = 1, ws = 0, ws_ph = 1;     // producer waits on the "empty" phase (starts free)
      int pre = sp.w_resident ? (p.num_kc < sp.sa ? p.num_kc : sp.sa) : 0;    // slabs of the first item issued in the prologue
      SubIter it;
      it.init(sp.num_super);
      for (int k = 0; k < my_items; ++k, it.next_item()) {
        const long long r0 = (long long)it.s * R - sp.halo;
        for (int kc = 0; kc < p.num_kc; ++kc) {
          if (pre > 0) {
            --pre;
            if (++slot == (uint32_t)sp.sa) { slot = 0; slot_ph ^= 1; }
            continue;
          }
          mbar_wait(&slab_empty[slot], slot_ph);
          mbar_expect_tx(&slab_full[slot], (uint32_t)sp.slab_bytes);
          for (int i = 0; i < sp.nboxes; ++i)
            tma_load_2d(slabs + (size_t)slot * sp.slab_bytes + (size_t)i * sp.box_rows * 128, &map_a, &slab_full[slot],
                        kc * 64, (int)(r0 + (long long)i * sp.box_rows));
          if (++slot == (uint32_t)sp.sa) { slot = 0; slot_ph ^= 1; }
          if (!sp.w_resident) {
            for (int t = 0; t < p.ntaps; ++t) {
              mbar_wait(&w_empty[ws], ws_ph);
              mbar_expect_tx(&w_full[ws], kWTileBytes);
              tma_load_2d(wbuf + (size_t)ws * kWTileBytes, &map_b, &w_full[ws], kc * 64, t * p.n_alloc + it.nt * BN);
              if (++ws == (uint32_t)sp.sw) { ws = 0; ws_ph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // One elected lane runs the whole MMA role (elect.sync lets ptxas keep every descriptor in
    // uniform registers: ~2 SASS instructions per tcgen05.mma).
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_f16(T16<T>::kFmt, kBlockM, BN);
      uint32_t slot = 0, slot_ph = 0;          // slab ring position / phase (wrap counters: no runtime div/mod in this thread)
      uint32_t ws = 0, ws_ph = 0;              // weight ring position / phase
      uint32_t aslot = 0, aph = 1;             // accumulator ring: wait for "empty" (starts free)
      RT_DECL(t_acc); RT_DECL(t_slab); RT_DECL(t_w); RT_DECL(t_total); RT_DECL(t_issue); RT_DECL(t_commit);
      RT_BEGIN(t_start);
      const uint32_t slab0_addr = smem_u32(slabs);
      const uint32_t wbuf_addr = smem_u32(wbuf);
      int tap_shift[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) tap_shift[t] = sp.halo + (p.ntaps == 9 ? (t / 3 - 1) * p.wp + (t % 3 - 1) : 0);
      if (sp.w_resident) { mbar_wait(&w_full[0], 0); }
      tc_fence_after();
      if (sp.desc_mode == 7 && sp.w_resident && p.num_kc == 1 && p.ntaps == 9) {
        // TIMING EXPERIMENT ONLY (BUCTD_DESC_MODE=7; results are wrong): the issue order of the dx-fused formulation
        //  -- for every (dy, sub-tile, k-step) three MMAs that share ONE A tile through the A collector
        // (fill / use / lastuse) with the weights of dx = -1, 0, +1.
        const int ksteps = p.k16 >> 4;
        for (int k = 0; k < my_items; ++k) {
          uint32_t d[MSUB], dslot[MSUB];
#pragma unroll
          for (int j = 0; j < MSUB; ++j) {
            mbar_wait(&acc_empty[aslot], aph);
            dslot[j] = aslot;
            d[j] = tmem_base + aslot * acc_stride<BN>();
            if (++aslot == NACC) { aslot = 0; aph ^= 1; }
          }
          { RT_BEGIN(t0); mbar_wait(&slab_full[slot], slot_ph); RT_ADD(t_slab, t0); }
          tc_fence_after();
          const uint32_t slab_addr = slab0_addr + slot * (uint32_t)sp.slab_bytes;
          RT_BEGIN(t_i0);
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
            for (int j = 0; j < MSUB; ++j) {
              const uint64_t adesc = umma_desc_sw128(slab_addr + (uint32_t)(tap_shift[dy * 3 + 1] + j * kBlockM) * 128);
              for (int ks = 0; ks < ksteps; ++ks) {
                const uint32_t first = (dy | ks) == 0 ? 0u : 1u;
                const uint64_t b0 = umma_desc_sw128(wbuf_addr + (dy * 3 + 0) * kWTileBytes) + (uint64_t)(2 * ks);
                const uint64_t b1 = umma_desc_sw128(wbuf_addr + (dy * 3 + 1) * kWTileBytes) + (uint64_t)(2 * ks);
                const uint64_t b2 = umma_desc_sw128(wbuf_addr + (dy * 3 + 2) * kWTileBytes) + (uint64_t)(2 * ks);
                tc_mma_f16_coll<1>(d[j], adesc + (uint64_t)(2 * ks), b0, idesc, first);
                tc_mma_f16_coll<2>(d[(j + 1) % MSUB], adesc + (uint64_t)(2 * ks), b1, idesc, first);
                tc_mma_f16_coll<3>(d[(j + 2) % MSUB], adesc + (uint64_t)(2 * ks), b2, idesc, first);
              }
            }
          }
          RT_ADD(t_issue, t_i0);
          tc_commit(&slab_empty[slot]);
          if (++slot == (uint32_t)sp.sa) { slot = 0; slot_ph ^= 1; }
#pragma unroll
          for (int j = 0; j < MSUB; ++j) tc_commit(&acc_full[dslot[j]]);
        }
      } else if (sp.jmajor) {
        // resident weights, one K chunk: sub-tile-major order.  Sub-tile j is finished (and handed to the epilogue)
        // after its own taps, so the epilogue lags the tensor core by ONE sub-tile, not one super-tile.
        const int ksteps = p.k16 >> 4;
        for (int k = 0; k < my_items; ++k) {
          { RT_BEGIN(t0); mbar_wait(&slab_full[slot], slot_ph); RT_ADD(t_slab, t0); }
          tc_fence_after();
          const uint32_t slab_addr = slab0_addr + slot * (uint32_t)sp.slab_bytes;
#pragma unroll
          for (int j = 0; j < MSUB; ++j) {
            { RT_BEGIN(t0); mbar_wait(&acc_empty[aslot], aph); RT_ADD(t_acc, t0); }
            tc_fence_after();
            const uint32_t d = tmem_base + aslot * acc_stride<BN>();
            RT_BEGIN(t_i0);
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              if (t < p.ntaps) {
                const uint64_t bdesc = umma_desc_sw128(wbuf_addr + t * kWTileBytes);
                const uint64_t adesc = umma_desc_sw128(slab_addr + (uint32_t)(tap_shift[t] + j * kBlockM) * 128);
                mma_ksteps_rt<BN>(ksteps, d, adesc, bdesc, idesc, t == 0 ? 0u : 1u);
              }
            }
            RT_ADD(t_issue, t_i0);
            tc_commit(&acc_full[aslot]);
            if (++aslot == NACC) { aslot = 0; aph ^= 1; }
          }
          tc_commit(&slab_empty[slot]);
          if (++slot == (uint32_t)sp.sa) { slot = 0; slot_ph ^= 1; }
        }
      } else {
        // streamed weights / several K chunks: tap-major order, every weight tile is used by all MSUB sub-tiles
        for (int k = 0; k < my_items; ++k) {
          uint32_t d[MSUB], dslot[MSUB];
#pragma unroll
          for (int j = 0; j < MSUB; ++j) {
            { RT_BEGIN(t0); mbar_wait(&acc_empty[aslot], aph); RT_ADD(t_acc, t0); }
            dslot[j] = aslot;
            d[j] = tmem_base + aslot * acc_stride<BN>();
            if (++aslot == NACC) { aslot = 0; aph ^= 1; }
          }
          tc_fence_after();
          uint32_t w_res_addr = wbuf_addr;         // resident weights: tile (kc, t) lives at (kc*ntaps + t) * kWTileBytes
          for (int kc = 0; kc < p.num_kc; ++kc) {
            { RT_BEGIN(t0); mbar_wait(&slab_full[slot], slot_ph); RT_ADD(t_slab, t0); }
            tc_fence_after();
            const uint32_t slab_addr = slab0_addr + slot * (uint32_t)sp.slab_bytes;
            const int ksteps = min(64, p.k16 - kc * 64) >> 4;
            RT_BEGIN(t_i0);
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              if (t < p.ntaps) {
                uint32_t w_addr;
                if (sp.w_resident) {
                  w_addr = w_res_addr;
                  w_res_addr += kWTileBytes;
                } else {
                  { RT_BEGIN(t0); mbar_wait(&w_full[ws], ws_ph); RT_ADD(t_w, t0); }
                  tc_fence_after();
                  w_addr = wbuf_addr + ws * kWTileBytes;
                }
                const uint64_t bdesc = umma_desc_sw128(w_addr);
                const uint64_t adesc0 = umma_desc_sw128(slab_addr + (uint32_t)tap_shift[t] * 128);
                const uint32_t first = (kc | t) == 0 ? 0u : 1u;
#pragma unroll
                for (int j = 0; j < MSUB; ++j)
                  mma_ksteps_rt<BN>(ksteps, d[j], adesc0 + (uint64_t)(j * (kBlockM * 128 / 16)), bdesc, idesc, first);
                if (!sp.w_resident) {
                  tc_commit(&w_empty[ws]);
                  if (++ws == (uint32_t)sp.sw) { ws = 0; ws_ph ^= 1; }
                }
              }
            }
            RT_ADD(t_issue, t_i0);
            tc_commit(&slab_empty[slot]);
            if (++slot == (uint32_t)sp.sa) { slot = 0; slot_ph ^= 1; }
          }
#pragma unroll
          for (int j = 0; j < MSUB; ++j) tc_commit(&acc_full[dslot[j]]);
        }
      }
      RT_ADD(t_total, t_start);
      RT_STORE(0, t_total); RT_STORE(1, t_acc); RT_STORE(2, t_slab); RT_STORE(3, t_w); RT_STORE(6, t_issue); RT_STORE(7, t_commit);
    }
    __syncwarp();
  } else {
    // ---- epilogue: warps 2..9 compute (TMEM -> stage), warp 10 moves tiles (residual TMA loads, TMA stores) ----
    const bool has_res = p.residual != nullptr;
    const bool head = p.out_mode == BUCTD_OUT_NCHW32;
    const int SE = sp.stage_bufs;
    const int total_u = my_items * MSUB;
    if (warp == 10) {
      if (!head && (!sp.direct || has_res) && elect_one()) {
        SubIter st_it, pr_it;
        st_it.init(sp.num_super);
        pr_it.init(sp.num_super);
        // prepare(u): the store that last used buffer u%SE has been read -> (re)fill it with the residual tile
        auto prepare = [&](int u) {
          const int e = SE == 2 ? (u & 1) : 0;
          if (has_res) {
            EpiTile t;
            t.m0 = pr_it.template m0<MSUB>();
            t.tb = t.ty0 = t.tx0 = 0;
            epi_load_residual<BN>(p, &map_res, t, pr_it.nt * BN, stage + (size_t)e * epi_stage_bytes<BN>(), &st_ready[e]);
          } else {
            mbar_arrive(&st_ready[e]);
          }
          pr_it.template next<MSUB>();
        };
        for (int u = 0; u < SE && u < total_u; ++u) prepare(u);
        for (int u = 0; u < total_u; ++u) {
          const int e = SE == 2 ? (u & 1) : 0;
          mbar_wait(&st_done[e], (SE == 2 ? (u >> 1) : u) & 1);
          EpiTile t;
          t.m0 = st_it.template m0<MSUB>();
          t.tb = t.ty0 = t.tx0 = 0;
          if (!sp.direct) epi_store<BN>(p, &map_out, t, st_it.nt * BN, stage + (size_t)e * epi_stage_bytes<BN>());
          st_it.template next<MSUB>();
          if (u + SE < total_u) {
            if (!sp.direct) tma_store_wait_read<0>();          // buffer e is free again once its store has read it
            prepare(u + SE);
          }
        }
        tma_store_wait_all();
      }
      __syncwarp();
    } else {
      RT_DECL(e_wait); RT_DECL(e_total); RT_DECL(e_res); RT_DECL(e_comp);
      RT_BEGIN(e_start);
      const int q = warp & 3;
      const int grp = (warp - 2) >> 2;          // 0: warps 2..5, 1: warps 6..9
      const int step = split ? 1 : 2;
      const int gtid = (threadIdx.x - 64) & (split ? 255 : 127);
      float* my_bias = s_bias + (split ? 0 : grp * BN);
      SubIter it;
      it.init(sp.num_super);
      int u = split ? 0 : grp;
      if (u == 1) it.template next<MSUB>();
      uint32_t aslot = (uint32_t)u, aph = 0;
      int cur_nt = -1;
      for (; u < total_u; u += step) {
        const int nt = it.nt;
        EpiTile tile;
        tile.m0 = it.template m0<MSUB>();
        tile.tb = tile.ty0 = tile.tx0 = 0;
        if (nt != cur_nt) {                      // (re)load the bias tile of this N tile
          if (split) { epi_sync<256>(); epi_load_bias<BN, 256>(p, nt * BN, my_bias, gtid); epi_sync<256>(); }
          else { epi_sync_grp(grp); epi_load_bias<BN, 128>(p, nt * BN, my_bias, gtid); epi_sync_grp(grp); }
          cur_nt = nt;
        }
        { RT_BEGIN(t0); mbar_wait(&acc_full[aslot], aph); RT_ADD(e_wait, t0); }
        tc_fence_after();
        const uint32_t acc = tmem_base + aslot * acc_stride<BN>();
        if (head) {
          if (!split || grp == 0) epi_head_nchw32<BN>(p, tile, nt * BN, acc, q, lane);
        } else if (sp.direct) {
          if (has_res) {       // direct stores; the residual tile arrives by TMA in this group's stage buffer
            uint8_t* st = stage + (size_t)grp * epi_stage_bytes<BN>();
            { RT_BEGIN(t0); mbar_wait(&st_ready[grp], (u >> 1) & 1); RT_ADD(e_res, t0); }
            { RT_BEGIN(t0); epi_compute_direct<T, BN, true>(p, tile, nt * BN, acc, st, my_bias, q, lane); RT_ADD(e_comp, t0); }
            __syncwarp();
            if (lane == 0) mbar_arrive(&st_done[grp]);
          } else {
            RT_BEGIN(t0);
            epi_compute_direct<T, BN, false>(p, tile, nt * BN, acc, stage, my_bias, q, lane);
            RT_ADD(e_comp, t0);
          }
        } else {
          const int e = split ? 0 : grp;
          uint8_t* st = stage + (size_t)e * epi_stage_bytes<BN>();
          { RT_BEGIN(t0); mbar_wait(&st_ready[e], (split ? u : (u >> 1)) & 1); RT_ADD(e_res, t0); }
          { RT_BEGIN(t0); epi_compute<T, BN>(p, tile, nt * BN, acc, st, my_bias, has_res, q, lane, split ? grp : 0, split ? 2 : 1); RT_ADD(e_comp, t0); }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&st_done[e]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[aslot]);
        it.template next<MSUB>();
        if (step == 2) it.template next<MSUB>();
        aslot += step;
        if (aslot >= NACC) { aslot -= NACC; aph ^= 1; }
      }
      RT_ADD(e_total, e_start);
      if (threadIdx.x == 64) { RT_STORE(4, e_total); RT_STORE(5, e_wait); RT_STORE(9, e_res); RT_STORE(10, e_comp); }
      if (threadIdx.x == 64 + 37) { RT_STORE(13, e_comp); }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, sp.tmem_cols);
#ifdef BUCTD_ROLE_TIMERS
  if (threadIdx.x == 0) { RT_DECL(t_all); RT_ADD(t_all, t_cta_start); RT_STORE(11, t_pro); RT_STORE(12, t_pdl); RT_STORE(14, t_all); }
#endif
}

// ---- host side --------------------------------------------------------------------------------
static int stage_bytes(int bn) { return (bn + 63) / 64 * kEpiBlockBytes; }
static size_t patch_smem_bytes(int bn) {
  const int s = bn <= 128 ? BUCTD_PATCH_STAGES : (BUCTD_PATCH_STAGES < 3 ? BUCTD_PATCH_STAGES : 3);
  return 1024 + (size_t)s * (kATileBytes + bn * 128) + stage_bytes(bn) + bn * 4 + 256;
}

template <typename T, int BN, int MSUB>
static int set_attr_slab() {
  if constexpr (MSUB <= slab_nacc<BN>()) {
    BUCTD_CUDA(cudaFuncSetAttribute(tapgemm_slab_kernel<T, BN, MSUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
  }
  return BUCTD_OK;
}
template <typename T, int BN>
static int set_attr() {
  BUCTD_CUDA(cudaFuncSetAttribute(tapgemm_patch_kernel<T, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)patch_smem_bytes(BN)));
  int rc;
  if ((rc = set_attr_slab<T, BN, 1>())) return rc;
  if ((rc = set_attr_slab<T, BN, 2>())) return rc;
  if ((rc = set_attr_slab<T, BN, 3>())) return rc;
  return set_attr_slab<T, BN, 4>();
}
template <typename T>
static int set_attrs() {
  int rc;
  if ((rc = set_attr<T, 32>())) return rc;
  if ((rc = set_attr<T, 48>())) return rc;
  if ((rc = set_attr<T, 64>())) return rc;
  if ((rc = set_attr<T, 96>())) return rc;
  if ((rc = set_attr<T, 128>())) return rc;
  if ((rc = set_attr<T, 192>())) return rc;
  return set_attr<T, 256>();
}
// called once per device from buctd_create (function attributes are per device)
int tapgemm_init_device() {
  int rc = set_attrs<__half>();
  return rc ? rc : set_attrs<__nv_bfloat16>();
}

thread_local int tl_slab_pair = 0;

static int pick_bn(int n_alloc) {
  static const int cand[] = {32, 48, 64, 96, 128, 192, 256};
  if (n_alloc <= 256) {
    for (int c : cand)
      if (c >= n_alloc) return c;
  }
  int best = 256, best_waste = 1 << 30;
  for (int c : {256, 192, 128}) {
    int waste = ceil_div(n_alloc, c) * c - n_alloc;
    if (waste < best_waste) { best = c; best_waste = waste; }
  }
  return best;
}

static void pick_patch_tile(int oh, int ow, int* tw, int* th) {
  // th x tw <= 128 output pixels per M tile, whole rows when they fit
  if (ow <= 128) {
    *tw = ow;
    int t = 128 / ow;
    if (t > oh) t = oh;
    int best = t;
    for (int c = t; c >= 1; --c)
      if (oh % c == 0) { best = (c * 2 > t) ? c : t; break; }
    *th = best;
  } else {
    *tw = 128;
    *th = 1;
  }
}

// ---- slab planning ----------------------------------------------------------------------------
// One candidate (bn, msub): shared-memory carve-up + an estimate of the kernel time from the three rates that
// bound the kernel (measured with tools/conv_sweep.py on B200):
//   ingest  : TMA fill of one SM's shared memory, ~36 B/clk (A slab incl. halo + streamed weight tiles),
//   tensor  : one 128 x bn x 16 SS-mode tcgen05.mma every ~max(60, 50 + 0.35 bn) clk (measured; the math alone is bn/2),
//   epilogue: ~700 + 9 bn clk per 128-row tile and group of 4 warps, two groups in parallel.
static bool plan_slab_one(const buctd_tapgemm_args& a, int bn, int num_kc, int k16, int msub, int ctas, SlabParams* out,
                          double* t_est) {
  const int nacc = (512 / acc_stride_of(bn)) < kMaxAcc ? (512 / acc_stride_of(bn)) : kMaxAcc;
  if (msub > nacc) return false;
  const int n_tiles = ceil_div(a.n_alloc, bn);
  const int halo = a.ntaps == 9 ? a.out_w + 2 : 0;        // (W+1) + 1 rows on each side
  const int wtile = (tl_slab_pair ? bn / 2 : bn) * 128;
  // direct epilogue: 16-bit rows, every N tile full (no channel tail), 16-byte aligned rows
  static const bool no_direct = getenv("BUCTD_NO_DIRECT") != nullptr;       // debug knob
  // (measured: a win for rows of <= 128 bytes, a loss for wide tiles whose rows are many partial-sector writes)
  // Opt-in (BUCTD_DIRECT_MAX_BN=256): isolated launches gain 3-18 % at bn <= 64, the whole forward does not
  // (14.15 vs 14.20 ms per 64 crops), so the staged TMA-store epilogue stays the default.
  static const int direct_max_bn = getenv("BUCTD_DIRECT_MAX_BN") ? atoi(getenv("BUCTD_DIRECT_MAX_BN")) : 0;
  bool direct = !no_direct && bn <= direct_max_bn && bn % 16 == 0 && a.out_mode != BUCTD_OUT_NCHW32 && a.n % bn == 0 && a.n == a.n_alloc &&
                a.out_cs % 16 == 0 && ((uintptr_t)a.out & 31) == 0;
  const bool stage_needed = !direct || a.residual != nullptr;      // direct mode still prefetches the residual tile by TMA
  const int stage = stage_needed ? stage_bytes(bn) : 0;
  const int fixed = 1024 + 2 * bn * 4 + 512;              // alignment slack + two bias tiles + barriers
  const long long total_w = (long long)a.ntaps * num_kc * wtile;
  // A CTA pair holds half of every weight tile per CTA: the 9 x 2 half tiles of a C = 96 layer (108 KB) stay resident
  // beside one slab set and one stage buffer, and nothing is streamed per work item (BUCTD_PAIR_RESIDENT=0: stream)
  const char* epr = getenv("BUCTD_PAIR_RESIDENT");
  const long long res_cap = tl_slab_pair && !(epr != nullptr && atoi(epr) == 0) ? 112 * 1024 : 80 * 1024;
  bool resident = n_tiles == 1 && total_w <= res_cap;
  const int rows = msub * 128 + 2 * halo;
  const int nboxes = ceil_div(rows, 256);
  const int box_rows = (ceil_div(rows, nboxes) + 7) / 8 * 8;   // every box starts on a 1024-byte swizzle atom
  const long long slab = (long long)nboxes * box_rows * 128;
  // preference order: two stage buffers, deep rings; then thinner weight ring; then one stage buffer
  int sa = 0, sw = 0, se = 0;
  long long need = 0;
  const int sa_max = num_kc > 1 ? 3 : 2;
  const char* esw = getenv("BUCTD_SW_MAX");               // experiment knobs: weight-ring depth (<= 8), stage buffers
  const char* ese = getenv("BUCTD_SE_MAX");
  // the weight ring is latency-bound (ring depth tiles per ~0.9 us TMA round trip: chain_bench with BUCTD_CHAIN_DEBUG=1):
  // as deep as fits beside two stage buffers and the slabs, up to 8 tiles (C = 96 chains: 4 -> 6 tiles, 25.8 -> 24.0 us
  // per convolution; whole forward 11.39 -> 11.25 ms)
  const int sw_max = esw ? (atoi(esw) < 2 ? 2 : atoi(esw) > 8 ? 8 : atoi(esw)) : 8;
  const int se_max = ese ? (atoi(ese) <= 1 ? 1 : 2) : 2;
  for (int pass = 0; pass < 2 && sa == 0; ++pass) {
    if (pass == 1) {
      // only the weight sets the pair rule newly admits fall back to streaming at the same super-tile; for the others
      // the caller tries a smaller super-tile first (resident weights are worth more than a long slab)
      if (!resident || total_w <= 80 * 1024) break;
      resident = false;
    }
    for (int cse = se_max; cse >= 1 && sa == 0; --cse)
      for (int csw = resident ? 0 : sw_max; (resident ? csw == 0 : csw >= 2) && sa == 0; --csw)
        for (int csa = sa_max; csa >= 2; --csa) {
          need = fixed + (long long)cse * stage + csa * slab + (resident ? total_w : (long long)csw * wtile);
          if (need <= kSmemBudget) { sa = csa; sw = csw; se = cse; break; }
        }
  }
  if (sa == 0) return false;
  if (direct && a.residual != nullptr && se != 2) direct = false;  // direct + residual needs one stage buffer per epilogue group
  SlabParams c;
  memset(&c, 0, sizeof(c));
  const long long supers = (a.m_rows + (long long)msub * 128 - 1) / ((long long)msub * 128);
  c.msub = msub; c.halo = halo; c.nboxes = nboxes; c.box_rows = box_rows; c.slab_bytes = (int)slab;
  c.sa = sa; c.sw = sw; c.stage_bufs = stage_needed ? se : 0; c.w_resident = resident; c.w_tiles = resident ? a.ntaps * num_kc : sw;
  c.num_super = (int)supers; c.num_items = (int)supers * n_tiles; c.smem_bytes = (int)need;
  // sub-tile-major order serialises 9*ksteps MMAs on ONE accumulator (measured 71 instead of 47 clk per MMA at bn=48):
  // kept behind a knob, off by default
  static const bool want_jmajor = getenv("BUCTD_JMAJOR") != nullptr && atoi(getenv("BUCTD_JMAJOR")) != 0;
  c.jmajor = want_jmajor && resident && num_kc == 1;
  c.direct = direct;
  const int cols = nacc * acc_stride_of(bn);
  c.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
  *out = c;
  // ---- time estimate (microseconds at 1.965 GHz) ----
  const double clk = 1965.0, fill_rate = 36.0 * clk;       // bytes per microsecond into one SM
  const int used = c.num_items < ctas ? c.num_items : ctas;
  const double rounds = (double)((c.num_items + used - 1) / used);
  const double ingest_item = ((double)nboxes * box_rows * k16 * 2 + (resident ? 0.0 : (double)a.ntaps * bn * k16 * 2)) / fill_rate;
  const double cyc_mma = fmax(60.0, 50.0 + 0.35 * bn);      // measured issue interval (tools/conv_sweep.py, role timers)
  const double mma_tile = a.ntaps * (k16 / 16) * cyc_mma / clk;
  const double epi_tile = (700.0 + 9.0 * bn) * (a.residual ? 1.2 : 1.0) / clk;
  const double mma_item = msub * mma_tile, epi_item = msub * epi_tile / 2.0;
  const bool overlapped = c.jmajor ? nacc >= 2 : nacc >= 2 * msub;
  const double steady = rounds * (overlapped ? fmax(fmax(ingest_item, mma_item), epi_item) : fmax(ingest_item, mma_item + epi_item));
  const double fill = ingest_item + (resident ? (double)total_w / fill_rate : 0.0) + (c.jmajor ? mma_tile : mma_item);
  const double drain = c.jmajor ? epi_tile : fmax(epi_tile, epi_item);
  *t_est = 3.0 + fill + steady + drain;
  return true;
}

// choose (bn, msub) for a FLAT tap-GEMM.  Measured on B200 over the whole (bn, msub) space for the W48 shapes
// (tools/slab_tune.py, profiles/r01_slab_tune.txt): one N tile whenever n_alloc <= 256 and, for wider layers, the
// widest exact tiling win everywhere (a 128 x bn x 16 tcgen05.mma costs ~50 + 0.35 bn clk, so wide tiles are the
// efficient ones); the best super-tile is the largest one that still leaves every CTA two work items, with double
// buffered accumulators (2 msub bn <= 512).  The time estimate of plan_slab_one only breaks ties / feeds the debug print.
static int plan_slab(const buctd_tapgemm_args& a, int num_kc, int k16, int ctas, int* bn_out, SlabParams* sp) {
  const char* force = getenv("BUCTD_FORCE_MSUB");         // test knobs: pin the super-tile size / the N tile
  const int forced = force ? atoi(force) : 0;
  const char* fbn = getenv("BUCTD_FORCE_BN");
  int bn = fbn ? atoi(fbn) : pick_bn(a.n_alloc);
  if (ceil_div(a.n_alloc, bn) > 1 && bn % 64 != 0) bn = pick_bn(a.n_alloc);   // store boxes of neighbouring N tiles must not overlap
  bool have = false;
  double t = 0.0;
  for (int msub = 4; msub >= 1; --msub) {
    if (forced > 0 ? msub > forced : 2 * msub * acc_stride_of(bn) > 512) continue;   // double-buffered accumulators
    SlabParams c;
    if (!plan_slab_one(a, bn, num_kc, k16, msub, ctas, &c, &t)) continue;
    *sp = c;
    have = true;
    if (forced > 0 || c.num_items >= 2 * ctas) break;     // largest super-tile that still gives every CTA >= 2 items
  }
  if (!have) {
    set_error("tapgemm: no shared-memory plan for n_alloc=%d kc=%d w=%d", a.n_alloc, num_kc, a.out_w);
    return BUCTD_ERR_INVALID;
  }
  *bn_out = bn;
  static const bool debug = getenv("BUCTD_PLAN_DEBUG") != nullptr;
  if (debug)
    fprintf(stderr, "[buctd] slab plan: rows %lld k %d n %d taps %d ctas %d -> bn %d msub %d sa %d sw %d stage_bufs %d resident %d items %d est %.1f us\n",
            (long long)a.m_rows, a.k, a.n, a.ntaps, ctas, bn, sp->msub, sp->sa, sp->sw, sp->stage_bufs, sp->w_resident,
            sp->num_items, t);
  return BUCTD_OK;
}

template <typename T, int BN, int MSUB>
static void launch_slab(const TapGemmLaunch& L, cudaStream_t st) {
  if constexpr (MSUB <= slab_nacc<BN>()) {
    launch_pdl(tapgemm_slab_kernel<T, BN, MSUB>, L.grid, dim3(kSlabThreads), L.sp.smem_bytes, st, L.map_a, L.map_b, L.map_out,
               L.map_res, L.kp, L.sp);
  }
}

template <typename T, int BN>
static int launch_one(const TapGemmLaunch& L, cudaStream_t st) {
  if (L.kp.mode == BUCTD_TG_PATCH) {
    launch_pdl(tapgemm_patch_kernel<T, BN>, L.grid, dim3(kThreads), patch_smem_bytes(BN), st, L.map_a, L.map_b, L.map_out, L.map_res,
               L.kp);
  } else {
    switch (L.sp.msub) {
      case 1: launch_slab<T, BN, 1>(L, st); break;
      case 2: launch_slab<T, BN, 2>(L, st); break;
      case 3: launch_slab<T, BN, 3>(L, st); break;
      default: launch_slab<T, BN, 4>(L, st); break;
    }
  }
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}

template <typename T>
static int dispatch_bn(const TapGemmLaunch& L, cudaStream_t st) {
  switch (L.bn) {
    case 32: return launch_one<T, 32>(L, st);
    case 48: return launch_one<T, 48>(L, st);
    case 64: return launch_one<T, 64>(L, st);
    case 96: return launch_one<T, 96>(L, st);
    case 128: return launch_one<T, 128>(L, st);
    case 192: return launch_one<T, 192>(L, st);
    case 256: return launch_one<T, 256>(L, st);
  }
  set_error("tapgemm: unsupported BN %d", L.bn);
  return BUCTD_ERR_INVALID;
}

int tapgemm_prepare(const buctd_tapgemm_args& a, TapGemmLaunch* L, int sm_count) {
  BUCTD_CHECK(a.dtype == BUCTD_F16 || a.dtype == BUCTD_BF16, BUCTD_ERR_INVALID, "tapgemm: bad dtype %d", a.dtype);
  BUCTD_CHECK(a.ntaps == 1 || a.ntaps == 9, BUCTD_ERR_INVALID, "tapgemm: ntaps must be 1 or 9, got %d", a.ntaps);
  BUCTD_CHECK(a.k > 0 && a.n > 0 && a.n_alloc >= a.n && a.n_alloc % 16 == 0, BUCTD_ERR_INVALID,
              "tapgemm: bad k/n/n_alloc %d/%d/%d", a.k, a.n, a.n_alloc);
  BUCTD_CHECK(a.a_cs % 8 == 0 && a.w_cs % 8 == 0 && a.w_cs >= a.k, BUCTD_ERR_INVALID,
              "tapgemm: channel strides must be multiples of 8 (a_cs %d, w_cs %d)", a.a_cs, a.w_cs);
  BUCTD_CHECK(a.a && a.w && a.out, BUCTD_ERR_INVALID, "tapgemm: null pointer");
  BUCTD_CHECK(((uintptr_t)a.a & 15) == 0 && ((uintptr_t)a.w & 15) == 0 && ((uintptr_t)a.out & 15) == 0 &&
                  ((uintptr_t)a.residual & 15) == 0,
              BUCTD_ERR_INVALID, "tapgemm: pointers must be 16-byte aligned");
  if (a.out_mode != BUCTD_OUT_NCHW32)
    BUCTD_CHECK(a.out_cs % 8 == 0 && (a.residual == nullptr || a.res_cs % 8 == 0), BUCTD_ERR_INVALID,
                "tapgemm: out_cs/res_cs must be multiples of 8 (%d/%d)", a.out_cs, a.res_cs);
  EncodeTiledFn enc = get_encode_tiled();
  BUCTD_CHECK(enc != nullptr, BUCTD_ERR_NO_DEVICE, "cuTensorMapEncodeTiled entry point not available");

  TapGemmKParams kp;
  memset(&kp, 0, sizeof(kp));
  const int k16 = (a.k + 15) / 16 * 16;
  kp.mode = a.mode; kp.ntaps = a.ntaps; kp.stride = a.stride > 0 ? a.stride : 1;
  kp.k16 = k16; kp.num_kc = (k16 + 63) / 64;
  kp.B = a.batch; kp.H = a.out_h; kp.W = a.out_w; kp.wp = a.out_w + 1;
  kp.m_rows = a.m_rows;
  kp.n = a.n; kp.n_alloc = a.n_alloc;
  kp.out_cs = a.out_cs; kp.res_cs = a.res_cs; kp.relu = a.relu; kp.out_mode = a.out_mode;
  kp.bias = a.bias; kp.residual = a.residual; kp.out = a.out;

  const CUtensorMapDataType dt = a.dtype == BUCTD_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  int bn = pick_bn(a.n_alloc);
  fastdiv_make((uint32_t)(a.out_w + 1), &kp.wp_mul, &kp.wp_sh);
  fastdiv_make((uint32_t)(a.out_h + 1), &kp.hp_mul, &kp.hp_sh);
  const cuuint64_t a_inner = (cuuint64_t)(a.a_cs < k16 ? a.a_cs : k16);
  CUresult cr;
  dim3 grid;
  memset(&L->sp, 0, sizeof(L->sp));
  L->pair = 0;
  if (pairgemm_wanted(a, kp.num_kc)) {
    BUCTD_CHECK(a.m_rows > 0 && a.m_rows < (1ll << 31), BUCTD_ERR_INVALID, "tapgemm: bad m_rows %lld", (long long)a.m_rows);
    L->kp = kp;
    return pairgemm_prepare(a, L, sm_count);
  }
  if (a.mode == BUCTD_TG_FLAT) {
    BUCTD_CHECK(a.m_rows > 0 && a.m_rows < (1ll << 31), BUCTD_ERR_INVALID, "tapgemm: bad m_rows %lld", (long long)a.m_rows);
    if (a.out_mode != BUCTD_OUT_ROWS16 || a.ntaps == 9)
      BUCTD_CHECK((long long)a.batch * (a.out_h + 1) * (a.out_w + 1) == a.m_rows, BUCTD_ERR_INVALID,
                  "tapgemm: m_rows %lld does not match act16 geometry %dx%dx%d", (long long)a.m_rows, a.batch, a.out_h, a.out_w);
    int rc = plan_slab(a, kp.num_kc, k16, sm_count, &bn, &L->sp);
    if (rc) return rc;
    const char* dm = getenv("BUCTD_DESC_MODE");        // debug knob for the row-shifted descriptor experiment
    L->sp.desc_mode = dm ? atoi(dm) : 0;
    cuuint64_t dims[2] = {a_inner, (cuuint64_t)a.m_rows};
    cuuint64_t strides[1] = {(cuuint64_t)a.a_cs * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)L->sp.box_rows};
    cuuint32_t es[2] = {1, 1};
    cr = enc(&L->map_a, dt, 2, const_cast<void*>(a.a), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    BUCTD_CHECK(cr == CUDA_SUCCESS, BUCTD_ERR_CUDA, "cuTensorMapEncodeTiled(A flat) failed: %d", (int)cr);
    kp.a_tile_bytes = (uint32_t)L->sp.slab_bytes;
    grid = dim3((unsigned)(L->sp.num_items < sm_count ? L->sp.num_items : sm_count), 1);
  } else {
    BUCTD_CHECK(a.out_mode == BUCTD_OUT_ACT16, BUCTD_ERR_INVALID, "tapgemm: PATCH mode writes act16 only");
    BUCTD_CHECK(a.in_h > 0 && a.in_w > 0, BUCTD_ERR_INVALID, "tapgemm: PATCH mode needs in_h/in_w");
    int tw, th;
    pick_patch_tile(a.out_h, a.out_w, &tw, &th);
    BUCTD_CHECK(tw * kp.stride <= 256 && th * kp.stride <= 256, BUCTD_ERR_INVALID, "tapgemm: patch box too large");
    kp.tw = tw; kp.th = th; kp.tiles_x = ceil_div(a.out_w, tw); kp.tiles_y = ceil_div(a.out_h, th);
    const cuuint64_t wp = a.in_w + 1, hp = a.in_h + 1;
    cuuint64_t dims[4] = {a_inner, wp, hp, (cuuint64_t)a.batch};
    cuuint64_t strides[3] = {(cuuint64_t)a.a_cs * 2, wp * a.a_cs * 2, hp * wp * a.a_cs * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)(tw * kp.stride), (cuuint32_t)(th * kp.stride), 1};
    cuuint32_t es[4] = {1, (cuuint32_t)kp.stride, (cuuint32_t)kp.stride, 1};
    cr = enc(&L->map_a, dt, 4, const_cast<void*>(a.a), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    BUCTD_CHECK(cr == CUDA_SUCCESS, BUCTD_ERR_CUDA, "cuTensorMapEncodeTiled(A patch) failed: %d", (int)cr);
    kp.a_tile_bytes = (uint32_t)(tw * th * 128);
    grid = dim3((unsigned)(a.batch * kp.tiles_x * kp.tiles_y), ceil_div(a.n_alloc, bn));
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)a.w_cs, (cuuint64_t)a.ntaps * a.n_alloc};
    cuuint64_t strides[1] = {(cuuint64_t)a.w_cs * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)(tl_slab_pair ? bn / 2 : bn)};
    cuuint32_t es[2] = {1, 1};
    cr = enc(&L->map_b, dt, 2, const_cast<void*>(a.w), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    BUCTD_CHECK(cr == CUDA_SUCCESS, BUCTD_ERR_CUDA, "cuTensorMapEncodeTiled(W) failed: %d", (int)cr);
  }
  // output / residual tile maps: [64 channels x 128 rows] SWIZZLE_128B boxes of the same row geometry as the tile
  L->map_out = L->map_a;
  L->map_res = L->map_a;
  if (a.out_mode != BUCTD_OUT_NCHW32) {
    for (int which = 0; which < 2; ++which) {
      void* base = which == 0 ? a.out : const_cast<void*>(a.residual);
      const int cs = which == 0 ? a.out_cs : a.res_cs;
      CUtensorMap* m = which == 0 ? &L->map_out : &L->map_res;
      if (base == nullptr) { *m = L->map_out; continue; }
      if (a.mode == BUCTD_TG_FLAT) {
        cuuint64_t dims[2] = {(cuuint64_t)cs, (cuuint64_t)a.m_rows};
        cuuint64_t strides[1] = {(cuuint64_t)cs * 2};
        cuuint32_t box[2] = {64, kBlockM};
        cuuint32_t es[2] = {1, 1};
        cr = enc(m, dt, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      } else {
        const cuuint64_t wp = a.out_w + 1, hp = a.out_h + 1;
        cuuint64_t dims[4] = {(cuuint64_t)cs, wp, hp, (cuuint64_t)a.batch};
        cuuint64_t strides[3] = {(cuuint64_t)cs * 2, wp * cs * 2, hp * wp * cs * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)kp.tw, (cuuint32_t)kp.th, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        cr = enc(m, dt, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      }
      BUCTD_CHECK(cr == CUDA_SUCCESS, BUCTD_ERR_CUDA, "cuTensorMapEncodeTiled(%s) failed: %d", which ? "residual" : "out", (int)cr);
    }
  }
  L->kp = kp; L->grid = grid; L->bn = bn; L->dtype = a.dtype;
  return BUCTD_OK;
}

int tapgemm_encode_rows_map(CUtensorMap* m, int dtype, const void* base, int inner, int cs, long long rows, int box_rows) {
  EncodeTiledFn enc = get_encode_tiled();
  BUCTD_CHECK(enc != nullptr, BUCTD_ERR_NO_DEVICE, "cuTensorMapEncodeTiled entry point not available");
  const CUtensorMapDataType dt = dtype == BUCTD_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cs * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult cr = enc(m, dt, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BUCTD_CHECK(cr == CUDA_SUCCESS, BUCTD_ERR_CUDA, "cuTensorMapEncodeTiled(rows map) failed: %d", (int)cr);
  return BUCTD_OK;
}

#ifdef BUCTD_ROLE_TIMERS
extern "C" __attribute__((visibility("default"))) int buctd_debug_role_cycles(unsigned long long* out, int n) {
  return (int)cudaMemcpyFromSymbol(out, g_role_cycles, sizeof(unsigned long long) * n);
}
#endif

int tapgemm_launch(const TapGemmLaunch& L, cudaStream_t st) {
  if (L.pair) return pairgemm_launch(L, st);
  if (L.dtype == BUCTD_F16) return dispatch_bn<__half>(L, st);
  return dispatch_bn<__nv_bfloat16>(L, st);
}

}  // namespace buctd
