// Chain tap-GEMM: up to BUCTD_MAX_CHAIN FLAT tap-GEMMs of IDENTICAL geometry (the eight 3x3 convolutions of the four
// BasicBlocks of one HRNet branch) in ONE persistent launch.
//
// Why: a slab tap-GEMM launch costs ~8 us around its ~29 us of tensor work at B = 64 (inter-kernel gap ~4 us even with
// programmatic dependent launch -- a CTA of the next kernel needs the whole SM -- plus prologue, pipeline fill and
// drain; tools/conv_sweep.py).  Here the rings (slabs, weights, TMEM accumulators, stage buffers) keep
// running across layers, and there is NO grid-wide barrier between them: the dependency of a 3x3 convolution is local,
// super-tile s of layer l+1 reads rows of the super-tiles s-1, s, s+1 of layer l (halo W+2 < 128 rows) and, as residual,
// super-tile s of an earlier layer.  Every layer has the same geometry, hence the same (super-tile -> CTA) assignment.
// The DMA warp publishes flag[l][s] (+1 per N tile) once the TMA stores of that work item are complete; the producer
// thread waits for the three flags of the previous layer before it loads a slab.  CTAs drift apart freely: a slow CTA
// only delays its neighbours.  Writes are safe by the same argument: whoever overwrites rows of super-tile s at layer
// l+1 has waited for every layer-l reader of those rows (buffers rotate as in a BasicBlock: in -> tmp -> out).
// All CTAs are co-resident by construction (grid <= SM count, one CTA per SM) and dependencies only point to lower
// layers, so the spins cannot deadlock; they trap after ~4 s instead of hanging the device.
//
// Same roles and pipeline as tapgemm_slab_kernel (tapgemm.cu); staged (TMA store) epilogue only.
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
};

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// wait until super-tiles [s_lo, s_hi] of layer `row` are complete; then order this thread's TMA (async proxy) reads after it
__device__ __forceinline__ void flags_wait(const unsigned int* row, int s_lo, int s_hi, unsigned int target) {
  const long long t0 = clock64();
  for (;;) {
    bool ok = true;
    for (int s = s_lo; s <= s_hi; ++s) ok = ok && ld_acquire(row + s) >= target;
    if (ok) break;
    __nanosleep(32);
    if (clock64() - t0 > (8ll << 30)) __trap();      // ~4 s: a CTA of this grid never became resident
  }
  fence_proxy_async_all();
}
// the TMA stores of one work item are complete: publish
__device__ __forceinline__ void flag_publish(unsigned int* f) {
  fence_proxy_async_all();
  __threadfence();
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(f) : "memory");
}
template <int N> __device__ __forceinline__ void tma_store_wait_complete() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

template <typename T, int BN, int MSUB>
__global__ void __launch_bounds__(kSlabThreads, 1)
tapgemm_chain_kernel(const __grid_constant__ ChainParams cp, const TapGemmKParams p0, const SlabParams sp) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kWTileBytes = BN * 128;
  constexpr int NACC = slab_nacc<BN>();
  static_assert(MSUB <= NACC, "a super-tile needs MSUB accumulators at once");
  uint8_t* slabs = smem;
  uint8_t* wbuf = slabs + (size_t)sp.sa * sp.slab_bytes;
  uint8_t* stage = wbuf + (size_t)sp.w_tiles * kWTileBytes;
  float* s_bias = reinterpret_cast<float*>(stage + (size_t)sp.stage_bufs * epi_stage_bytes<BN>());   // [2][BN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + 2 * BN);
  uint64_t* slab_full = bars;
  uint64_t* slab_empty = slab_full + 4;
  uint64_t* w_full = slab_empty + 4;
  uint64_t* w_empty = w_full + 8;
  uint64_t* acc_full = w_empty + 8;
  uint64_t* acc_empty = acc_full + kMaxAcc;
  uint64_t* st_ready = acc_empty + kMaxAcc;
  uint64_t* st_done = st_ready + 2;
  uint64_t* w_free = st_done + 2;             // resident weights: the MMAs of the layer have read them
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_free + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool split = sp.stage_bufs == 1;
  const int NL = cp.n_layers;

  if (warp == 0 && lane == 0) {
    for (int l = 0; l < NL; ++l) {
      tma_prefetch_desc(&cp.m[l].a);
      tma_prefetch_desc(&cp.m[l].b);
    }
    for (int s = 0; s < 4; ++s) { mbar_init(&slab_full[s], 1); mbar_init(&slab_empty[s], 1); }
    for (int s = 0; s < 8; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int s = 0; s < kMaxAcc; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], split ? 8 : 4); }
    for (int s = 0; s < 2; ++s) { mbar_init(&st_ready[s], 1); mbar_init(&st_done[s], split ? 8 : 4); }
    mbar_init(w_free, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, sp.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;
  constexpr int R = MSUB * kBlockM;
  const int my_items = (sp.num_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total_u = my_items * MSUB;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t slot = 0, slot_ph = 1, ws = 0, ws_ph = 1;
      for (int l = 0; l < NL; ++l) {
        const CUtensorMap* map_a = &cp.m[l].a;
        const CUtensorMap* map_b = &cp.m[l].b;
        if (sp.w_resident) {
          if (l > 0) mbar_wait(w_free, (uint32_t)(l - 1) & 1u);
          mbar_expect_tx(&w_full[0], (uint32_t)(sp.w_tiles * kWTileBytes));
          for (int kc = 0; kc < p0.num_kc; ++kc)
            for (int t = 0; t < p0.ntaps; ++t)
              tma_load_2d(wbuf + (size_t)(kc * p0.ntaps + t) * kWTileBytes, map_b, &w_full[0], kc * 64, t * p0.n_alloc);
        }
        SubIter it;
        it.init(sp.num_super);
        for (int k = 0; k < my_items; ++k, it.next_item()) {
          if (l > 0) {      // the rows this slab covers: super-tiles s-1 .. s+1 of the previous layer (s alone without halo)
            const int reach = (sp.halo + R - 1) / R;
            flags_wait(cp.flags + (size_t)(l - 1) * sp.num_super, max(it.s - reach, 0), min(it.s + reach, sp.num_super - 1),
                       (unsigned)cp.n_tiles);
          }
          const long long r0 = (long long)it.s * R - sp.halo;
          for (int kc = 0; kc < p0.num_kc; ++kc) {
            mbar_wait(&slab_empty[slot], slot_ph);
            mbar_expect_tx(&slab_full[slot], (uint32_t)sp.slab_bytes);
            for (int i = 0; i < sp.nboxes; ++i)
              tma_load_2d(slabs + (size_t)slot * sp.slab_bytes + (size_t)i * sp.box_rows * 128, map_a, &slab_full[slot], kc * 64,
                          (int)(r0 + (long long)i * sp.box_rows));
            if (++slot == (uint32_t)sp.sa) { slot = 0; slot_ph ^= 1; }
            if (!sp.w_resident) {
              for (int t = 0; t < p0.ntaps; ++t) {
                mbar_wait(&w_empty[ws], ws_ph);
                mbar_expect_tx(&w_full[ws], kWTileBytes);
                tma_load_2d(wbuf + (size_t)ws * kWTileBytes, map_b, &w_full[ws], kc * 64, t * p0.n_alloc + it.nt * BN);
                if (++ws == (uint32_t)sp.sw) { ws = 0; ws_ph ^= 1; }
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_f16(T16<T>::kFmt, kBlockM, BN);
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
          uint32_t d[MSUB], dslot[MSUB];
#pragma unroll
          for (int j = 0; j < MSUB; ++j) {
            mbar_wait(&acc_empty[aslot], aph);
            dslot[j] = aslot;
            d[j] = tmem_base + aslot * acc_stride<BN>();
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
                  mma_ksteps_rt<BN>(ksteps, d[j], adesc0 + (uint64_t)(j * (kBlockM * 128 / 16)), bdesc, idesc, first);
                if (!sp.w_resident) {
                  tc_commit(&w_empty[ws]);
                  if (++ws == (uint32_t)sp.sw) { ws = 0; ws_ph ^= 1; }
                }
              }
            }
            tc_commit(&slab_empty[slot]);
            if (++slot == (uint32_t)sp.sa) { slot = 0; slot_ph ^= 1; }
          }
#pragma unroll
          for (int j = 0; j < MSUB; ++j) tc_commit(&acc_full[dslot[j]]);
        }
        if (sp.w_resident) tc_commit(w_free);
      }
    }
    __syncwarp();
  } else {
    const int SE = sp.stage_bufs;
    // every role numbers the sub-tiles of the WHOLE chain: ug = l * total_u + u.  Stage buffer, its barrier phase and
    // the epilogue group of a sub-tile are functions of ug, so the rings simply keep turning across layers.
    if (warp == 10) {
      if (elect_one()) {
        int ubase = 0;
        int pending_s = -1, pending2_s = -1;    // super-tiles whose stores are committed but not yet published
        for (int l = 0; l < NL; ++l, ubase += total_u) {
          TapGemmKParams p = p0;
          p.residual = cp.has_res[l] ? reinterpret_cast<const void*>(1) : nullptr;      // epi_load_residual only reads geometry
          const bool has_res = cp.has_res[l] != 0;
          const CUtensorMap* map_out = &cp.m[l].out;
          const CUtensorMap* map_res = &cp.m[l].res;
          // the residual tile of super-tile s is an earlier layer's output of super-tile s: written, completed and
          // published by THIS thread (same super-tile -> CTA assignment in every layer)
          SubIter st_it, pr_it;
          st_it.init(sp.num_super);
          pr_it.init(sp.num_super);
          auto prepare = [&](int u) {
            const int ug = ubase + u;
            const int e = SE == 2 ? (ug & 1) : 0;
            if (has_res) {
              EpiTile t;
              t.m0 = pr_it.template m0<MSUB>();
              t.tb = t.ty0 = t.tx0 = 0;
              epi_load_residual<BN>(p, map_res, t, pr_it.nt * BN, stage + (size_t)e * epi_stage_bytes<BN>(), &st_ready[e]);
            } else {
              mbar_arrive(&st_ready[e]);
            }
            pr_it.template next<MSUB>();
          };
          for (int u = 0; u < SE && u < total_u; ++u) prepare(u);
          for (int u = 0; u < total_u; ++u) {
            const int ug = ubase + u;
            const int e = SE == 2 ? (ug & 1) : 0;
            mbar_wait(&st_done[e], (uint32_t)(SE == 2 ? (ug >> 1) : ug) & 1u);
            EpiTile t;
            t.m0 = st_it.template m0<MSUB>();
            t.tb = t.ty0 = t.tx0 = 0;
            epi_store<BN>(p, map_out, t, st_it.nt * BN, stage + (size_t)e * epi_stage_bytes<BN>());
            const int s_now = st_it.s;
            const bool item_end = st_it.j == MSUB - 1;
            st_it.template next<MSUB>();
            if (item_end && l + 1 < NL) {
              // publish the item before the previous one: its MSUB store groups precede the 2 MSUB most recent ones, which
              // stay in flight (waiting for the newest stores would stall this thread, and with it the stage buffers)
              if (pending2_s >= 0) {
                tma_store_wait_complete<2 * MSUB>();
                flag_publish(cp.flags + (size_t)l * sp.num_super + pending2_s);
              }
              pending2_s = pending_s;
              pending_s = s_now;
            }
            if (u + SE < total_u) {
              tma_store_wait_read<0>();
              prepare(u + SE);
            }
          }
          tma_store_wait_all();                 // this CTA's layer-l tiles are in global memory
          if (pending2_s >= 0) flag_publish(cp.flags + (size_t)l * sp.num_super + pending2_s);
          if (pending_s >= 0) flag_publish(cp.flags + (size_t)l * sp.num_super + pending_s);
          pending_s = pending2_s = -1;
        }
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
      int ubase = 0;
      for (int l = 0; l < NL; ++l, ubase += total_u) {
        TapGemmKParams p = p0;
        p.relu = cp.relu[l];
        p.bias = cp.bias[l];
        const bool has_res = cp.has_res[l] != 0;
        SubIter it;
        it.init(sp.num_super);
        int u = ug - ubase;                     // 0 or 1 at a layer start
        if (u == 1) it.template next<MSUB>();
        int cur_nt = -1;
        for (; u < total_u; u += step, ug += step) {
          const int nt = it.nt;
          EpiTile tile;
          tile.m0 = it.template m0<MSUB>();
          tile.tb = tile.ty0 = tile.tx0 = 0;
          if (nt != cur_nt) {
            if (split) { epi_sync<256>(); epi_load_bias<BN, 256>(p, nt * BN, my_bias, gtid); epi_sync<256>(); }
            else { epi_sync_grp(grp); epi_load_bias<BN, 128>(p, nt * BN, my_bias, gtid); epi_sync_grp(grp); }
            cur_nt = nt;
          }
          mbar_wait(&acc_full[aslot], aph);
          tc_fence_after();
          const uint32_t acc = tmem_base + aslot * acc_stride<BN>();
          const int e = split ? 0 : grp;
          uint8_t* st = stage + (size_t)e * epi_stage_bytes<BN>();
          mbar_wait(&st_ready[e], (uint32_t)(split ? ug : (ug >> 1)) & 1u);
          epi_compute<T, BN>(p, tile, nt * BN, acc, st, my_bias, has_res, q, lane, split ? grp : 0, split ? 2 : 1);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&st_done[e]);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[aslot]);
          it.template next<MSUB>();
          if (step == 2) it.template next<MSUB>();
          aslot += step;
          if (aslot >= NACC) { aslot -= NACC; aph ^= 1; }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, sp.tmem_cols);
}

// ---- host side --------------------------------------------------------------------------------
struct ChainLaunch {
  ChainParams cp;
  TapGemmKParams kp;
  SlabParams sp;
  dim3 grid;
  int bn, dtype;
};

template <typename T, int BN, int MSUB>
static int chain_set_attr() {
  if constexpr (MSUB <= slab_nacc<BN>()) {
    BUCTD_CUDA(cudaFuncSetAttribute(tapgemm_chain_kernel<T, BN, MSUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
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

size_t tapgemm_chain_flag_count(const ChainLaunch& C) { return (size_t)C.cp.n_layers * C.sp.num_super; }

int tapgemm_chain_prepare(const buctd_tapgemm_chain_args& a, ChainLaunch* C, int sm_count, unsigned int* counter) {
  BUCTD_CHECK(a.n_layers >= 1 && a.n_layers <= BUCTD_MAX_CHAIN, BUCTD_ERR_INVALID, "tapgemm_chain: n_layers %d out of range [1,%d]",
              a.n_layers, BUCTD_MAX_CHAIN);
  BUCTD_CHECK(counter != nullptr, BUCTD_ERR_INVALID, "tapgemm_chain: no barrier counter");
  const buctd_tapgemm_args& f = a.layer[0];
  memset(&C->cp, 0, sizeof(C->cp));
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
    } else {
      BUCTD_CHECK(L.bn == C->bn && L.sp.msub == C->sp.msub && L.sp.sa == C->sp.sa && L.sp.sw == C->sp.sw &&
                      L.sp.stage_bufs == C->sp.stage_bufs && L.sp.smem_bytes == C->sp.smem_bytes,
                  BUCTD_ERR_INVALID, "tapgemm_chain: layer %d planned differently from layer 0", l);
    }
    BUCTD_CHECK(L.sp.direct == 0, BUCTD_ERR_INVALID, "tapgemm_chain: direct epilogue is not supported in chains (unset BUCTD_DIRECT_MAX_BN)");
    C->cp.m[l].a = L.map_a; C->cp.m[l].b = L.map_b; C->cp.m[l].out = L.map_out; C->cp.m[l].res = L.map_res;
    C->cp.bias[l] = g.bias;
    C->cp.has_res[l] = g.residual != nullptr;
    C->cp.relu[l] = g.relu;
    if (g.residual == nullptr) C->cp.m[l].res = L.map_out;
  }
  C->kp.bias = nullptr; C->kp.residual = nullptr; C->kp.out = nullptr; C->kp.relu = 0;
  C->cp.flags = counter;
  C->cp.n_layers = a.n_layers;
  C->cp.n_tiles = C->sp.num_items / C->sp.num_super;
  return BUCTD_OK;
}

template <typename T, int BN, int MSUB>
static void chain_launch_one(const ChainLaunch& C, cudaStream_t st) {
  if constexpr (MSUB <= slab_nacc<BN>()) {
    launch_pdl(tapgemm_chain_kernel<T, BN, MSUB>, C.grid, dim3(kSlabThreads), C.sp.smem_bytes, st, C.cp, C.kp, C.sp);
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
