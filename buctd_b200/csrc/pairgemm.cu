// pairgemm: out[M][N] = act(A[M][K] W[N][K]^T + bias[N]) for the LARGE nn.Linear layers of the path -- the channel
// attention's fc_o, Linear(H*W, H*W) applied to the [B*C, H*W] planar matrix (lib/models/self_attention.py:147-160 via
// lib/models/pose_hrnet_coam.py:700-716: 6912 x 6912 on branch 0 of W48 at 384x288, 293 GFLOP per 64 crops) -- on a CTA
// PAIR: `tcgen05.mma.cta_group::2` with M = 256 (128 rows of A per CTA), N = 256 (128 rows of W per CTA).
//
// Why a pair.  As a one-CTA 128 x 256 tile (the slab kernel's flat plan) this product moves 5.3 MB of operands per
// tile through L2 -> shared memory for 30 us of tensor math: 3.4 GB per launch, ~7.6 TB/s, i.e. the launch runs at the
// L2 fabric's speed (454 us, 646 TFLOP/s), not the tensor core's.  In a pair each CTA loads HALF of the W tile and the
// tensor cores of both SMs read both halves: 3.5 MB per CTA and 128 x 256 outputs (2/3 of the L2 traffic), and each
// MMA's shared-memory operand fetch drops from 128 + 256 to 128 + 128 rows of 128 B.
//
// Layout per CTA: warp 0 = TMA producer (its A half and its W half of every 64-wide K chunk into a 5-stage ring;
// `cp.async.bulk.tensor...cta_group::2` completes on the LEADER's full barrier), warp 1 = TMEM allocator and, in the
// leader CTA (cluster rank 0), the one thread that issues every MMA of the pair and commits with
// `.multicast::cluster` to the barriers of both CTAs; warps 2..5 = epilogue (TMEM -> + bias -> ReLU -> 16 bit ->
// SWIZZLE_128B stage -> TMA store, double-buffered 64-column blocks).  Persistent over 256 x 256 tiles, two
// accumulator slots of 256 TMEM columns: the epilogue of tile i overlaps the K loop of tile i + 1.
#include "common.cuh"
#include "epilogue.cuh"
#include "internal.h"

namespace buctd {

constexpr int kPgThreads = 192;
constexpr int kPgStages = 5;
constexpr int kPgHalfBytes = 128 * 128;                 // one operand half of a stage: 128 rows x 64 k x 2 B
constexpr int kPgStageBytes = 2 * kPgHalfBytes;        // A half + W half
constexpr int kPgSmemBytes = 1024 + kPgStages * kPgStageBytes + 2 * kEpiBlockBytes + 256 * 4 + 256;

__device__ __forceinline__ void pg_tmem_ld32(uint32_t taddr, uint32_t* r) {
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
__global__ void __launch_bounds__(kPgThreads, 1)
pairgemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                const __grid_constant__ CUtensorMap map_out, const PairGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;                                                    // [stages][A half | W half]
  uint8_t* stage = ring + kPgStages * kPgStageBytes;                       // [2][128 rows x 128 B] output blocks
  float* s_bias = reinterpret_cast<float*>(stage + 2 * kEpiBlockBytes);    // [256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + 256);
  uint64_t* full_bar = bars;                   // [stages]  leader only: both CTAs' halves of the stage have landed
  uint64_t* empty_bar = full_bar + 8;          // [stages]  per CTA: the pair's MMAs have read the stage
  uint64_t* acc_full = empty_bar + 8;          // [2]       per CTA: the accumulator slot holds a finished tile
  uint64_t* acc_empty = acc_full + 2;          // [2]       leader only: both CTAs' epilogues have read the slot
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    tma_prefetch_desc(&map_out);
    for (int i = 0; i < kPgStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }      // 4 epilogue warps x 2 CTAs
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();                      // the allocator's shared-memory write, CTA-wide (also what racecheck understands)
  cluster_sync_all();                   // barriers of both CTAs are initialised, TMEM of both is allocated
  tc_fence_after();
  pdl_launch_dependents();
  const uint32_t tmem_base = *tmem_slot;
  const int cluster_id = (int)blockIdx.x >> 1, nclusters = (int)gridDim.x >> 1;

  if (warp == 0) {
    if (elect_one()) {
      pdl_wait();                                   // A is produced by the previous kernel of the stream
      const uint32_t full_addr0 = mapa_u32(&full_bar[0], 0);       // the leader's barriers (consecutive 8-byte words)
      uint32_t slot = 0, ph = 1;
      for (int t = cluster_id; t < p.num_tiles; t += nclusters) {
        const int nt = t / p.m_tiles, mt = t - nt * p.m_tiles;
        const int am = mt * 256 + (int)rank * 128;          // this CTA's 128 rows of A
        const int bn = nt * 256 + (int)rank * 128;          // this CTA's 128 rows of W (columns of the output)
        for (int kc = 0; kc < p.num_kc; ++kc) {
          mbar_wait(&empty_bar[slot], ph);
          if (leader) mbar_expect_tx(&full_bar[slot], 2 * kPgStageBytes);
          uint8_t* sa = ring + (size_t)slot * kPgStageBytes;
          tma_load_2d_pair(sa, &map_a, full_addr0 + slot * 8, kc * 64, am);
          tma_load_2d_pair(sa + kPgHalfBytes, &map_b, full_addr0 + slot * 8, kc * 64, bn);
          if (++slot == kPgStages) { slot = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (leader && elect_one()) {
      const uint32_t idesc = umma_idesc_f16(T16<T>::kFmt, 256, 256);
      const uint32_t ring_addr = smem_u32(ring);
      uint32_t slot = 0, ph = 0, as = 0, aph = 1;
      for (int t = cluster_id; t < p.num_tiles; t += nclusters) {
        mbar_wait_cluster(&acc_empty[as], aph);
        tc_fence_after();
        const uint32_t d = tmem_base + as * 256;
        for (int kc = 0; kc < p.num_kc; ++kc) {
          mbar_wait(&full_bar[slot], ph);
          tc_fence_after();
          const uint32_t sa = ring_addr + slot * kPgStageBytes;
          const uint64_t ad = umma_desc_sw128(sa), bd = umma_desc_sw128(sa + kPgHalfBytes);
          const int ksteps = min(64, p.k16 - kc * 64) >> 4;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            if (ks < ksteps) tc_mma_f16_pair(d, ad + (uint64_t)(2 * ks), bd + (uint64_t)(2 * ks), idesc, (uint32_t)((kc | ks) != 0));
          tc_commit_pair(&empty_bar[slot]);
          if (++slot == kPgStages) { slot = 0; ph ^= 1; }
        }
        tc_commit_pair(&acc_full[as]);
        if (++as == 2) { as = 0; aph ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // epilogue: thread = output row (TMEM lane) of this CTA's half of the tile
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int etid = (int)threadIdx.x - 64;
    const uint32_t acc_empty_addr0 = mapa_u32(&acc_empty[0], 0), acc_empty_addr1 = mapa_u32(&acc_empty[1], 0);
    uint32_t as = 0, aph = 0;
    int nstore = 0;                                  // stores issued so far by the store thread (stage buffer = nstore & 1)
    for (int t = cluster_id; t < p.num_tiles; t += nclusters) {
      const int nt = t / p.m_tiles, mt = t - nt * p.m_tiles;
      const int m0 = mt * 256 + (int)rank * 128, n0 = nt * 256;
      epi_sync<128>();                               // the previous tile's readers of s_bias are done
      for (int i = etid; i < 256; i += 128) s_bias[i] = (p.bias != nullptr && n0 + i < p.n) ? p.bias[n0 + i] : 0.f;
      mbar_wait(&acc_full[as], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * 256;
      const int nblk = min(4, (p.n_store - n0 + 63) >> 6);
      for (int blk = 0; blk < 4; ++blk) {
        uint32_t v[64];
        pg_tmem_ld32(taddr + blk * 64, v);
        pg_tmem_ld32(taddr + blk * 64 + 32, v + 32);
        tmem_ld_wait();
        if (blk == 3) {                              // the accumulator slot is free once all 4 warps of both CTAs got here
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(as == 0 ? acc_empty_addr0 : acc_empty_addr1);
        }
        if (blk >= nblk) continue;                   // columns past the output's width (uniform over the CTA)
        uint8_t* sb = stage + (size_t)(nstore & 1) * kEpiBlockBytes;
        if (etid == 0) tma_store_wait_read<1>();     // the store that last read this buffer (two blocks ago) is done with it
        epi_sync<128>();
        const float* bb = s_bias + blk * 64;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            f[i] = __uint_as_float(v[c * 8 + i]) + bb[c * 8 + i];
            if (p.relu) f[i] = fmaxf(f[i], 0.f);
          }
          uint4 u;
          u.x = T16<T>::pack2(f[0], f[1]); u.y = T16<T>::pack2(f[2], f[3]);
          u.z = T16<T>::pack2(f[4], f[5]); u.w = T16<T>::pack2(f[6], f[7]);
          *reinterpret_cast<uint4*>(sb + row * 128 + ((c ^ (row & 7)) << 4)) = u;      // SWIZZLE_128B
        }
        fence_proxy_async_smem();
        epi_sync<128>();
        if (etid == 0) {
          tma_store_2d(&map_out, sb, n0 + blk * 64, m0);
          tma_store_commit();
        }
        ++nstore;
      }
      if (++as == 2) { as = 0; aph ^= 1; }
    }
    if (etid == 0) tma_store_wait_all();
  }
  tc_fence_before();
  cluster_sync_all();                   // no CTA of the pair leaves (or frees TMEM) while the other may still touch it
  if (warp == 1) tmem_dealloc_pair(tmem_base, 512);
}

bool pairgemm_wanted(const buctd_tapgemm_args& a, int num_kc) {
  const char* e = getenv("BUCTD_PAIRGEMM");          // read at prepare time: 0 = the one-CTA slab plan (tests, A/B timing)
  if (e != nullptr && atoi(e) == 0) return false;
  return a.mode == BUCTD_TG_FLAT && a.ntaps == 1 && a.out_mode == BUCTD_OUT_ROWS16 && a.residual == nullptr && a.n_alloc >= 256 &&
         num_kc >= 4 && a.m_rows >= 256;
}

template <typename T>
static int pairgemm_max_clusters(int* out) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2); cfg.blockDim = dim3(kPgThreads); cfg.dynamicSmemBytes = kPgSmemBytes;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int n = 0;
  BUCTD_CUDA(cudaOccupancyMaxActiveClusters(&n, pairgemm_kernel<T>, &cfg));
  *out = n;
  return BUCTD_OK;
}

int pair_max_clusters(int* out) {
  static int cached = 0;
  if (cached == 0) {
    int rc = pairgemm_max_clusters<__half>(&cached);
    if (rc) return rc;
    BUCTD_CHECK(cached > 0, BUCTD_ERR_CUDA, "no CTA pair fits this device");
  }
  *out = cached;
  return BUCTD_OK;
}

int pairgemm_init_device() {
  BUCTD_CUDA(cudaFuncSetAttribute(pairgemm_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPgSmemBytes));
  BUCTD_CUDA(cudaFuncSetAttribute(pairgemm_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPgSmemBytes));
  return BUCTD_OK;
}

// `a` has passed tapgemm_prepare's checks; ctas = the SMs this op may use
int pairgemm_prepare(const buctd_tapgemm_args& a, TapGemmLaunch* L, int ctas) {
  const int k16 = (a.k + 15) / 16 * 16;
  PairGemmParams& p = L->pp;
  memset(&p, 0, sizeof(p));
  p.k16 = k16; p.num_kc = (k16 + 63) / 64;
  p.n = a.n;
  p.n_store = a.out_cs < a.n_alloc ? a.out_cs : a.n_alloc;       // columns that exist in the output rows
  p.relu = a.relu; p.bias = a.bias;
  p.m_tiles = (int)((a.m_rows + 255) / 256);
  const int n_tiles = (p.n_store + 255) / 256;
  p.num_tiles = p.m_tiles * n_tiles;
  const int inner = a.a_cs < k16 ? a.a_cs : k16;
  int rc = tapgemm_encode_rows_map(&L->map_a, a.dtype, a.a, inner, a.a_cs, a.m_rows, 128);
  if (rc) return rc;
  rc = tapgemm_encode_rows_map(&L->map_b, a.dtype, a.w, a.w_cs, a.w_cs, a.n_alloc, 128);
  if (rc) return rc;
  rc = tapgemm_encode_rows_map(&L->map_out, a.dtype, a.out, a.out_cs, a.out_cs, a.m_rows, 128);
  if (rc) return rc;
  int mc = 0;
  if ((rc = pair_max_clusters(&mc))) return rc;
  int clusters = ctas / 2 < mc ? ctas / 2 : mc;
  if (clusters < 1) clusters = 1;
  if (clusters > p.num_tiles) clusters = p.num_tiles;
  L->grid = dim3((unsigned)(2 * clusters));
  L->pair = 1;
  L->dtype = a.dtype;
  return BUCTD_OK;
}

int pairgemm_launch(const TapGemmLaunch& L, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = L.grid; cfg.blockDim = dim3(kPgThreads); cfg.dynamicSmemBytes = kPgSmemBytes; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  static const bool no_pdl = getenv("BUCTD_NO_PDL") != nullptr;
  cfg.attrs = attr; cfg.numAttrs = (tl_use_pdl && !no_pdl) ? 2 : 1;
  cudaError_t e;
  if (L.dtype == BUCTD_F16) e = cudaLaunchKernelEx(&cfg, pairgemm_kernel<__half>, L.map_a, L.map_b, L.map_out, L.pp);
  else e = cudaLaunchKernelEx(&cfg, pairgemm_kernel<__nv_bfloat16>, L.map_a, L.map_b, L.map_out, L.pp);
  BUCTD_CHECK(e == cudaSuccess, BUCTD_ERR_CUDA, "pairgemm launch failed: %s", cudaGetErrorString(e));
  return BUCTD_OK;
}

}  // namespace buctd
