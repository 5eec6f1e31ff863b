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

namespace buctd {

constexpr int kBlockM = 128;
constexpr int kATileBytes = kBlockM * 128;
constexpr int kThreads = 192;
constexpr int kSmemBudget = 227 * 1024;

template <int BN> __host__ __device__ constexpr int patch_stages() { return BN <= 128 ? 4 : 3; }

// =================================================================================================
// PATCH kernel (stride-2 3x3): one tile per CTA
// =================================================================================================
template <typename T, int BN>
__global__ void __launch_bounds__(kThreads)
tapgemm_patch_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                     const TapGemmKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int S = patch_stages<BN>();
  constexpr int kBTileBytes = BN * 128;
  constexpr int kTmemCols = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + S * kATileBytes;
  uint8_t* stage = smem_b + S * kBTileBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stage + epi_stage_bytes<BN>());
  uint64_t* empty_bar = full_bar + S;
  uint64_t* tmem_full_bar = empty_bar + S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * BN;

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
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int iters = p.ntaps * p.num_kc;

  if (warp == 0) {
    if (lane == 0) {
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
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(T16<T>::kFmt, kBlockM, BN);
      for (int it = 0; it < iters; ++it) {
        const int s = it % S;
        const int kc = it % p.num_kc;
        mbar_wait(&full_bar[s], (it / S) & 1);
        tc_fence_after();
        const uint64_t adesc = umma_desc_sw128(smem_u32(smem_a + s * kATileBytes));
        const uint64_t bdesc = umma_desc_sw128(smem_u32(smem_b + s * kBTileBytes));
        const int ksteps = min(64, p.k16 - kc * 64) >> 4;
        for (int k = 0; k < ksteps; ++k)
          tc_mma_f16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (it | k) != 0);
        tc_commit(&empty_bar[s]);
      }
      tc_commit(tmem_full_bar);
    }
  } else {
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    epilogue_tile<T, BN>(p, tile, n0, tmem_base, stage, threadIdx.x - 64, warp & 3, lane);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// =================================================================================================
// SLAB kernel (FLAT mode): persistent, A-slab + row-shifted tap views, double-buffered accumulators
// =================================================================================================
template <typename T, int BN>
__global__ void __launch_bounds__(kThreads, 1)
tapgemm_slab_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const TapGemmKParams p, const SlabParams sp) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kWTileBytes = BN * 128;
  uint8_t* slabs = smem;                                         // [sa][slab_bytes]
  uint8_t* wbuf = slabs + (size_t)sp.sa * sp.slab_bytes;         // resident: [num_kc*ntaps] tiles; ring: [sw] tiles
  uint8_t* stage = wbuf + (size_t)sp.w_tiles * kWTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage + epi_stage_bytes<BN>());
  uint64_t* slab_full = bars;                 // [sa]
  uint64_t* slab_empty = slab_full + 4;       // [sa]
  uint64_t* w_full = slab_empty + 4;          // [sw] (resident: [0] only)
  uint64_t* w_empty = w_full + 8;             // [sw]
  uint64_t* acc_full = w_empty + 8;           // [2]
  uint64_t* acc_empty = acc_full + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int s = 0; s < 4; ++s) { mbar_init(&slab_full[s], 1); mbar_init(&slab_empty[s], 1); }
    for (int s = 0; s < 8; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, sp.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int R = sp.msub * kBlockM;

  if (warp == 0) {
    if (lane == 0) {
      if (sp.w_resident) {
        // all weight tiles of this CTA's N tile... resident mode is only chosen when there is one N tile
        mbar_expect_tx(&w_full[0], (uint32_t)(sp.w_tiles * kWTileBytes));
        for (int kc = 0; kc < p.num_kc; ++kc)
          for (int t = 0; t < p.ntaps; ++t)
            tma_load_2d(wbuf + (size_t)(kc * p.ntaps + t) * kWTileBytes, &map_b, &w_full[0], kc * 64, t * p.n_alloc);
      }
      uint32_t sa_it = 0, sw_it = 0;
      for (int item = blockIdx.x; item < sp.num_items; item += gridDim.x) {
        const int s = item % sp.num_super, nt = item / sp.num_super;
        const long long r0 = (long long)s * R - sp.halo;
        for (int kc = 0; kc < p.num_kc; ++kc) {
          const uint32_t slot = sa_it % sp.sa;
          mbar_wait(&slab_empty[slot], ((sa_it / sp.sa) & 1) ^ 1);
          mbar_expect_tx(&slab_full[slot], (uint32_t)sp.slab_bytes);
          for (int i = 0; i < sp.nboxes; ++i)
            tma_load_2d(slabs + (size_t)slot * sp.slab_bytes + (size_t)i * sp.box_rows * 128, &map_a, &slab_full[slot],
                        kc * 64, (int)(r0 + (long long)i * sp.box_rows));
          ++sa_it;
          if (!sp.w_resident) {
            for (int t = 0; t < p.ntaps; ++t) {
              const uint32_t ws = sw_it % sp.sw;
              mbar_wait(&w_empty[ws], ((sw_it / sp.sw) & 1) ^ 1);
              mbar_expect_tx(&w_full[ws], kWTileBytes);
              tma_load_2d(wbuf + (size_t)ws * kWTileBytes, &map_b, &w_full[ws], kc * 64, t * p.n_alloc + nt * BN);
              ++sw_it;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(T16<T>::kFmt, kBlockM, BN);
      uint32_t sa_it = 0, sw_it = 0, it_local = 0;
      if (sp.w_resident) { mbar_wait(&w_full[0], 0); }
      for (int item = blockIdx.x; item < sp.num_items; item += gridDim.x, ++it_local) {
        const uint32_t buf = it_local & 1;
        mbar_wait(&acc_empty[buf], ((it_local >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int kc = 0; kc < p.num_kc; ++kc) {
          const uint32_t slot = sa_it % sp.sa;
          mbar_wait(&slab_full[slot], (sa_it / sp.sa) & 1);
          tc_fence_after();
          const uint32_t slab_addr = smem_u32(slabs + (size_t)slot * sp.slab_bytes);
          const int ksteps = min(64, p.k16 - kc * 64) >> 4;
          for (int t = 0; t < p.ntaps; ++t) {
            uint32_t w_addr;
            uint32_t ws = 0;
            if (sp.w_resident) {
              w_addr = smem_u32(wbuf + (size_t)(kc * p.ntaps + t) * kWTileBytes);
            } else {
              ws = sw_it % sp.sw;
              mbar_wait(&w_full[ws], (sw_it / sp.sw) & 1);
              tc_fence_after();
              w_addr = smem_u32(wbuf + (size_t)ws * kWTileBytes);
            }
            const uint64_t bdesc = umma_desc_sw128(w_addr);
            int shift = sp.halo;
            if (p.ntaps == 9) shift += (t / 3 - 1) * p.wp + (t % 3 - 1);
            for (int j = 0; j < sp.msub; ++j) {
              const uint32_t row_off = (uint32_t)(shift + j * kBlockM);
              uint64_t adesc = umma_desc_sw128(slab_addr + row_off * 128);
              if (sp.desc_mode == 1) adesc |= (uint64_t)(row_off & 7) << 49;
              const uint32_t d = tmem_base + (buf * sp.msub + j) * BN;
              for (int k = 0; k < ksteps; ++k)
                tc_mma_f16(d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kc | t | k) != 0);
            }
            if (!sp.w_resident) { tc_commit(&w_empty[ws]); ++sw_it; }
          }
          tc_commit(&slab_empty[slot]);
          ++sa_it;
        }
        tc_commit(&acc_full[buf]);
      }
    }
  } else {
    const int etid = threadIdx.x - 64;
    const int q = warp & 3;
    uint32_t it_local = 0;
    for (int item = blockIdx.x; item < sp.num_items; item += gridDim.x, ++it_local) {
      const int s = item % sp.num_super, nt = item / sp.num_super;
      const uint32_t buf = it_local & 1;
      mbar_wait(&acc_full[buf], (it_local >> 1) & 1);
      tc_fence_after();
      for (int j = 0; j < sp.msub; ++j) {
        EpiTile tile;
        tile.m0 = (long long)s * R + (long long)j * kBlockM;
        tile.tb = tile.ty0 = tile.tx0 = 0;
        if (tile.m0 >= p.m_rows) break;                                    // uniform
        epilogue_tile<T, BN>(p, tile, nt * BN, tmem_base + (buf * sp.msub + j) * BN, stage, etid, q, lane);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, sp.tmem_cols);
}

// ---- host side --------------------------------------------------------------------------------
static size_t patch_smem_bytes(int bn) {
  const int s = bn <= 128 ? 4 : 3;
  return 1024 + (size_t)s * (kATileBytes + bn * 128) + 128 * (bn * 2 + 16) + 256;
}

template <typename T, int BN>
static int set_attr() {
  BUCTD_CUDA(cudaFuncSetAttribute(tapgemm_patch_kernel<T, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)patch_smem_bytes(BN)));
  BUCTD_CUDA(cudaFuncSetAttribute(tapgemm_slab_kernel<T, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
  return BUCTD_OK;
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

// choose super-tile size, ring depths and the shared-memory carve-up of the slab kernel
static int plan_slab(const buctd_tapgemm_args& a, int bn, int n_tiles, int num_kc, int sm_count, SlabParams* sp) {
  const int halo = a.ntaps == 9 ? a.out_w + 2 : 0;        // (W+1) + 1 rows on each side
  const int wtile = bn * 128;
  const int fixed = 1024 + 128 * (bn * 2 + 16) + 512;     // alignment slack + epilogue stage + barriers
  const long long total_w = (long long)a.ntaps * num_kc * wtile;
  const bool resident = n_tiles == 1 && total_w <= 80 * 1024;
  bool have = false;
  SlabParams best;
  memset(&best, 0, sizeof(best));
  const char* force = getenv("BUCTD_FORCE_MSUB");         // test knob: pin the super-tile size
  const int forced = force ? atoi(force) : 0;
  for (int msub = 4; msub >= 1; --msub) {
    if (2 * msub * bn > 512) continue;                    // two accumulator sets in TMEM
    if (forced > 0 && msub > forced) continue;
    const int rows = msub * 128 + 2 * halo;
    const int nboxes = ceil_div(rows, 256);
    const int box_rows = (ceil_div(rows, nboxes) + 7) / 8 * 8;   // every box starts on a 1024-byte swizzle atom
    const long long slab = (long long)nboxes * box_rows * 128;
    int sw = resident ? 0 : 4;
    int sa = 0;
    long long need = 0;
    for (; sa == 0 && (resident || sw >= 2); --sw) {
      for (int c = (num_kc > 1 ? 3 : 2); c >= 2; --c) {
        need = fixed + c * slab + (resident ? total_w : (long long)sw * wtile);
        if (need <= kSmemBudget) { sa = c; break; }
      }
      if (resident) break;
    }
    if (sa == 0) continue;
    if (!resident) ++sw;                                  // undo the loop's last decrement
    SlabParams c;
    memset(&c, 0, sizeof(c));
    const long long supers = (a.m_rows + (long long)msub * 128 - 1) / ((long long)msub * 128);
    c.msub = msub; c.halo = halo; c.nboxes = nboxes; c.box_rows = box_rows; c.slab_bytes = (int)slab;
    c.sa = sa; c.sw = sw; c.w_resident = resident; c.w_tiles = resident ? a.ntaps * num_kc : sw;
    c.num_super = (int)supers; c.num_items = (int)supers * n_tiles; c.smem_bytes = (int)need;
    const int cols = 2 * msub * bn;
    c.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
    best = c;
    have = true;
    if (forced > 0 || c.num_items >= 2 * sm_count) break; // largest super-tile that still gives every SM >= 2 items
  }
  if (!have) {
    set_error("tapgemm: no shared-memory plan for bn=%d kc=%d halo=%d", bn, num_kc, halo);
    return BUCTD_ERR_INVALID;
  }
  *sp = best;
  return BUCTD_OK;
}

template <typename T, int BN>
static int launch_one(const TapGemmLaunch& L, cudaStream_t st) {
  if (L.kp.mode == BUCTD_TG_PATCH) {
    tapgemm_patch_kernel<T, BN><<<L.grid, kThreads, patch_smem_bytes(BN), st>>>(L.map_a, L.map_b, L.kp);
  } else {
    tapgemm_slab_kernel<T, BN><<<L.grid, kThreads, L.sp.smem_bytes, st>>>(L.map_a, L.map_b, L.kp, L.sp);
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
  const int bn = pick_bn(a.n_alloc);
  const int ntiles = ceil_div(a.n_alloc, bn);
  const cuuint64_t a_inner = (cuuint64_t)(a.a_cs < k16 ? a.a_cs : k16);
  CUresult cr;
  dim3 grid;
  memset(&L->sp, 0, sizeof(L->sp));
  if (a.mode == BUCTD_TG_FLAT) {
    BUCTD_CHECK(a.m_rows > 0 && a.m_rows < (1ll << 31), BUCTD_ERR_INVALID, "tapgemm: bad m_rows %lld", (long long)a.m_rows);
    if (a.out_mode != BUCTD_OUT_ROWS16 || a.ntaps == 9)
      BUCTD_CHECK((long long)a.batch * (a.out_h + 1) * (a.out_w + 1) == a.m_rows, BUCTD_ERR_INVALID,
                  "tapgemm: m_rows %lld does not match act16 geometry %dx%dx%d", (long long)a.m_rows, a.batch, a.out_h, a.out_w);
    int rc = plan_slab(a, bn, ntiles, kp.num_kc, sm_count, &L->sp);
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
    grid = dim3((unsigned)(a.batch * kp.tiles_x * kp.tiles_y), ntiles);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)a.w_cs, (cuuint64_t)a.ntaps * a.n_alloc};
    cuuint64_t strides[1] = {(cuuint64_t)a.w_cs * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)bn};
    cuuint32_t es[2] = {1, 1};
    cr = enc(&L->map_b, dt, 2, const_cast<void*>(a.w), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    BUCTD_CHECK(cr == CUDA_SUCCESS, BUCTD_ERR_CUDA, "cuTensorMapEncodeTiled(W) failed: %d", (int)cr);
  }
  L->kp = kp; L->grid = grid; L->bn = bn; L->dtype = a.dtype;
  return BUCTD_OK;
}

int tapgemm_launch(const TapGemmLaunch& L, cudaStream_t st) {
  if (L.dtype == BUCTD_F16) return dispatch_bn<__half>(L, st);
  return dispatch_bn<__nv_bfloat16>(L, st);
}

}  // namespace buctd
