// Tap-GEMM: implicit-GEMM convolutions and linear layers on the sm_100a tensor cores.
//
//   out[m, n] = act( sum_t sum_k A[m (+) tap_t, k] * W[t][n][k] + bias[n] + residual[m, n] )
//
// One CTA computes one 128 x BN output tile.  Warp 0 is the TMA producer, warp 1 allocates TMEM
// and issues tcgen05.mma (one elected lane), warps 2..5 run the epilogue (TMEM -> registers ->
// bias / residual / ReLU / pad mask -> global).  A and W tiles travel global -> shared memory
// with cp.async.bulk.tensor (SWIZZLE_128B, 64 channels = 128 B per row), the accumulator lives
// in TMEM.  The convolution is im2col-free: in FLAT mode a 3x3 tap is a row shift of the padded
// NHWC matrix (a plain 2-D TMA box at row m0 + dy*(W+1) + dx; out-of-range rows are zero-filled
// by TMA), in PATCH mode (stride 2) the tile is a th x tw output patch and the A box is a 4-D
// TMA box traversed with element stride 2.
//
// Replaces: nn.Conv2d/BatchNorm2d/ReLU/residual of lib/models/pose_hrnet_coam.py:44-60, 81-101,
// 200-242, 405-432, 565, 635-641, 666-670 and nn.Linear of lib/models/self_attention.py:74-76,87,159.
#include "common.cuh"
#include "internal.h"

namespace buctd {

constexpr int kBlockM = 128;
constexpr int kStages = 4;
constexpr int kATileBytes = kBlockM * 128;
constexpr int kThreads = 192;


template <typename T, int BN>
__global__ void __launch_bounds__(kThreads)
tapgemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const TapGemmKParams p) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kBTileBytes = BN * 128;
  constexpr int kTmemCols = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kATileBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + kStages * kBTileBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * BN;

  // tile origin
  long long m0 = 0;
  int tb = 0, ty0 = 0, tx0 = 0;
  if (p.mode == BUCTD_TG_FLAT) {
    m0 = (long long)blockIdx.x * kBlockM;
  } else {
    int t = blockIdx.x;
    int per_img = p.tiles_x * p.tiles_y;
    tb = t / per_img;
    int r = t - tb * per_img;
    ty0 = (r / p.tiles_x) * p.th;
    tx0 = (r % p.tiles_x) * p.tw;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int s = 0; s < kStages; ++s) {
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
        const int s = it % kStages;
        const int t = it / p.num_kc;
        const int kc = it - t * p.num_kc;
        mbar_wait(&empty_bar[s], ((it / kStages) & 1) ^ 1);
        mbar_expect_tx(&full_bar[s], p.a_tile_bytes + kBTileBytes);
        int dy = 0, dx = 0;
        if (p.ntaps == 9) { dy = t / 3 - 1; dx = t % 3 - 1; }
        if (p.mode == BUCTD_TG_FLAT) {
          long long row = m0 + (long long)dy * p.wp + dx;
          tma_load_2d(smem_a + s * kATileBytes, &map_a, &full_bar[s], kc * 64, (int)row);
        } else {
          tma_load_4d(smem_a + s * kATileBytes, &map_a, &full_bar[s], kc * 64, tx0 * p.stride + dx,
                      ty0 * p.stride + dy, tb);
        }
        tma_load_2d(smem_b + s * kBTileBytes, &map_b, &full_bar[s], kc * 64, t * p.n_alloc + n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(T16<T>::kFmt, kBlockM, BN);
      for (int it = 0; it < iters; ++it) {
        const int s = it % kStages;
        const int kc = it % p.num_kc;
        mbar_wait(&full_bar[s], (it / kStages) & 1);
        tc_fence_after();
        const uint64_t adesc = umma_desc_sw128(smem_u32(smem_a + s * kATileBytes));
        const uint64_t bdesc = umma_desc_sw128(smem_u32(smem_b + s * kBTileBytes));
        const int ksteps = min(64, p.k16 - kc * 64) >> 4;
        for (int k = 0; k < ksteps; ++k) {
          // advance 16 elements (32 B) along K inside the 128-byte swizzle row: +2 in the >>4 address field
          tc_mma_f16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (it | k) != 0);
        }
        tc_commit(&empty_bar[s]);   // frees the smem stage when these MMAs retire
      }
      tc_commit(tmem_full_bar);     // accumulator complete
    }
  } else {
    // ---- epilogue: warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32) -------------------------
    const int q = warp & 3;
    const int row = q * 32 + lane;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();

    const int Hp = p.H + 1, Wp = p.W + 1;
    bool valid = false, pad = false;
    long long r = 0;
    int ob = 0, oy = 0, ox = 0;
    if (p.mode == BUCTD_TG_FLAT) {
      r = m0 + row;
      valid = r < p.m_rows;
      if (p.out_mode != BUCTD_OUT_ROWS16 && valid) {
        ox = (int)(r % Wp);
        long long ry = r / Wp;
        oy = (int)(ry % Hp);
        ob = (int)(ry / Hp);
        pad = (ox == p.W) || (oy == p.H);
      }
    } else {
      int ty = row / p.tw, tx = row - ty * p.tw;
      oy = ty0 + ty; ox = tx0 + tx; ob = tb;
      valid = (row < p.tw * p.th) && oy < p.H && ox < p.W;
      r = ((long long)tb * Hp + oy) * Wp + ox;
    }
    if (p.out_mode == BUCTD_OUT_NCHW32 && pad) valid = false;

    const T* res = reinterpret_cast<const T*>(p.residual);
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      const int nb = n0 + c0;
      if (nb >= p.n && !(p.out_mode == BUCTD_OUT_ACT16 && nb < p.out_cs)) break;   // warp-uniform
      uint32_t acc[16];
      tmem_ld16(taddr + c0, acc);
      tmem_ld_wait();
      if (!valid) continue;
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        v[j] = __uint_as_float(acc[j]);
        if (p.bias != nullptr && nb + j < p.n) v[j] += __ldg(p.bias + nb + j);
      }
      if (res != nullptr && !pad) {
        const T* rp = res + r * p.res_cs + nb;
        if (nb + 16 <= p.n && (p.res_cs & 7) == 0) {
          uint4 u0 = *reinterpret_cast<const uint4*>(rp);
          uint4 u1 = *reinterpret_cast<const uint4*>(rp + 8);
          uint32_t w[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float2 f = T16<T>::unpack2(w[j]);
            v[2 * j] += f.x;
            v[2 * j + 1] += f.y;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (nb + j < p.n) v[j] += T16<T>::to_f(rp[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (p.relu) v[j] = fmaxf(v[j], 0.f);
        if (pad || nb + j >= p.n) v[j] = 0.f;
      }
      if (p.out_mode == BUCTD_OUT_NCHW32) {
        float* o = reinterpret_cast<float*>(p.out);
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (nb + j < p.n) o[(((long long)ob * p.n + nb + j) * p.H + oy) * p.W + ox] = v[j];
      } else {
        T* o = reinterpret_cast<T*>(p.out) + r * p.out_cs + nb;
        if (nb + 16 <= p.out_cs && (p.out_cs & 7) == 0) {
          uint4 u0, u1;
          u0.x = T16<T>::pack2(v[0], v[1]);   u0.y = T16<T>::pack2(v[2], v[3]);
          u0.z = T16<T>::pack2(v[4], v[5]);   u0.w = T16<T>::pack2(v[6], v[7]);
          u1.x = T16<T>::pack2(v[8], v[9]);   u1.y = T16<T>::pack2(v[10], v[11]);
          u1.z = T16<T>::pack2(v[12], v[13]); u1.w = T16<T>::pack2(v[14], v[15]);
          *reinterpret_cast<uint4*>(o) = u0;
          *reinterpret_cast<uint4*>(o + 8) = u1;
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (nb + j < p.out_cs) o[j] = T16<T>::from_f(v[j]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// ---- host side --------------------------------------------------------------------------------
static size_t tg_smem_bytes(int bn) { return 1024 + (size_t)kStages * (kATileBytes + bn * 128) + 128; }

template <typename T, int BN>
static int set_attr() {
  BUCTD_CUDA(cudaFuncSetAttribute(tapgemm_kernel<T, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)tg_smem_bytes(BN)));
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
    // prefer a th that divides oh to avoid ragged tiles
    int best = t;
    for (int c = t; c >= 1; --c)
      if (oh % c == 0) { best = (c * 2 > t) ? c : t; break; }
    *th = best;
  } else {
    *tw = 128;
    *th = 1;
  }
}

template <typename T, int BN>
static int launch_tg(const CUtensorMap& ma, const CUtensorMap& mb, const TapGemmKParams& kp, dim3 grid,
                     cudaStream_t st) {
  const size_t smem = tg_smem_bytes(BN);
  tapgemm_kernel<T, BN><<<grid, kThreads, smem, st>>>(ma, mb, kp);
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}

template <typename T>
static int dispatch_bn(int bn, const CUtensorMap& ma, const CUtensorMap& mb, const TapGemmKParams& kp, dim3 grid,
                       cudaStream_t st) {
  switch (bn) {
    case 32: return launch_tg<T, 32>(ma, mb, kp, grid, st);
    case 48: return launch_tg<T, 48>(ma, mb, kp, grid, st);
    case 64: return launch_tg<T, 64>(ma, mb, kp, grid, st);
    case 96: return launch_tg<T, 96>(ma, mb, kp, grid, st);
    case 128: return launch_tg<T, 128>(ma, mb, kp, grid, st);
    case 192: return launch_tg<T, 192>(ma, mb, kp, grid, st);
    case 256: return launch_tg<T, 256>(ma, mb, kp, grid, st);
  }
  set_error("tapgemm: unsupported BN %d", bn);
  return BUCTD_ERR_INVALID;
}

int tapgemm_prepare(const buctd_tapgemm_args& a, TapGemmLaunch* L) {
  BUCTD_CHECK(a.dtype == BUCTD_F16 || a.dtype == BUCTD_BF16, BUCTD_ERR_INVALID, "tapgemm: bad dtype %d", a.dtype);
  BUCTD_CHECK(a.ntaps == 1 || a.ntaps == 9, BUCTD_ERR_INVALID, "tapgemm: ntaps must be 1 or 9, got %d", a.ntaps);
  BUCTD_CHECK(a.k > 0 && a.n > 0 && a.n_alloc >= a.n && a.n_alloc % 16 == 0, BUCTD_ERR_INVALID,
              "tapgemm: bad k/n/n_alloc %d/%d/%d", a.k, a.n, a.n_alloc);
  BUCTD_CHECK(a.a_cs % 8 == 0 && a.w_cs % 8 == 0 && a.w_cs >= a.k, BUCTD_ERR_INVALID,
              "tapgemm: channel strides must be multiples of 8 (a_cs %d, w_cs %d)", a.a_cs, a.w_cs);
  BUCTD_CHECK(a.a && a.w && a.out, BUCTD_ERR_INVALID, "tapgemm: null pointer");
  BUCTD_CHECK(((uintptr_t)a.a & 15) == 0 && ((uintptr_t)a.w & 15) == 0 && ((uintptr_t)a.out & 15) == 0,
              BUCTD_ERR_INVALID, "tapgemm: pointers must be 16-byte aligned");
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
  if (a.mode == BUCTD_TG_FLAT) {
    BUCTD_CHECK(a.m_rows > 0 && a.m_rows < (1ll << 31), BUCTD_ERR_INVALID, "tapgemm: bad m_rows %lld", (long long)a.m_rows);
    if (a.out_mode != BUCTD_OUT_ROWS16 || a.ntaps == 9)
      BUCTD_CHECK((long long)a.batch * (a.out_h + 1) * (a.out_w + 1) == a.m_rows, BUCTD_ERR_INVALID,
                  "tapgemm: m_rows %lld does not match act16 geometry %dx%dx%d", (long long)a.m_rows, a.batch, a.out_h, a.out_w);
    cuuint64_t dims[2] = {a_inner, (cuuint64_t)a.m_rows};
    cuuint64_t strides[1] = {(cuuint64_t)a.a_cs * 2};
    cuuint32_t box[2] = {64, kBlockM};
    cuuint32_t es[2] = {1, 1};
    cr = enc(&L->map_a, dt, 2, const_cast<void*>(a.a), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    BUCTD_CHECK(cr == CUDA_SUCCESS, BUCTD_ERR_CUDA, "cuTensorMapEncodeTiled(A flat) failed: %d", (int)cr);
    kp.a_tile_bytes = kATileBytes;
    grid = dim3((unsigned)((a.m_rows + kBlockM - 1) / kBlockM), ntiles);
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
  if (L.dtype == BUCTD_F16) return dispatch_bn<__half>(L.bn, L.map_a, L.map_b, L.kp, L.grid, st);
  return dispatch_bn<__nv_bfloat16>(L.bn, L.map_a, L.map_b, L.kp, L.grid, st);
}

}  // namespace buctd
