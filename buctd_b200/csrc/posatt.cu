// CoAM position attention core: o[t] = softmax_s(q[t].k[s] / sqrt(C)) v[s] over the H*W valid pixels
// of each crop, flash style: the [HW x HW] score matrix (191 MB fp32 per crop on the 96x72 branch in
// the reference) is never materialised.  ScaledDotProductAttention.forward,
// lib/models/self_attention.py:78-86 (h = 1, eval mode).
//
// Round-1 implementation: warp-level mma.sync.m16n8k16 (legacy tensor path) with fp32 online
// soft-max; 64 queries x 64 keys per step, K/V tiles double-buffered with cp.async.  q/k/v are
// act16 tensors; the sequence of one crop is its (H+1)(W+1) padded rows, pad rows are masked out
// as keys (additive -inf) and never stored as queries.  A tcgen05/TMEM version is the follow-up.
#include "common.cuh"
#include "internal.h"

namespace buctd {

constexpr int kBM = 64, kBN = 64;

template <typename T> struct MmaOp;
template <> struct MmaOp<__half> {
  __device__ static __forceinline__ void mma(float* c, const uint32_t* a, const uint32_t* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
};
template <> struct MmaOp<__nv_bfloat16> {
  __device__ static __forceinline__ void mma(float* c, const uint32_t* a, const uint32_t* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
};

// 2 exponentials per MUFU op, result already in the 16-bit type the P.V MMA consumes
template <typename T> __device__ __forceinline__ uint32_t ex2_pair(float a, float b);
template <> __device__ __forceinline__ uint32_t ex2_pair<__half>(float a, float b) {
  uint32_t h = T16<__half>::pack2(a, b), r;
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(r) : "r"(h));
  return r;
}
template <> __device__ __forceinline__ uint32_t ex2_pair<__nv_bfloat16>(float a, float b) {
  uint32_t h = T16<__nv_bfloat16>::pack2(a, b), r;
  asm("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(r) : "r"(h));
  return r;
}

__device__ __forceinline__ void ldsm_x4(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool pred) {
  const int sz = pred ? 16 : 0;    // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

template <typename T, int D>
__global__ void __launch_bounds__(128) pos_att_kernel(buctd_pos_att_args a) {
  constexpr int PITCH = D * 2 + 16;            // bytes; odd number of 16-B chunks => conflict-free ldmatrix
  constexpr int TILE = kBN * PITCH;
  constexpr int CHUNKS = D / 8;                // 16-byte chunks per row
  extern __shared__ __align__(16) uint8_t smem[];
  uint8_t* s_q = smem;                         // [64][PITCH]
  uint8_t* s_k = s_q + kBM * PITCH;            // [2][64][PITCH]
  uint8_t* s_v = s_k + 2 * TILE;               // [2][64][PITCH]
  float* s_bias = reinterpret_cast<float*>(s_v + 2 * TILE);   // [2][64] additive key mask

  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y;
  const int Wp = a.w + 1, Hp = a.h + 1;
  const int nrows = Hp * Wp;
  const int m0 = blockIdx.x * kBM;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long base = (long long)b * nrows;
  const T* gq = reinterpret_cast<const T*>(a.q) + base * a.q_cs;
  const T* gk = reinterpret_cast<const T*>(a.k) + base * a.k_cs;
  const T* gv = reinterpret_cast<const T*>(a.v) + base * a.v_cs;

  // ---- async loads --------------------------------------------------------------------------
  for (int i = tid; i < kBM * CHUNKS; i += 128) {
    const int r = i / CHUNKS, c = i % CHUNKS;
    const bool ok = m0 + r < nrows;
    cp_async16(smem_u32(s_q + r * PITCH + c * 16), gq + (long long)(ok ? m0 + r : 0) * a.q_cs + c * 8, ok);
  }
  auto load_kv = [&](int tile, int buf) {
    const int n0 = tile * kBN;
    for (int i = tid; i < kBN * CHUNKS; i += 128) {
      const int r = i / CHUNKS, c = i % CHUNKS;
      const bool ok = n0 + r < nrows;
      const long long row = ok ? n0 + r : 0;
      cp_async16(smem_u32(s_k + buf * TILE + r * PITCH + c * 16), gk + row * a.k_cs + c * 8, ok);
      cp_async16(smem_u32(s_v + buf * TILE + r * PITCH + c * 16), gv + row * a.v_cs + c * 8, ok);
    }
    if (tid < kBN) {
      const int r = n0 + tid;
      const bool valid = r < nrows && (r % Wp) != a.w && (r / Wp) != a.h;
      s_bias[buf * kBN + tid] = valid ? 0.f : -INFINITY;
    }
  };
  const int ntiles = (nrows + kBN - 1) / kBN;
  load_kv(0, 0);
  cp_async_commit();

  // ---- per-thread state -----------------------------------------------------------------------
  constexpr int KS = D / 16;                   // k-steps of QK^T
  constexpr int NT_O = D / 8;                  // n-tiles of the output
  uint32_t qf[KS][4];
  float o[NT_O][4];
#pragma unroll
  for (int i = 0; i < NT_O; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float row_m[2] = {-INFINITY, -INFINITY};
  // Row sums come out of the tensor cores: a constant extra B fragment ("ones column" appended to V) makes
  // o_l[0] / o_l[2] (lanes with lane%4 == 0) accumulate sum_k P[row][k] with exactly the rounded P that P.V uses.
  float o_l[4] = {0.f, 0.f, 0.f, 0.f};
  uint32_t ones_frag[2];
  ones_frag[0] = ones_frag[1] = (lane >> 2) == 0 ? T16<T>::pack2(1.f, 1.f) : 0u;
  const float sc = rsqrtf((float)a.c) * 1.4426950408889634f;   // 1/sqrt(C) * log2(e)

  cp_async_wait<0>();
  __syncthreads();
  {
    const int r = warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
      ldsm_x4(qf[ks], smem_u32(s_q + r * PITCH + (ks * 16 + 8 * (lane >> 4)) * 2));
  }

  for (int j = 0; j < ntiles; ++j) {
    const int buf = j & 1;
    if (j + 1 < ntiles) load_kv(j + 1, buf ^ 1);
    cp_async_commit();
    // tile j was committed one group earlier (or before the loop) -> at most 1 group may stay pending
    cp_async_wait<1>();
    __syncthreads();

    const uint8_t* kb = s_k + buf * TILE;
    const uint8_t* vb = s_v + buf * TILE;
    float s[kBN / 8][4];
#pragma unroll
    for (int nt = 0; nt < kBN / 8; ++nt) { s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int np = 0; np < kBN / 16; ++np) {
        uint32_t kf[4];
        const int kr = np * 16 + (lane & 7) + 8 * (lane >> 4);
        ldsm_x4(kf, smem_u32(kb + kr * PITCH + (ks * 16 + 8 * ((lane >> 3) & 1)) * 2));
        MmaOp<T>::mma(s[2 * np], qf[ks], kf);
        MmaOp<T>::mma(s[2 * np + 1], qf[ks], kf + 2);
      }
    }
    // ---- online soft-max (rows g and g+8 of this warp's 16) ------------------------------------
    const float* bias = s_bias + buf * kBN;
    float mx[2] = {row_m[0], row_m[1]};
#pragma unroll
    for (int nt = 0; nt < kBN / 8; ++nt) {
      const float b0 = bias[nt * 8 + 2 * (lane & 3)], b1 = bias[nt * 8 + 2 * (lane & 3) + 1];
      s[nt][0] = s[nt][0] * sc + b0; s[nt][1] = s[nt][1] * sc + b1;
      s[nt][2] = s[nt][2] * sc + b0; s[nt][3] = s[nt][3] * sc + b1;
      mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffff, mx[h], 1));
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffff, mx[h], 2));
    }
    float corr[2], msafe[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      msafe[h] = (mx[h] == -INFINITY) ? 0.f : mx[h];          // a tile of pads only
      corr[h] = exp2f(row_m[h] - msafe[h]);                   // row_m = -inf on the first tile -> 0
      row_m[h] = mx[h];
    }
    uint32_t pf[kBN / 16][4];
#pragma unroll
    for (int nt = 0; nt < kBN / 8; ++nt) {
      // C fragment of n-tiles (2k, 2k+1) -> A fragment of k-step k
      pf[nt >> 1][(nt & 1) * 2 + 0] = ex2_pair<T>(s[nt][0] - msafe[0], s[nt][1] - msafe[0]);
      pf[nt >> 1][(nt & 1) * 2 + 1] = ex2_pair<T>(s[nt][2] - msafe[1], s[nt][3] - msafe[1]);
    }
#pragma unroll
    for (int nt = 0; nt < NT_O; ++nt) {
      o[nt][0] *= corr[0]; o[nt][1] *= corr[0];
      o[nt][2] *= corr[1]; o[nt][3] *= corr[1];
    }
    o_l[0] *= corr[0]; o_l[1] *= corr[0]; o_l[2] *= corr[1]; o_l[3] *= corr[1];
    // ---- O += P V -------------------------------------------------------------------------------
#pragma unroll
    for (int ks = 0; ks < kBN / 16; ++ks) {
      MmaOp<T>::mma(o_l, pf[ks], ones_frag);
#pragma unroll
      for (int np = 0; np < NT_O / 2; ++np) {
        uint32_t vf[4];
        const int vr = ks * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
        ldsm_x4_t(vf, smem_u32(vb + vr * PITCH + (np * 16 + 8 * (lane >> 4)) * 2));
        MmaOp<T>::mma(o[2 * np], pf[ks], vf);
        MmaOp<T>::mma(o[2 * np + 1], pf[ks], vf + 2);
      }
    }
    __syncthreads();      // everyone done with buffer `buf` before it is refilled two steps later
  }

  // ---- finalise: divide by the row sum (reduced over the 4 lanes of a row) and store ----------
  float row_l[2];
  row_l[0] = __shfl_sync(0xffffffff, o_l[0], lane & ~3);      // column 0 of the ones tile lives in lane%4 == 0
  row_l[1] = __shfl_sync(0xffffffff, o_l[2], lane & ~3);
  T* go = reinterpret_cast<T*>(a.o) + base * a.o_cs;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int r = m0 + warp * 16 + (lane >> 2) + 8 * h;
    if (r >= nrows) continue;
    const bool pad = (r % Wp) == a.w || (r / Wp) == a.h;
    const float inv = pad ? 0.f : 1.f / row_l[h];
#pragma unroll
    for (int nt = 0; nt < NT_O; ++nt) {
      const uint32_t u = T16<T>::pack2(o[nt][2 * h] * inv, o[nt][2 * h + 1] * inv);
      *reinterpret_cast<uint32_t*>(go + (long long)r * a.o_cs + nt * 8 + 2 * (lane & 3)) = u;
    }
  }
}

template <int D> static size_t pa_smem() { return (size_t)(kBM + 4 * kBN) * (D * 2 + 16) + 2 * kBN * sizeof(float); }

template <typename T, int D>
static int pa_launch(const buctd_pos_att_args& a, cudaStream_t st) {
  const int nrows = (a.h + 1) * (a.w + 1);
  dim3 grid((nrows + kBM - 1) / kBM, a.batch);
  launch_pdl(pos_att_kernel<T, D>, grid, dim3(128), pa_smem<D>(), st, a);
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}

template <typename T, int D> static int pa_attr() {
  BUCTD_CUDA(cudaFuncSetAttribute(pos_att_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pa_smem<D>()));
  return BUCTD_OK;
}
template <typename T> static int pa_attrs() {
  int rc;
  if ((rc = pa_attr<T, 32>())) return rc;
  if ((rc = pa_attr<T, 48>())) return rc;
  if ((rc = pa_attr<T, 64>())) return rc;
  if ((rc = pa_attr<T, 96>())) return rc;
  if ((rc = pa_attr<T, 128>())) return rc;
  return pa_attr<T, 192>();
}
int pos_att_init_device() {
  int rc = pa_attrs<__half>();
  return rc ? rc : pa_attrs<__nv_bfloat16>();
}

template <typename T>
static int pa_dispatch(const buctd_pos_att_args& a, cudaStream_t st) {
  switch (a.c) {
    case 32: return pa_launch<T, 32>(a, st);
    case 48: return pa_launch<T, 48>(a, st);
    case 64: return pa_launch<T, 64>(a, st);
    case 96: return pa_launch<T, 96>(a, st);
    case 128: return pa_launch<T, 128>(a, st);
    case 192: return pa_launch<T, 192>(a, st);
  }
  set_error("pos_att: unsupported head dim C=%d (supported: 32 48 64 96 128 192)", a.c);
  return BUCTD_ERR_INVALID;
}

int launch_pos_att(const buctd_pos_att_args& a, cudaStream_t st) {
  BUCTD_CHECK(a.q && a.k && a.v && a.o, BUCTD_ERR_INVALID, "pos_att: null pointer");
  BUCTD_CHECK(a.q_cs % 8 == 0 && a.k_cs % 8 == 0 && a.v_cs % 8 == 0 && a.o_cs % 2 == 0, BUCTD_ERR_INVALID,
              "pos_att: channel strides must be multiples of 8");
  BUCTD_CHECK((((uintptr_t)a.q | (uintptr_t)a.k | (uintptr_t)a.v) & 15) == 0, BUCTD_ERR_INVALID,
              "pos_att: q/k/v must be 16-byte aligned");
  if (a.batch == 0) return BUCTD_OK;
  if (pos_att_tc_supported(a)) return launch_pos_att_tc(a, st);       // tcgen05 path (C <= 64)
  return a.dtype == BUCTD_F16 ? pa_dispatch<__half>(a, st) : pa_dispatch<__nv_bfloat16>(a, st);
}

}  // namespace buctd
