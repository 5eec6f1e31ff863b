// CUDA-core kernels of the path: everything that is HBM-bound or too small for tensor cores.
// stem conv1, exchange/fuse sum, antialiased condition resize, CoAM channel attention, DAModule
// combine, layout converters.  (Decode + rasteriser live in decode.cu, attention in posatt.cu.)
#include "common.cuh"
#include "internal.h"

namespace buctd {

// =================================================================================================
// Stem conv1: 3x3 s2, 3 -> 64, folded BN + ReLU; nchw32 in, act16 out.
// lib/models/pose_hrnet_coam.py:505-507.  One thread = one output pixel x 64 channels.
// =================================================================================================
constexpr int kStemTile = 16;   // 16x16 output pixels per block

template <typename T>
__global__ void __launch_bounds__(256) stem_kernel(buctd_stem_args a) {
  __shared__ float s_w[27][64];
  __shared__ float s_b[64];
  __shared__ float s_in[3][2 * kStemTile + 1][2 * kStemTile + 2];
  const int oh = a.h / 2, ow = a.w / 2;
  const int b = blockIdx.z;
  const int oy0 = blockIdx.y * kStemTile, ox0 = blockIdx.x * kStemTile;
  const int tid = threadIdx.x;
  for (int i = tid; i < 27 * 64; i += 256) {
    int co = i / 27, t = i % 27;          // weight [co][ci*9 + ky*3 + kx]
    s_w[t][co] = a.weight[i];
  }
  if (tid < 64) s_b[tid] = a.bias[tid];
  const int iy0 = 2 * oy0 - 1, ix0 = 2 * ox0 - 1;
  constexpr int PH = 2 * kStemTile + 1, PW = 2 * kStemTile + 1;
  for (int i = tid; i < 3 * PH * PW; i += 256) {
    int c = i / (PH * PW), r = i % (PH * PW);
    int yy = r / PW, xx = r % PW;
    int iy = iy0 + yy, ix = ix0 + xx;
    float v = 0.f;
    if (iy >= 0 && iy < a.h && ix >= 0 && ix < a.w) {
      const long long o = (((long long)b * a.x_channels + c) * a.h + iy) * a.w + ix;
      v = a.x_half ? __half2float(reinterpret_cast<const __half*>(a.x)[o]) : reinterpret_cast<const float*>(a.x)[o];
    }
    s_in[c][yy][xx] = v;
  }
  __syncthreads();
  const int ty = tid / kStemTile, tx = tid % kStemTile;
  const int oy = oy0 + ty, ox = ox0 + tx;
  if (oy >= oh || ox >= ow) return;
  float in[27];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) in[c * 9 + ky * 3 + kx] = s_in[c][2 * ty + ky][2 * tx + kx];
  T* o = reinterpret_cast<T*>(a.out) + (((long long)b * (oh + 1) + oy) * (ow + 1) + ox) * a.out_cs;
#pragma unroll 1
  for (int c0 = 0; c0 < 64; c0 += 8) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = s_b[c0 + j];
#pragma unroll
    for (int t = 0; t < 27; ++t) {
      const float4 w0 = *reinterpret_cast<const float4*>(&s_w[t][c0]);
      const float4 w1 = *reinterpret_cast<const float4*>(&s_w[t][c0 + 4]);
      acc[0] = fmaf(in[t], w0.x, acc[0]); acc[1] = fmaf(in[t], w0.y, acc[1]);
      acc[2] = fmaf(in[t], w0.z, acc[2]); acc[3] = fmaf(in[t], w0.w, acc[3]);
      acc[4] = fmaf(in[t], w1.x, acc[4]); acc[5] = fmaf(in[t], w1.y, acc[5]);
      acc[6] = fmaf(in[t], w1.z, acc[6]); acc[7] = fmaf(in[t], w1.w, acc[7]);
    }
    const float lo = a.no_relu ? -INFINITY : 0.f;
    uint4 u;
    u.x = T16<T>::pack2(fmaxf(acc[0], lo), fmaxf(acc[1], lo));
    u.y = T16<T>::pack2(fmaxf(acc[2], lo), fmaxf(acc[3], lo));
    u.z = T16<T>::pack2(fmaxf(acc[4], lo), fmaxf(acc[5], lo));
    u.w = T16<T>::pack2(fmaxf(acc[6], lo), fmaxf(acc[7], lo));
    *reinterpret_cast<uint4*>(o + c0) = u;
  }
}

int launch_stem(const buctd_stem_args& a, cudaStream_t st) {
  BUCTD_CHECK(a.h % 2 == 0 && a.w % 2 == 0 && a.out_cs >= 64 && a.out_cs % 8 == 0, BUCTD_ERR_INVALID,
              "stem: need even H,W and out_cs>=64 (got %dx%d, cs %d)", a.h, a.w, a.out_cs);
  BUCTD_CHECK(a.x && a.weight && a.bias && a.out && a.x_channels >= 3, BUCTD_ERR_INVALID, "stem: bad pointers/channels");
  dim3 grid(ceil_div(a.w / 2, kStemTile), ceil_div(a.h / 2, kStemTile), a.batch);
  if (a.dtype == BUCTD_F16) stem_kernel<__half><<<grid, 256, 0, st>>>(a);
  else stem_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(a);
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}

// =================================================================================================
// Exchange/fuse: out = ReLU(sum_j nearest_up(term_j)), 8 channels (16 B) per thread.
// lib/models/pose_hrnet_coam.py:257-266.
// =================================================================================================
// One block = one padded row (b, y) of the output; thread = (pixel lane, 8-channel group), fixed for the whole row, so
// there is no index division per element and the terms' row bases are computed once (the round-2 ncu capture showed the
// flat-index version issue-bound: 73 % issue slots busy at 30 % of the DRAM bandwidth).
template <typename T>
__global__ void __launch_bounds__(256) fuse_kernel(buctd_fuse_args a) {
  pdl_launch_dependents();
  pdl_wait();
  const unsigned cv = (unsigned)a.c / 8;
  const unsigned Hp = a.h + 1, Wp = a.w + 1;
  const unsigned px_lanes = blockDim.x / cv;              // pixels per step
  const unsigned pl = threadIdx.x / cv, c8 = threadIdx.x - pl * cv;
  if (pl >= px_lanes) return;
  const int nt = a.nterms;
  for (unsigned row = blockIdx.x; row < (unsigned)a.batch * Hp; row += gridDim.x) {
    const unsigned b = row / Hp, y = row - b * Hp;
    T* out = reinterpret_cast<T*>(a.out) + (size_t)row * Wp * a.out_cs + c8 * 8;
    if (y == (unsigned)a.h) {                               // bottom pad row
      for (unsigned x = pl; x < Wp; x += px_lanes) *reinterpret_cast<uint4*>(out + (size_t)x * a.out_cs) = make_uint4(0, 0, 0, 0);
      continue;
    }
    const T* tb[4];
    int tcs[4], tsh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j < nt) {
        const int sh = a.term_shift[j];
        const unsigned sw = ((unsigned)a.w >> sh) + 1, shh = ((unsigned)a.h >> sh) + 1;
        tb[j] = reinterpret_cast<const T*>(a.term[j]) + (size_t)(b * shh + (y >> sh)) * sw * a.term_cs[j] + c8 * 8;
        tcs[j] = a.term_cs[j]; tsh[j] = sh;
      }
    }
    for (unsigned x = pl; x < Wp; x += px_lanes) {
      uint4 u = make_uint4(0, 0, 0, 0);
      if (x < (unsigned)a.w) {
        uint4 t[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)                         // all loads first: the terms' latencies overlap
          if (j < nt) t[j] = *reinterpret_cast<const uint4*>(tb[j] + (size_t)(x >> tsh[j]) * tcs[j]);
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (j < nt) {
            const uint32_t w[4] = {t[j].x, t[j].y, t[j].z, t[j].w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float2 f = T16<T>::unpack2(w[q]);
              acc[2 * q] += f.x;
              acc[2 * q + 1] += f.y;
            }
          }
        }
        if (a.relu) {
#pragma unroll
          for (int q = 0; q < 8; ++q) acc[q] = fmaxf(acc[q], 0.f);
        }
        u.x = T16<T>::pack2(acc[0], acc[1]); u.y = T16<T>::pack2(acc[2], acc[3]);
        u.z = T16<T>::pack2(acc[4], acc[5]); u.w = T16<T>::pack2(acc[6], acc[7]);
      }
      *reinterpret_cast<uint4*>(out + (size_t)x * a.out_cs) = u;
    }
  }
}

int launch_fuse(const buctd_fuse_args& a, cudaStream_t st) {
  BUCTD_CHECK(a.nterms >= 1 && a.nterms <= 4, BUCTD_ERR_INVALID, "fuse: nterms %d out of range", a.nterms);
  BUCTD_CHECK(a.c % 8 == 0 && a.out_cs % 8 == 0, BUCTD_ERR_INVALID, "fuse: channels must be a multiple of 8");
  for (int j = 0; j < a.nterms; ++j) {
    BUCTD_CHECK(a.term[j] != nullptr && a.term_cs[j] % 8 == 0, BUCTD_ERR_INVALID, "fuse: bad term %d", j);
    BUCTD_CHECK((a.h % (1 << a.term_shift[j])) == 0 && (a.w % (1 << a.term_shift[j])) == 0, BUCTD_ERR_INVALID,
                "fuse: %dx%d not divisible by 2^%d", a.h, a.w, a.term_shift[j]);
  }
  BUCTD_CHECK(a.c / 8 <= 256 && (long long)a.batch * (a.h + 1) < (1ll << 31), BUCTD_ERR_INVALID, "fuse: C <= 2048 and B (H+1) < 2^31");
  const long long rows_total = (long long)a.batch * (a.h + 1);
  const int blocks = (int)(rows_total < 148 * 16 ? rows_total : 148 * 16);
  if (a.dtype == BUCTD_F16) launch_pdl(fuse_kernel<__half>, dim3(blocks), dim3(256), 0, st, a);
  else launch_pdl(fuse_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, st, a);
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}

// =================================================================================================
// Condition resize (antialiased bilinear = triangle filter of support `scale`), nchw32 -> act16.
// TF.resize(cond_hm, (H_i, W_i)), lib/models/pose_hrnet_coam.py:755; closed form SURVEY.md App. B:
// centre = scale*(o+0.5); taps j in [max(0,int(centre-scale+0.5)), min(in,int(centre+scale+0.5)));
// weight max(0, 1-|j-centre+0.5|/scale), normalised.  antialias=0: support 1 (2-tap bilinear).
// Horizontal pass first, then vertical, both in fp32 (ATen's separable order).
// =================================================================================================
constexpr int kRzRows = 4;      // output rows per block
constexpr int kRzMaxW = 512;
constexpr int kRzMaxOw = 128;

__device__ __forceinline__ void aa_span(float scale, float support, int o, int n_in, int* lo, int* hi, float* center) {
  *center = scale * (o + 0.5f);
  *lo = max(0, (int)(*center - support + 0.5f));
  *hi = min(n_in, (int)(*center + support + 0.5f));
}

template <typename T>
__global__ void __launch_bounds__(256) resize_kernel(buctd_resize_args a) {
  // block = kRzRows output rows of one crop, ALL channels: per channel, pass 1 (threads over input columns,
  // coalesced) = vertical triangle filter into shared memory, pass 2 = horizontal filter into a channel-last
  // shared tile; the tile leaves with coalesced 16-byte stores (act16 rows are out_cs*2 bytes of contiguous channels).
  __shared__ float s_v[kRzRows][kRzMaxW];
  extern __shared__ __align__(16) uint8_t s_out[];            // [kRzRows][ow][out_cs] 16-bit
  const int b = blockIdx.y;
  const int oy0 = blockIdx.x * kRzRows;
  const int nrows = min(kRzRows, a.oh - oy0);
  const float sy = (float)a.h / a.oh, sx = (float)a.w / a.ow;
  const bool aay = a.antialias && sy > 1.f, aax = a.antialias && sx > 1.f;
  const float supy = aay ? sy : 1.f, supx = aax ? sx : 1.f;
  const float invy = aay ? 1.f / sy : 1.f, invx = aax ? 1.f / sx : 1.f;
  T* so = reinterpret_cast<T*>(s_out);
  for (int i = threadIdx.x; i < nrows * a.ow * a.out_cs; i += blockDim.x) so[i] = T16<T>::from_f(0.f);
  for (int c = 0; c < a.cc; ++c) {
    const float* src = a.x + (((long long)b * a.x_channels + a.c_off + c) * a.h) * a.w;
    for (int g = 0; g < nrows; ++g) {
      int lo, hi; float center;
      aa_span(sy, supy, oy0 + g, a.h, &lo, &hi, &center);
      float tot = 0.f;
      for (int j = lo; j < hi; ++j) tot += fmaxf(0.f, 1.f - fabsf((j - center + 0.5f) * invy));
      const float rt = 1.f / tot;
      for (int x = threadIdx.x; x < a.w; x += blockDim.x) {
        float acc = 0.f;
        for (int j = lo; j < hi; ++j)
          acc = fmaf(fmaxf(0.f, 1.f - fabsf((j - center + 0.5f) * invy)) * rt, __ldg(src + (long long)j * a.w + x), acc);
        s_v[g][x] = acc;
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nrows * a.ow; i += blockDim.x) {
      const int g = i / a.ow, ox = i - g * a.ow;
      int lo, hi; float center;
      aa_span(sx, supx, ox, a.w, &lo, &hi, &center);
      float tot = 0.f, acc = 0.f;
      for (int j = lo; j < hi; ++j) {
        const float w = fmaxf(0.f, 1.f - fabsf((j - center + 0.5f) * invx));
        tot += w;
        acc = fmaf(w, s_v[g][j], acc);
      }
      so[(g * a.ow + ox) * a.out_cs + c] = T16<T>::from_f(acc / tot);
    }
    __syncthreads();
  }
  // rows of the tile are contiguous in the act16 tensor: (ow * out_cs * 2) bytes per output row
  const int chunks_per_row = a.ow * a.out_cs / 8;
  for (int i = threadIdx.x; i < nrows * chunks_per_row; i += blockDim.x) {
    const int g = i / chunks_per_row, ch = i - g * chunks_per_row;
    T* o = reinterpret_cast<T*>(a.out) + (((long long)b * (a.oh + 1) + oy0 + g) * (a.ow + 1)) * a.out_cs;
    reinterpret_cast<uint4*>(o)[ch] = reinterpret_cast<const uint4*>(so + (size_t)g * a.ow * a.out_cs)[ch];
  }
}

// Fast path: antialiased, integer scale S (4 / 8 / 16 for the W48 and W32 families), W % 4 == 0.
// Block = kRzRows output rows x 8 channels x the full width; thread = (channel, 4 consecutive input columns).
// Vertical pass: every thread streams the (kRzRows+1)*S input rows of its band with 128-bit loads (all
// independent, compile-time trip count, running pointer) into 4x4 accumulators -- input row i feeds output row
// i/S with tap i%S and output row i/S-1 with tap i%S+S.  Horizontal pass from shared memory, then one
// 16-byte (8-channel) store per output pixel.
template <int S> __device__ __forceinline__ float aa_w(int t) { return 1.f - fabsf(t - S + 0.5f) * (1.f / S); }

template <typename T, int S>
__global__ void __launch_bounds__(576) resize_int_kernel(buctd_resize_args a) {
  extern __shared__ __align__(16) float s_rz[];            // [8 ch][kRzRows][w] fp32, then [rows][ow][8] 16-bit
  const int b = blockIdx.z, c0 = blockIdx.y * 8;
  const int oy0 = blockIdx.x * kRzRows;
  const int nrows = min(kRzRows, a.oh - oy0);
  const int w4 = a.w >> 2;
  float* s_v = s_rz;
  T* so = reinterpret_cast<T*>(s_rz + 8 * kRzRows * a.w);
  const int base = S * oy0 - S / 2;
  // normalisation of the vertical taps (same for every column): sum of the weights of the rows that exist
  float rtot[kRzRows];
#pragma unroll
  for (int g = 0; g < kRzRows; ++g) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 2 * S; ++k) {
      const int jrow = base + g * S + k;
      t += (jrow >= 0 && jrow < a.h) ? aa_w<S>(k) : 0.f;
    }
    rtot[g] = t > 0.f ? 1.f / t : 0.f;
  }
  for (int i = threadIdx.x; i < 8 * w4; i += blockDim.x) {
    const int cc = i / w4, x4 = i - cc * w4;
    float4 acc[kRzRows];
#pragma unroll
    for (int g = 0; g < kRzRows; ++g) acc[g] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c0 + cc < a.cc) {
      const float4* ptr = reinterpret_cast<const float4*>(a.x + (((long long)b * a.x_channels + a.c_off + c0 + cc) * a.h) * a.w) +
                          (long long)base * w4 + x4;
#pragma unroll
      for (int r = 0; r < (kRzRows + 1) * S; ++r) {
        const bool ok = (unsigned)(base + r) < (unsigned)a.h;
        const float4 v = ok ? __ldg(ptr) : make_float4(0.f, 0.f, 0.f, 0.f);
        ptr += w4;
        const int g = r / S, t = r % S;                   // compile-time after unrolling
        if (g < kRzRows) {
          const float wt = aa_w<S>(t);
          acc[g].x = fmaf(wt, v.x, acc[g].x); acc[g].y = fmaf(wt, v.y, acc[g].y);
          acc[g].z = fmaf(wt, v.z, acc[g].z); acc[g].w = fmaf(wt, v.w, acc[g].w);
        }
        if (g >= 1) {
          const float wt = aa_w<S>(t + S);
          acc[g - 1].x = fmaf(wt, v.x, acc[g - 1].x); acc[g - 1].y = fmaf(wt, v.y, acc[g - 1].y);
          acc[g - 1].z = fmaf(wt, v.z, acc[g - 1].z); acc[g - 1].w = fmaf(wt, v.w, acc[g - 1].w);
        }
      }
    }
#pragma unroll
    for (int g = 0; g < kRzRows; ++g) {
      float4 o = acc[g];
      o.x *= rtot[g]; o.y *= rtot[g]; o.z *= rtot[g]; o.w *= rtot[g];
      reinterpret_cast<float4*>(s_v + (cc * kRzRows + g) * a.w)[x4] = o;
    }
  }
  __syncthreads();
  // horizontal pass: one thread per (row, output pixel), all 8 channels -> one 16-byte chunk
  for (int i = threadIdx.x; i < nrows * a.ow; i += blockDim.x) {
    const int g = i / a.ow, ox = i - g * a.ow;
    const int lo = S * ox - S / 2;
    float tot = 0.f;
#pragma unroll
    for (int t = 0; t < 2 * S; ++t) tot += ((unsigned)(lo + t) < (unsigned)a.w) ? aa_w<S>(t) : 0.f;
    const float rt = 1.f / tot;
    float o[8];
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) {
      const float* row = s_v + (cc * kRzRows + g) * a.w;
      float acc = 0.f;
#pragma unroll
      for (int t = 0; t < 2 * S; ++t) {
        const int jx = lo + t;
        acc = fmaf(aa_w<S>(t), ((unsigned)jx < (unsigned)a.w) ? row[jx] : 0.f, acc);
      }
      o[cc] = (c0 + cc < a.cc) ? acc * rt : 0.f;
    }
    uint4 u;
    u.x = T16<T>::pack2(o[0], o[1]); u.y = T16<T>::pack2(o[2], o[3]);
    u.z = T16<T>::pack2(o[4], o[5]); u.w = T16<T>::pack2(o[6], o[7]);
    reinterpret_cast<uint4*>(so)[i] = u;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nrows * a.ow; i += blockDim.x) {
    const int g = i / a.ow, ox = i - g * a.ow;
    T* o = reinterpret_cast<T*>(a.out) + (((long long)b * (a.oh + 1) + oy0 + g) * (a.ow + 1) + ox) * a.out_cs + c0;
    *reinterpret_cast<uint4*>(o) = reinterpret_cast<const uint4*>(so)[i];
  }
}

template <typename T>
static int resize_dispatch(const buctd_resize_args& a, cudaStream_t st) {
  const bool int_scale = a.antialias && a.h % a.oh == 0 && a.w % a.ow == 0 && a.h / a.oh == a.w / a.ow && a.w <= kRzMaxW &&
                         a.ow <= kRzMaxOw && a.out_cs % 8 == 0 && a.w % 4 == 0 && ((uintptr_t)a.x & 15) == 0;
  const int sc = int_scale ? a.h / a.oh : 0;
  dim3 gi(ceil_div(a.oh, kRzRows), ceil_div(a.cc, 8), a.batch);     // channel groups beyond cc stay zero (workspace is zero-filled)
  const size_t smi = (size_t)8 * kRzRows * a.w * 4 + (size_t)kRzRows * a.ow * 16;
  if (sc == 4 || sc == 8 || sc == 16) {
    if (sc == 4) {
      BUCTD_CUDA(cudaFuncSetAttribute(resize_int_kernel<T, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smi));
      resize_int_kernel<T, 4><<<gi, 576, smi, st>>>(a);
    } else if (sc == 8) {
      BUCTD_CUDA(cudaFuncSetAttribute(resize_int_kernel<T, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smi));
      resize_int_kernel<T, 8><<<gi, 576, smi, st>>>(a);
    } else {
      BUCTD_CUDA(cudaFuncSetAttribute(resize_int_kernel<T, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smi));
      resize_int_kernel<T, 16><<<gi, 576, smi, st>>>(a);
    }
    return BUCTD_OK;
  }
  dim3 grid(ceil_div(a.oh, kRzRows), a.batch);
  const size_t smem = (size_t)kRzRows * a.ow * a.out_cs * 2;
  BUCTD_CHECK(smem <= 40 * 1024, BUCTD_ERR_INVALID, "resize: output tile too large (%zu B)", smem);
  resize_kernel<T><<<grid, 256, smem, st>>>(a);
  return BUCTD_OK;
}

// =================================================================================================
// Fused condition pyramid: the three CoAM scales (1/4, 1/8, 1/16) of the antialiased resize in ONE pass over the
// condition planes -- the three separate launches read the 7.5 MB per crop 3.75 times, this one 1.5 times.
// Block = a band of 32 input rows (8 / 4 / 2 output rows of the three scales) x 8 channels x the full width;
// thread = (channel, 4 consecutive input columns).  The vertical triangle filters of all three scales are
// accumulated in registers while the 48 rows [32 i - 8, 32 i + 40) of the band stream by (128-bit loads, compile-time
// tap weights); then, scale by scale, the vertical result goes through shared memory for the horizontal pass.
// Same arithmetic (vertical in fp32, normalise, horizontal in fp32, normalise) as resize_int_kernel.
// =================================================================================================
constexpr int kPyrBand = 32;

template <int S, int NR>
__device__ __forceinline__ void pyr_accum(float4 (&acc)[NR], const float4& v, int r) {
  // input row r (relative to band start - 8) feeds output row g of scale S with tap k = r - (S*g - S/2 + 8), 0 <= k < 2S
#pragma unroll
  for (int g = 0; g < NR; ++g) {
    const int k = r - (S * g - S / 2 + 8);
    if (k >= 0 && k < 2 * S) {
      const float wt = aa_w<S>(k);
      acc[g].x = fmaf(wt, v.x, acc[g].x); acc[g].y = fmaf(wt, v.y, acc[g].y);
      acc[g].z = fmaf(wt, v.z, acc[g].z); acc[g].w = fmaf(wt, v.w, acc[g].w);
    }
  }
}

template <typename T, int S, int NR>
__device__ __forceinline__ void pyr_emit(const buctd_resize3_args& a, int which, const float4 (&acc)[NR], float* s_v, T* so, int band,
                                         int b, int c0, int cc, int x4, bool active) {
  const int oh = a.h / S, ow = a.w / S;
  const int oy0 = band * NR;
  const int nrows = min(NR, oh - oy0);
  // vertical normalisation: weights of the input rows that exist
  if (active) {
#pragma unroll
    for (int g = 0; g < NR; ++g) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 2 * S; ++k) {
        const int jrow = S * (oy0 + g) - S / 2 + k;
        t += (jrow >= 0 && jrow < a.h) ? aa_w<S>(k) : 0.f;
      }
      const float rt = t > 0.f ? 1.f / t : 0.f;
      float4 o = acc[g];
      o.x *= rt; o.y *= rt; o.z *= rt; o.w *= rt;
      reinterpret_cast<float4*>(s_v + (cc * NR + g) * a.w)[x4] = o;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nrows * ow; i += blockDim.x) {
    const int g = i / ow, ox = i - g * ow;
    const int lo = S * ox - S / 2;
    float tot = 0.f;
#pragma unroll
    for (int t = 0; t < 2 * S; ++t) tot += ((unsigned)(lo + t) < (unsigned)a.w) ? aa_w<S>(t) : 0.f;
    const float rt = 1.f / tot;
    float o[8];
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
      const float* row = s_v + (ch * NR + g) * a.w;
      float sum = 0.f;
#pragma unroll
      for (int t = 0; t < 2 * S; ++t) {
        const int jx = lo + t;
        sum = fmaf(aa_w<S>(t), ((unsigned)jx < (unsigned)a.w) ? row[jx] : 0.f, sum);
      }
      o[ch] = (c0 + ch < a.cc) ? sum * rt : 0.f;
    }
    uint4 u;
    u.x = T16<T>::pack2(o[0], o[1]); u.y = T16<T>::pack2(o[2], o[3]);
    u.z = T16<T>::pack2(o[4], o[5]); u.w = T16<T>::pack2(o[6], o[7]);
    T* dst = reinterpret_cast<T*>(a.out[which]) + (((long long)b * (oh + 1) + oy0 + g) * (ow + 1) + ox) * a.out_cs[which] + c0;
    *reinterpret_cast<uint4*>(dst) = u;
  }
  __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(576) resize_pyramid_kernel(buctd_resize3_args a) {
  extern __shared__ __align__(16) float s_pyr[];             // [8 ch][8 rows][w] fp32 (the largest of the three scales)
  const int b = blockIdx.z, c0 = blockIdx.y * 8, band = blockIdx.x;
  const int w4 = a.w >> 2;
  const int cc = threadIdx.x / w4, x4 = threadIdx.x - cc * w4;
  const bool active = threadIdx.x < 8 * w4;
  float4 a4[8], a8[4], a16[2];
#pragma unroll
  for (int g = 0; g < 8; ++g) a4[g] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int g = 0; g < 4; ++g) a8[g] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int g = 0; g < 2; ++g) a16[g] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (active && c0 + cc < a.cc) {
    const int base = kPyrBand * band - 8;
    const float4* ptr = reinterpret_cast<const float4*>(a.x + (((long long)b * a.x_channels + a.c_off + c0 + cc) * a.h) * a.w) +
                        (long long)base * w4 + x4;
#pragma unroll
    for (int r = 0; r < kPyrBand + 16; ++r) {
      const bool ok = (unsigned)(base + r) < (unsigned)a.h;
      const float4 v = ok ? __ldg(ptr) : make_float4(0.f, 0.f, 0.f, 0.f);
      ptr += w4;
      pyr_accum<4, 8>(a4, v, r);
      pyr_accum<8, 4>(a8, v, r);
      pyr_accum<16, 2>(a16, v, r);
    }
  }
  T* so = nullptr;
  pyr_emit<T, 4, 8>(a, 0, a4, s_pyr, so, band, b, c0, cc, x4, active);
  pyr_emit<T, 8, 4>(a, 1, a8, s_pyr, so, band, b, c0, cc, x4, active);
  pyr_emit<T, 16, 2>(a, 2, a16, s_pyr, so, band, b, c0, cc, x4, active);
}

int launch_resize3(const buctd_resize3_args& a, cudaStream_t st) {
  BUCTD_CHECK(a.x && a.out[0] && a.out[1] && a.out[2], BUCTD_ERR_INVALID, "resize3: null pointer");
  BUCTD_CHECK(a.cc > 0 && a.h % 16 == 0 && a.w % 16 == 0 && a.w <= kRzMaxW && ((uintptr_t)a.x & 15) == 0 && (a.w / 4) * 8 <= 576,
              BUCTD_ERR_INVALID, "resize3: needs H, W multiples of 16, W <= %d, 16-byte aligned input (got %dx%d)", 288, a.h, a.w);
  for (int i = 0; i < 3; ++i)
    BUCTD_CHECK(a.out_cs[i] % 8 == 0 && a.cc <= a.out_cs[i], BUCTD_ERR_INVALID, "resize3: cc %d vs out_cs[%d] %d", a.cc, i, a.out_cs[i]);
  if (a.batch == 0) return BUCTD_OK;
  const size_t smem = (size_t)8 * 8 * a.w * 4;
  dim3 grid(ceil_div(a.h, kPyrBand), ceil_div(a.cc, 8), a.batch);
  if (a.dtype == BUCTD_F16) {
    BUCTD_CUDA(cudaFuncSetAttribute(resize_pyramid_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    resize_pyramid_kernel<__half><<<grid, 576, smem, st>>>(a);
  } else {
    BUCTD_CUDA(cudaFuncSetAttribute(resize_pyramid_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    resize_pyramid_kernel<__nv_bfloat16><<<grid, 576, smem, st>>>(a);
  }
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}

int launch_resize(const buctd_resize_args& a, cudaStream_t st) {
  BUCTD_CHECK(a.cc > 0 && a.cc <= a.out_cs && a.out_cs % 8 == 0, BUCTD_ERR_INVALID, "resize: cc %d vs out_cs %d", a.cc, a.out_cs);
  BUCTD_CHECK(a.w <= kRzMaxW && a.ow <= kRzMaxOw, BUCTD_ERR_INVALID, "resize: width %d -> %d out of range", a.w, a.ow);
  BUCTD_CHECK(a.x && a.out, BUCTD_ERR_INVALID, "resize: null pointer");
  if (a.batch == 0) return BUCTD_OK;
  int rc = a.dtype == BUCTD_F16 ? resize_dispatch<__half>(a, st) : resize_dispatch<__nv_bfloat16>(a, st);
  if (rc) return rc;
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}

// =================================================================================================
// CoAM channel attention (SimplifiedScaledDotProductAttention, h = 1):
//   S[c1][c2] = sum_p yc[p][c1] * y[p][c2] / sqrt(HW)   (pad rows are zero in both, so the sum runs
//   over all (H+1)(W+1) rows), att = softmax_c2(S), out[c1][p] = sum_c2 att[c1][c2] * y[p][c2].
// lib/models/self_attention.py:146-158; lib/models/pose_hrnet_coam.py:673-689.
// =================================================================================================
// S is computed in kCaTile x kCaTile output tiles per block by 8x8 threads, each holding a PER x PER register tile
// (6x6 for the W48 family: 36 FMAs per 12 shared-memory words); kCaTile is 48 for C = 48/96/192 and 32 for C = 32/64/128.
constexpr int kCaRows = 32;      // rows staged per step

template <typename T, int kCaTile>
__global__ void __launch_bounds__(64) chan_scores_kernel(buctd_chan_att_args a, int rows_per_chunk) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int PER = kCaTile / 8;
  __shared__ __align__(8) float s_q[kCaRows][kCaTile + 2];
  __shared__ __align__(8) float s_k[kCaRows][kCaTile + 2];
  const int b = blockIdx.z;
  const int tiles = a.c / kCaTile;
  const int t1 = blockIdx.y / tiles, t2 = blockIdx.y % tiles;
  const long long rows = (long long)(a.h + 1) * (a.w + 1);
  const long long r0 = (long long)blockIdx.x * rows_per_chunk;
  const long long r1 = min(rows, r0 + rows_per_chunk);
  const T* yc = reinterpret_cast<const T*>(a.yc) + (long long)b * rows * a.yc_cs + t1 * kCaTile;
  const T* y = reinterpret_cast<const T*>(a.y) + (long long)b * rows * a.y_cs + t2 * kCaTile;
  const int tx = threadIdx.x % 8, ty = threadIdx.x / 8;
  float acc[PER][PER] = {};
  for (long long r = r0; r < r1; r += kCaRows) {
    // 8 channels (16 bytes) per load
    for (int i = threadIdx.x; i < kCaRows * (kCaTile / 8); i += 64) {
      const int rr = i / (kCaTile / 8), c8 = (i % (kCaTile / 8)) * 8;
      uint4 uq = make_uint4(0, 0, 0, 0), uk = make_uint4(0, 0, 0, 0);
      if (r + rr < r1) {
        uq = *reinterpret_cast<const uint4*>(yc + (r + rr) * a.yc_cs + c8);
        uk = *reinterpret_cast<const uint4*>(y + (r + rr) * a.y_cs + c8);
      }
      const uint32_t wq[4] = {uq.x, uq.y, uq.z, uq.w}, wk[4] = {uk.x, uk.y, uk.z, uk.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 fq = T16<T>::unpack2(wq[k]), fk = T16<T>::unpack2(wk[k]);
        *reinterpret_cast<float2*>(&s_q[rr][c8 + 2 * k]) = fq;
        *reinterpret_cast<float2*>(&s_k[rr][c8 + 2 * k]) = fk;
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int rr = 0; rr < kCaRows; ++rr) {
      float q[PER], k[PER];
#pragma unroll
      for (int i = 0; i < PER; i += 2) {
        const float2 fq = *reinterpret_cast<const float2*>(&s_q[rr][ty * PER + i]);
        const float2 fk = *reinterpret_cast<const float2*>(&s_k[rr][tx * PER + i]);
        q[i] = fq.x; q[i + 1] = fq.y; k[i] = fk.x; k[i + 1] = fk.y;
      }
#pragma unroll
      for (int i = 0; i < PER; ++i)
#pragma unroll
        for (int j = 0; j < PER; ++j) acc[i][j] = fmaf(q[i], k[j], acc[i][j]);
    }
    __syncthreads();
  }
  // deterministic: every row chunk owns its own partial [C][C] (summed in fixed order by the soft-max kernel)
  float* S = a.scratch + ((long long)b * gridDim.x + blockIdx.x) * a.c * a.c;
#pragma unroll
  for (int i = 0; i < PER; ++i)
#pragma unroll
    for (int j = 0; j < PER; ++j)
      S[(long long)(t1 * kCaTile + ty * PER + i) * a.c + t2 * kCaTile + tx * PER + j] = acc[i][j];
}

template <typename T>
__global__ void chan_softmax_kernel(float* S, int c, float scale, int nrows, int nchunks, int nsum, T* p16, int ld16, long long p16_bstride) {
  pdl_launch_dependents();
  pdl_wait();
  // one warp per row (b, c1) of S; first reduce the row-chunk partials [b][chunk][c1][:] into chunk 0, in order
  const int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x & 31;
  if (row >= nrows) return;
  const int b = row / c, c1 = row - b * c;
  float* p = S + ((long long)b * nchunks * c + c1) * c;
  for (int i = lane; i < c; i += 32) {
    float t = p[i];
    for (int k = 1; k < nsum; ++k) t += p[(long long)k * c * c + i];        // the first nsum chunks hold partial scores
    p[i] = t;
  }
  __syncwarp();
  float m = -INFINITY;
  for (int i = lane; i < c; i += 32) m = fmaxf(m, p[i] * scale);
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffff, m, o));
  float s = 0.f;
  for (int i = lane; i < c; i += 32) s += expf(p[i] * scale - m);
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffff, s, o);
  const float inv = 1.f / s;
  T* q = p16 ? p16 + (long long)b * p16_bstride + (long long)c1 * ld16 : nullptr;      // [b][c1][ld16]: the A operand of the tensor-core apply
  for (int i = lane; i < c; i += 32) {
    const float v = expf(p[i] * scale - m) * inv;
    p[i] = v;
    if (q) q[i] = T16<T>::from_f(v);
  }
  if (q)
    for (int i = c + lane; i < ld16; i += 32) q[i] = T16<T>::from_f(0.f);
}

constexpr int kCaPix = 64;
template <typename T, int kCaTile>
__global__ void __launch_bounds__(128) chan_apply_kernel(buctd_chan_att_args a) {
  pdl_launch_dependents();
  pdl_wait();
  // block: 64 valid pixels x kCaTile output channels (c1 tile); thread = 4 pixels x G channels.
  // dynamic smem: att[tile][C] f32, y[64][C+1] f32, out[tile][64] 16-bit
  constexpr int G = kCaTile / 8;
  extern __shared__ float s_dyn[];
  float* s_att = s_dyn;                         // [tile][C]
  float* s_y = s_dyn + kCaTile * a.c;           // [64][C+1]
  const int b = blockIdx.z, t1 = blockIdx.y;
  const int hw = a.h * a.w;
  const int ps = a.out_stride > 0 ? a.out_stride : hw;      // planar row stride
  const int p0 = blockIdx.x * kCaPix;
  const int cs1 = a.c + 1;
  const int nchunks = (int)(((long long)(a.h + 1) * (a.w + 1) + BUCTD_CHAN_ATT_ROWS_PER_CHUNK - 1) / BUCTD_CHAN_ATT_ROWS_PER_CHUNK);
  const float* att = a.scratch + ((long long)b * nchunks * a.c + t1 * kCaTile) * a.c;     // probabilities live in chunk 0
  if (a.drop_p > 0.f) {
    // training: nn.Dropout on the probabilities (hash mask keyed by (seed, crop, c1, c2), kept values / (1 - p))
    const uint32_t thr = drop_threshold(a.drop_p);
    const float ks = 1.f / (1.f - a.drop_p);
    const uint32_t seed = a.drop_seed ^ (a.seed_xor ? *a.seed_xor : 0u);
    for (int i = threadIdx.x; i < kCaTile * a.c; i += 128) {
      const int c1 = t1 * kCaTile + i / a.c, c2 = i % a.c;
      s_att[i] = drop_hash(seed, (uint32_t)b, (uint32_t)c1, (uint32_t)c2) >= thr ? att[i] * ks : 0.f;
    }
  } else {
    for (int i = threadIdx.x; i < kCaTile * a.c; i += 128) s_att[i] = att[i];
  }
  const T* y = reinterpret_cast<const T*>(a.y) + (long long)b * (a.h + 1) * (a.w + 1) * a.y_cs;
  for (int i = threadIdx.x; i < kCaPix * (a.c / 8); i += 128) {
    const int pp = i / (a.c / 8), c8 = (i % (a.c / 8)) * 8;
    const int p = p0 + pp;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (p < hw) {
      const int py = p / a.w, px = p % a.w;
      u = *reinterpret_cast<const uint4*>(y + ((long long)py * (a.w + 1) + px) * a.y_cs + c8);
    }
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = T16<T>::unpack2(w[k]);
      s_y[pp * cs1 + c8 + 2 * k] = f.x;
      s_y[pp * cs1 + c8 + 2 * k + 1] = f.y;
    }
  }
  __syncthreads();
  const int pg = threadIdx.x % 16, cg = threadIdx.x / 16;      // 16 pixel groups x 8 channel groups
  float acc[4][G] = {};
  for (int c2 = 0; c2 < a.c; ++c2) {
    float yv[4], av[G];
#pragma unroll
    for (int i = 0; i < 4; ++i) yv[i] = s_y[(pg * 4 + i) * cs1 + c2];
#pragma unroll
    for (int j = 0; j < G; ++j) av[j] = s_att[(cg * G + j) * a.c + c2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < G; ++j) acc[i][j] = fmaf(av[j], yv[i], acc[i][j]);
  }
  // 4 consecutive pixels of one channel = 8 bytes
#pragma unroll
  for (int j = 0; j < G; ++j) {
    const int p = p0 + pg * 4;
    T* o = reinterpret_cast<T*>(a.out) + ((long long)b * a.c + t1 * kCaTile + cg * G + j) * ps + p;
    if (a.out_padded) {                 // pixel p = (py, px) lands at py (W + 1) + px
      T* ob = o - p;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (p + i < hw) ob[(p + i) / a.w * (a.w + 1) + (p + i) % a.w] = T16<T>::from_f(acc[i][j]);
    } else if (p + 4 <= hw && ((ps & 3) == 0)) {
      uint2 u;
      u.x = T16<T>::pack2(acc[0][j], acc[1][j]);
      u.y = T16<T>::pack2(acc[2][j], acc[3][j]);
      *reinterpret_cast<uint2*>(o) = u;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (p + i < hw) o[i] = T16<T>::from_f(acc[i][j]);
    }
  }
}

template <typename T, int TILE>
static int chan_att_run(const buctd_chan_att_args& a, cudaStream_t st) {
  const long long rows = (long long)(a.h + 1) * (a.w + 1);
  const int tiles = a.c / TILE;
  const int chunk = BUCTD_CHAN_ATT_ROWS_PER_CHUNK;
  dim3 g1((unsigned)((rows + chunk - 1) / chunk), tiles * tiles, a.batch);
  const size_t smem = sizeof(float) * ((size_t)TILE * a.c + (size_t)kCaPix * (a.c + 1));
  dim3 g3(ceil_div(a.h * a.w, kCaPix), tiles, a.batch);
  const float scale = 1.f / sqrtf((float)(a.h * a.w));
  // scores: S[c1][c2] = sum over the pixel rows of yc[r][c1] y[r][c2] -- a contraction over the rows of two act16 tensors,
  // i.e. buctd_gemm16 with both operands MN-major (tcgen05; the CUDA-core chan_scores_kernel took 75 us on branch 0).  The
  // rows are split over nsplit CTAs per crop whose float32 partials land in the first chunks of the scratch and are summed
  // in index order by the soft-max kernel (deterministic).  BUCTD_CHAN_SCORES_TC=0: the CUDA-core kernel.
  static const bool use_tc = !(getenv("BUCTD_CHAN_SCORES_TC") && atoi(getenv("BUCTD_CHAN_SCORES_TC")) == 0);
  int nsum = (int)g1.x;
  if (use_tc && a.y_cs % 8 == 0 && a.yc_cs % 8 == 0 && ((((uintptr_t)a.y | (uintptr_t)a.yc)) & 15) == 0) {
    int nsplit = 148 / a.batch;
    const int cap = (int)g1.x > 1 ? (int)g1.x - 1 : 1;          // leave the last chunk free (16-bit probabilities of the apply)
    nsplit = nsplit < 1 ? 1 : (nsplit > cap ? cap : nsplit);
    buctd_gemm16_args g;
    memset(&g, 0, sizeof(g));
    g.dtype = a.dtype; g.a_mn_major = 1; g.b_mn_major = 1;
    g.m = a.c; g.n = a.c; g.k = (int)rows;
    g.nbatch = a.batch; g.nsplit = nsplit; g.ntaps = 1; g.b_rows = (int)rows;
    g.lda = a.yc_cs; g.ldb = a.y_cs; g.ldd = a.c; g.out_f32 = 1; g.accumulate = 0;
    g.a_batch_stride = rows * a.yc_cs; g.b_batch_stride = rows * a.y_cs;
    g.d_t_stride = 0; g.d_s_stride = (long long)a.c * a.c; g.d_b_stride = (long long)g1.x * a.c * a.c;
    g.alpha = 1.f;
    g.a = a.yc; g.b = a.y; g.d = a.scratch;
    const int rc = launch_gemm16(g, st);
    if (rc) return rc;
    nsum = nsplit;
  } else {
    launch_pdl(chan_scores_kernel<T, TILE>, g1, dim3(64), 0, st, a, chunk);
  }
  // apply: out[c1][r] = sum_c2 att[c1][c2] y[r][c2] over the PADDED pixel rows r -- a plain K-major x K-major product per crop
  // (buctd_gemm16 again) once the probabilities exist in 16 bit; they go into the last scratch chunk (free: nsum < nchunks)
  const int nchunks = (int)g1.x;
  const int ld16 = (a.c + 7) / 8 * 8;
  const bool tc_apply = a.out_padded && nsum < nchunks && a.drop_p <= 0.f && a.out_stride >= (int)rows && a.out_stride % 8 == 0 &&
                        (((uintptr_t)a.out) & 15) == 0 && 2 * a.c >= ld16 && !(getenv("BUCTD_CHAN_APPLY_TC") && atoi(getenv("BUCTD_CHAN_APPLY_TC")) == 0);
  T* p16 = tc_apply ? reinterpret_cast<T*>(a.scratch + (size_t)(nchunks - 1) * a.c * a.c) : nullptr;
  launch_pdl(chan_softmax_kernel<T>, dim3(ceil_div(a.batch * a.c, 8)), dim3(256), 0, st, a.scratch, a.c, scale, a.batch * a.c, nchunks, nsum,
             p16, ld16, (long long)2 * nchunks * a.c * a.c);
  if (tc_apply) {
    buctd_gemm16_args g;
    memset(&g, 0, sizeof(g));
    g.dtype = a.dtype; g.a_mn_major = 0; g.b_mn_major = 0;
    g.m = a.c; g.n = (int)rows; g.k = a.c;
    g.nbatch = a.batch; g.nsplit = 1; g.ntaps = 1;
    g.lda = ld16; g.ldb = a.y_cs; g.ldd = a.out_stride; g.out_f32 = 0; g.accumulate = 0;
    g.a_batch_stride = (long long)2 * nchunks * a.c * a.c; g.b_batch_stride = rows * a.y_cs;
    g.d_b_stride = (long long)a.c * a.out_stride;
    g.alpha = 1.f;
    g.a = p16; g.b = a.y; g.d = a.out;
    const int rc = launch_gemm16(g, st);
    if (rc) return rc;
    BUCTD_CUDA(cudaGetLastError());
    return BUCTD_OK;
  }
  BUCTD_CUDA(cudaFuncSetAttribute(chan_apply_kernel<T, TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  launch_pdl(chan_apply_kernel<T, TILE>, g3, dim3(128), smem, st, a);
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}

int launch_chan_att(const buctd_chan_att_args& a, cudaStream_t st, int* nlaunch) {
  BUCTD_CHECK(a.c % 48 == 0 || a.c % 32 == 0, BUCTD_ERR_INVALID, "chan_att: C=%d must be a multiple of 48 or 32", a.c);
  BUCTD_CHECK(a.y && a.yc && a.scratch && a.out, BUCTD_ERR_INVALID, "chan_att: null pointer");
  BUCTD_CHECK(a.c <= 384, BUCTD_ERR_INVALID, "chan_att: C=%d too large for the shared-memory tile", a.c);
  BUCTD_CHECK(!a.out_padded || a.out_stride >= (a.h + 1) * (a.w + 1), BUCTD_ERR_INVALID, "chan_att: out_padded needs out_stride >= (H+1)(W+1)");
  if (nlaunch) *nlaunch = 3;
  const bool t48 = a.c % 48 == 0;
  if (a.dtype == BUCTD_F16) return t48 ? chan_att_run<__half, 48>(a, st) : chan_att_run<__half, 32>(a, st);
  return t48 ? chan_att_run<__nv_bfloat16, 48>(a, st) : chan_att_run<__nv_bfloat16, 32>(a, st);
}

// =================================================================================================
// DAModule combine: out = x + (p_out + c_out)   (channel_only: out = x * c_out)
// lib/models/pose_hrnet_coam.py:716-724.  c_out is planar [B][C][HW]: 32x32 smem transpose.
// =================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) da_combine_kernel(buctd_da_combine_args a) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_c[32][33];
  const int b = blockIdx.z;
  const int hw = a.h * a.w;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;   // 32 x 8
  const int ps = a.c_stride > 0 ? a.c_stride : hw;          // planar row stride of c_out
  const T* cout = reinterpret_cast<const T*>(a.c_out) + (long long)b * a.c * ps;
  for (int cc = ty; cc < 32; cc += 8) {
    int c = c0 + cc, p = p0 + tx;
    s_c[cc][tx] = (c < a.c && p < hw) ? T16<T>::to_f(cout[(long long)c * ps + p]) : 0.f;
  }
  __syncthreads();
  for (int pp = ty; pp < 32; pp += 8) {
    const int p = p0 + pp, c = c0 + tx;
    if (p >= hw || c >= a.c) continue;
    const int py = p / a.w, px = p % a.w;
    const long long r = ((long long)b * (a.h + 1) + py) * (a.w + 1) + px;
    const float xv = T16<T>::to_f(reinterpret_cast<const T*>(a.x)[r * a.x_cs + c]);
    float v;
    if (a.channel_only) v = xv * s_c[tx][pp];
    else v = xv + (T16<T>::to_f(reinterpret_cast<const T*>(a.p_out)[r * a.p_cs + c]) + s_c[tx][pp]);
    reinterpret_cast<T*>(a.out)[r * a.out_cs + c] = T16<T>::from_f(v);
  }
}

// Vector variant (C, channel strides multiples of 8): tile = 64 pixels x 32 channels; the planar c_out tile goes through
// shared memory, x / p_out / out move as 16-byte groups of 8 channels.
template <typename T>
__global__ void __launch_bounds__(256) da_combine_vec_kernel(buctd_da_combine_args a) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_c[32][65];
  const int b = blockIdx.z;
  const int hw = a.h * a.w;
  const int p0 = blockIdx.x * 64, c0 = blockIdx.y * 32;
  const int ps = a.c_stride > 0 ? a.c_stride : hw;
  const T* cout = reinterpret_cast<const T*>(a.c_out) + (size_t)b * a.c * ps;
  for (int i = threadIdx.x; i < 32 * 64; i += 256) {
    const int cc = i >> 6, pp = i & 63;
    const int c = c0 + cc, p = p0 + pp;
    s_c[cc][pp] = (c < a.c && p < hw) ? T16<T>::to_f(cout[(size_t)c * ps + p]) : 0.f;
  }
  __syncthreads();
  const int pp = threadIdx.x >> 2, cg = threadIdx.x & 3;
  const int p = p0 + pp, c = c0 + cg * 8;
  if (p >= hw || c >= a.c) return;
  const int py = p / a.w, px = p - py * a.w;
  const size_t r = ((size_t)b * (a.h + 1) + py) * (a.w + 1) + px;
  const uint4 ux = *reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(a.x) + r * a.x_cs + c);
  const uint32_t wx[4] = {ux.x, ux.y, ux.z, ux.w};
  float v[8];
  if (a.channel_only) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = T16<T>::unpack2(wx[k]);
      v[2 * k] = f.x * s_c[cg * 8 + 2 * k][pp];
      v[2 * k + 1] = f.y * s_c[cg * 8 + 2 * k + 1][pp];
    }
  } else {
    const uint4 up = *reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(a.p_out) + r * a.p_cs + c);
    const uint32_t wp[4] = {up.x, up.y, up.z, up.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = T16<T>::unpack2(wx[k]), g = T16<T>::unpack2(wp[k]);
      v[2 * k] = f.x + (g.x + s_c[cg * 8 + 2 * k][pp]);
      v[2 * k + 1] = f.y + (g.y + s_c[cg * 8 + 2 * k + 1][pp]);
    }
  }
  uint4 uo;
  uo.x = T16<T>::pack2(v[0], v[1]); uo.y = T16<T>::pack2(v[2], v[3]);
  uo.z = T16<T>::pack2(v[4], v[5]); uo.w = T16<T>::pack2(v[6], v[7]);
  *reinterpret_cast<uint4*>(reinterpret_cast<T*>(a.out) + r * a.out_cs + c) = uo;
}

int launch_da_combine(const buctd_da_combine_args& a, cudaStream_t st) {
  BUCTD_CHECK(a.x && a.c_out && a.out && (a.channel_only || a.p_out), BUCTD_ERR_INVALID, "da_combine: null pointer");
  const bool vec = a.c % 8 == 0 && a.x_cs % 8 == 0 && a.out_cs % 8 == 0 && (a.channel_only || a.p_cs % 8 == 0) &&
                   ((((uintptr_t)a.x | (uintptr_t)a.out | (uintptr_t)a.p_out) & 15) == 0);
  if (vec) {
    dim3 grid(ceil_div(a.h * a.w, 64), ceil_div(a.c, 32), a.batch);
    if (a.dtype == BUCTD_F16) launch_pdl(da_combine_vec_kernel<__half>, grid, dim3(256), 0, st, a);
    else launch_pdl(da_combine_vec_kernel<__nv_bfloat16>, grid, dim3(256), 0, st, a);
    BUCTD_CUDA(cudaGetLastError());
    return BUCTD_OK;
  }
  dim3 grid(ceil_div(a.h * a.w, 32), ceil_div(a.c, 32), a.batch);
  if (a.dtype == BUCTD_F16) launch_pdl(da_combine_kernel<__half>, grid, dim3(256), 0, st, a);
  else launch_pdl(da_combine_kernel<__nv_bfloat16>, grid, dim3(256), 0, st, a);
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}

// =================================================================================================
// Layout converters (tests, and the nchw32 -> act16 entry of tensors produced outside the plan)
// =================================================================================================
template <typename T>
__global__ void act16_to_nchw32_kernel(const T* act, int B, int H, int W, int C, int cs, float* out) {
  const long long total = (long long)B * C * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int x = (int)(i % W);
    long long t = i / W;
    int y = (int)(t % H); t /= H;
    int c = (int)(t % C);
    int b = (int)(t / C);
    out[i] = T16<T>::to_f(act[(((long long)b * (H + 1) + y) * (W + 1) + x) * cs + c]);
  }
}
template <typename T>
__global__ void nchw32_to_act16_kernel(const float* x, int B, int xc, int c_off, int C, int H, int W, int cs, T* out) {
  const long long total = (long long)B * (H + 1) * (W + 1) * cs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % cs);
    long long t = i / cs;
    int xx = (int)(t % (W + 1)); t /= (W + 1);
    int y = (int)(t % (H + 1));
    int b = (int)(t / (H + 1));
    float v = 0.f;
    if (c < C && xx < W && y < H) v = x[(((long long)b * xc + c_off + c) * H + y) * W + xx];
    out[i] = T16<T>::from_f(v);
  }
}

int launch_act16_to_nchw32(int dtype, const void* act, int B, int H, int W, int C, int cs, float* out, cudaStream_t st) {
  BUCTD_CHECK(act && out, BUCTD_ERR_INVALID, "act16_to_nchw32: null pointer");
  const long long total = (long long)B * C * H * W;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  if (dtype == BUCTD_F16) act16_to_nchw32_kernel<<<blocks, 256, 0, st>>>((const __half*)act, B, H, W, C, cs, out);
  else act16_to_nchw32_kernel<<<blocks, 256, 0, st>>>((const __nv_bfloat16*)act, B, H, W, C, cs, out);
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}
int launch_nchw32_to_act16(int dtype, const float* x, int B, int xc, int c_off, int C, int H, int W, int cs, void* out,
                           cudaStream_t st) {
  BUCTD_CHECK(x && out && C <= cs, BUCTD_ERR_INVALID, "nchw32_to_act16: bad args");
  const long long total = (long long)B * (H + 1) * (W + 1) * cs;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  if (dtype == BUCTD_F16) nchw32_to_act16_kernel<<<blocks, 256, 0, st>>>(x, B, xc, c_off, C, H, W, cs, (__half*)out);
  else nchw32_to_act16_kernel<<<blocks, 256, 0, st>>>(x, B, xc, c_off, C, H, W, cs, (__nv_bfloat16*)out);
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}

}  // namespace buctd

namespace buctd {
// =================================================================================================
// preNet (lib/models/pose_hrnet.py:431-442, 452-458): composed 9x9 (rgb) + 7x7 (cond) 3->3 convs.
// Block = 32x8 output pixels; rgb/cond halo tiles in shared memory; the (warp-uniform almost
// everywhere) weight variant is read through the read-only cache.
// =================================================================================================
__device__ __forceinline__ int prenet_variant(int p, int n) {
  const int lo = 3 - p;             // rows cut at the top/left
  if (lo > 0) return lo;
  const int hi = p + 3 - (n - 1);   // rows cut at the bottom/right
  if (hi > 0) return 3 + hi;
  return 0;
}

__global__ void __launch_bounds__(256) prenet_kernel(buctd_prenet_args a) {
  constexpr int TW = 32, TH = 8;
  __shared__ float s_rgb[3][TH + 8][TW + 8];
  __shared__ float s_cond[3][TH + 6][TW + 6 + 2];
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const float* xb = a.x + (long long)b * a.x_channels * a.h * a.w;
  for (int i = threadIdx.x; i < 3 * (TH + 8) * (TW + 8); i += 256) {
    const int c = i / ((TH + 8) * (TW + 8)), r = i % ((TH + 8) * (TW + 8));
    const int yy = r / (TW + 8), xx = r % (TW + 8);
    const int gy = y0 + yy - 4, gx = x0 + xx - 4;
    s_rgb[c][yy][xx] = (gy >= 0 && gy < a.h && gx >= 0 && gx < a.w) ? xb[((long long)c * a.h + gy) * a.w + gx] : 0.f;
  }
  for (int i = threadIdx.x; i < 3 * (TH + 6) * (TW + 6); i += 256) {
    const int c = i / ((TH + 6) * (TW + 6)), r = i % ((TH + 6) * (TW + 6));
    const int yy = r / (TW + 6), xx = r % (TW + 6);
    const int gy = y0 + yy - 3, gx = x0 + xx - 3;
    s_cond[c][yy][xx] = (gy >= 0 && gy < a.h && gx >= 0 && gx < a.w) ? xb[((long long)(3 + c) * a.h + gy) * a.w + gx] : 0.f;
  }
  __syncthreads();
  const int tx = threadIdx.x % TW, ty = threadIdx.x / TW;
  const int x = x0 + tx, y = y0 + ty;
  if (x >= a.w || y >= a.h) return;
  const int v = prenet_variant(y, a.h) * 7 + prenet_variant(x, a.w);
  const float* we = a.w_eff + (long long)v * 3 * 3 * 81;
  float acc[3];
#pragma unroll
  for (int co = 0; co < 3; ++co) acc[co] = __ldg(a.b_eff + v * 3 + co);
  for (int ci = 0; ci < 3; ++ci)
    for (int ky = 0; ky < 9; ++ky)
#pragma unroll
      for (int kx = 0; kx < 9; ++kx) {
        const float in = s_rgb[ci][ty + ky][tx + kx];
        const int o = ci * 81 + ky * 9 + kx;
        acc[0] = fmaf(in, __ldg(we + o), acc[0]);
        acc[1] = fmaf(in, __ldg(we + 243 + o), acc[1]);
        acc[2] = fmaf(in, __ldg(we + 486 + o), acc[2]);
      }
  for (int ci = 0; ci < 3; ++ci)
    for (int ky = 0; ky < 7; ++ky)
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) {
        const float in = s_cond[ci][ty + ky][tx + kx];
        const int o = ci * 49 + ky * 7 + kx;
        acc[0] = fmaf(in, __ldg(a.wc + o), acc[0]);
        acc[1] = fmaf(in, __ldg(a.wc + 147 + o), acc[1]);
        acc[2] = fmaf(in, __ldg(a.wc + 294 + o), acc[2]);
      }
#pragma unroll
  for (int co = 0; co < 3; ++co) a.out[(((long long)b * 3 + co) * a.h + y) * a.w + x] = acc[co];
}

int launch_prenet(const buctd_prenet_args& a, cudaStream_t st) {
  BUCTD_CHECK(a.x && a.w_eff && a.b_eff && a.wc && a.out, BUCTD_ERR_INVALID, "prenet: null pointer");
  BUCTD_CHECK(a.x_channels >= 6 && a.h >= 7 && a.w >= 7, BUCTD_ERR_INVALID, "prenet: need >= 6 input channels and H,W >= 7");
  if (a.batch == 0) return BUCTD_OK;
  prenet_kernel<<<dim3(ceil_div(a.w, 32), ceil_div(a.h, 8), a.batch), 256, 0, st>>>(a);
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}
}  // namespace buctd
