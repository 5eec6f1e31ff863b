// Heat-map decode (argmax + quarter-pixel refine + inverse crop transform) and the condition
// rasteriser.  Both are HBM-bound: decode reads each heat-map exactly once, the rasteriser writes
// each condition plane exactly once.
#include "common.cuh"
#include "internal.h"

namespace buctd {

// =================================================================================================
// decode: get_max_preds + get_final_preds, lib/core/inference.py:19-87; transform_preds,
// lib/utils/transforms.py:78-118 (inv=1, rot=0: a similarity with factor 200*scale[0]/W).
// One block per (b,k).  Ties: the FIRST flat index wins (np.argmax).  NaN inputs are not defined.
// =================================================================================================
__global__ void __launch_bounds__(256) decode_kernel(buctd_decode_args a) {
  const int bk = blockIdx.x;
  const int hw = a.h * a.w;
  const float* hm = a.heatmaps + (long long)bk * hw;
  float best = -INFINITY;
  int bidx = 0x7fffffff;
  // 128-bit loads when the plane is 16-byte aligned
  if ((hw & 3) == 0 && ((reinterpret_cast<uintptr_t>(hm) & 15) == 0)) {
    const float4* h4 = reinterpret_cast<const float4*>(hm);
    for (int i = threadIdx.x; i < hw / 4; i += blockDim.x) {
      const float4 v = h4[i];
      const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (e[j] > best) { best = e[j]; bidx = 4 * i + j; }   // strict >: first index within the thread's ascending scan
    }
  } else {
    for (int i = threadIdx.x; i < hw; i += blockDim.x) {
      const float v = hm[i];
      if (v > best) { best = v; bidx = i; }
    }
  }
  // warp then block reduction on (value desc, index asc)
  for (int o = 16; o; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffff, best, o);
    const int oi = __shfl_xor_sync(0xffffffff, bidx, o);
    if (ov > best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
  }
  __shared__ float s_v[8];
  __shared__ int s_i[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_v[warp] = best; s_i[warp] = bidx; }
  __syncthreads();
  if (threadIdx.x != 0) return;
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
    if (s_v[w] > best || (s_v[w] == best && s_i[w] < bidx)) { best = s_v[w]; bidx = s_i[w]; }
  if (bidx == 0x7fffffff) bidx = 0;      // all -inf / NaN plane: np.argmax gives 0

  float cx = (float)(bidx % a.w);
  float cy = floorf((float)bidx / (float)a.w);
  if (!(best > 0.0f)) { cx = 0.f; cy = 0.f; }             // pred_mask = maxvals > 0
  if (a.idx) a.idx[bk] = bidx;
  if (a.coords) { a.coords[2 * bk] = cx; a.coords[2 * bk + 1] = cy; }
  a.maxvals[bk] = best;
  if (a.post_process) {
    const int px = (int)floorf(cx + 0.5f), py = (int)floorf(cy + 0.5f);
    if (1 < px && px < a.w - 1 && 1 < py && py < a.h - 1) {
      const float dx = hm[py * a.w + px + 1] - hm[py * a.w + px - 1];
      const float dy = hm[(py + 1) * a.w + px] - hm[(py - 1) * a.w + px];
      cx += (dx > 0.f ? 0.25f : (dx < 0.f ? -0.25f : 0.f));
      cy += (dy > 0.f ? 0.25f : (dy < 0.f ? -0.25f : 0.f));
    }
  }
  if (a.center != nullptr && a.scale != nullptr) {
    const int b = bk / a.k;
    // the reference builds the 2x3 matrix in float64 from float32 points, applies it in float64,
    // then stores float32
    const double f = (double)(a.scale[2 * b] * 200.0f) / (double)a.w;
    const double tx = (double)a.center[2 * b] - f * (double)a.w * 0.5;
    const double ty = (double)a.center[2 * b + 1] - f * (double)a.h * 0.5;
    a.preds[2 * bk] = (float)(f * (double)cx + tx);
    a.preds[2 * bk + 1] = (float)(f * (double)cy + ty);
  } else {
    a.preds[2 * bk] = cx;
    a.preds[2 * bk + 1] = cy;
  }
}

int launch_decode(const buctd_decode_args& a, cudaStream_t st) {
  BUCTD_CHECK(a.batch >= 0 && a.k > 0 && a.h > 0 && a.w > 0, BUCTD_ERR_INVALID, "decode: bad shape");
  if (a.batch == 0) return BUCTD_OK;      // empty batch: nothing to do (pointers may be null)
  BUCTD_CHECK(a.heatmaps && a.preds && a.maxvals, BUCTD_ERR_INVALID, "decode: null pointer");
  BUCTD_CHECK((a.center == nullptr) == (a.scale == nullptr), BUCTD_ERR_INVALID, "decode: center and scale go together");
  decode_kernel<<<a.batch * a.k, 256, 0, st>>>(a);
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}

// =================================================================================================
// Head + decode fused (when the heat-maps themselves are not wanted): final_layer (1x1, C -> K, bias; pose_hrnet_coam.py:565)
// evaluated on CUDA cores straight from the act16 feature map, running (max, first index) per joint, then the quarter-pixel
// offset from the four re-evaluated neighbours and the crop transform -- the [B][K][H][W] float32 heat-maps (470 KB per
// crop written and read back) never exist.  Two launches: a scan (thread = pixel, all K joints, float32 accumulation in
// channel order, per-block partial arg-max) and a finish (thread = joint).  Same decode arithmetic as decode_kernel.  The tap-GEMM head accumulates the same 16-bit products in the
// tensor core's own order, so values agree to float32 rounding, not bit for bit (documented at the entry point).
// =================================================================================================
constexpr int kHdMaxK = 32;
constexpr int kHdPix = 256;              // pixels per block of the scan kernel (one per thread)

template <typename T, int CN>
__device__ __forceinline__ void hd_load_row(const T* row, float (&x)[CN]) {
#pragma unroll
  for (int c0 = 0; c0 < CN; c0 += 8) {
    const uint4 u = *reinterpret_cast<const uint4*>(row + c0);
    float2 t;
    t = T16<T>::unpack2(u.x); x[c0] = t.x; x[c0 + 1] = t.y;
    t = T16<T>::unpack2(u.y); x[c0 + 2] = t.x; x[c0 + 3] = t.y;
    t = T16<T>::unpack2(u.z); x[c0 + 4] = t.x; x[c0 + 5] = t.y;
    t = T16<T>::unpack2(u.w); x[c0 + 6] = t.x; x[c0 + 7] = t.y;
  }
}
template <int CN>
__device__ __forceinline__ float hd_dot(const float (&x)[CN], const float* w, float bias) {
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < CN; c += 4) {
    const float4 w4 = *reinterpret_cast<const float4*>(w + c);
    acc = fmaf(x[c], w4.x, acc); acc = fmaf(x[c + 1], w4.y, acc);
    acc = fmaf(x[c + 2], w4.z, acc); acc = fmaf(x[c + 3], w4.w, acc);
  }
  return acc + bias;
}

// pass 1: grid (ceil(HW / 256), B); thread = pixel; per joint the block's (max, first index) -> scratch[b][blk][k]
template <typename T, int CN>
__global__ void __launch_bounds__(kHdPix) head_scan_kernel(buctd_head_decode_args a) {
  __shared__ __align__(16) float s_w[kHdMaxK][CN];
  __shared__ float s_b[kHdMaxK];
  __shared__ float s_v[kHdPix / 32][kHdMaxK];
  __shared__ int s_i[kHdPix / 32][kHdMaxK];
  const int b = blockIdx.y, K = a.k;
  for (int i = threadIdx.x; i < K * CN; i += blockDim.x) {
    const int k = i / CN, c = i - k * CN;
    s_w[k][c] = T16<T>::to_f(reinterpret_cast<const T*>(a.weight)[(long long)k * a.w_cs + c]);
  }
  if ((int)threadIdx.x < K) s_b[threadIdx.x] = a.bias ? a.bias[threadIdx.x] : 0.f;
  __syncthreads();
  const int hw = a.h * a.w;
  const int p = blockIdx.x * kHdPix + threadIdx.x;
  const bool live = p < hw;
  float x[CN];
  if (live) {
    const int py = p / a.w, px = p - py * a.w;
    hd_load_row<T, CN>(reinterpret_cast<const T*>(a.feat) + (((long long)b * (a.h + 1) + py) * (a.w + 1) + px) * a.a_cs, x);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = 0; k < K; ++k) {
    float bv = live ? hd_dot<CN>(x, s_w[k], s_b[k]) : -INFINITY;
    int bi = live ? p : 0x7fffffff;
    for (int o = 16; o; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffff, bv, o);
      const int oi = __shfl_xor_sync(0xffffffff, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { s_v[warp][k] = bv; s_i[warp][k] = bi; }
  }
  __syncthreads();
  if ((int)threadIdx.x < K) {
    const int k = threadIdx.x;
    float bv = s_v[0][k];
    int bi = s_i[0][k];
    for (int w = 1; w < kHdPix / 32; ++w)
      if (s_v[w][k] > bv || (s_v[w][k] == bv && s_i[w][k] < bi)) { bv = s_v[w][k]; bi = s_i[w][k]; }
    float2* out = reinterpret_cast<float2*>(a.scratch) + ((long long)b * gridDim.x + blockIdx.x) * K + k;
    *out = make_float2(bv, __int_as_float(bi));
  }
}

// pass 2: one block per crop, thread = joint: reduce the block partials in order, re-evaluate the four neighbours, decode
template <typename T, int CN>
__global__ void __launch_bounds__(32) head_finish_kernel(buctd_head_decode_args a, int nblk) {
  const int b = blockIdx.x, k = threadIdx.x;
  if (k >= a.k) return;
  const float2* part = reinterpret_cast<const float2*>(a.scratch) + (long long)b * nblk * a.k + k;
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int j = 0; j < nblk; ++j) {
    const float2 v = part[(long long)j * a.k];
    const int oi = __float_as_int(v.y);
    if (v.x > bv || (v.x == bv && oi < bi)) { bv = v.x; bi = oi; }
  }
  if (bi == 0x7fffffff) bi = 0;
  const int bk = b * a.k + k;
  float cx = (float)(bi % a.w);
  float cy = floorf((float)bi / (float)a.w);
  if (!(bv > 0.0f)) { cx = 0.f; cy = 0.f; }
  if (a.idx) a.idx[bk] = bi;
  a.maxvals[bk] = bv;
  if (a.post_process) {
    const int px = (int)floorf(cx + 0.5f), py = (int)floorf(cy + 0.5f);
    if (1 < px && px < a.w - 1 && 1 < py && py < a.h - 1) {
      float w[CN];
#pragma unroll
      for (int c = 0; c < CN; ++c) w[c] = T16<T>::to_f(reinterpret_cast<const T*>(a.weight)[(long long)k * a.w_cs + c]);
      const float bias = a.bias ? a.bias[k] : 0.f;
      const T* feat = reinterpret_cast<const T*>(a.feat) + (long long)b * (a.h + 1) * (a.w + 1) * a.a_cs;
      auto eval = [&](int yy, int xx) {
        float x[CN];
        hd_load_row<T, CN>(feat + ((long long)yy * (a.w + 1) + xx) * a.a_cs, x);
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < CN; ++c) acc = fmaf(x[c], w[c], acc);
        return acc + bias;
      };
      const float dx = eval(py, px + 1) - eval(py, px - 1);
      const float dy = eval(py + 1, px) - eval(py - 1, px);
      cx += (dx > 0.f ? 0.25f : (dx < 0.f ? -0.25f : 0.f));
      cy += (dy > 0.f ? 0.25f : (dy < 0.f ? -0.25f : 0.f));
    }
  }
  if (a.center != nullptr && a.scale != nullptr) {
    const double f = (double)(a.scale[2 * b] * 200.0f) / (double)a.w;
    const double tx = (double)a.center[2 * b] - f * (double)a.w * 0.5;
    const double ty = (double)a.center[2 * b + 1] - f * (double)a.h * 0.5;
    a.preds[2 * bk] = (float)(f * (double)cx + tx);
    a.preds[2 * bk + 1] = (float)(f * (double)cy + ty);
  } else {
    a.preds[2 * bk] = cx;
    a.preds[2 * bk + 1] = cy;
  }
}

template <typename T, int CN>
static void head_decode_launch(const buctd_head_decode_args& a, cudaStream_t st) {
  const int nblk = ceil_div(a.h * a.w, kHdPix);
  head_scan_kernel<T, CN><<<dim3(nblk, a.batch), kHdPix, 0, st>>>(a);
  head_finish_kernel<T, CN><<<a.batch, 32, 0, st>>>(a, nblk);
}

int launch_head_decode(const buctd_head_decode_args& a, cudaStream_t st) {
  BUCTD_CHECK(a.batch >= 0 && a.k > 0 && a.k <= kHdMaxK && (a.c == 32 || a.c == 48 || a.c == 64) && a.h > 0 && a.w > 0, BUCTD_ERR_INVALID,
              "head_decode: K <= %d and C in {32, 48, 64} (got K %d, C %d)", kHdMaxK, a.k, a.c);
  if (a.batch == 0) return BUCTD_OK;
  BUCTD_CHECK(a.feat && a.weight && a.preds && a.maxvals && a.scratch && a.a_cs % 8 == 0 && a.w_cs >= a.c, BUCTD_ERR_INVALID,
              "head_decode: bad pointers / strides");
  BUCTD_CHECK((a.center == nullptr) == (a.scale == nullptr), BUCTD_ERR_INVALID, "head_decode: center and scale go together");
  const bool h16 = a.dtype == BUCTD_F16;
  switch (a.c) {
    case 32: if (h16) head_decode_launch<__half, 32>(a, st); else head_decode_launch<__nv_bfloat16, 32>(a, st); break;
    case 48: if (h16) head_decode_launch<__half, 48>(a, st); else head_decode_launch<__nv_bfloat16, 48>(a, st); break;
    default: if (h16) head_decode_launch<__half, 64>(a, st); else head_decode_launch<__nv_bfloat16, 64>(a, st); break;
  }
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}

// =================================================================================================
// rasteriser: lib/dataset/JointsDataset.py:457-543.  cv2.GaussianBlur(15x15, sigma 0 -> 2.6,
// BORDER_REFLECT_101) of an impulse image is separable; with reflection the 1-D response of an
// impulse at p is  g_p[x] = sum_t k[t] * [reflect101(x + t - 7) == p].  Arithmetic in float64 like
// the reference (np.zeros is float64), stored as float32.
// =================================================================================================
__device__ __forceinline__ int reflect101(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

__device__ void gauss15(double* k) {
  const double sigma = 0.3 * ((15 - 1) * 0.5 - 1) + 0.8;
  double s = 0;
  for (int i = 0; i < 15; ++i) {
    const double r = i - 7;
    k[i] = exp(-(r * r) / (2.0 * sigma * sigma));
    s += k[i];
  }
  for (int i = 0; i < 15; ++i) k[i] /= s;
}

__device__ __forceinline__ double folded(const double* k, int x, int p, int n) {
  double v = 0;
  for (int t = 0; t < 15; ++t)
    if (reflect101(x + t - 7, n) == p) v += k[t];
  return v;
}

__device__ __forceinline__ bool kpt_pixel(const float* kp, int w, int h, int* px, int* py) {
  const int x = (int)kp[0], y = (int)kp[1];       // .astype(int): truncation toward zero
  if (0 < x && x < w && 0 < y && y < h) { *px = x - 1; *py = y - 1; return true; }
  return false;
}

// mode 0 (stacked): grid (ceil(H/32), K, B); each block writes 32 rows of one plane.
__global__ void __launch_bounds__(256) raster_stacked_kernel(buctd_raster_args a) {
  __shared__ double s_k[15];
  __shared__ double s_gx[512];
  __shared__ double s_gy[32];
  __shared__ double s_mx, s_my;
  const int k = blockIdx.y, b = blockIdx.z;
  const int y0 = blockIdx.x * 32;
  float* out = a.out + (((long long)b * a.out_channels + a.c_off + k) * a.h) * a.w;
  int px, py;
  const bool on = kpt_pixel(a.kpts + ((long long)b * a.k + k) * a.kstride, a.w, a.h, &px, &py);
  if (threadIdx.x == 0) gauss15(s_k);
  __syncthreads();
  if (on) {
    for (int x = threadIdx.x; x < a.w; x += blockDim.x) s_gx[x] = folded(s_k, x, px, a.w);
    if (threadIdx.x < 32) s_gy[threadIdx.x] = (y0 + threadIdx.x < a.h) ? folded(s_k, y0 + threadIdx.x, py, a.h) : 0.0;
    if (threadIdx.x == 0) {
      // peak of the blurred impulse = max_x gx * max_y gy, attained inside the 15-tap window
      double mx = 0, my = 0;
      for (int t = -7; t <= 7; ++t) {
        if (px + t >= 0 && px + t < a.w) mx = fmax(mx, folded(s_k, px + t, px, a.w));
        if (py + t >= 0 && py + t < a.h) my = fmax(my, folded(s_k, py + t, py, a.h));
      }
      s_mx = mx; s_my = my;
    }
  }
  __syncthreads();
  const double am = on ? (s_my * (s_mx * 255.0)) : 0.0;
  const double den = am / 255.0;
  for (int i = threadIdx.x; i < 32 * a.w; i += blockDim.x) {
    const int yy = i / a.w, x = i % a.w;
    if (y0 + yy >= a.h) break;
    float v = 0.f;
    if (on) {
      const double g = s_gy[yy] * (s_gx[x] * 255.0);     // column pass of (row pass of the impulse)
      v = (float)(g / den);
    }
    out[(long long)(y0 + yy) * a.w + x] = v;
  }
}

// modes 1/2: pass 1 accumulates the un-normalised image maximum per crop (as ordered int64 bits of
// a non-negative double) into scratch; pass 2 recomputes and normalises.
template <bool WRITE>
__global__ void __launch_bounds__(256) raster_sum_kernel(buctd_raster_args a, unsigned long long* scratch) {
  __shared__ double s_k[15];
  __shared__ int s_px[64], s_py[64];
  __shared__ float s_amp[64][3];
  __shared__ int s_n;
  const int b = blockIdx.z;
  const int y = blockIdx.x;
  if (threadIdx.x == 0) {
    gauss15(s_k);
    int n = 0;
    for (int k = 0; k < a.k && k < 64; ++k) {
      int px, py;
      if (!kpt_pixel(a.kpts + ((long long)b * a.k + k) * a.kstride, a.w, a.h, &px, &py)) continue;
      // later keypoints overwrite earlier ones at the same pixel
      int slot = n;
      for (int j = 0; j < n; ++j)
        if (s_px[j] == px && s_py[j] == py) slot = j;
      s_px[slot] = px; s_py[slot] = py;
      for (int c = 0; c < 3; ++c) s_amp[slot][c] = a.mode == 1 ? a.colors[k * 3 + c] : 255.f;
      if (slot == n) ++n;
    }
    s_n = n;
  }
  __syncthreads();
  const int nch = a.mode == 1 ? 3 : 1;
  double lmax = 0;
  double den = 1.0;
  if (WRITE) {
    const double am = __longlong_as_double((long long)scratch[b]);
    den = am / 255.0;
  }
  for (int x = threadIdx.x; x < a.w; x += blockDim.x) {
    double v[3] = {0, 0, 0};
    for (int j = 0; j < s_n; ++j) {
      if (abs(y - s_py[j]) > 14 || abs(x - s_px[j]) > 14) continue;
      const double gy = folded(s_k, y, s_py[j], a.h);
      if (gy == 0) continue;
      const double gx = folded(s_k, x, s_px[j], a.w);
      for (int c = 0; c < nch; ++c) v[c] += gy * (gx * (double)s_amp[j][c]);
    }
    if (!WRITE) {
      for (int c = 0; c < nch; ++c) lmax = fmax(lmax, v[c]);
    } else {
      for (int c = 0; c < 3; ++c) {
        double r = v[a.mode == 1 ? c : 0];
        if (den != 0) r = r / den;
        float f = (float)r;
        if (a.mode == 2) f = (float)(long long)r;           // .astype(int)
        a.out[(((long long)b * a.out_channels + a.c_off + c) * a.h + y) * a.w + x] = f;
      }
    }
  }
  if (!WRITE) {
    for (int o = 16; o; o >>= 1) lmax = fmax(lmax, __shfl_xor_sync(0xffffffff, lmax, o));
    if ((threadIdx.x & 31) == 0 && lmax > 0)
      atomicMax(reinterpret_cast<unsigned long long*>(&scratch[b]), (unsigned long long)__double_as_longlong(lmax));
  }
}

// =================================================================================================
// Condition pyramid straight from the proposals (stacked mode): rasterise (above) composed with the antialiased
// resize to the three CoAM scales (ops.cu: triangle filter of 2S taps starting at S*o - S/2, weight 1 - |t - S + 0.5| / S,
// normalised over the taps inside the image).  Both are separable, so per keypoint and scale there are two short
// vectors ry[Y] = sum_y wY[Y][y] gy[y], rx[X] = sum_x wX[X][x] gx[x] (float64) and
// out[b][Y][X][k] = (255 / (max gy max gx)) ry[Y] rx[X].  Grid (2 row slices, 3 scales, B); every CTA recomputes the
// vectors of its scale (a few thousand FMAs) and writes its slice of rows with 16-byte stores.
// lib/dataset/JointsDataset.py:457-498 + lib/models/pose_hrnet_coam.py:750-757.
// =================================================================================================
constexpr int kCkMaxK = 32;       // keypoint channels (act16 channel stride of the condition maps is 32 for K = 17)
constexpr int kCkMaxDim = 160;    // largest H/4 or W/4 (640-pixel crops)

__device__ __forceinline__ double aa_resampled(const double* k, int o, int S, int p, int n) {
  // sum over the 2S taps of output index o of w[t] * g_p[S*o - S/2 + t], normalised by the in-range weights
  const int lo = S * o - S / 2;
  // g_p[j] != 0 needs |j - p| <= 7 (a reflected tap only folds back onto pixels within 7 of the border it left through)
  if (lo > p + 7 || lo + 2 * S - 1 < p - 7) return 0.0;
  double acc = 0, tot = 0;
  for (int t = 0; t < 2 * S; ++t) {
    const int j = lo + t;
    if (j < 0 || j >= n) continue;
    const double wt = 1.0 - fabs((double)t - S + 0.5) * (1.0 / S);
    tot += wt;
    if (abs(j - p) <= 7) acc += wt * folded(k, j, p, n);
  }
  return tot > 0 ? acc / tot : 0.0;
}

template <typename T>
__global__ void __launch_bounds__(256) cond_from_kpts_kernel(buctd_cond_kpts_args a) {
  __shared__ double s_k[15];
  __shared__ float s_rx[kCkMaxK][kCkMaxDim];
  __shared__ float s_ry[kCkMaxK][kCkMaxDim];
  __shared__ int s_px[kCkMaxK], s_py[kCkMaxK];
  __shared__ float s_amp[kCkMaxK];
  const int b = blockIdx.z, sc = blockIdx.y;
  const int S = 4 << sc;
  const int oh = a.h / S, ow = a.w / S;
  const int cs = a.out_cs[sc];
  // rows of this slice
  const int nrow = (oh + 1) * (ow + 1);
  const int per = (nrow + gridDim.x - 1) / gridDim.x;
  const int r0 = blockIdx.x * per, r1 = min(nrow, r0 + per);
  if (r0 >= r1) return;
  if (threadIdx.x == 0) gauss15(s_k);
  if (threadIdx.x < kCkMaxK) {
    int px = -1, py = -1;
    bool on = false;
    if (threadIdx.x < a.k) on = kpt_pixel(a.kpts + ((long long)b * a.k + threadIdx.x) * a.kstride, a.w, a.h, &px, &py);
    s_px[threadIdx.x] = on ? px : -1;
    s_py[threadIdx.x] = py;
  }
  __syncthreads();
  if (threadIdx.x < a.k && s_px[threadIdx.x] >= 0) {
    const int px = s_px[threadIdx.x], py = s_py[threadIdx.x];
    double mx = 0, my = 0;
    for (int t = -7; t <= 7; ++t) {
      if (px + t >= 0 && px + t < a.w) mx = fmax(mx, folded(s_k, px + t, px, a.w));
      if (py + t >= 0 && py + t < a.h) my = fmax(my, folded(s_k, py + t, py, a.h));
    }
    s_amp[threadIdx.x] = (float)(1.0 / (mx * my));           // the rx vector carries the 255
  }
  for (int i = threadIdx.x; i < a.k * (ow + oh); i += blockDim.x) {
    const int k = i / (ow + oh), o = i - k * (ow + oh);
    if (s_px[k] < 0) continue;
    if (o < ow) s_rx[k][o] = (float)(aa_resampled(s_k, o, S, s_px[k], a.w) * 255.0);
    else s_ry[k][o - ow] = (float)aa_resampled(s_k, o - ow, S, s_py[k], a.h);
  }
  __syncthreads();
  T* out = reinterpret_cast<T*>(a.out[sc]) + (long long)b * nrow * cs;
  const int groups = cs >> 3;
  for (int i = r0 * groups + threadIdx.x; i < r1 * groups; i += blockDim.x) {
    const int r = i / groups, c0 = (i - r * groups) << 3;
    const int Y = r / (ow + 1), X = r - Y * (ow + 1);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = c0 + j;
      v[j] = (Y < oh && X < ow && k < a.k && s_px[k] >= 0) ? s_amp[k] * s_ry[k][Y] * s_rx[k][X] : 0.f;
    }
    uint4 u;
    u.x = T16<T>::pack2(v[0], v[1]); u.y = T16<T>::pack2(v[2], v[3]);
    u.z = T16<T>::pack2(v[4], v[5]); u.w = T16<T>::pack2(v[6], v[7]);
    *reinterpret_cast<uint4*>(out + (long long)r * cs + c0) = u;
  }
}

int launch_cond_from_kpts(const buctd_cond_kpts_args& a, cudaStream_t st) {
  BUCTD_CHECK(a.kpts && a.out[0] && a.out[1] && a.out[2], BUCTD_ERR_INVALID, "cond_from_kpts: null pointer");
  BUCTD_CHECK(a.k >= 1 && a.k <= kCkMaxK && a.kstride >= 2, BUCTD_ERR_INVALID, "cond_from_kpts: K=%d out of range [1,%d] / kstride", a.k, kCkMaxK);
  BUCTD_CHECK(a.h % 16 == 0 && a.w % 16 == 0 && a.h / 4 <= kCkMaxDim && a.w / 4 <= kCkMaxDim, BUCTD_ERR_INVALID,
              "cond_from_kpts: needs H, W multiples of 16 and <= %d (got %dx%d)", 4 * kCkMaxDim, a.h, a.w);
  for (int i = 0; i < 3; ++i)
    BUCTD_CHECK(a.out_cs[i] % 8 == 0 && a.k <= a.out_cs[i] && (((uintptr_t)a.out[i]) & 15) == 0, BUCTD_ERR_INVALID,
                "cond_from_kpts: K %d vs out_cs[%d] %d / alignment", a.k, i, a.out_cs[i]);
  if (a.batch == 0) return BUCTD_OK;
  dim3 grid(2, 3, a.batch);      // every CTA recomputes the separable vectors of its (crop, scale): few row slices
  if (a.dtype == BUCTD_F16) cond_from_kpts_kernel<__half><<<grid, 256, 0, st>>>(a);
  else cond_from_kpts_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(a);
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}

int launch_raster(const buctd_raster_args& a, cudaStream_t st, int* nlaunch) {
  BUCTD_CHECK(a.kpts && a.out, BUCTD_ERR_INVALID, "raster: null pointer");
  BUCTD_CHECK(a.kstride >= 2 && a.w <= 512 && a.k >= 1, BUCTD_ERR_INVALID, "raster: bad kstride/width/k");
  BUCTD_CHECK(a.mode >= 0 && a.mode <= 2, BUCTD_ERR_INVALID, "raster: bad mode %d", a.mode);
  if (nlaunch) *nlaunch = 1;
  if (a.batch == 0) return BUCTD_OK;
  if (a.mode == 0) {
    BUCTD_CHECK(a.c_off + a.k <= a.out_channels, BUCTD_ERR_INVALID, "raster: channels out of range");
    raster_stacked_kernel<<<dim3(ceil_div(a.h, 32), a.k, a.batch), 256, 0, st>>>(a);
  } else {
    BUCTD_CHECK(a.c_off + 3 <= a.out_channels && a.k <= 64, BUCTD_ERR_INVALID, "raster: channels out of range / K > 64");
    BUCTD_CHECK(a.mode != 1 || a.colors != nullptr, BUCTD_ERR_INVALID, "raster: coloured mode needs colors");
    BUCTD_CHECK(a.scratch != nullptr, BUCTD_ERR_INVALID, "raster: modes 1/2 need 8*B bytes of scratch");
    unsigned long long* d_scratch = reinterpret_cast<unsigned long long*>(a.scratch);
    BUCTD_CUDA(cudaMemsetAsync(d_scratch, 0, sizeof(unsigned long long) * a.batch, st));
    raster_sum_kernel<false><<<dim3(a.h, 1, a.batch), 128, 0, st>>>(a, d_scratch);
    raster_sum_kernel<true><<<dim3(a.h, 1, a.batch), 128, 0, st>>>(a, d_scratch);
    if (nlaunch) *nlaunch = 2;
  }
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}

}  // namespace buctd
