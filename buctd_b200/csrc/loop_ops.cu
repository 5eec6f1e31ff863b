// Kernels of the callers either side of the forward (SURVEY.md section 8 rows a-L, a-T and f-1):
//   * flip-test input flip / flip_back + SHIFT_HEATMAP + average  (lib/core/function.py:213-236,
//     lib/utils/transforms.py:16-58) -- the reference round-trips the heat-maps through host NumPy for this
//   * JointsMSELoss forward + gradient                            (lib/core/loss.py:17-41)
//   * Adam step over one flat parameter buffer                     (lib/utils/utils.py:268-272 -> torch.optim.Adam)
// All HBM-bound streaming kernels: coalesced float4 accesses where the geometry allows, grids sized in multiples of
// the SM count, fixed-order (bit-reproducible) reductions.
#include "common.cuh"
#include "internal.h"

namespace buctd {

// =================================================================================================
// out[b][c][y][x] = x[b][perm[c]][y][W-1-x]
// =================================================================================================
__global__ void __launch_bounds__(256) flip_channels_kernel(buctd_flip_args a) {
  // one block per (b, c, group of rows); a warp mirrors one row through registers
  const int plane = blockIdx.y;                 // b * C + c
  const int b = plane / a.channels, c = plane - b * a.channels;
  const float* src = a.x + ((long long)b * a.channels + a.perm[c]) * a.h * a.w;
  float* dst = a.out + (long long)plane * a.h * a.w;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int y = blockIdx.x * 8 + warp; y < a.h; y += gridDim.x * 8) {
    const float* s = src + (long long)y * a.w;
    float* d = dst + (long long)y * a.w;
    for (int x = lane; x < a.w; x += 32) d[x] = __ldg(s + (a.w - 1 - x));
  }
}

int launch_flip_channels(const buctd_flip_args& a, cudaStream_t st) {
  BUCTD_CHECK(a.x && a.out && a.x != a.out, BUCTD_ERR_INVALID, "flip_channels: null or aliased pointers");
  BUCTD_CHECK(a.channels > 0 && a.channels <= BUCTD_MAX_FLIP_CHANNELS, BUCTD_ERR_INVALID, "flip_channels: %d channels (max %d)",
              a.channels, BUCTD_MAX_FLIP_CHANNELS);
  for (int c = 0; c < a.channels; ++c)
    BUCTD_CHECK(a.perm[c] >= 0 && a.perm[c] < a.channels, BUCTD_ERR_INVALID, "flip_channels: perm[%d] = %d out of range", c, a.perm[c]);
  if (a.batch == 0) return BUCTD_OK;
  dim3 grid(ceil_div(a.h, 8 * 4), a.batch * a.channels);
  flip_channels_kernel<<<grid, 256, 0, st>>>(a);
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}

// =================================================================================================
// out = (hm + S) * 0.5,  F[b][k][y][x] = hm_flipped[b][perm[k]][y][W-1-x],  S = F shifted right by one column
// (column 0 keeps F[0]) when shift != 0, else S = F.
// =================================================================================================
__global__ void __launch_bounds__(256) flip_merge_kernel(buctd_flip_merge_args a) {
  const int plane = blockIdx.y;
  const int b = plane / a.k, k = plane - b * a.k;
  const float* o = a.hm + (long long)plane * a.h * a.w;
  const float* f = a.hm_flipped + ((long long)b * a.k + a.perm[k]) * a.h * a.w;
  float* d = a.out + (long long)plane * a.h * a.w;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int y = blockIdx.x * 8 + warp; y < a.h; y += gridDim.x * 8) {
    const long long ro = (long long)y * a.w;
    for (int x = lane; x < a.w; x += 32) {
      const int xs = a.shift ? (x > 0 ? x - 1 : 0) : x;          // column of F that lands on x
      const float fv = __ldg(f + ro + (a.w - 1 - xs));
      d[ro + x] = (o[ro + x] + fv) * 0.5f;
    }
  }
}

int launch_flip_merge(const buctd_flip_merge_args& a, cudaStream_t st) {
  BUCTD_CHECK(a.hm && a.hm_flipped && a.out, BUCTD_ERR_INVALID, "flip_merge: null pointer");
  BUCTD_CHECK(a.out != a.hm_flipped, BUCTD_ERR_INVALID, "flip_merge: out must not alias hm_flipped");
  BUCTD_CHECK(a.k > 0 && a.k <= BUCTD_MAX_FLIP_CHANNELS, BUCTD_ERR_INVALID, "flip_merge: %d joints (max %d)", a.k, BUCTD_MAX_FLIP_CHANNELS);
  for (int c = 0; c < a.k; ++c)
    BUCTD_CHECK(a.perm[c] >= 0 && a.perm[c] < a.k, BUCTD_ERR_INVALID, "flip_merge: perm[%d] = %d out of range", c, a.perm[c]);
  if (a.batch == 0) return BUCTD_OK;
  dim3 grid(ceil_div(a.h, 8 * 4), a.batch * a.k);
  flip_merge_kernel<<<grid, 256, 0, st>>>(a);
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}

// =================================================================================================
// JointsMSELoss: loss = 0.5 / (K B HW) * sum_{b,k,p} (w_bk (o - t))^2 ; grad = w_bk^2 (o - t) / (K B HW)
// pass 1: one block per (b,k) row -> partial sum (double); pass 2: one block adds the B*K partials in index order
// =================================================================================================
__global__ void __launch_bounds__(256) joints_mse_rows_kernel(buctd_joints_mse_args a) {
  const int row = blockIdx.x;                    // b * K + k
  const float w = a.use_target_weight ? a.target_weight[row] : 1.f;
  const float* o = a.output + (long long)row * a.hw;
  const float* t = a.target + (long long)row * a.hw;
  float* g = a.grad ? a.grad + (long long)row * a.hw : nullptr;
  const float gscale = w * w / ((float)a.k * (float)a.batch * (float)a.hw) * (a.grad_scale != 0.f ? a.grad_scale : 1.f);
  double acc = 0.0;
  for (int i = threadIdx.x; i < a.hw; i += 256) {
    const float d = o[i] - t[i];
    const float wd = w * d;
    acc += (double)wd * (double)wd;
    if (g) g[i] = gscale * d;
  }
  __shared__ double s[256];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int n = 128; n > 0; n >>= 1) {
    if (threadIdx.x < n) s[threadIdx.x] += s[threadIdx.x + n];
    __syncthreads();
  }
  if (threadIdx.x == 0) a.scratch[row] = s[0];
}

__global__ void __launch_bounds__(256) joints_mse_final_kernel(buctd_joints_mse_args a) {
  __shared__ double s[256];
  const int rows = a.batch * a.k;
  double acc = 0.0;
  for (int i = threadIdx.x; i < rows; i += 256) acc += a.scratch[i];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int n = 128; n > 0; n >>= 1) {
    if (threadIdx.x < n) s[threadIdx.x] += s[threadIdx.x + n];
    __syncthreads();
  }
  if (threadIdx.x == 0) *a.loss = (float)(0.5 * s[0] / ((double)a.k * (double)a.batch * (double)a.hw));
}

int launch_joints_mse(const buctd_joints_mse_args& a, cudaStream_t st) {
  BUCTD_CHECK(a.output && a.target && a.loss && a.scratch, BUCTD_ERR_INVALID, "joints_mse: null pointer");
  BUCTD_CHECK(!a.use_target_weight || a.target_weight, BUCTD_ERR_INVALID, "joints_mse: use_target_weight without target_weight");
  BUCTD_CHECK(a.batch > 0 && a.k > 0 && a.hw > 0, BUCTD_ERR_INVALID, "joints_mse: empty input %dx%dx%d", a.batch, a.k, a.hw);
  joints_mse_rows_kernel<<<a.batch * a.k, 256, 0, st>>>(a);
  joints_mse_final_kernel<<<1, 256, 0, st>>>(a);
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}

// =================================================================================================
// Adam (torch.optim.Adam defaults the reference uses: betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad),
// in torch's own operation order so the step matches it to rounding:
//   m = m + (g - m) (1 - b1);  v = v b2 + (1 - b2) g g;  p -= (lr / (1 - b1^t)) m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// =================================================================================================
struct AdamScalars {
  float beta2, omb1, omb2, eps, step_size, bc2_sqrt;
};
__global__ void __launch_bounds__(256) adam_kernel(buctd_adam_args a, AdamScalars c) {
  const float omb1 = c.omb1, omb2 = c.omb2, step_size = c.step_size, bc2_sqrt = c.bc2_sqrt;
  const long long n4 = a.n >> 2;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  auto upd = [&](float& p, float g, float& m, float& v) {
    m = m + (g - m) * omb1;
    v = v * c.beta2 + omb2 * g * g;
    const float denom = sqrtf(v) / bc2_sqrt + c.eps;
    p = p - step_size * (m / denom);
  };
  float4* P = reinterpret_cast<float4*>(a.param);
  const float4* G = reinterpret_cast<const float4*>(a.grad);
  float4* M = reinterpret_cast<float4*>(a.exp_avg);
  float4* V = reinterpret_cast<float4*>(a.exp_avg_sq);
  for (long long i = tid; i < n4; i += nth) {
    float4 p = P[i], m = M[i], v = V[i];
    const float4 g = G[i];
    upd(p.x, g.x, m.x, v.x); upd(p.y, g.y, m.y, v.y); upd(p.z, g.z, m.z, v.z); upd(p.w, g.w, m.w, v.w);
    P[i] = p; M[i] = m; V[i] = v;
  }
  for (long long i = (n4 << 2) + tid; i < a.n; i += nth) upd(a.param[i], a.grad[i], a.exp_avg[i], a.exp_avg_sq[i]);
}

int launch_adam(const buctd_adam_args& a, int sm_count, cudaStream_t st) {
  BUCTD_CHECK(a.param && a.grad && a.exp_avg && a.exp_avg_sq, BUCTD_ERR_INVALID, "adam: null pointer");
  BUCTD_CHECK(a.step >= 1, BUCTD_ERR_INVALID, "adam: step must be >= 1 (got %d)", a.step);
  BUCTD_CHECK((((uintptr_t)a.param | (uintptr_t)a.grad | (uintptr_t)a.exp_avg | (uintptr_t)a.exp_avg_sq) & 15) == 0,
              BUCTD_ERR_INVALID, "adam: buffers must be 16-byte aligned");
  if (a.n <= 0) return BUCTD_OK;
  // scalars in double on the host, rounded to float once, exactly as torch's Python-side scalar arithmetic does
  const double bc1 = 1.0 - pow(a.beta1, (double)a.step), bc2 = 1.0 - pow(a.beta2, (double)a.step);
  AdamScalars c;
  c.beta2 = (float)a.beta2; c.omb1 = (float)(1.0 - a.beta1); c.omb2 = (float)(1.0 - a.beta2); c.eps = (float)a.eps;
  c.step_size = (float)(a.lr / bc1); c.bc2_sqrt = (float)sqrt(bc2);
  long long want = (a.n / 4 + 255) / 256;
  const long long cap = (long long)sm_count * 8;               // 8 resident 256-thread blocks per SM, grid-stride beyond
  const int blocks = (int)(want < 1 ? 1 : (want < cap ? want : cap));
  adam_kernel<<<blocks, 256, 0, st>>>(a, c);
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}

// ---- training targets (JointsDataset.generate_target) ---------------------------------------------
// One block per (crop, joint): zero the [H][W] plane and paste the clipped (6 sigma + 1)^2 stamp.  The stamp table is
// computed by the caller with the reference's own float32 numpy expression, so the planes are bit-identical to the
// reference's whatever libm the host has.
__global__ void gen_target_kernel(buctd_gen_target_args a) {
  const int bk = blockIdx.x, k = bk % a.k;
  float* plane = a.target + (size_t)bk * a.h * a.w;
  const double* j = a.joints + (size_t)bk * 3;
  float wgt = (float)a.joints_vis[(size_t)bk * 3];
  // mu = int(joint / feat_stride + 0.5): float64 arithmetic, truncation toward zero
  const int mu_x = (int)(j[0] / a.feat_stride_x + 0.5), mu_y = (int)(j[1] / a.feat_stride_y + 0.5);
  const int tmp = a.tmp_size, size = 2 * tmp + 1;
  const int ul_x = mu_x - tmp, ul_y = mu_y - tmp, br_x = mu_x + tmp + 1, br_y = mu_y + tmp + 1;
  const bool outside = ul_x >= a.w || ul_y >= a.h || br_x < 0 || br_y < 0;
  if (outside) wgt = 0.f;
  const bool paste = !outside && wgt > 0.5f;
  const int x0 = max(0, ul_x), x1 = min(br_x, a.w), y0 = max(0, ul_y), y1 = min(br_y, a.h);
  for (int i = threadIdx.x; i < a.h * a.w; i += blockDim.x) {
    const int y = i / a.w, x = i - y * a.w;
    float v = 0.f;
    if (paste && x >= x0 && x < x1 && y >= y0 && y < y1) v = a.stamp[(y - ul_y) * size + (x - ul_x)];
    plane[i] = v;
  }
  if (threadIdx.x == 0) a.target_weight[bk] = a.joints_weight != nullptr ? wgt * a.joints_weight[k] : wgt;
}

int launch_gen_target(const buctd_gen_target_args& a, cudaStream_t st) {
  if (a.batch == 0) return BUCTD_OK;
  BUCTD_CHECK(a.joints && a.joints_vis && a.stamp && a.target && a.target_weight, BUCTD_ERR_INVALID, "gen_target: null pointer");
  BUCTD_CHECK(a.batch >= 0 && a.k > 0 && a.h > 0 && a.w > 0 && a.tmp_size >= 0, BUCTD_ERR_INVALID, "gen_target: bad shape");
  BUCTD_CHECK(a.feat_stride_x > 0 && a.feat_stride_y > 0, BUCTD_ERR_INVALID, "gen_target: bad feat_stride");
  if (a.batch == 0) return BUCTD_OK;
  gen_target_kernel<<<a.batch * a.k, 256, 0, st>>>(a);
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}

// ---- PCK accuracy on arg-max coordinates (core.evaluate.accuracy after its two get_max_preds) ---------
// One block, one thread per joint; the reference's float64 arithmetic without fused multiply-adds:
//   dist = || pred / (h/10, w/10) - target / (h/10, w/10) ||   when target x > 1 and y > 1, else -1
//   acc[1 + k] = #(dist < thr) / #(dist != -1)  (or -1), acc[0] = mean of the acc[1 + k] >= 0, cnt = how many.
__global__ void pck_kernel(buctd_pck_args a) {
  __shared__ double s_acc[BUCTD_MAX_FLIP_CHANNELS];
  const int k = threadIdx.x;
  if (k < a.k) {
    const double nx = (double)a.h / 10.0, ny = (double)a.w / 10.0;     // norm = ones((B, 2)) * [h, w] / 10
    int hit = 0, n = 0;
    for (int b = 0; b < a.batch; ++b) {
      const float* p = a.pred + ((size_t)b * a.k + k) * 2;
      const float* t = a.target + ((size_t)b * a.k + k) * 2;
      if (t[0] > 1.f && t[1] > 1.f) {
        const double dx = __dsub_rn((double)p[0] / nx, (double)t[0] / nx), dy = __dsub_rn((double)p[1] / ny, (double)t[1] / ny);
        const double d = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
        ++n;
        if (d < a.thr) ++hit;
      }
    }
    s_acc[k] = n > 0 ? (double)hit * 1.0 / (double)n : -1.0;
    a.acc[1 + k] = s_acc[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double sum = 0.0;
    int cnt = 0;
    for (int i = 0; i < a.k; ++i)
      if (s_acc[i] >= 0.0) { sum = __dadd_rn(sum, s_acc[i]); ++cnt; }
    a.acc[0] = cnt != 0 ? sum / (double)cnt : 0.0;
    *a.cnt = cnt;
  }
}

int launch_pck(const buctd_pck_args& a, cudaStream_t st) {
  BUCTD_CHECK(a.pred && a.target && a.acc && a.cnt, BUCTD_ERR_INVALID, "pck_accuracy: null pointer");
  BUCTD_CHECK(a.k > 0 && a.k <= BUCTD_MAX_FLIP_CHANNELS && a.batch >= 0, BUCTD_ERR_INVALID, "pck_accuracy: k %d out of range [1,%d]", a.k,
              BUCTD_MAX_FLIP_CHANNELS);
  pck_kernel<<<1, BUCTD_MAX_FLIP_CHANNELS, 0, st>>>(a);
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}

// ---- crop warp + normalise (cv2.warpAffine INTER_LINEAR, BORDER_CONSTANT 0; ToTensor; Normalize) -------------
// OpenCV's 8-bit affine warp is fixed-point (imgwarp.cpp, WarpAffineInvoker + remapBilinear): the FORWARD matrix is
// inverted in float64, source coordinates are built in 1/1024 px (round half even) + 16, reduced to 1/32 px, and the
// four taps are combined with integer weights (32-fx)(32-fy)*32 ... that sum to 2^15, + 2^14, >> 15.  Restated here
// operation for operation (no fused multiply-adds), so the 8-bit crop is bit-identical to cv2.warpAffine's, and the
// float image after x/255, (x - mean)/std is bit-identical to torchvision's ToTensor + Normalize.
__device__ __forceinline__ int cv_round(double v) {           // cv::saturate_cast<int>(double): lrint + saturation
  v = fmin(fmax(v, -2147483648.0), 2147483647.0);
  return __double2int_rn(v);
}
__global__ void crop_warp_kernel(buctd_crop_warp_args a) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
  if (x >= a.w) return;
  const double* m = a.inv_trans + (size_t)b * 6;
  const int h_img = a.img_h[b], w_img = a.img_w[b];
  const long long stride = a.img_stride[b];
  const uint8_t* img = a.images[b];
  const int X0 = cv_round(__dmul_rn(__dadd_rn(__dmul_rn(m[1], (double)y), m[2]), 1024.0)) + 16;
  const int Y0 = cv_round(__dmul_rn(__dadd_rn(__dmul_rn(m[4], (double)y), m[5]), 1024.0)) + 16;
  const int X = (X0 + cv_round(__dmul_rn(__dmul_rn(m[0], (double)x), 1024.0))) >> 5;
  const int Y = (Y0 + cv_round(__dmul_rn(__dmul_rn(m[3], (double)x), 1024.0))) >> 5;
  const int sx = X >> 5, sy = Y >> 5, fx = X & 31, fy = Y & 31;
  const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32, w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
  const bool x0 = sx >= 0 && sx < w_img, x1 = sx + 1 >= 0 && sx + 1 < w_img;
  const bool y0 = sy >= 0 && sy < h_img, y1 = sy + 1 >= 0 && sy + 1 < h_img;
  const uint8_t* r0 = img + (long long)sy * stride + (long long)sx * 3;
  const uint8_t* r1 = r0 + stride;
  float* out = a.out + ((size_t)b * a.out_channels + a.c_off) * a.h * a.w + (size_t)y * a.w + x;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int p00 = (y0 && x0) ? r0[c] : 0, p01 = (y0 && x1) ? r0[3 + c] : 0;
    const int p10 = (y1 && x0) ? r1[c] : 0, p11 = (y1 && x1) ? r1[3 + c] : 0;
    const int v = (p00 * w00 + p01 * w01 + p10 * w10 + p11 * w11 + (1 << 14)) >> 15;
    const int oc = a.swap_rb ? 2 - c : c;
    if (a.out_u8 != nullptr) a.out_u8[(((size_t)b * a.h + y) * a.w + x) * 3 + oc] = (uint8_t)v;
    out[(size_t)oc * a.h * a.w] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, 255.f), a.mean[oc]), a.std[oc]);
  }
}

int launch_crop_warp(const buctd_crop_warp_args& a, cudaStream_t st) {
  if (a.batch == 0) return BUCTD_OK;
  BUCTD_CHECK(a.images && a.img_h && a.img_w && a.img_stride && a.inv_trans && a.out, BUCTD_ERR_INVALID, "crop_warp: null pointer");
  BUCTD_CHECK(a.batch >= 0 && a.h > 0 && a.w > 0 && a.out_channels >= 3 && a.c_off >= 0 && a.c_off + 3 <= a.out_channels,
              BUCTD_ERR_INVALID, "crop_warp: bad shape (batch %d, %dx%d, channels %d, c_off %d)", a.batch, a.h, a.w, a.out_channels, a.c_off);
  BUCTD_CHECK(a.std[0] != 0.f && a.std[1] != 0.f && a.std[2] != 0.f, BUCTD_ERR_INVALID, "crop_warp: zero std");
  BUCTD_CHECK(a.batch <= 65535 && a.h <= 65535, BUCTD_ERR_INVALID, "crop_warp: batch / height above the grid limit");
  dim3 grid((a.w + 127) / 128, a.h, a.batch);
  crop_warp_kernel<<<grid, 128, 0, st>>>(a);
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}

}  // namespace buctd
