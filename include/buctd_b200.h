/* buctd_b200 -- C ABI of the B200 (sm_100a) implementation of the BUCTD conditional top-down
 * pose path.  Plain pointers and sizes only; no torch / C++ types.  Every entry point returns
 * 0 on success or a negative BUCTD_ERR_* code; buctd_last_error() returns the message of the
 * last failure on the calling thread.  All device work is enqueued on the caller's stream
 * (stream-ordered, no implicit device synchronisation, no hidden allocation): the caller owns
 * every buffer, the library never frees caller memory.  Handles are per device; calls on
 * different handles may run concurrently from different host threads (the reference's
 * torch.nn.DataParallel drives one host thread per GPU, tools/test.py:134).
 *
 * The reference (amathislab/BUCTD) is pure Python on top of cuDNN/cuBLAS/ATen/NumPy, so there
 * is no existing FFI to mirror; each entry point below names the reference function it replaces
 * (paths relative to the reference root), and INTEGRATION.md shows the ctypes stubs a reference
 * maintainer would add.
 *
 * Data layouts
 *   "act16"  : padded NHWC 16-bit activation  [B][H+1][W+1][Cs]  (fp16 or bf16, see buctd_dtype);
 *              row y==H and column x==W are zero ("shared pad": a 3x3 tap that leaves the image
 *              lands on a zero).  Flat row index r = (b*(H+1) + y)*(W+1) + x; Cs = channel stride
 *              in elements (multiple of 8).
 *   "planar" : [B][C][H*W] 16-bit, unpadded (the layout of the CoAM channel attention and of its
 *              HW x HW projection).
 *   "nchw32" : [B][C][H][W] float32, unpadded: the reference's own tensors (model input, heat-maps).
 */
#ifndef BUCTD_B200_H_
#define BUCTD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BUCTD_ABI_VERSION 3

#define BUCTD_OK 0
#define BUCTD_ERR_INVALID (-1)   /* bad argument / unsupported shape */
#define BUCTD_ERR_CUDA (-2)      /* a CUDA runtime / driver call failed */
#define BUCTD_ERR_NO_DEVICE (-3) /* no sm_100 device, or driver entry point missing */

typedef enum { BUCTD_F16 = 0, BUCTD_BF16 = 1 } buctd_dtype;

typedef struct buctd_ctx buctd_ctx;   /* per-device handle */
typedef struct buctd_plan buctd_plan; /* recorded sequence of launches with pre-encoded tensor maps */
typedef void* buctd_stream;           /* cudaStream_t */

int buctd_abi_version(void);
const char* buctd_last_error(void);
int buctd_create(int device, buctd_ctx** out);
int buctd_destroy(buctd_ctx* ctx);
/* number of kernels launched through this handle since creation (bench.py's gpu_launches) */
long long buctd_launch_count(const buctd_ctx* ctx);

/* ------------------------------------------------------------------------------------------
 * Tap-GEMM: every dense contraction of the path on the tcgen05 tensor cores.
 *   out[m, n] = act( sum_t sum_k A[m (+) tap t, k] * Wt[t][n][k] + bias[n] + residual[m, n] )
 * mode FLAT : A is an act16 tensor seen as a 2-D matrix [rows, Cs]; tap t of a 3x3 stride-1
 *             convolution is the same matrix shifted by dy*(W+1)+dx rows (zero pads make the shift
 *             exact), ntaps = 9; ntaps = 1 is a 1x1 convolution / nn.Linear.
 * mode PATCH: stride-2 3x3 convolution; M-tile = th x tw output pixels of one image, A tiles are
 *             loaded with a strided 4-D TMA box.
 * Replaces nn.Conv2d + nn.BatchNorm2d(eval) + ReLU (+ residual) in
 *   lib/models/pose_hrnet_coam.py:44-60 (BasicBlock), :81-101 (Bottleneck), :200-242 (fuse convs),
 *   :405-432 (transitions), :565 (final_layer), :635-641,666-670 (CoAM convs) and nn.Linear in
 *   lib/models/self_attention.py:74-76,87,159.
 * Weights are packed by the caller: BN folded in fp32, cast to 16 bit, layout [ntaps][n_alloc][k_cs]
 * (K contiguous); bias is the folded BN shift / conv bias in fp32.
 * ------------------------------------------------------------------------------------------ */
enum { BUCTD_TG_FLAT = 0, BUCTD_TG_PATCH = 1 };
enum { BUCTD_OUT_ACT16 = 0,   /* act16 with pad rows/cols written as zero */
       BUCTD_OUT_ROWS16 = 1,  /* plain [rows][out_cs] 16-bit, no pad handling (linear layers) */
       BUCTD_OUT_NCHW32 = 2   /* float32 [B][N][H][W] unpadded (the returned heat-maps) */ };

typedef struct {
  int32_t dtype;        /* buctd_dtype of A, W, residual and 16-bit outputs */
  int32_t mode;         /* BUCTD_TG_FLAT | BUCTD_TG_PATCH */
  int32_t ntaps;        /* 1 or 9 */
  int32_t stride;       /* PATCH: 2 (1 allowed); FLAT: 1 */
  int32_t batch;        /* geometry of the OUTPUT image grid (B, H, W); for ROWS16 only m_rows is used */
  int32_t out_h, out_w;
  int32_t in_h, in_w;   /* PATCH: input image size (in_h = stride*out_h) */
  int64_t m_rows;       /* FLAT: number of A/out rows (act16: B*(H+1)*(W+1)) */
  int32_t k;            /* valid input channels (cin) */
  int32_t a_cs;         /* A channel stride (elements) */
  int32_t n;            /* valid output channels (cout) */
  int32_t n_alloc;      /* rows per tap in W (>= n, multiple of 16); rows n..n_alloc-1 of every tap MUST be zero */
  int32_t w_cs;         /* W row stride in elements (>= k, multiple of 8) */
  int32_t out_cs;       /* out channel stride (ACT16/ROWS16) */
  int32_t res_cs;       /* residual channel stride */
  int32_t relu;
  int32_t out_mode;
  int32_t reserved;
  const void* a;
  const void* w;
  const float* bias;    /* [n] or NULL */
  const void* residual; /* same row indexing as out, or NULL */
  void* out;
} buctd_tapgemm_args;
int buctd_tapgemm(buctd_ctx* ctx, const buctd_tapgemm_args* args, buctd_stream stream);

/* Chain of up to BUCTD_MAX_CHAIN FLAT tap-GEMMs of identical geometry (same dtype, taps, rows, K, N, strides; bias,
 * relu, residual and the buffers differ) in ONE persistent launch with a grid-wide barrier between layers: the eight
 * 3x3 convolutions of the four BasicBlocks of one HRNet branch, lib/models/pose_hrnet_coam.py:31-60,142-165,254-255.
 * Layer l+1 may read (as A or residual) anything layers <= l wrote.  act16 output only. */
#define BUCTD_MAX_CHAIN 8
typedef struct {
  int32_t n_layers, reserved;
  buctd_tapgemm_args layer[BUCTD_MAX_CHAIN];
} buctd_tapgemm_chain_args;
int buctd_tapgemm_chain(buctd_ctx* ctx, const buctd_tapgemm_chain_args* args, buctd_stream stream);

/* Stem conv1: 3x3 stride 2, 3 -> 64, BN folded, ReLU, reading the nchw32 model input directly
 * (channels [0,3) of x[B][x_channels][H][W]) and writing act16 [B][H/2+1][W/2+1][64].
 * Replaces conv1/bn1/relu, lib/models/pose_hrnet_coam.py:505-507.  w: float32 [64][27] (ci,ky,kx). */
typedef struct {
  int32_t dtype, batch, x_channels, h, w, out_cs;
  int32_t x_half;       /* 0: x is float32 (the reference's tensor); 1: x is IEEE half [B][x_channels][H][W] (halves the H2D copy) */
  int32_t no_relu;      /* training: raw convolution output (BatchNorm + ReLU follow as separate passes) */
  const void* x; const float* weight; const float* bias; void* out;
} buctd_stem_args;
int buctd_stem_conv(buctd_ctx* ctx, const buctd_stem_args* args, buctd_stream stream);

/* Exchange/fuse sum: out = ReLU(sum_j term_j), term_j an act16 tensor at the output resolution or
 * 2^shift times coarser (nearest up-sampling).  Replaces the sum + nn.Upsample + ReLU of
 * HighResolutionModule.forward, lib/models/pose_hrnet_coam.py:257-266 (the 1x1 / 3x3 convs of
 * the terms are tap-GEMMs). */
typedef struct {
  int32_t dtype, batch, h, w, c, out_cs, nterms, relu;
  const void* term[4]; int32_t term_cs[4]; int32_t term_shift[4];
  void* out;
} buctd_fuse_args;
int buctd_hr_fuse(buctd_ctx* ctx, const buctd_fuse_args* args, buctd_stream stream);

/* Condition resize: antialiased bilinear (triangle filter, support = scale) from channels
 * [c_off, c_off+cc) of the nchw32 model input to act16 [B][oh+1][ow+1][out_cs] (channels >= cc
 * zero).  Replaces TF.resize(cond_hm, (H_i, W_i)), lib/models/pose_hrnet_coam.py:755 as
 * torchvision >= 0.17 computes it (antialias=True); antialias=0 gives the 2-tap legacy form. */
typedef struct {
  int32_t dtype, batch, x_channels, c_off, cc, h, w, oh, ow, out_cs, antialias, reserved;
  const float* x; void* out;
} buctd_resize_args;
int buctd_cond_resize(buctd_ctx* ctx, const buctd_resize_args* args, buctd_stream stream);

/* The three CoAM scales of buctd_cond_resize (1/4, 1/8, 1/16 of an H x W input, H and W multiples of 16, antialiased)
 * in one pass over the condition planes: out[i] is act16 [B][H/s_i + 1][W/s_i + 1][out_cs[i]], s = {4, 8, 16}.
 * Same arithmetic and results as three buctd_cond_resize calls (CoAMBlock.forward, lib/models/pose_hrnet_coam.py:750-757). */
typedef struct {
  int32_t dtype, batch, x_channels, c_off, cc, h, w, reserved;
  int32_t out_cs[3]; int32_t reserved2;
  const float* x; void* out[3];
} buctd_resize3_args;
int buctd_cond_resize3(buctd_ctx* ctx, const buctd_resize3_args* args, buctd_stream stream);

/* Condition pyramid straight from the proposals (stacked mode): what buctd_rasterize_condition (mode 0) followed by
 * buctd_cond_resize3 produce, without the full-resolution planes ever existing.  The blurred impulse of keypoint k is
 * separable, g_k[y][x] = gy_k[y] gx_k[x] 255 / (max gy_k max gx_k) (15-tap sigma-2.6 response with BORDER_REFLECT_101
 * folding), and so is the antialiased resize, hence out_i[b][Y][X][k] = 255 / peak_k * (sum_y wY_i[Y][y] gy_k[y]) *
 * (sum_x wX_i[X][x] gx_k[x]) -- two short vectors per keypoint and scale (float64), one 16-bit store per output element.
 * kpts float32 [B][K][kstride] crop-pixel (x, y, ...), same drawing rule as the rasteriser (truncation, 0 < x < W,
 * 0 < y < H, pixel (y-1, x-1)); out[i]: act16 [B][H/s_i + 1][W/s_i + 1][out_cs[i]], s = {4, 8, 16}, channels >= K zero.
 * Replaces get_stacked_condition + the H2D copy of 17 float32 planes per crop + TF.resize x 3,
 * lib/dataset/JointsDataset.py:471-498, lib/models/pose_hrnet_coam.py:750-757 (481 MB -> 13 KB per 64 crops over PCIe). */
typedef struct {
  int32_t dtype, batch, k, kstride, h, w;
  int32_t out_cs[3]; int32_t reserved;
  const float* kpts; void* out[3];
} buctd_cond_kpts_args;
int buctd_cond_from_kpts(buctd_ctx* ctx, const buctd_cond_kpts_args* args, buctd_stream stream);

/* CoAM channel attention core: att = softmax_c2( yc^T y / sqrt(HW) ) [C x C] per crop, then
 * out[c1][p] = sum_c2 att[c1][c2] y[p][c2], written planar [B][C][HW] (the A operand of the
 * HW x HW fc_o tap-GEMM).  y, yc: act16 [B][H+1][W+1][cs].  scratch: float32 [B][nchunks][C][C] with
 * nchunks = ceil((H+1)(W+1) / BUCTD_CHAN_ATT_ROWS_PER_CHUNK) (per-chunk partial sums, reduced in a fixed order:
 * the result is bit-reproducible); after the call [b][0][:][:] holds the soft-max probabilities.
 * out_padded != 0: out is written in PADDED pixel order, out[b][c1][y (W+1) + x] with row stride out_stride >=
 * (H+1)(W+1) (pad positions: zeros on the tensor-core path, left untouched by the CUDA-core kernels -- allocate zeroed): the consumer is a fc_o whose weight matrix carries zero columns at the pad
 * positions.  In that form (no dropout, >= 2 scratch chunks) both contractions run on the tensor cores (buctd_gemm16's
 * kernel: scores over the pixel rows with MN-major operands, then att (16 bit) x y^T); otherwise CUDA-core kernels.
 * Replaces SimplifiedScaledDotProductAttention.forward up to fc_o,
 * lib/models/self_attention.py:146-158 as called from ChannelAttentionModule.forward,
 * lib/models/pose_hrnet_coam.py:673-689. */
#define BUCTD_CHAN_ATT_ROWS_PER_CHUNK 256
typedef struct {
  int32_t dtype, batch, h, w, c, y_cs, yc_cs;
  int32_t out_stride;   /* planar row stride of out in elements (0 = H*W); a multiple of 8 lets the HW x HW fc_o read it by TMA */
  float drop_p; uint32_t drop_seed;   /* training: dropout of the probabilities (see "Dropout" below); 0: none */
  int32_t out_padded, reserved;
  const void* y; const void* yc; float* scratch; void* out;
  const uint32_t* seed_xor;           /* optional device word XOR-ed into drop_seed (the step counter of a replayed CUDA graph) */
} buctd_chan_att_args;
int buctd_coam_channel_att(buctd_ctx* ctx, const buctd_chan_att_args* args, buctd_stream stream);

/* CoAM position attention core (flash style, the [HW x HW] score matrix never leaves the SM):
 * o[t] = softmax_s( q[t].k[s] / sqrt(C) ) v[s] over the H*W valid pixels of each crop.
 * q, k, v, o: act16 [B][H+1][W+1][cs] with C valid channels (q/k/v may be slices of one tensor); C a multiple of 16 in
 * [16, 384] (every branch of W32 / W48 at every ATT_MODULES placement); tcgen05 for all of them.
 * Replaces ScaledDotProductAttention.forward between the projections,
 * lib/models/self_attention.py:78-86 (eval mode: dropout is the identity). */
typedef struct {
  int32_t dtype, batch, h, w, c, q_cs, k_cs, v_cs, o_cs, reserved;
  float drop_p; uint32_t drop_seed;   /* training: dropout of the probabilities; 0: none */
  const void* q; const void* k; const void* v; void* o;
  float* lse2;          /* optional out [B][(H+1)(W+1)]: log2-sum-exp of every query row's scaled scores (for the backward) */
  const uint32_t* seed_xor;           /* optional device word XOR-ed into drop_seed */
} buctd_pos_att_args;
int buctd_coam_position_att(buctd_ctx* ctx, const buctd_pos_att_args* args, buctd_stream stream);

/* DAModule output: out = x + (p_out + c_out), p_out act16, c_out planar [B][C][HW]; with
 * channel_only: out = x * c_out.  Replaces DAModule.forward, lib/models/pose_hrnet_coam.py:716-724. */
typedef struct {
  int32_t dtype, batch, h, w, c, x_cs, p_cs, out_cs, channel_only;
  int32_t c_stride;     /* planar row stride of c_out in elements (0 = H*W) */
  const void* x; const void* p_out; const void* c_out; void* out;
} buctd_da_combine_args;
int buctd_da_combine(buctd_ctx* ctx, const buctd_da_combine_args* args, buctd_stream stream);

/* preNet: x' = BN(conv7x7(BN(conv3x3(rgb)))) + BN(conv7x7(cond)) at full resolution; nchw32 in
 * (rgb = channels [0,3), cond = channels [3,6) of x), nchw32 3-channel out.  There is no
 * non-linearity between the 3x3 and the 7x7 convolution, so the caller composes them (with both
 * BatchNorms folded) into one 9x9 3->3 kernel.  The 64-channel intermediate is zero-padded by the
 * 7x7 convolution, so near the border the composition only includes the taps that fall inside the
 * image: w_eff holds the 7x7 = 49 variants [vy*7+vx][co][ci][9*9], vy/vx = 0 (interior), 1..3 (that
 * many rows/cols cut at the top/left), 4..6 (1..3 cut at the bottom/right); b_eff [49][3] is the
 * matching constant (b2 + bc + sum of w2*b1 over the included taps).  wc [3][3][7*7] is the
 * BN-folded cond kernel.  Replaces rgb_preNet / cond_preNet, lib/models/pose_hrnet.py:431-442,452-458. */
typedef struct {
  int32_t batch, x_channels, h, w;
  const float* x; const float* w_eff; const float* b_eff; const float* wc; float* out;
} buctd_prenet_args;
int buctd_prenet(buctd_ctx* ctx, const buctd_prenet_args* args, buctd_stream stream);

/* Heat-map decode: per (b,k) flat argmax (first maximum), max value, quarter-pixel offset from
 * the sign of the neighbour differences (strict 1<px<W-1, 1<py<H-1), zeroing when maxval<=0, and
 * the inverse crop transform x_img = cx + (x - W/2) * 200*scale[0]/W (likewise y with H/2).
 * heatmaps nchw32 [B][K][H][W]; center, scale float32 [B][2] (may be NULL: then preds = heat-map
 * coordinates); outputs preds [B][K][2], maxvals [B][K], optional coords [B][K][2] (the
 * get_max_preds result before the offset) and idx [B][K] (int32 flat argmax).
 * Replaces get_max_preds / get_final_preds, lib/core/inference.py:19-87 and transform_preds,
 * lib/utils/transforms.py:78-118. */
typedef struct {
  int32_t batch, k, h, w, post_process, reserved;
  const float* heatmaps; const float* center; const float* scale;
  float* preds; float* maxvals; float* coords; int32_t* idx;
} buctd_decode_args;
int buctd_decode_heatmaps(buctd_ctx* ctx, const buctd_decode_args* args, buctd_stream stream);

/* final_layer + decode in one kernel, for callers that only want the keypoints: feat act16 [B][H+1][W+1][a_cs] (C channels,
 * the last stage's branch-0 tensor), weight = the packed 16-bit final_layer weights [K rows][w_cs] (1x1), bias float32 [K];
 * outputs as buctd_decode_heatmaps.  The heat-map values are the same 16-bit products accumulated in float32 in channel
 * order instead of the tensor core's order: maxvals agree with the unfused path to float32 rounding (~1e-6 relative), the
 * arg-max wherever the top-2 margin exceeds that.  Replaces final_layer + get_final_preds,
 * lib/models/pose_hrnet_coam.py:565 -> lib/core/inference.py:51-87, without the heat-map round trip. */
typedef struct {
  int32_t dtype, batch, k, h, w, c, a_cs, w_cs, post_process, reserved;
  const void* feat; const void* weight; const float* bias; const float* center; const float* scale;
  float* preds; float* maxvals; int32_t* idx;
  void* scratch;        /* B * ceil(H W / 256) * K * 8 bytes */
} buctd_head_decode_args;
int buctd_head_decode(buctd_ctx* ctx, const buctd_head_decode_args* args, buctd_stream stream);

/* Condition rasteriser: keypoint proposals -> blurred impulse maps written into channels
 * [c_off, c_off+Cc) of an nchw32 tensor out[B][out_channels][H][W].
 *  mode 0 stacked: Cc = K, one channel per keypoint, each scaled to peak 255;
 *  mode 1 coloured: Cc = 3, impulses carry colors[k][3], later keypoints overwrite, one global scale;
 *  mode 2 plain:   Cc = 3, one map of all keypoints, repeated, truncated to integer values.
 * kpts float32 [B][K][kstride] (x, y first; truncated toward zero like .astype(int)); a keypoint
 * is drawn only if 0 < x < W and 0 < y < H, at pixel (y-1, x-1); 15x15 Gaussian, sigma 2.6,
 * BORDER_REFLECT_101.  Replaces get_stacked_condition / get_condition_image_colored /
 * get_condition_image + generate_heatmap, lib/dataset/JointsDataset.py:457-543. */
typedef struct {
  int32_t batch, k, kstride, h, w, mode, out_channels, c_off;
  const float* kpts; const float* colors; /* [K][3] for mode 1, else NULL */
  float* out;
  void* scratch;        /* modes 1/2: >= 8*B bytes of device scratch (per-crop maximum); mode 0: unused */
} buctd_raster_args;
int buctd_rasterize_condition(buctd_ctx* ctx, const buctd_raster_args* args, buctd_stream stream);

/* ------------------------------------------------------------------------------------------
 * The callers either side of the forward (a-L, a-T).
 * ------------------------------------------------------------------------------------------ */
#define BUCTD_MAX_FLIP_CHANNELS 64
/* Mirror + channel permutation of an nchw32 tensor: out[b][c][y][x] = x[b][perm[c]][y][W-1-x].
 * With perm = identity on the RGB planes and the dataset's flip_pairs swapped on the stacked condition planes
 * this is the flip-test input of validate(): input[:, :3].flip(3) ++ flip_hm(input[:, 3:]),
 * lib/core/function.py:213-217 and lib/utils/transforms.py:49-55 (which goes through host NumPy). */
typedef struct {
  int32_t batch, channels, h, w;
  int32_t perm[BUCTD_MAX_FLIP_CHANNELS];
  const float* x; float* out;
} buctd_flip_args;
int buctd_flip_channels(buctd_ctx* ctx, const buctd_flip_args* args, buctd_stream stream);

/* Flip-test merge: out = (hm + S) * 0.5 with F = flip_back(hm_flipped, flip_pairs) (mirror + joint swap,
 * perm[k] = partner of k) and S = F shifted one column to the right when shift != 0 (column 0 keeps F[0]:
 * "output_flipped[:, :, :, 1:] = output_flipped.clone()[:, :, :, 0:-1]").  out may alias hm.
 * Replaces flip_back + SHIFT_HEATMAP + average, lib/core/function.py:226-236, lib/utils/transforms.py:16-30
 * (a D2H + H2D round trip of the heat-maps in the reference). */
typedef struct {
  int32_t batch, k, h, w, shift, reserved;
  int32_t perm[BUCTD_MAX_FLIP_CHANNELS];
  const float* hm; const float* hm_flipped; float* out;
} buctd_flip_merge_args;
int buctd_flip_merge(buctd_ctx* ctx, const buctd_flip_merge_args* args, buctd_stream stream);

/* JointsMSELoss: *loss = 0.5 / (K B HW) * sum (w[b][k] (output - target))^2 (w = 1 without use_target_weight)
 * and, if grad != NULL, grad = d loss / d output = w^2 (output - target) / (K B HW).  output, target, grad:
 * float32 [B][K][HW]; target_weight float32 [B][K]; scratch: B*K doubles.  Fixed-order reduction in float64
 * (bit-reproducible).  Replaces JointsMSELoss.forward (+ its autograd backward), lib/core/loss.py:17-41. */
typedef struct {
  int32_t batch, k, hw, use_target_weight;
  float grad_scale;     /* static loss scale applied to `grad` only (0 is read as 1): 16-bit gradient tensors downstream */
  int32_t reserved;
  const float* output; const float* target; const float* target_weight;
  float* loss; float* grad; double* scratch;
} buctd_joints_mse_args;
int buctd_joints_mse(buctd_ctx* ctx, const buctd_joints_mse_args* args, buctd_stream stream);

/* One Adam step over a flat float32 buffer of n parameters (torch.optim.Adam as get_optimizer builds it,
 * lib/utils/utils.py:268-272: only lr is passed, so betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad):
 *   m += (g - m)(1 - beta1);  v = v beta2 + (1 - beta2) g^2;  p -= lr/(1 - beta1^step) * m / (sqrt(v)/sqrt(1 - beta2^step) + eps)
 * step counts from 1.  All four buffers 16-byte aligned. */
typedef struct {
  int64_t n;
  double lr, beta1, beta2, eps;
  int32_t step, reserved;
  float* param; const float* grad; float* exp_avg; float* exp_avg_sq;
} buctd_adam_args;
int buctd_adam_step(buctd_ctx* ctx, const buctd_adam_args* args, buctd_stream stream);

/* Crop warp + normalise: for every crop b the 8-bit source image images[b]
 * ([img_h][img_w][3], row pitch img_stride bytes, device memory) is warped to [h][w] exactly as
 * cv2.warpAffine(img, trans, (w, h), flags=cv2.INTER_LINEAR) does (fixed-point bilinear, constant border 0), then
 * ToTensor + Normalize: out[b][c_off + c][y][x] = (v / 255 - mean[c]) / std[c] in float32 (nchw32, out_channels
 * planes per crop: the condition planes of the same tensor are written by buctd_rasterize_condition).
 * inv_trans: float64 [B][6], the INVERSE of the reference's `trans` (cv2 inverts the forward matrix itself:
 * cv::invertAffineTransform arithmetic, done once per crop by the caller, buctd_b200/crop.py).  swap_rb != 0 reads
 * BGR and writes RGB planes (cv2.cvtColor(..., COLOR_BGR2RGB) before the warp, DATASET.COLOR_RGB).  out_u8 (optional):
 * the 8-bit crop [B][h][w][3] itself.  Replaces cv2.warpAffine + transforms.ToTensor + Normalize,
 * lib/dataset/JointsDataset.py:262-288, tools/inference.py:86-107 (CPU, DataLoader workers). */
typedef struct {
  int32_t batch, h, w, out_channels, c_off, swap_rb;
  float mean[3], std[3];
  const uint8_t* const* images; const int32_t* img_h; const int32_t* img_w; const int64_t* img_stride;
  const double* inv_trans;
  float* out; uint8_t* out_u8;
} buctd_crop_warp_args;
int buctd_crop_warp(buctd_ctx* ctx, const buctd_crop_warp_args* args, buctd_stream stream);

/* Training targets: target[b][k] = the (2 tmp_size + 1)^2 Gaussian stamp centred on
 * mu = int(joint / feat_stride + 0.5), clipped to the [h][w] plane, pasted only when the (possibly zeroed) weight
 * is > 0.5; target_weight[b][k] = joints_vis[b][k][0], zeroed when the stamp lies completely outside, times
 * joints_weight[k] when given.  joints, joints_vis: float64 [B][K][3] (the reference keeps them in float64);
 * stamp: float32 [(2 tmp_size + 1)^2] = exp(-((x - x0)^2 + (y - y0)^2) / (2 sigma^2)) as the caller's numpy computes
 * it.  Replaces JointsDataset.generate_target, lib/dataset/JointsDataset.py:397-453 (CPU, DataLoader workers). */
typedef struct {
  int32_t batch, k, h, w, tmp_size, reserved;
  double feat_stride_x, feat_stride_y;
  const double* joints; const double* joints_vis; const float* stamp; const float* joints_weight; /* [K] or NULL */
  float* target; float* target_weight;
} buctd_gen_target_args;
int buctd_generate_target(buctd_ctx* ctx, const buctd_gen_target_args* args, buctd_stream stream);

/* PCK accuracy of core.evaluate.accuracy (lib/core/evaluate.py:14-70) on the arg-max coordinates of the output and
 * target heat-maps (buctd_decode_heatmaps, post_process = 0, `coords`): acc float64 [1 + K] (acc[0] = average over
 * the joints that have at least one labelled crop, acc[1 + k] = fraction of crops with normalised distance < thr,
 * -1 when joint k is labelled nowhere), *cnt = number of such joints.  pred, target: float32 [B][K][2]; h, w: the
 * heat-map size (normalisation (h/10, w/10) applied to (x, y) as the reference does). */
typedef struct {
  int32_t batch, k, h, w;
  double thr;
  const float* pred; const float* target;
  double* acc; int32_t* cnt;
} buctd_pck_args;
int buctd_pck_accuracy(buctd_ctx* ctx, const buctd_pck_args* args, buctd_stream stream);

/* ------------------------------------------------------------------------------------------
 * The training step (SURVEY.md section 8 row a-T): lib/core/function.py:113-138 -- model.train() forward with batch
 * statistics and attention dropout, JointsMSELoss, backward, Adam.  The data gradients of every convolution / linear
 * layer are tap-GEMMs with transposed (and tap-flipped) packed weights (buctd_tapgemm; a stride-2 convolution through the
 * zero-stuffed output gradient); everything else the autograd of the reference does is below.  All 16-bit gradient
 * tensors are act16 with zero pads, like the activations.  `inv_scale` = 1 / (static loss scale): parameter gradients
 * are written unscaled in float32 (torch layouts), straight into the optimizer's flat gradient buffer.
 * ------------------------------------------------------------------------------------------ */

/* D[M][N] = alpha * sum_k A[m][k] B[n][k] on tcgen05; each operand K-major ([MN rows][K contiguous], row stride ld) or
 * MN-major ([K rows][MN contiguous]).  z = (tap t, split s, batch b): batch b offsets both operands by *_batch_stride
 * elements (K ranges never cross a batch: 3-D tensor maps), split s contracts k in [s Kc, (s+1) Kc), tap t adds
 * tap_off[t] to the K-row index of an MN-major B (rows outside [0, b_rows) read as zero).  D element offset
 * = t d_t_stride + s d_s_stride + b d_b_stride + m ldd + n; float32 or 16-bit; accumulate != 0 adds to what is there.
 * Uses: weight gradients dW_t[n][k] = sum_r dY[r][n] X[r + off_t][k] (A, B MN-major act16 tensors, split partials),
 * nn.Linear weight gradients, and the products of the attention backward (torch.matmul autograd). */
typedef struct {
  int32_t dtype, a_mn_major, b_mn_major, m, n, k;
  int32_t nbatch, nsplit, ntaps, b_rows;       /* b_rows: K rows of one batch of an MN-major B (0: k) */
  int32_t lda, ldb, ldd, out_f32, accumulate;
  int32_t tap_group;    /* 3: the taps come in groups of three with row offsets o, o+1, o+2 (the dx taps of a kernel row): one CTA computes
                           a group from one A tile and one widened B stage (a third of the operand traffic); 0 / 1: one tap per CTA */
  int64_t a_batch_stride, b_batch_stride, d_t_stride, d_s_stride, d_b_stride;
  int32_t tap_off[9]; float alpha;
  const void* a; const void* b; void* d;
} buctd_gemm16_args;
int buctd_gemm16(buctd_ctx* ctx, const buctd_gemm16_args* args, buctd_stream stream);

/* Split partials [nsplit][ntaps][n][k] float32 (buctd_gemm16) -> grad[n][k][ntaps] float32 (nn.Conv2d weight layout
 * [cout][cin][kh][kw]; nn.Linear [out][in] with ntaps = 1), summed over the splits in index order, times inv_scale. */
typedef struct {
  int32_t n, k, ntaps, nsplit, accumulate, reserved; float inv_scale, pad_;
  const float* partial; float* grad;
} buctd_wgrad_reduce_args;
int buctd_wgrad_reduce(buctd_ctx* ctx, const buctd_wgrad_reduce_args* args, buctd_stream stream);

/* nn.BatchNorm2d(training=True) statistics of an act16 tensor z (C channels, the B*H*W valid pixels): mean_rstd[0][c] =
 * mean, [1][c] = 1 / sqrt(biased var + eps) (float64 accumulation, fixed order); running_mean / running_var updated in
 * place with `momentum` and the UNBIASED variance (lib/models/pose_hrnet_coam.py:20 BN_MOMENTUM = 0.1) when not NULL.
 * scratch: 2 * C * BUCTD_BN_BLOCKS doubles. */
#define BUCTD_BN_BLOCKS 296
typedef struct {
  int32_t dtype, batch, h, w, c, z_cs; float eps, momentum;
  const void* z; float* mean_rstd; float* running_mean; float* running_var; double* scratch;
} buctd_bn_stats_args;
int buctd_bn_stats(buctd_ctx* ctx, const buctd_bn_stats_args* args, buctd_stream stream);

/* y = [ReLU]( gamma (z - mean) rstd + beta [+ residual] ), act16 in / out (pads written as zero). */
typedef struct {
  int32_t dtype, batch, h, w, c, z_cs, res_cs, out_cs, relu, reserved;
  const void* z; const float* mean_rstd; const float* gamma; const float* beta; const void* residual; void* out;
} buctd_bn_apply_args;
int buctd_bn_apply(buctd_ctx* ctx, const buctd_bn_apply_args* args, buctd_stream stream);

/* Backward of the above: g = dy * [y > 0] (relu != 0); dbeta = sum g, dgamma = sum g zhat (zhat = (z - mean) rstd), both
 * times inv_scale; dz = gamma rstd (g - mean(g) - zhat mean(g zhat)); dres (optional) = g (res_accumulate: += g), the
 * gradient of the residual input.  y is only read when relu != 0.  scratch: 2 * C * (BUCTD_BN_BLOCKS + 1) doubles. */
typedef struct {
  int32_t dtype, batch, h, w, c, dy_cs, y_cs, z_cs, dz_cs, dres_cs, relu, res_accumulate;
  float inv_scale, pad_;
  const void* dy; const void* y; const void* z; const float* mean_rstd; const float* gamma;
  float* dgamma; float* dbeta; double* scratch; void* dz; void* dres;
} buctd_bn_bwd_args;
int buctd_bn_bwd(buctd_ctx* ctx, const buctd_bn_bwd_args* args, buctd_stream stream);

/* dbias[c] (+)= inv_scale * sum over rows of dy[r][c]  (16-bit [rows][cs] matrix; pads are zero).  scratch: C * BUCTD_BN_BLOCKS doubles */
typedef struct {
  int32_t dtype, c, cs, accumulate; int64_t rows; float inv_scale, pad_;
  const void* dy; float* dbias; double* scratch;
} buctd_colsum_args;
int buctd_colsum(buctd_ctx* ctx, const buctd_colsum_args* args, buctd_stream stream);

/* Zero-stuffing: out [B][2h+1][2w+1][cs] with out[b][2y][2x] = in[b][y][x] (in: act16 [B][h+1][w+1][cs]) and zeros elsewhere:
 * a stride-2 3x3 convolution's data gradient is the stride-1 transposed convolution of this, and its weight gradient the
 * stride-1 weight gradient against it. */
typedef struct { int32_t dtype, batch, h, w, cs, reserved; const void* in; void* out; } buctd_zero_stuff_args;
int buctd_zero_stuff(buctd_ctx* ctx, const buctd_zero_stuff_args* args, buctd_stream stream);

/* Backward of buctd_hr_fuse: g = dout * [out > 0]; term j at the output resolution gets g, a term 2^shift coarser the
 * sum of g over its 2^shift x 2^shift block (nearest up-sampling); term_accumulate[j] adds to dterm[j]. */
typedef struct {
  int32_t dtype, batch, h, w, c, cs, nterms, relu;
  const void* dout; const void* out;
  void* dterm[4]; int32_t term_cs[4]; int32_t term_shift[4]; int32_t term_accumulate[4];
} buctd_fuse_bwd_args;
int buctd_hr_fuse_bwd(buctd_ctx* ctx, const buctd_fuse_bwd_args* args, buctd_stream stream);

/* out (+)= a for act16 tensors of one geometry (gradient accumulation where no kernel epilogue can do it). */
typedef struct { int32_t dtype, cs_out, cs_a, c; int64_t rows; int32_t accumulate, reserved; const void* a; void* out; } buctd_add16_args;
int buctd_add16(buctd_ctx* ctx, const buctd_add16_args* args, buctd_stream stream);

/* planar [B][C][p_stride] <-> act16 [B][H+1][W+1][cs] (valid pixels; act16 pads / spare channels written as zero). */
typedef struct { int32_t dtype, batch, h, w, c, cs, p_stride, accumulate; const void* in; void* out; } buctd_relayout_args;
int buctd_planar_to_act16(buctd_ctx* ctx, const buctd_relayout_args* args, buctd_stream stream);
int buctd_act16_to_planar(buctd_ctx* ctx, const buctd_relayout_args* args, buctd_stream stream);

/* Dropout of the attention probabilities (nn.Dropout(0.1), lib/models/self_attention.py:38,84,115,156) is a counter-based
 * hash of (seed, crop, row, column) so that the test oracle reproduces the mask:
 *   x = seed ^ (row * 0x9E3779B1 + col * 0x85EBCA77 + crop * 0xC2B2AE3D);  x ^= x >> 16; x *= 0x7FEB352D; x ^= x >> 15;
 *   x *= 0x846CA68B; x ^= x >> 16;   keep iff x >= floor(p * 2^32)   (all uint32 arithmetic), kept values are scaled by 1 / (1 - p). */

/* Softmax backward on the [C x C] channel-attention matrices: probs [B][C][C] (buctd_coam_channel_att leaves them in
 * scratch chunk 0, stride prob_stride floats per crop), da_d = gradient w.r.t. the DROPPED probabilities;
 * ds = probs (da - rowsum(da probs)) * scale with da = mask da_d / (1 - p); writes ds and the dropped probabilities as
 * 16-bit [B][C][ld] matrices (GEMM operands). */
typedef struct {
  int32_t dtype, batch, c, ld; int64_t prob_stride; float scale, drop_p; uint32_t seed; int32_t reserved;
  const float* probs; const float* da_d; void* ds; void* probs_d; const uint32_t* seed_xor;
} buctd_chan_softmax_bwd_args;
int buctd_chan_softmax_bwd(buctd_ctx* ctx, const buctd_chan_softmax_bwd_args* args, buctd_stream stream);

/* Position-attention backward, element-wise part over the [rows x rows] matrices of every crop (rows = (H+1)(W+1), row
 * stride ld): s = log2-domain scores q.k * log2(e)/sqrt(C) and dpd = dO V^T (float32, buctd_gemm16), lse2[b][r] = the
 * forward's log2-sum-exp, delta[b][r] = sum_c dO O.  p = exp2(s - lse2) on valid (query, key) pixels, else 0;
 * pd = mask p / (1 - drop_p); ds = p (mask dpd / (1 - drop_p) - delta) / sqrt(C).  Writes pd and ds as 16-bit. */
typedef struct {
  int32_t dtype, batch, h, w, ld, reserved; float inv_sqrt_c, drop_p; uint32_t seed; int32_t reserved2;
  const float* s; const float* dpd; const float* lse2; const float* delta; void* pd; void* ds; const uint32_t* seed_xor;
} buctd_pos_softmax_bwd_args;
int buctd_pos_softmax_bwd(buctd_ctx* ctx, const buctd_pos_softmax_bwd_args* args, buctd_stream stream);

/* The same result without the float32 score matrices: per (query tile, key tile) both products S = Q K^T and dP = dO V^T are
 * accumulated in TMEM and the soft-max backward runs in the epilogue (2 x 6.4 GB of float32 less at batch 32, 96x72).
 * q, k, v, d_o: act16 [B][(H+1)(W+1)][cs]; pd, ds: 16-bit [B][rows][ld] as above. */
typedef struct {
  int32_t dtype, batch, h, w, c, q_cs, k_cs, v_cs, do_cs, ld; float drop_p; uint32_t seed;
  const void* q; const void* k; const void* v; const void* d_o; const float* lse2; const float* delta; void* pd; void* ds;
  const uint32_t* seed_xor;
} buctd_pos_bwd_scores_args;
int buctd_pos_att_bwd_scores(buctd_ctx* ctx, const buctd_pos_bwd_scores_args* args, buctd_stream stream);

/* delta[r] = sum_c a[r][c] b[r][c] (float32) over 16-bit [rows][cs] matrices */
typedef struct { int32_t dtype, c, a_cs, b_cs; int64_t rows; const void* a; const void* b; float* out; } buctd_rowdot_args;
int buctd_rowdot(buctd_ctx* ctx, const buctd_rowdot_args* args, buctd_stream stream);

/* Weight gradient of the stem conv1 (3x3 stride 2, 3 -> 64, float32 nchw input): partial[blk][64][27] float32 with
 * blk < BUCTD_BN_BLOCKS (reduce with buctd_wgrad_reduce: n 64, k 27, ntaps 1, nsplit BUCTD_BN_BLOCKS). */
typedef struct { int32_t dtype, batch, x_channels, h, w, dz_cs, x_half, reserved; const void* x; const void* dz; float* partial; } buctd_stem_wgrad_args;
int buctd_stem_wgrad(buctd_ctx* ctx, const buctd_stem_wgrad_args* args, buctd_stream stream);

/* Re-pack float32 master weights into the 16-bit tap-GEMM layouts after an optimizer step, table driven (one launch):
 * entry e: src [cout][cin][taps] float32 at param + src_off -> dst at w16 + dst_off, forward layout [tap][n_alloc][k_cs]
 * (transposed = 0) or data-gradient layout [taps-1-tap][cin_alloc][cout_cs] holding W[co][ci][tap] at [..][ci][co]
 * (transposed = 1).  table: int64 [n_entries][8] = {src_off, dst_off, cout, cin, taps (<= 9), rows_alloc, cs, transposed};
 * units: int32 [n_units][4] = {entry, cout tile, cin tile, 0}, one per 32 x 32 tile of every entry.  Padding rows / columns of
 * w16 are not written: allocate it zeroed. */
typedef struct { int32_t dtype, n_entries, n_units, reserved; const float* param; void* w16; const int64_t* table; const int32_t* units; } buctd_repack_args;
int buctd_repack_weights(buctd_ctx* ctx, const buctd_repack_args* args, buctd_stream stream);

/* Utility kernels used by the host mirror: zero-fill, and act16 -> nchw32 export for tests. */
int buctd_act16_to_nchw32(buctd_ctx* ctx, int dtype, const void* act, int batch, int h, int w, int c, int cs,
                          float* out, buctd_stream stream);
int buctd_nchw32_to_act16(buctd_ctx* ctx, int dtype, const float* x, int batch, int x_channels, int c_off, int c,
                          int h, int w, int cs, void* out, buctd_stream stream);

/* ------------------------------------------------------------------------------------------
 * Plans: the host mirror records one forward pass (a few hundred launches) once per (config,
 * batch) and replays it with a single call; tensor maps are encoded at record time.
 * kind = one of BUCTD_OP_*, args = the matching struct above (copied).
 * ------------------------------------------------------------------------------------------ */
enum { BUCTD_OP_TAPGEMM = 1, BUCTD_OP_STEM = 2, BUCTD_OP_FUSE = 3, BUCTD_OP_RESIZE = 4, BUCTD_OP_CHAN_ATT = 5,
       BUCTD_OP_POS_ATT = 6, BUCTD_OP_DA_COMBINE = 7, BUCTD_OP_PRENET = 8, BUCTD_OP_DECODE = 9,
       BUCTD_OP_RASTER = 10, BUCTD_OP_CHAIN = 11 };
#define BUCTD_MAX_LANES 8
int buctd_plan_create(buctd_ctx* ctx, buctd_plan** out);
int buctd_plan_destroy(buctd_plan* plan);
int buctd_plan_add(buctd_plan* plan, int kind, const void* args, size_t args_size);
/* Lanes: HRNet's branches (and the terms of an exchange/fuse step) are independent chains.  An op added
 * with lane L runs in program order with the other ops of lane L and concurrently with other lanes;
 * dependencies BETWEEN lanes are derived by the library from the device byte ranges every op reads and
 * writes (RAW, WAW and WAR against the last conflicting op of each other lane), so the result is the
 * one the serial program order gives.  max_ctas (> 0) caps the CTAs of a persistent tap-GEMM: its share
 * of the SMs while other lanes run beside it.  buctd_plan_add == lane 0, all SMs. */
int buctd_plan_add_lane(buctd_plan* plan, int kind, const void* args, size_t args_size, int lane, int max_ctas);
int buctd_plan_num_lanes(const buctd_plan* plan);
/* Capture the whole plan once into a CUDA graph (lanes become parallel branches); buctd_plan_run then
 * costs one cudaGraphLaunch on the caller's stream.  The plan is frozen afterwards. */
int buctd_plan_instantiate(buctd_plan* plan);
int buctd_plan_size(const buctd_plan* plan);
/* all ops: the instantiated graph if there is one, else the lanes on internal streams forked from and
 * joined back into `stream` */
int buctd_plan_run(buctd_plan* plan, buctd_stream stream);
/* run ops [first, last) only, serially in program order on `stream` (stage-boundary parity tests) */
int buctd_plan_run_range(buctd_plan* plan, int first, int last, buctd_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* BUCTD_B200_H_ */
