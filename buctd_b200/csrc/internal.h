// Internal (non-ABI) declarations shared by the translation units of libbuctd_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "../../include/buctd_b200.h"

struct buctd_ctx {
  int device = 0;
  int sm_count = 0;
  std::atomic<long long> launches{0};
  unsigned int* barriers = nullptr;          // device: grid-barrier counters of the chain tap-GEMMs (owned by the handle)
  int next_barrier = 0;
  std::mutex barrier_mutex;
  static constexpr int kBarriers = 1 << 21;   // 8 MB of flags: a W48 forward at B = 64 uses ~0.15 M
};

namespace buctd {

struct TapGemmKParams {
  int mode, ntaps, stride, num_kc, k16, wp;
  long long m_rows;
  int B, H, W;
  int tw, th, tiles_x, tiles_y;
  int n, n_alloc;
  int out_cs, res_cs, relu, out_mode;
  uint32_t a_tile_bytes;
  uint32_t wp_mul, wp_sh, hp_mul, hp_sh;   // fastdiv pairs for (W+1) and (H+1)
  const float* bias;
  const void* residual;
  void* out;
};

// slab (FLAT-mode) kernel plan: super-tile, rings and shared-memory carve-up
struct SlabParams {
  int msub;          // 128-row sub-tiles per super-tile
  int halo;          // slab rows loaded before/after the super-tile (W+2 for 3x3, 0 for 1x1)
  int nboxes, box_rows, slab_bytes;
  int sa, sw;        // slab ring / weight ring depth (stages)
  int tg;            // taps per weight-ring stage
  int stage_bufs;    // epilogue stage buffers (2: residual prefetch + store overlap)
  int w_resident, w_tiles;
  int num_super, num_items;
  int tmem_cols, smem_bytes;
  int desc_mode;
  int direct;        // epilogue writes rows straight to global memory (no stage buffers, no DMA warp)
  int jmajor;        // resident weights + one K chunk: sub-tile-major MMA order (per-sub-tile hand-off to the epilogue)
};

// (mul, shift) with n / d == (umulhi(n, mul) + n) >> shift for every n < 2^31 (round-up method, 33-bit magic)
inline void fastdiv_make(uint32_t d, uint32_t* mul, uint32_t* shift) {
  uint32_t s = 0;
  while ((1ull << s) < d) ++s;
  *mul = (uint32_t)(((1ull << 32) * ((1ull << s) - d)) / d + 1);
  *shift = s;
}

// CTA-pair GEMM (pairgemm.cu): the plan of a large nn.Linear
struct PairGemmParams {
  int k16, num_kc;
  int n, n_store;            // true output columns (bias range) / columns that exist in an output row
  int m_tiles, num_tiles;    // 256 x 256 tiles, m fastest
  int relu;
  const float* bias;
};

// everything a tap-GEMM launch needs, computed (tensor maps encoded) once
struct TapGemmLaunch {
  alignas(64) CUtensorMap map_a;
  alignas(64) CUtensorMap map_b;
  alignas(64) CUtensorMap map_out;   // 16-bit output tile store (unused for the fp32 head)
  alignas(64) CUtensorMap map_res;   // residual tile load (copy of map_out when there is no residual)
  TapGemmKParams kp;
  SlabParams sp;
  dim3 grid;
  int bn;
  int dtype;
  int pair;                          // the launch is a pairgemm (cta_group::2) launch: map_a / map_b / map_out + pp
  PairGemmParams pp;
};

int tapgemm_prepare(const buctd_tapgemm_args& a, TapGemmLaunch* L, int sm_count);
int tapgemm_launch(const TapGemmLaunch& L, cudaStream_t st);
// [rows][cs] 16-bit matrix seen through SWIZZLE_128B boxes of 64 channels x box_rows rows (inner = channels that exist)
int tapgemm_encode_rows_map(CUtensorMap* m, int dtype, const void* base, int inner, int cs, long long rows, int box_rows);
int tapgemm_init_device();
bool pairgemm_wanted(const buctd_tapgemm_args& a, int num_kc);
int pairgemm_prepare(const buctd_tapgemm_args& a, TapGemmLaunch* L, int ctas);
int pairgemm_launch(const TapGemmLaunch& L, cudaStream_t st);
int pairgemm_init_device();
int pair_max_clusters(int* out);       // CTA pairs of a one-CTA-per-SM kernel that can be co-resident on the device
// set around tapgemm_prepare by a caller that plans for a CTA pair: every weight tile is split between the two CTAs
// (shared-memory carve-up and the weight tensor map's box use bn / 2 rows)
extern thread_local int tl_slab_pair;

// immediate launchers of the CUDA-core kernels (each returns a BUCTD_* code)
int launch_stem(const buctd_stem_args& a, cudaStream_t st);
int launch_fuse(const buctd_fuse_args& a, cudaStream_t st);
int launch_resize(const buctd_resize_args& a, cudaStream_t st);
int launch_cond_from_kpts(const buctd_cond_kpts_args& a, cudaStream_t st);
int launch_resize3(const buctd_resize3_args& a, cudaStream_t st);
int launch_chan_att(const buctd_chan_att_args& a, cudaStream_t st, int* nlaunch);
int launch_pos_att(const buctd_pos_att_args& a, cudaStream_t st);
int launch_da_combine(const buctd_da_combine_args& a, cudaStream_t st);
int launch_prenet(const buctd_prenet_args& a, cudaStream_t st);
int launch_decode(const buctd_decode_args& a, cudaStream_t st);
int launch_head_decode(const buctd_head_decode_args& a, cudaStream_t st);
int launch_raster(const buctd_raster_args& a, cudaStream_t st, int* nlaunch);
int launch_act16_to_nchw32(int dtype, const void* act, int B, int H, int W, int C, int cs, float* out, cudaStream_t st);
int launch_nchw32_to_act16(int dtype, const float* x, int B, int xc, int c_off, int C, int H, int W, int cs, void* out,
                           cudaStream_t st);
int launch_flip_channels(const buctd_flip_args& a, cudaStream_t st);
int launch_flip_merge(const buctd_flip_merge_args& a, cudaStream_t st);
int launch_joints_mse(const buctd_joints_mse_args& a, cudaStream_t st);
int launch_adam(const buctd_adam_args& a, int sm_count, cudaStream_t st);
int launch_gen_target(const buctd_gen_target_args& a, cudaStream_t st);
int launch_pck(const buctd_pck_args& a, cudaStream_t st);
int launch_crop_warp(const buctd_crop_warp_args& a, cudaStream_t st);
int gemm16_init_device();
int launch_gemm16(const buctd_gemm16_args& a, cudaStream_t st);
int launch_wgrad_reduce(const buctd_wgrad_reduce_args& a, cudaStream_t st);
int launch_bn_stats(const buctd_bn_stats_args& a, cudaStream_t st);
int launch_bn_apply(const buctd_bn_apply_args& a, cudaStream_t st);
int launch_bn_bwd(const buctd_bn_bwd_args& a, cudaStream_t st);
int launch_colsum(const buctd_colsum_args& a, cudaStream_t st);
int launch_zero_stuff(const buctd_zero_stuff_args& a, cudaStream_t st);
int launch_fuse_bwd(const buctd_fuse_bwd_args& a, cudaStream_t st);
int launch_add16(const buctd_add16_args& a, cudaStream_t st);
int launch_planar_to_act16(const buctd_relayout_args& a, cudaStream_t st);
int launch_act16_to_planar(const buctd_relayout_args& a, cudaStream_t st);
int launch_chan_softmax_bwd(const buctd_chan_softmax_bwd_args& a, cudaStream_t st);
int launch_pos_softmax_bwd(const buctd_pos_softmax_bwd_args& a, cudaStream_t st);
int pos_att_bwd_init_device();
int launch_pos_bwd_scores(const buctd_pos_bwd_scores_args& a, cudaStream_t st);
int launch_rowdot(const buctd_rowdot_args& a, cudaStream_t st);
int launch_stem_wgrad(const buctd_stem_wgrad_args& a, cudaStream_t st);
int launch_repack(const buctd_repack_args& a, cudaStream_t st);
struct ChainLaunch;
int tapgemm_chain_init_device();
int tapgemm_chain_prepare(const buctd_tapgemm_chain_args& a, ChainLaunch* C, int sm_count, unsigned int* counter);
int tapgemm_chain_launch(const ChainLaunch& C, cudaStream_t st);
ChainLaunch* tapgemm_chain_new();
size_t tapgemm_chain_flag_count(const ChainLaunch& C);
void tapgemm_chain_set_flags(ChainLaunch* C, unsigned int* flags);
void tapgemm_chain_delete(ChainLaunch* c);
int pos_att_tc_init_device();
bool pos_att_tc_supported(const buctd_pos_att_args& a);
int launch_pos_att_tc(const buctd_pos_att_args& a, cudaStream_t st);

}  // namespace buctd
