// C ABI of libbuctd_b200.so: handles, immediate launches, and recorded plans.
#include <memory>
#include <string>

#include "common.cuh"
#include "internal.h"

namespace buctd {

static thread_local std::string g_err;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

thread_local bool tl_use_pdl = true;

struct Span {
  uintptr_t lo, hi;     // [lo, hi) device bytes
};

struct PlanOp {
  int kind = 0;
  int nlaunch = 1;
  int lane = 0;                 // concurrent lane (internal stream / parallel graph branch)
  int ndeps = 0;                // cross-lane dependencies: the last conflicting op of every other lane
  int deps[BUCTD_MAX_LANES];
  bool record = false;          // some op on another lane waits for this one
  TapGemmLaunch tg;     // kind == TAPGEMM
  buctd_tapgemm_args tga;
  std::shared_ptr<ChainLaunch> chain;                       // kind == CHAIN
  std::shared_ptr<buctd_tapgemm_chain_args> chain_args;
  union {
    buctd_stem_args stem;
    buctd_fuse_args fuse;
    buctd_resize_args resize;
    buctd_chan_att_args chan;
    buctd_pos_att_args pos;
    buctd_da_combine_args da;
    buctd_prenet_args prenet;
    buctd_decode_args decode;
    buctd_raster_args raster;
  } u;
};

static int run_op(PlanOp& op, cudaStream_t st) {
  switch (op.kind) {
    case BUCTD_OP_TAPGEMM: return tapgemm_launch(op.tg, st);
    case BUCTD_OP_CHAIN: return tapgemm_chain_launch(*op.chain, st);
    case BUCTD_OP_STEM: return launch_stem(op.u.stem, st);
    case BUCTD_OP_FUSE: return launch_fuse(op.u.fuse, st);
    case BUCTD_OP_RESIZE: return launch_resize(op.u.resize, st);
    case BUCTD_OP_CHAN_ATT: return launch_chan_att(op.u.chan, st, &op.nlaunch);
    case BUCTD_OP_POS_ATT: return launch_pos_att(op.u.pos, st);
    case BUCTD_OP_DA_COMBINE: return launch_da_combine(op.u.da, st);
    case BUCTD_OP_PRENET: return launch_prenet(op.u.prenet, st);
    case BUCTD_OP_DECODE: return launch_decode(op.u.decode, st);
    case BUCTD_OP_RASTER: return launch_raster(op.u.raster, st, &op.nlaunch);
  }
  set_error("plan: unknown op kind %d", op.kind);
  return BUCTD_ERR_INVALID;
}

template <typename A>
static int copy_args(A* dst, const void* src, size_t size, const char* name) {
  BUCTD_CHECK(size == sizeof(A), BUCTD_ERR_INVALID, "%s: args struct size %zu, library expects %zu (ABI mismatch)", name,
              size, sizeof(A));
  memcpy(dst, src, sizeof(A));
  return BUCTD_OK;
}

static int make_op(int kind, const void* args, size_t size, PlanOp* op, int sm_count, buctd_ctx* ctx) {
  BUCTD_CHECK(args != nullptr, BUCTD_ERR_INVALID, "null args");
  op->kind = kind;
  switch (kind) {
    case BUCTD_OP_CHAIN: {
      op->chain_args = std::make_shared<buctd_tapgemm_chain_args>();
      int rc = copy_args(op->chain_args.get(), args, size, "tapgemm_chain");
      if (rc) return rc;
      BUCTD_CHECK(ctx != nullptr && ctx->barriers != nullptr, BUCTD_ERR_INVALID, "tapgemm_chain: handle has no barrier buffer");
      op->chain = std::shared_ptr<ChainLaunch>(tapgemm_chain_new(), tapgemm_chain_delete);
      op->nlaunch = 1;
      rc = tapgemm_chain_prepare(*op->chain_args, op->chain.get(), sm_count, ctx->barriers);
      if (rc) return rc;
      // a private range of flags [n_layers][num_super] out of the handle's pool (ranges are reused round-robin: chain
      // launches of one handle are stream-ordered and each zeroes its range first)
      const int need = (int)tapgemm_chain_flag_count(*op->chain);
      BUCTD_CHECK(need <= buctd_ctx::kBarriers, BUCTD_ERR_INVALID, "tapgemm_chain: %d flags exceed the pool", need);
      int start;
      {
        std::lock_guard<std::mutex> lock(ctx->barrier_mutex);
        start = ctx->next_barrier;
        if (start + need > buctd_ctx::kBarriers) start = 0;
        ctx->next_barrier = start + need;
      }
      tapgemm_chain_set_flags(op->chain.get(), ctx->barriers + start);
      return BUCTD_OK;
    }
    case BUCTD_OP_TAPGEMM: {
      int rc = copy_args(&op->tga, args, size, "tapgemm");
      return rc ? rc : tapgemm_prepare(op->tga, &op->tg, sm_count);
    }
    case BUCTD_OP_STEM: return copy_args(&op->u.stem, args, size, "stem");
    case BUCTD_OP_FUSE: return copy_args(&op->u.fuse, args, size, "fuse");
    case BUCTD_OP_RESIZE: return copy_args(&op->u.resize, args, size, "resize");
    case BUCTD_OP_CHAN_ATT: return copy_args(&op->u.chan, args, size, "chan_att");
    case BUCTD_OP_POS_ATT: return copy_args(&op->u.pos, args, size, "pos_att");
    case BUCTD_OP_DA_COMBINE: return copy_args(&op->u.da, args, size, "da_combine");
    case BUCTD_OP_PRENET: return copy_args(&op->u.prenet, args, size, "prenet");
    case BUCTD_OP_DECODE: return copy_args(&op->u.decode, args, size, "decode");
    case BUCTD_OP_RASTER: return copy_args(&op->u.raster, args, size, "raster");
  }
  set_error("unknown op kind %d", kind);
  return BUCTD_ERR_INVALID;
}

// ---- what an op reads and writes (byte ranges; over-approximation is safe) ---------------------
static Span span(const void* p, long long bytes) {
  Span s;
  s.lo = reinterpret_cast<uintptr_t>(p);
  s.hi = s.lo + (uintptr_t)(bytes > 0 ? bytes : 0);
  return s;
}
static long long act_bytes(int b, int h, int w, int cs) { return 2ll * b * (h + 1) * (w + 1) * cs; }

static void op_spans(const PlanOp& op, std::vector<Span>* rd, std::vector<Span>* wr) {
  rd->clear();
  wr->clear();
  switch (op.kind) {
    case BUCTD_OP_CHAIN: {
      for (int l = 0; l < op.chain_args->n_layers; ++l) {
        const buctd_tapgemm_args& a = op.chain_args->layer[l];
        rd->push_back(span(a.a, 2ll * a.m_rows * a.a_cs));
        if (a.residual) rd->push_back(span(a.residual, 2ll * a.m_rows * a.res_cs));
        wr->push_back(span(a.out, 2ll * a.m_rows * a.out_cs));
      }
      break;
    }
    case BUCTD_OP_TAPGEMM: {
      const buctd_tapgemm_args& a = op.tga;
      const bool flat = a.mode == BUCTD_TG_FLAT;
      rd->push_back(span(a.a, flat ? 2ll * a.m_rows * a.a_cs : act_bytes(a.batch, a.in_h, a.in_w, a.a_cs)));
      const long long orows = flat ? a.m_rows : (long long)a.batch * (a.out_h + 1) * (a.out_w + 1);
      if (a.residual) rd->push_back(span(a.residual, 2ll * orows * a.res_cs));
      if (a.out_mode == BUCTD_OUT_NCHW32) wr->push_back(span(a.out, 4ll * a.batch * a.n * a.out_h * a.out_w));
      else wr->push_back(span(a.out, 2ll * orows * a.out_cs));
      break;
    }
    case BUCTD_OP_FUSE: {
      const buctd_fuse_args& a = op.u.fuse;
      for (int j = 0; j < a.nterms; ++j)
        rd->push_back(span(a.term[j], act_bytes(a.batch, a.h >> a.term_shift[j], a.w >> a.term_shift[j], a.term_cs[j])));
      wr->push_back(span(a.out, act_bytes(a.batch, a.h, a.w, a.out_cs)));
      break;
    }
    case BUCTD_OP_CHAN_ATT: {
      const buctd_chan_att_args& a = op.u.chan;
      rd->push_back(span(a.y, act_bytes(a.batch, a.h, a.w, a.y_cs)));
      rd->push_back(span(a.yc, act_bytes(a.batch, a.h, a.w, a.yc_cs)));
      const long long nchunks = ((long long)(a.h + 1) * (a.w + 1) + BUCTD_CHAN_ATT_ROWS_PER_CHUNK - 1) / BUCTD_CHAN_ATT_ROWS_PER_CHUNK;
      wr->push_back(span(a.scratch, 4ll * a.batch * nchunks * a.c * a.c));
      wr->push_back(span(a.out, 2ll * a.batch * a.c * (a.out_stride > 0 ? a.out_stride : a.h * a.w)));
      break;
    }
    case BUCTD_OP_POS_ATT: {
      const buctd_pos_att_args& a = op.u.pos;
      rd->push_back(span(a.q, act_bytes(a.batch, a.h, a.w, a.q_cs)));
      rd->push_back(span(a.k, act_bytes(a.batch, a.h, a.w, a.k_cs)));
      rd->push_back(span(a.v, act_bytes(a.batch, a.h, a.w, a.v_cs)));
      wr->push_back(span(a.o, act_bytes(a.batch, a.h, a.w, a.o_cs)));
      break;
    }
    case BUCTD_OP_DA_COMBINE: {
      const buctd_da_combine_args& a = op.u.da;
      rd->push_back(span(a.x, act_bytes(a.batch, a.h, a.w, a.x_cs)));
      if (a.p_out) rd->push_back(span(a.p_out, act_bytes(a.batch, a.h, a.w, a.p_cs)));
      rd->push_back(span(a.c_out, 2ll * a.batch * a.c * (a.c_stride > 0 ? a.c_stride : a.h * a.w)));
      wr->push_back(span(a.out, act_bytes(a.batch, a.h, a.w, a.out_cs)));
      break;
    }
    default:
      // ops that touch caller tensors of unknown extent (stem, resize, prenet, decode, raster) order against
      // everything: the whole address space is both read and written
      rd->push_back(Span{0, ~(uintptr_t)0});
      wr->push_back(Span{0, ~(uintptr_t)0});
  }
}

static bool overlap(const std::vector<Span>& a, const std::vector<Span>& b) {
  for (const Span& x : a)
    for (const Span& y : b)
      if (x.lo < y.hi && y.lo < x.hi) return true;
  return false;
}

// For every op: the LAST conflicting (RAW / WAW / WAR) earlier op of every other lane.  Ops of one lane run in
// program order on one stream, so earlier conflicts of that lane are implied.
static void analyse_deps(std::vector<PlanOp>& ops) {
  const int n = (int)ops.size();
  std::vector<std::vector<Span>> rd(n), wr(n);
  for (int i = 0; i < n; ++i) {
    op_spans(ops[i], &rd[i], &wr[i]);
    ops[i].ndeps = 0;
    ops[i].record = false;
  }
  for (int i = 0; i < n; ++i) {
    bool have[BUCTD_MAX_LANES] = {false};
    have[ops[i].lane] = true;
    int found = 1;
    for (int j = i - 1; j >= 0 && found < BUCTD_MAX_LANES; --j) {
      const int lj = ops[j].lane;
      if (have[lj]) continue;
      if (overlap(wr[j], rd[i]) || overlap(wr[j], wr[i]) || overlap(rd[j], wr[i])) {
        have[lj] = true;
        ++found;
        ops[i].deps[ops[i].ndeps++] = j;
        ops[j].record = true;
      }
    }
  }
}

}  // namespace buctd

struct buctd_plan {
  buctd_ctx* ctx = nullptr;
  std::vector<buctd::PlanOp> ops;
  int nlanes = 1;
  bool analysed = false;
  std::vector<cudaEvent_t> events;                 // one per op that another lane waits for, + fork/join
  cudaStream_t lane_stream[BUCTD_MAX_LANES] = {};
  cudaEvent_t fork_ev = nullptr, join_ev[BUCTD_MAX_LANES] = {};
  cudaStream_t capture_stream = nullptr;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  long long graph_launches = 0;
  ~buctd_plan() {
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    for (cudaEvent_t e : events) if (e) cudaEventDestroy(e);
    if (fork_ev) cudaEventDestroy(fork_ev);
    for (int l = 0; l < BUCTD_MAX_LANES; ++l) {
      if (join_ev[l]) cudaEventDestroy(join_ev[l]);
      if (lane_stream[l]) cudaStreamDestroy(lane_stream[l]);
    }
    if (capture_stream) cudaStreamDestroy(capture_stream);
  }
};

using namespace buctd;

#define BUCTD_API extern "C" __attribute__((visibility("default")))

BUCTD_API int buctd_abi_version(void) { return BUCTD_ABI_VERSION; }
BUCTD_API const char* buctd_last_error(void) { return g_err.c_str(); }

BUCTD_API int buctd_create(int device, buctd_ctx** out) {
  BUCTD_CHECK(out != nullptr, BUCTD_ERR_INVALID, "buctd_create: out is null");
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    set_error("buctd_create: no CUDA device visible");
    return BUCTD_ERR_NO_DEVICE;
  }
  BUCTD_CHECK(device >= 0 && device < n, BUCTD_ERR_INVALID, "buctd_create: device %d out of range [0,%d)", device, n);
  cudaDeviceProp prop;
  BUCTD_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("buctd_create: device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major, prop.minor);
    return BUCTD_ERR_NO_DEVICE;
  }
  int prev = 0;
  BUCTD_CUDA(cudaGetDevice(&prev));
  BUCTD_CUDA(cudaSetDevice(device));
  int rc = tapgemm_init_device();
  if (!rc) rc = tapgemm_chain_init_device();
  if (!rc) rc = pos_att_tc_init_device();
  if (!rc) rc = gemm16_init_device();
  if (!rc) rc = pos_att_bwd_init_device();
  if (!rc) rc = pairgemm_init_device();
  if (!rc && get_encode_tiled() == nullptr) {
    set_error("buctd_create: cuTensorMapEncodeTiled driver entry point not found");
    rc = BUCTD_ERR_NO_DEVICE;
  }
  unsigned int* barriers = nullptr;
  if (!rc && cudaMalloc(&barriers, sizeof(unsigned int) * buctd_ctx::kBarriers) != cudaSuccess) {
    set_error("buctd_create: cudaMalloc of the barrier counters failed");
    rc = BUCTD_ERR_CUDA;
  }
  cudaSetDevice(prev);
  if (rc) return rc;
  buctd_ctx* c = new buctd_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->barriers = barriers;
  *out = c;
  return BUCTD_OK;
}

BUCTD_API int buctd_destroy(buctd_ctx* ctx) {
  if (ctx && ctx->barriers) cudaFree(ctx->barriers);
  delete ctx;
  return BUCTD_OK;
}

BUCTD_API long long buctd_launch_count(const buctd_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }

static int immediate(buctd_ctx* ctx, int kind, const void* args, size_t size, buctd_stream stream) {
  BUCTD_CHECK(ctx != nullptr, BUCTD_ERR_INVALID, "null ctx");
  PlanOp op;
  int rc = make_op(kind, args, size, &op, ctx->sm_count, ctx);
  if (rc) return rc;
  rc = run_op(op, static_cast<cudaStream_t>(stream));
  if (!rc) ctx->launches += op.nlaunch;
  return rc;
}

BUCTD_API int buctd_tapgemm(buctd_ctx* c, const buctd_tapgemm_args* a, buctd_stream s) {
  return immediate(c, BUCTD_OP_TAPGEMM, a, sizeof(*a), s);
}
BUCTD_API int buctd_tapgemm_chain(buctd_ctx* c, const buctd_tapgemm_chain_args* a, buctd_stream s) {
  return immediate(c, BUCTD_OP_CHAIN, a, sizeof(*a), s);
}
BUCTD_API int buctd_stem_conv(buctd_ctx* c, const buctd_stem_args* a, buctd_stream s) {
  return immediate(c, BUCTD_OP_STEM, a, sizeof(*a), s);
}
BUCTD_API int buctd_hr_fuse(buctd_ctx* c, const buctd_fuse_args* a, buctd_stream s) {
  return immediate(c, BUCTD_OP_FUSE, a, sizeof(*a), s);
}
BUCTD_API int buctd_cond_resize(buctd_ctx* c, const buctd_resize_args* a, buctd_stream s) {
  return immediate(c, BUCTD_OP_RESIZE, a, sizeof(*a), s);
}
BUCTD_API int buctd_coam_channel_att(buctd_ctx* c, const buctd_chan_att_args* a, buctd_stream s) {
  return immediate(c, BUCTD_OP_CHAN_ATT, a, sizeof(*a), s);
}
BUCTD_API int buctd_coam_position_att(buctd_ctx* c, const buctd_pos_att_args* a, buctd_stream s) {
  return immediate(c, BUCTD_OP_POS_ATT, a, sizeof(*a), s);
}
BUCTD_API int buctd_da_combine(buctd_ctx* c, const buctd_da_combine_args* a, buctd_stream s) {
  return immediate(c, BUCTD_OP_DA_COMBINE, a, sizeof(*a), s);
}
BUCTD_API int buctd_prenet(buctd_ctx* c, const buctd_prenet_args* a, buctd_stream s) {
  return immediate(c, BUCTD_OP_PRENET, a, sizeof(*a), s);
}
BUCTD_API int buctd_decode_heatmaps(buctd_ctx* c, const buctd_decode_args* a, buctd_stream s) {
  return immediate(c, BUCTD_OP_DECODE, a, sizeof(*a), s);
}
BUCTD_API int buctd_rasterize_condition(buctd_ctx* c, const buctd_raster_args* a, buctd_stream s) {
  return immediate(c, BUCTD_OP_RASTER, a, sizeof(*a), s);
}
template <typename A, typename F>
static int simple_call(buctd_ctx* c, const A* a, int nlaunch, F&& fn) {
  BUCTD_CHECK(c != nullptr && a != nullptr, BUCTD_ERR_INVALID, "null ctx or args");
  int rc = fn(*a);
  if (!rc) c->launches += nlaunch;
  return rc;
}
BUCTD_API int buctd_cond_resize3(buctd_ctx* c, const buctd_resize3_args* a, buctd_stream s) {
  return simple_call(c, a, 1, [&](const buctd_resize3_args& x) { return launch_resize3(x, static_cast<cudaStream_t>(s)); });
}
BUCTD_API int buctd_head_decode(buctd_ctx* c, const buctd_head_decode_args* a, buctd_stream s) {
  return simple_call(c, a, 1, [&](const buctd_head_decode_args& x) { return launch_head_decode(x, static_cast<cudaStream_t>(s)); });
}
BUCTD_API int buctd_cond_from_kpts(buctd_ctx* c, const buctd_cond_kpts_args* a, buctd_stream s) {
  return simple_call(c, a, 1, [&](const buctd_cond_kpts_args& x) { return launch_cond_from_kpts(x, static_cast<cudaStream_t>(s)); });
}
BUCTD_API int buctd_flip_channels(buctd_ctx* c, const buctd_flip_args* a, buctd_stream s) {
  return simple_call(c, a, 1, [&](const buctd_flip_args& x) { return launch_flip_channels(x, static_cast<cudaStream_t>(s)); });
}
BUCTD_API int buctd_flip_merge(buctd_ctx* c, const buctd_flip_merge_args* a, buctd_stream s) {
  return simple_call(c, a, 1, [&](const buctd_flip_merge_args& x) { return launch_flip_merge(x, static_cast<cudaStream_t>(s)); });
}
BUCTD_API int buctd_joints_mse(buctd_ctx* c, const buctd_joints_mse_args* a, buctd_stream s) {
  return simple_call(c, a, 2, [&](const buctd_joints_mse_args& x) { return launch_joints_mse(x, static_cast<cudaStream_t>(s)); });
}
BUCTD_API int buctd_adam_step(buctd_ctx* c, const buctd_adam_args* a, buctd_stream s) {
  return simple_call(c, a, 1, [&](const buctd_adam_args& x) { return launch_adam(x, c->sm_count, static_cast<cudaStream_t>(s)); });
}
BUCTD_API int buctd_crop_warp(buctd_ctx* c, const buctd_crop_warp_args* a, buctd_stream s) {
  return simple_call(c, a, 1, [&](const buctd_crop_warp_args& x) { return launch_crop_warp(x, static_cast<cudaStream_t>(s)); });
}
BUCTD_API int buctd_generate_target(buctd_ctx* c, const buctd_gen_target_args* a, buctd_stream s) {
  return simple_call(c, a, 1, [&](const buctd_gen_target_args& x) { return launch_gen_target(x, static_cast<cudaStream_t>(s)); });
}
BUCTD_API int buctd_pck_accuracy(buctd_ctx* c, const buctd_pck_args* a, buctd_stream s) {
  return simple_call(c, a, 1, [&](const buctd_pck_args& x) { return launch_pck(x, static_cast<cudaStream_t>(s)); });
}
BUCTD_API int buctd_gemm16(buctd_ctx* c, const buctd_gemm16_args* a, buctd_stream s) {
  return simple_call(c, a, 1, [&](const buctd_gemm16_args& x) { return launch_gemm16(x, static_cast<cudaStream_t>(s)); });
}
BUCTD_API int buctd_wgrad_reduce(buctd_ctx* c, const buctd_wgrad_reduce_args* a, buctd_stream s) {
  return simple_call(c, a, 1, [&](const buctd_wgrad_reduce_args& x) { return launch_wgrad_reduce(x, static_cast<cudaStream_t>(s)); });
}
BUCTD_API int buctd_bn_stats(buctd_ctx* c, const buctd_bn_stats_args* a, buctd_stream s) {
  return simple_call(c, a, 2, [&](const buctd_bn_stats_args& x) { return launch_bn_stats(x, static_cast<cudaStream_t>(s)); });
}
BUCTD_API int buctd_bn_apply(buctd_ctx* c, const buctd_bn_apply_args* a, buctd_stream s) {
  return simple_call(c, a, 1, [&](const buctd_bn_apply_args& x) { return launch_bn_apply(x, static_cast<cudaStream_t>(s)); });
}
BUCTD_API int buctd_bn_bwd(buctd_ctx* c, const buctd_bn_bwd_args* a, buctd_stream s) {
  return simple_call(c, a, 3, [&](const buctd_bn_bwd_args& x) { return launch_bn_bwd(x, static_cast<cudaStream_t>(s)); });
}
BUCTD_API int buctd_colsum(buctd_ctx* c, const buctd_colsum_args* a, buctd_stream s) {
  return simple_call(c, a, 2, [&](const buctd_colsum_args& x) { return launch_colsum(x, static_cast<cudaStream_t>(s)); });
}
BUCTD_API int buctd_zero_stuff(buctd_ctx* c, const buctd_zero_stuff_args* a, buctd_stream s) {
  return simple_call(c, a, 1, [&](const buctd_zero_stuff_args& x) { return launch_zero_stuff(x, static_cast<cudaStream_t>(s)); });
}
BUCTD_API int buctd_hr_fuse_bwd(buctd_ctx* c, const buctd_fuse_bwd_args* a, buctd_stream s) {
  return simple_call(c, a, 2, [&](const buctd_fuse_bwd_args& x) { return launch_fuse_bwd(x, static_cast<cudaStream_t>(s)); });
}
BUCTD_API int buctd_add16(buctd_ctx* c, const buctd_add16_args* a, buctd_stream s) {
  return simple_call(c, a, 1, [&](const buctd_add16_args& x) { return launch_add16(x, static_cast<cudaStream_t>(s)); });
}
BUCTD_API int buctd_planar_to_act16(buctd_ctx* c, const buctd_relayout_args* a, buctd_stream s) {
  return simple_call(c, a, 1, [&](const buctd_relayout_args& x) { return launch_planar_to_act16(x, static_cast<cudaStream_t>(s)); });
}
BUCTD_API int buctd_act16_to_planar(buctd_ctx* c, const buctd_relayout_args* a, buctd_stream s) {
  return simple_call(c, a, 1, [&](const buctd_relayout_args& x) { return launch_act16_to_planar(x, static_cast<cudaStream_t>(s)); });
}
BUCTD_API int buctd_chan_softmax_bwd(buctd_ctx* c, const buctd_chan_softmax_bwd_args* a, buctd_stream s) {
  return simple_call(c, a, 1, [&](const buctd_chan_softmax_bwd_args& x) { return launch_chan_softmax_bwd(x, static_cast<cudaStream_t>(s)); });
}
BUCTD_API int buctd_pos_softmax_bwd(buctd_ctx* c, const buctd_pos_softmax_bwd_args* a, buctd_stream s) {
  return simple_call(c, a, 1, [&](const buctd_pos_softmax_bwd_args& x) { return launch_pos_softmax_bwd(x, static_cast<cudaStream_t>(s)); });
}
BUCTD_API int buctd_pos_att_bwd_scores(buctd_ctx* c, const buctd_pos_bwd_scores_args* a, buctd_stream s) {
  return simple_call(c, a, 1, [&](const buctd_pos_bwd_scores_args& x) { return launch_pos_bwd_scores(x, static_cast<cudaStream_t>(s)); });
}
BUCTD_API int buctd_rowdot(buctd_ctx* c, const buctd_rowdot_args* a, buctd_stream s) {
  return simple_call(c, a, 1, [&](const buctd_rowdot_args& x) { return launch_rowdot(x, static_cast<cudaStream_t>(s)); });
}
BUCTD_API int buctd_stem_wgrad(buctd_ctx* c, const buctd_stem_wgrad_args* a, buctd_stream s) {
  return simple_call(c, a, 1, [&](const buctd_stem_wgrad_args& x) { return launch_stem_wgrad(x, static_cast<cudaStream_t>(s)); });
}
BUCTD_API int buctd_repack_weights(buctd_ctx* c, const buctd_repack_args* a, buctd_stream s) {
  return simple_call(c, a, 1, [&](const buctd_repack_args& x) { return launch_repack(x, static_cast<cudaStream_t>(s)); });
}
BUCTD_API int buctd_act16_to_nchw32(buctd_ctx* c, int dtype, const void* act, int batch, int h, int w, int ch, int cs,
                                    float* out, buctd_stream s) {
  BUCTD_CHECK(c != nullptr, BUCTD_ERR_INVALID, "null ctx");
  int rc = launch_act16_to_nchw32(dtype, act, batch, h, w, ch, cs, out, static_cast<cudaStream_t>(s));
  if (!rc) c->launches += 1;
  return rc;
}
BUCTD_API int buctd_nchw32_to_act16(buctd_ctx* c, int dtype, const float* x, int batch, int xc, int c_off, int ch, int h,
                                    int w, int cs, void* out, buctd_stream s) {
  BUCTD_CHECK(c != nullptr, BUCTD_ERR_INVALID, "null ctx");
  int rc = launch_nchw32_to_act16(dtype, x, batch, xc, c_off, ch, h, w, cs, out, static_cast<cudaStream_t>(s));
  if (!rc) c->launches += 1;
  return rc;
}

BUCTD_API int buctd_plan_create(buctd_ctx* ctx, buctd_plan** out) {
  BUCTD_CHECK(ctx != nullptr && out != nullptr, BUCTD_ERR_INVALID, "buctd_plan_create: null argument");
  buctd_plan* p = new buctd_plan();
  p->ctx = ctx;
  *out = p;
  return BUCTD_OK;
}
BUCTD_API int buctd_plan_destroy(buctd_plan* plan) {
  delete plan;
  return BUCTD_OK;
}
BUCTD_API int buctd_plan_add_lane(buctd_plan* plan, int kind, const void* args, size_t size, int lane, int max_ctas) {
  BUCTD_CHECK(plan != nullptr, BUCTD_ERR_INVALID, "buctd_plan_add: null plan");
  BUCTD_CHECK(lane >= 0 && lane < BUCTD_MAX_LANES, BUCTD_ERR_INVALID, "buctd_plan_add_lane: lane %d out of range [0,%d)", lane,
              BUCTD_MAX_LANES);
  BUCTD_CHECK(plan->exec == nullptr, BUCTD_ERR_INVALID, "buctd_plan_add: plan already instantiated");
  PlanOp op;
  const int sms = plan->ctx->sm_count;
  int rc = make_op(kind, args, size, &op, max_ctas > 0 && max_ctas < sms ? max_ctas : sms, plan->ctx);
  if (rc) return rc;
  op.lane = lane;
  plan->ops.push_back(op);
  if (lane + 1 > plan->nlanes) plan->nlanes = lane + 1;
  plan->analysed = false;
  return BUCTD_OK;
}
BUCTD_API int buctd_plan_add(buctd_plan* plan, int kind, const void* args, size_t size) {
  return buctd_plan_add_lane(plan, kind, args, size, 0, 0);
}

// Enqueue every op on its lane: lane 0 = `origin`, the others are internal streams forked from / joined back
// into `origin` with events.  Works both directly on a caller stream and under stream capture.
static int run_lanes(buctd_plan* plan, cudaStream_t origin) {
  std::vector<PlanOp>& ops = plan->ops;
  if (!plan->analysed) {
    analyse_deps(ops);
    size_t need = 0;
    for (const PlanOp& op : ops) need += op.record ? 1 : 0;
    while (plan->events.size() < need) {
      cudaEvent_t e;
      BUCTD_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      plan->events.push_back(e);
    }
    if (!plan->fork_ev) BUCTD_CUDA(cudaEventCreateWithFlags(&plan->fork_ev, cudaEventDisableTiming));
    for (int l = 1; l < plan->nlanes; ++l) {
      if (!plan->lane_stream[l]) BUCTD_CUDA(cudaStreamCreateWithFlags(&plan->lane_stream[l], cudaStreamNonBlocking));
      if (!plan->join_ev[l]) BUCTD_CUDA(cudaEventCreateWithFlags(&plan->join_ev[l], cudaEventDisableTiming));
    }
    plan->analysed = true;
  }
  const int n = (int)ops.size();
  std::vector<cudaEvent_t> ev_of(n, nullptr);
  size_t next_ev = 0;
  auto stream_of = [&](int lane) { return lane == 0 ? origin : plan->lane_stream[lane]; };
  if (plan->nlanes > 1) {
    BUCTD_CUDA(cudaEventRecord(plan->fork_ev, origin));
    for (int l = 1; l < plan->nlanes; ++l) BUCTD_CUDA(cudaStreamWaitEvent(plan->lane_stream[l], plan->fork_ev, 0));
  }
  long long launches = 0;
  bool after_kernel[BUCTD_MAX_LANES] = {false};   // the previous node of this lane's stream is a kernel of this run
  for (int i = 0; i < n; ++i) {
    PlanOp& op = ops[i];
    cudaStream_t st = stream_of(op.lane);
    for (int d = 0; d < op.ndeps; ++d) BUCTD_CUDA(cudaStreamWaitEvent(st, ev_of[op.deps[d]], 0));
    // programmatic dependent launch only kernel-after-kernel on one stream; event waits are full dependencies
    tl_use_pdl = op.ndeps == 0 && after_kernel[op.lane];
    int rc = run_op(op, st);
    tl_use_pdl = true;
    if (rc) return rc;
    launches += op.nlaunch;
    after_kernel[op.lane] = true;
    if (op.record) {
      ev_of[i] = plan->events[next_ev++];
      BUCTD_CUDA(cudaEventRecord(ev_of[i], st));
      after_kernel[op.lane] = false;
    }
  }
  for (int l = 1; l < plan->nlanes; ++l) {
    BUCTD_CUDA(cudaEventRecord(plan->join_ev[l], plan->lane_stream[l]));
    BUCTD_CUDA(cudaStreamWaitEvent(origin, plan->join_ev[l], 0));
  }
  plan->graph_launches = launches;
  return BUCTD_OK;
}

BUCTD_API int buctd_plan_instantiate(buctd_plan* plan) {
  BUCTD_CHECK(plan != nullptr, BUCTD_ERR_INVALID, "buctd_plan_instantiate: null plan");
  if (plan->exec) return BUCTD_OK;
  if (!plan->capture_stream) BUCTD_CUDA(cudaStreamCreateWithFlags(&plan->capture_stream, cudaStreamNonBlocking));
  BUCTD_CUDA(cudaStreamBeginCapture(plan->capture_stream, cudaStreamCaptureModeThreadLocal));
  int rc = run_lanes(plan, plan->capture_stream);
  cudaGraph_t g = nullptr;
  cudaError_t e = cudaStreamEndCapture(plan->capture_stream, &g);
  if (rc) {
    if (g) cudaGraphDestroy(g);
    return rc;
  }
  if (e != cudaSuccess) {
    set_error("buctd_plan_instantiate: cudaStreamEndCapture failed: %s", cudaGetErrorString(e));
    return BUCTD_ERR_CUDA;
  }
  plan->graph = g;
  BUCTD_CUDA(cudaGraphInstantiate(&plan->exec, g, 0));
  return BUCTD_OK;
}
BUCTD_API int buctd_plan_num_lanes(const buctd_plan* plan) { return plan ? plan->nlanes : 0; }
BUCTD_API int buctd_plan_size(const buctd_plan* plan) { return plan ? (int)plan->ops.size() : 0; }

BUCTD_API int buctd_plan_run_range(buctd_plan* plan, int first, int last, buctd_stream stream) {
  BUCTD_CHECK(plan != nullptr, BUCTD_ERR_INVALID, "buctd_plan_run: null plan");
  BUCTD_CHECK(first >= 0 && last <= (int)plan->ops.size() && first <= last, BUCTD_ERR_INVALID,
              "buctd_plan_run_range: [%d,%d) out of range (plan has %zu ops)", first, last, plan->ops.size());
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long n = 0;
  for (int i = first; i < last; ++i) {
    int rc = run_op(plan->ops[i], st);
    if (rc) return rc;
    n += plan->ops[i].nlaunch;
  }
  plan->ctx->launches += n;
  return BUCTD_OK;
}
BUCTD_API int buctd_plan_run(buctd_plan* plan, buctd_stream stream) {
  BUCTD_CHECK(plan != nullptr, BUCTD_ERR_INVALID, "buctd_plan_run: null plan");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (plan->exec) {                       // instantiated: one graph launch, lanes are parallel branches
    BUCTD_CUDA(cudaGraphLaunch(plan->exec, st));
    plan->ctx->launches += plan->graph_launches;
    return BUCTD_OK;
  }
  if (plan->nlanes > 1) {                 // multi-lane, not instantiated: internal streams forked from the caller's
    int rc = run_lanes(plan, st);
    if (!rc) plan->ctx->launches += plan->graph_launches;
    return rc;
  }
  return buctd_plan_run_range(plan, 0, (int)plan->ops.size(), stream);
}
