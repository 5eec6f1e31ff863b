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

struct PlanOp {
  int kind = 0;
  int nlaunch = 1;
  TapGemmLaunch tg;     // kind == TAPGEMM
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

static int make_op(int kind, const void* args, size_t size, PlanOp* op, int sm_count) {
  BUCTD_CHECK(args != nullptr, BUCTD_ERR_INVALID, "null args");
  op->kind = kind;
  switch (kind) {
    case BUCTD_OP_TAPGEMM: {
      buctd_tapgemm_args a;
      int rc = copy_args(&a, args, size, "tapgemm");
      return rc ? rc : tapgemm_prepare(a, &op->tg, sm_count);
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

}  // namespace buctd

struct buctd_plan {
  buctd_ctx* ctx = nullptr;
  std::vector<buctd::PlanOp> ops;
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
  if (!rc) rc = pos_att_init_device();
  if (!rc) rc = pos_att_tc_init_device();
  if (!rc && get_encode_tiled() == nullptr) {
    set_error("buctd_create: cuTensorMapEncodeTiled driver entry point not found");
    rc = BUCTD_ERR_NO_DEVICE;
  }
  cudaSetDevice(prev);
  if (rc) return rc;
  buctd_ctx* c = new buctd_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  *out = c;
  return BUCTD_OK;
}

BUCTD_API int buctd_destroy(buctd_ctx* ctx) {
  delete ctx;
  return BUCTD_OK;
}

BUCTD_API long long buctd_launch_count(const buctd_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }

static int immediate(buctd_ctx* ctx, int kind, const void* args, size_t size, buctd_stream stream) {
  BUCTD_CHECK(ctx != nullptr, BUCTD_ERR_INVALID, "null ctx");
  PlanOp op;
  int rc = make_op(kind, args, size, &op, ctx->sm_count);
  if (rc) return rc;
  rc = run_op(op, static_cast<cudaStream_t>(stream));
  if (!rc) ctx->launches += op.nlaunch;
  return rc;
}

BUCTD_API int buctd_tapgemm(buctd_ctx* c, const buctd_tapgemm_args* a, buctd_stream s) {
  return immediate(c, BUCTD_OP_TAPGEMM, a, sizeof(*a), s);
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
BUCTD_API int buctd_plan_add(buctd_plan* plan, int kind, const void* args, size_t size) {
  BUCTD_CHECK(plan != nullptr, BUCTD_ERR_INVALID, "buctd_plan_add: null plan");
  PlanOp op;
  int rc = make_op(kind, args, size, &op, plan->ctx->sm_count);
  if (rc) return rc;
  plan->ops.push_back(op);
  return BUCTD_OK;
}
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
  return buctd_plan_run_range(plan, 0, plan ? (int)plan->ops.size() : 0, stream);
}
