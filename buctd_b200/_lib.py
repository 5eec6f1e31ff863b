"""ctypes binding of libbuctd_b200.so (the C ABI declared in include/buctd_b200.h).

The product path fails loudly when the CUDA library is missing or cannot be loaded: there is no
CPU or PyTorch fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BUCTD_LIB") or os.path.join(_HERE, "libbuctd_b200.so")   # BUCTD_LIB: debug builds (role timers)

ABI_VERSION = 3
F16, BF16 = 0, 1
TG_FLAT, TG_PATCH = 0, 1
OUT_ACT16, OUT_ROWS16, OUT_NCHW32 = 0, 1, 2
CHAN_ATT_ROWS_PER_CHUNK = 256
MAX_LANES = 8


def chan_att_scratch_floats(batch: int, h: int, w: int, c: int) -> int:
    """float32 elements of scratch buctd_coam_channel_att needs (per-chunk partial [C][C] score matrices)"""
    nchunks = ((h + 1) * (w + 1) + CHAN_ATT_ROWS_PER_CHUNK - 1) // CHAN_ATT_ROWS_PER_CHUNK
    return batch * nchunks * c * c
(OP_TAPGEMM, OP_STEM, OP_FUSE, OP_RESIZE, OP_CHAN_ATT, OP_POS_ATT, OP_DA_COMBINE, OP_PRENET, OP_DECODE,
 OP_RASTER, OP_CHAIN) = range(1, 12)
MAX_CHAIN = 8

i32, i64, vp, fp = C.c_int32, C.c_int64, C.c_void_p, C.c_void_p


class TapGemmArgs(C.Structure):
    _fields_ = [("dtype", i32), ("mode", i32), ("ntaps", i32), ("stride", i32), ("batch", i32),
                ("out_h", i32), ("out_w", i32), ("in_h", i32), ("in_w", i32), ("m_rows", i64),
                ("k", i32), ("a_cs", i32), ("n", i32), ("n_alloc", i32), ("w_cs", i32), ("out_cs", i32),
                ("res_cs", i32), ("relu", i32), ("out_mode", i32), ("reserved", i32),
                ("a", vp), ("w", vp), ("bias", fp), ("residual", vp), ("out", vp)]


class TapGemmChainArgs(C.Structure):
    _fields_ = [("n_layers", i32), ("reserved", i32), ("layer", TapGemmArgs * MAX_CHAIN)]


class StemArgs(C.Structure):
    _fields_ = [("dtype", i32), ("batch", i32), ("x_channels", i32), ("h", i32), ("w", i32), ("out_cs", i32),
                ("x_half", i32), ("no_relu", i32), ("x", fp), ("weight", fp), ("bias", fp), ("out", vp)]


class FuseArgs(C.Structure):
    _fields_ = [("dtype", i32), ("batch", i32), ("h", i32), ("w", i32), ("c", i32), ("out_cs", i32),
                ("nterms", i32), ("relu", i32), ("term", vp * 4), ("term_cs", i32 * 4), ("term_shift", i32 * 4),
                ("out", vp)]


class ResizeArgs(C.Structure):
    _fields_ = [("dtype", i32), ("batch", i32), ("x_channels", i32), ("c_off", i32), ("cc", i32), ("h", i32),
                ("w", i32), ("oh", i32), ("ow", i32), ("out_cs", i32), ("antialias", i32), ("reserved", i32),
                ("x", fp), ("out", vp)]


class Resize3Args(C.Structure):
    _fields_ = [("dtype", i32), ("batch", i32), ("x_channels", i32), ("c_off", i32), ("cc", i32), ("h", i32), ("w", i32),
                ("reserved", i32), ("out_cs", i32 * 3), ("reserved2", i32), ("x", fp), ("out", vp * 3)]


class CondKptsArgs(C.Structure):
    _fields_ = [("dtype", i32), ("batch", i32), ("k", i32), ("kstride", i32), ("h", i32), ("w", i32),
                ("out_cs", i32 * 3), ("reserved", i32), ("kpts", fp), ("out", vp * 3)]


class ChanAttArgs(C.Structure):
    _fields_ = [("dtype", i32), ("batch", i32), ("h", i32), ("w", i32), ("c", i32), ("y_cs", i32), ("yc_cs", i32),
                ("out_stride", i32), ("drop_p", C.c_float), ("drop_seed", C.c_uint32), ("out_padded", i32), ("reserved", i32),
                ("y", vp), ("yc", vp), ("scratch", fp), ("out", vp), ("seed_xor", vp)]


class PosAttArgs(C.Structure):
    _fields_ = [("dtype", i32), ("batch", i32), ("h", i32), ("w", i32), ("c", i32), ("q_cs", i32), ("k_cs", i32),
                ("v_cs", i32), ("o_cs", i32), ("reserved", i32), ("drop_p", C.c_float), ("drop_seed", C.c_uint32),
                ("q", vp), ("k", vp), ("v", vp), ("o", vp), ("lse2", fp), ("seed_xor", vp)]


class DaCombineArgs(C.Structure):
    _fields_ = [("dtype", i32), ("batch", i32), ("h", i32), ("w", i32), ("c", i32), ("x_cs", i32), ("p_cs", i32),
                ("out_cs", i32), ("channel_only", i32), ("c_stride", i32),
                ("x", vp), ("p_out", vp), ("c_out", vp), ("out", vp)]


class PrenetArgs(C.Structure):
    _fields_ = [("batch", i32), ("x_channels", i32), ("h", i32), ("w", i32),
                ("x", fp), ("w_eff", fp), ("b_eff", fp), ("wc", fp), ("out", fp)]


class DecodeArgs(C.Structure):
    _fields_ = [("batch", i32), ("k", i32), ("h", i32), ("w", i32), ("post_process", i32), ("reserved", i32),
                ("heatmaps", fp), ("center", fp), ("scale", fp), ("preds", fp), ("maxvals", fp), ("coords", fp),
                ("idx", vp)]


class HeadDecodeArgs(C.Structure):
    _fields_ = [("dtype", i32), ("batch", i32), ("k", i32), ("h", i32), ("w", i32), ("c", i32), ("a_cs", i32), ("w_cs", i32),
                ("post_process", i32), ("reserved", i32), ("feat", vp), ("weight", vp), ("bias", fp), ("center", fp), ("scale", fp),
                ("preds", fp), ("maxvals", fp), ("idx", vp), ("scratch", vp)]


class RasterArgs(C.Structure):
    _fields_ = [("batch", i32), ("k", i32), ("kstride", i32), ("h", i32), ("w", i32), ("mode", i32),
                ("out_channels", i32), ("c_off", i32), ("kpts", fp), ("colors", fp), ("out", fp), ("scratch", vp)]


MAX_FLIP_CHANNELS = 64


class FlipArgs(C.Structure):
    _fields_ = [("batch", i32), ("channels", i32), ("h", i32), ("w", i32), ("perm", i32 * MAX_FLIP_CHANNELS),
                ("x", fp), ("out", fp)]


class FlipMergeArgs(C.Structure):
    _fields_ = [("batch", i32), ("k", i32), ("h", i32), ("w", i32), ("shift", i32), ("reserved", i32),
                ("perm", i32 * MAX_FLIP_CHANNELS), ("hm", fp), ("hm_flipped", fp), ("out", fp)]


class JointsMseArgs(C.Structure):
    _fields_ = [("batch", i32), ("k", i32), ("hw", i32), ("use_target_weight", i32), ("grad_scale", C.c_float), ("reserved", i32),
                ("output", fp), ("target", fp), ("target_weight", fp), ("loss", fp), ("grad", fp), ("scratch", vp)]


class AdamArgs(C.Structure):
    _fields_ = [("n", i64), ("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double),
                ("step", i32), ("reserved", i32), ("param", fp), ("grad", fp), ("exp_avg", fp), ("exp_avg_sq", fp)]


class CropWarpArgs(C.Structure):
    _fields_ = [("batch", i32), ("h", i32), ("w", i32), ("out_channels", i32), ("c_off", i32), ("swap_rb", i32),
                ("mean", C.c_float * 3), ("std", C.c_float * 3),
                ("images", vp), ("img_h", vp), ("img_w", vp), ("img_stride", vp), ("inv_trans", vp), ("out", fp), ("out_u8", vp)]


class GenTargetArgs(C.Structure):
    _fields_ = [("batch", i32), ("k", i32), ("h", i32), ("w", i32), ("tmp_size", i32), ("reserved", i32),
                ("feat_stride_x", C.c_double), ("feat_stride_y", C.c_double),
                ("joints", vp), ("joints_vis", vp), ("stamp", fp), ("joints_weight", fp), ("target", fp), ("target_weight", fp)]


class PckArgs(C.Structure):
    _fields_ = [("batch", i32), ("k", i32), ("h", i32), ("w", i32), ("thr", C.c_double),
                ("pred", fp), ("target", fp), ("acc", vp), ("cnt", vp)]


f32c, u32, f64c = C.c_float, C.c_uint32, C.c_double


class Gemm16Args(C.Structure):
    _fields_ = [("dtype", i32), ("a_mn_major", i32), ("b_mn_major", i32), ("m", i32), ("n", i32), ("k", i32),
                ("nbatch", i32), ("nsplit", i32), ("ntaps", i32), ("b_rows", i32),
                ("lda", i32), ("ldb", i32), ("ldd", i32), ("out_f32", i32), ("accumulate", i32), ("tap_group", i32),
                ("a_batch_stride", i64), ("b_batch_stride", i64), ("d_t_stride", i64), ("d_s_stride", i64), ("d_b_stride", i64),
                ("tap_off", i32 * 9), ("alpha", f32c), ("a", vp), ("b", vp), ("d", vp)]


class WgradReduceArgs(C.Structure):
    _fields_ = [("n", i32), ("k", i32), ("ntaps", i32), ("nsplit", i32), ("accumulate", i32), ("reserved", i32),
                ("inv_scale", f32c), ("pad_", f32c), ("partial", fp), ("grad", fp)]


BN_BLOCKS = 296


class BnStatsArgs(C.Structure):
    _fields_ = [("dtype", i32), ("batch", i32), ("h", i32), ("w", i32), ("c", i32), ("z_cs", i32), ("eps", f32c), ("momentum", f32c),
                ("z", vp), ("mean_rstd", fp), ("running_mean", fp), ("running_var", fp), ("scratch", vp)]


class BnApplyArgs(C.Structure):
    _fields_ = [("dtype", i32), ("batch", i32), ("h", i32), ("w", i32), ("c", i32), ("z_cs", i32), ("res_cs", i32), ("out_cs", i32),
                ("relu", i32), ("reserved", i32), ("z", vp), ("mean_rstd", fp), ("gamma", fp), ("beta", fp), ("residual", vp), ("out", vp)]


class BnBwdArgs(C.Structure):
    _fields_ = [("dtype", i32), ("batch", i32), ("h", i32), ("w", i32), ("c", i32), ("dy_cs", i32), ("y_cs", i32), ("z_cs", i32),
                ("dz_cs", i32), ("dres_cs", i32), ("relu", i32), ("res_accumulate", i32), ("inv_scale", f32c), ("pad_", f32c),
                ("dy", vp), ("y", vp), ("z", vp), ("mean_rstd", fp), ("gamma", fp), ("dgamma", fp), ("dbeta", fp), ("scratch", vp),
                ("dz", vp), ("dres", vp)]


class ColsumArgs(C.Structure):
    _fields_ = [("dtype", i32), ("c", i32), ("cs", i32), ("accumulate", i32), ("rows", i64), ("inv_scale", f32c), ("pad_", f32c),
                ("dy", vp), ("dbias", fp), ("scratch", vp)]


class ZeroStuffArgs(C.Structure):
    _fields_ = [("dtype", i32), ("batch", i32), ("h", i32), ("w", i32), ("cs", i32), ("reserved", i32), ("in_", vp), ("out", vp)]


class FuseBwdArgs(C.Structure):
    _fields_ = [("dtype", i32), ("batch", i32), ("h", i32), ("w", i32), ("c", i32), ("cs", i32), ("nterms", i32), ("relu", i32),
                ("dout", vp), ("out", vp), ("dterm", vp * 4), ("term_cs", i32 * 4), ("term_shift", i32 * 4), ("term_accumulate", i32 * 4)]


class Add16Args(C.Structure):
    _fields_ = [("dtype", i32), ("cs_out", i32), ("cs_a", i32), ("c", i32), ("rows", i64), ("accumulate", i32), ("reserved", i32),
                ("a", vp), ("out", vp)]


class RelayoutArgs(C.Structure):
    _fields_ = [("dtype", i32), ("batch", i32), ("h", i32), ("w", i32), ("c", i32), ("cs", i32), ("p_stride", i32), ("accumulate", i32),
                ("in_", vp), ("out", vp)]


class ChanSoftmaxBwdArgs(C.Structure):
    _fields_ = [("dtype", i32), ("batch", i32), ("c", i32), ("ld", i32), ("prob_stride", i64), ("scale", f32c), ("drop_p", f32c),
                ("seed", u32), ("reserved", i32), ("probs", fp), ("da_d", fp), ("ds", vp), ("probs_d", vp), ("seed_xor", vp)]


class PosSoftmaxBwdArgs(C.Structure):
    _fields_ = [("dtype", i32), ("batch", i32), ("h", i32), ("w", i32), ("ld", i32), ("reserved", i32), ("inv_sqrt_c", f32c),
                ("drop_p", f32c), ("seed", u32), ("reserved2", i32), ("s", fp), ("dpd", fp), ("lse2", fp), ("delta", fp), ("pd", vp), ("ds", vp), ("seed_xor", vp)]


class PosBwdScoresArgs(C.Structure):
    _fields_ = [("dtype", i32), ("batch", i32), ("h", i32), ("w", i32), ("c", i32), ("q_cs", i32), ("k_cs", i32), ("v_cs", i32),
                ("do_cs", i32), ("ld", i32), ("drop_p", f32c), ("seed", u32), ("q", vp), ("k", vp), ("v", vp), ("d_o", vp),
                ("lse2", fp), ("delta", fp), ("pd", vp), ("ds", vp), ("seed_xor", vp)]


class RowdotArgs(C.Structure):
    _fields_ = [("dtype", i32), ("c", i32), ("a_cs", i32), ("b_cs", i32), ("rows", i64), ("a", vp), ("b", vp), ("out", fp)]


class StemWgradArgs(C.Structure):
    _fields_ = [("dtype", i32), ("batch", i32), ("x_channels", i32), ("h", i32), ("w", i32), ("dz_cs", i32), ("x_half", i32), ("reserved", i32),
                ("x", vp), ("dz", vp), ("partial", fp)]


class RepackArgs(C.Structure):
    _fields_ = [("dtype", i32), ("n_entries", i32), ("n_units", i32), ("reserved", i32), ("param", fp), ("w16", vp), ("table", vp), ("units", vp)]


TRAIN_EXPORTS = ["buctd_gemm16", "buctd_wgrad_reduce", "buctd_bn_stats", "buctd_bn_apply", "buctd_bn_bwd", "buctd_colsum", "buctd_zero_stuff",
                 "buctd_hr_fuse_bwd", "buctd_add16", "buctd_planar_to_act16", "buctd_act16_to_planar", "buctd_chan_softmax_bwd",
                 "buctd_pos_softmax_bwd", "buctd_pos_att_bwd_scores", "buctd_rowdot", "buctd_stem_wgrad", "buctd_repack_weights"]

OP_STRUCT = {OP_TAPGEMM: TapGemmArgs, OP_STEM: StemArgs, OP_FUSE: FuseArgs, OP_RESIZE: ResizeArgs,
             OP_CHAN_ATT: ChanAttArgs, OP_POS_ATT: PosAttArgs, OP_DA_COMBINE: DaCombineArgs, OP_PRENET: PrenetArgs,
             OP_DECODE: DecodeArgs, OP_RASTER: RasterArgs, OP_CHAIN: TapGemmChainArgs}

# every symbol include/buctd_b200.h declares (tests/test_abi.py checks the header against this list)
EXPORTS = ["buctd_abi_version", "buctd_last_error", "buctd_create", "buctd_destroy", "buctd_launch_count",
           "buctd_tapgemm", "buctd_tapgemm_chain", "buctd_stem_conv", "buctd_hr_fuse", "buctd_cond_resize", "buctd_cond_resize3", "buctd_cond_from_kpts", "buctd_coam_channel_att",
           "buctd_coam_position_att", "buctd_da_combine", "buctd_prenet", "buctd_decode_heatmaps", "buctd_head_decode",
           "buctd_rasterize_condition", "buctd_flip_channels", "buctd_flip_merge", "buctd_joints_mse", "buctd_adam_step",
           "buctd_crop_warp", "buctd_generate_target", "buctd_pck_accuracy", "buctd_act16_to_nchw32", "buctd_nchw32_to_act16", "buctd_plan_create",
           "buctd_plan_destroy", "buctd_plan_add", "buctd_plan_add_lane", "buctd_plan_num_lanes", "buctd_plan_instantiate",
           "buctd_plan_size", "buctd_plan_run", "buctd_plan_run_range"] + TRAIN_EXPORTS

_IMMEDIATE = {OP_TAPGEMM: "buctd_tapgemm", OP_STEM: "buctd_stem_conv", OP_FUSE: "buctd_hr_fuse",
              OP_RESIZE: "buctd_cond_resize", OP_CHAN_ATT: "buctd_coam_channel_att",
              OP_POS_ATT: "buctd_coam_position_att", OP_DA_COMBINE: "buctd_da_combine", OP_PRENET: "buctd_prenet",
              OP_DECODE: "buctd_decode_heatmaps", OP_RASTER: "buctd_rasterize_condition", OP_CHAIN: "buctd_tapgemm_chain"}


class BuctdError(RuntimeError):
    pass


_lib = None
_lock = threading.Lock()


def load():
    """dlopen the library (once).  Raises BuctdError if it has not been built."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise BuctdError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        lib.buctd_last_error.restype = C.c_char_p
        lib.buctd_launch_count.restype = C.c_longlong
        lib.buctd_launch_count.argtypes = [vp]
        lib.buctd_create.argtypes = [C.c_int, C.POINTER(vp)]
        lib.buctd_destroy.argtypes = [vp]
        for name in list(_IMMEDIATE.values()) + ["buctd_cond_resize3", "buctd_cond_from_kpts", "buctd_head_decode", "buctd_flip_channels", "buctd_flip_merge", "buctd_joints_mse", "buctd_adam_step",
                                                     "buctd_crop_warp", "buctd_generate_target", "buctd_pck_accuracy"] + TRAIN_EXPORTS:
            getattr(lib, name).argtypes = [vp, vp, vp]
        lib.buctd_act16_to_nchw32.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]
        lib.buctd_nchw32_to_act16.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                              C.c_int, vp, vp]
        lib.buctd_plan_create.argtypes = [vp, C.POINTER(vp)]
        lib.buctd_plan_destroy.argtypes = [vp]
        lib.buctd_plan_add.argtypes = [vp, C.c_int, vp, C.c_size_t]
        lib.buctd_plan_add_lane.argtypes = [vp, C.c_int, vp, C.c_size_t, C.c_int, C.c_int]
        lib.buctd_plan_instantiate.argtypes = [vp]
        lib.buctd_plan_num_lanes.argtypes = [vp]
        lib.buctd_plan_size.argtypes = [vp]
        lib.buctd_plan_run.argtypes = [vp, vp]
        lib.buctd_plan_run_range.argtypes = [vp, C.c_int, C.c_int, vp]
        if lib.buctd_abi_version() != ABI_VERSION:
            raise BuctdError("libbuctd_b200.so ABI %d != binding ABI %d: rebuild" % (lib.buctd_abi_version(), ABI_VERSION))
        _lib = lib
        return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().buctd_last_error().decode("utf-8", "replace")
        raise BuctdError("%s failed (%d): %s" % (what or "buctd call", rc, msg))


class Context:
    """Per-device handle (buctd_create / buctd_destroy)."""
    _by_device = {}

    def __init__(self, device: int):
        self.lib = load()
        h = vp()
        check(self.lib.buctd_create(int(device), C.byref(h)), "buctd_create(%d)" % device)
        self.handle = h
        self.device = int(device)

    @classmethod
    def get(cls, device: int) -> "Context":
        with _lock:
            ctx = cls._by_device.get(device)
        if ctx is None:
            ctx = cls(device)
            with _lock:
                cls._by_device[device] = ctx
        return ctx

    def launch_count(self) -> int:
        return int(self.lib.buctd_launch_count(self.handle))

    def run(self, kind: int, args: C.Structure, stream: int):
        fn = getattr(self.lib, _IMMEDIATE[kind])
        check(fn(self.handle, C.byref(args), vp(stream)), _IMMEDIATE[kind])

    def call(self, name: str, args: C.Structure, stream: int):
        """immediate entry points outside the plan op table (flip / loss / Adam)"""
        check(getattr(self.lib, name)(self.handle, C.byref(args), vp(stream)), name)

    def act16_to_nchw32(self, dtype, act_ptr, b, h, w, c, cs, out_ptr, stream):
        check(self.lib.buctd_act16_to_nchw32(self.handle, dtype, vp(act_ptr), b, h, w, c, cs, vp(out_ptr), vp(stream)),
              "buctd_act16_to_nchw32")

    def nchw32_to_act16(self, dtype, x_ptr, b, xc, c_off, c, h, w, cs, out_ptr, stream):
        check(self.lib.buctd_nchw32_to_act16(self.handle, dtype, vp(x_ptr), b, xc, c_off, c, h, w, cs, vp(out_ptr),
                                             vp(stream)), "buctd_nchw32_to_act16")


class Plan:
    """Recorded launch sequence (buctd_plan_*)."""

    def __init__(self, ctx: Context):
        self.ctx = ctx
        h = vp()
        check(ctx.lib.buctd_plan_create(ctx.handle, C.byref(h)), "buctd_plan_create")
        self.handle = h
        self.labels = []
        self.lanes = []

    def add(self, kind: int, args: C.Structure, label: str = "", lane: int = 0, max_ctas: int = 0):
        """lane: concurrent lane of the op (cross-lane dependencies are derived by the library from the buffers
        the ops touch); max_ctas: SM share of a persistent tap-GEMM while other lanes run beside it (0 = all)."""
        check(self.ctx.lib.buctd_plan_add_lane(self.handle, kind, C.byref(args), C.sizeof(args), int(lane), int(max_ctas)),
              "buctd_plan_add_lane(%s)" % (label or kind))
        self.labels.append(label)
        self.lanes.append(int(lane))

    def instantiate(self):
        """capture the plan into a CUDA graph; run() then costs one graph launch"""
        check(self.ctx.lib.buctd_plan_instantiate(self.handle), "buctd_plan_instantiate")

    @property
    def num_lanes(self) -> int:
        return int(self.ctx.lib.buctd_plan_num_lanes(self.handle))

    def __len__(self):
        return int(self.ctx.lib.buctd_plan_size(self.handle))

    def run(self, stream: int, first: int = 0, last: int | None = None):
        if first == 0 and last is None:
            check(self.ctx.lib.buctd_plan_run(self.handle, vp(stream)), "buctd_plan_run")
            return
        last = len(self) if last is None else last
        check(self.ctx.lib.buctd_plan_run_range(self.handle, first, last, vp(stream)), "buctd_plan_run_range")

    def __del__(self):
        try:
            if self.handle:
                self.ctx.lib.buctd_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass
