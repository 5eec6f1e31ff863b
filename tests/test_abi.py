"""not gpu: the C-ABI library loads and exports every symbol include/buctd_b200.h declares; the ctypes
structs match the header field for field (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import pytest

from buctd_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "buctd_b200.h")


def _header():
    with open(HEADER) as f:
        return f.read()


def test_header_symbols_match_binding_list():
    declared = set(re.findall(r"^(?:int|long long|const char\*)\s+(buctd_\w+)\s*\(", _header(), re.M))
    assert declared == set(L.EXPORTS), declared ^ set(L.EXPORTS)


@pytest.mark.skipif(not os.path.exists(L.LIB_PATH), reason="library not built (run __graft_entry__.build())")
def test_library_exports_every_symbol():
    lib = ctypes.CDLL(L.LIB_PATH)
    for name in L.EXPORTS:
        assert hasattr(lib, name), name
    assert lib.buctd_abi_version() == L.ABI_VERSION


def test_struct_fields_match_header():
    h = _header()
    pairs = {"buctd_tapgemm_args": L.TapGemmArgs, "buctd_stem_args": L.StemArgs, "buctd_fuse_args": L.FuseArgs,
             "buctd_resize_args": L.ResizeArgs, "buctd_chan_att_args": L.ChanAttArgs, "buctd_pos_att_args": L.PosAttArgs,
             "buctd_da_combine_args": L.DaCombineArgs, "buctd_prenet_args": L.PrenetArgs, "buctd_decode_args": L.DecodeArgs,
             "buctd_raster_args": L.RasterArgs, "buctd_flip_args": L.FlipArgs, "buctd_flip_merge_args": L.FlipMergeArgs,
             "buctd_joints_mse_args": L.JointsMseArgs, "buctd_adam_args": L.AdamArgs, "buctd_crop_warp_args": L.CropWarpArgs,
             "buctd_gen_target_args": L.GenTargetArgs, "buctd_pck_args": L.PckArgs, "buctd_cond_kpts_args": L.CondKptsArgs,
             "buctd_resize3_args": L.Resize3Args, "buctd_head_decode_args": L.HeadDecodeArgs,
             # the training step
             "buctd_gemm16_args": L.Gemm16Args, "buctd_wgrad_reduce_args": L.WgradReduceArgs, "buctd_bn_stats_args": L.BnStatsArgs,
             "buctd_bn_apply_args": L.BnApplyArgs, "buctd_bn_bwd_args": L.BnBwdArgs, "buctd_colsum_args": L.ColsumArgs,
             "buctd_zero_stuff_args": L.ZeroStuffArgs, "buctd_fuse_bwd_args": L.FuseBwdArgs, "buctd_add16_args": L.Add16Args,
             "buctd_relayout_args": L.RelayoutArgs, "buctd_chan_softmax_bwd_args": L.ChanSoftmaxBwdArgs,
             "buctd_pos_softmax_bwd_args": L.PosSoftmaxBwdArgs, "buctd_pos_bwd_scores_args": L.PosBwdScoresArgs, "buctd_rowdot_args": L.RowdotArgs,
             "buctd_stem_wgrad_args": L.StemWgradArgs, "buctd_repack_args": L.RepackArgs}
    for cname, st in pairs.items():
        body = re.search(r"typedef struct \{([^}]*)\}\s*%s;" % cname, h).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            # "int32_t a, b, c" / "const void* term[4]" / "const float* x"
            decl = re.sub(r"^(const\s+)?(int32_t|int64_t|uint32_t|uint8_t|float|void|double)\s*\**(\s*const\s*\*)?", "", decl)
            for part in decl.split(","):
                names.append(re.sub(r"\[.*\]", "", part).replace("*", "").strip())
        mine = [f[0] if f[0] != "in_" else "in" for f in st._fields_]       # `in` is a Python keyword
        assert names == mine, (cname, names, mine)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(L.BuctdError, match="no CPU fallback"):
        L.load()
