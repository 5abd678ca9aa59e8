"""ctypes binding of libmtbc.so (the C ABI declared in include/mtbc.h).

There is deliberately no fallback: if the shared library is missing or the device is not sm_100, every entry point
raises.  PyTorch is only used by callers for device memory and streams; nothing here touches torch.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libmtbc.so"

MAX_VIEWS = 8
MAX_SEGS = 48

c_i32, c_i64, c_f32, c_vp, c_f64 = C.c_int32, C.c_int64, C.c_float, C.c_void_p, C.c_double


class ActView(C.Structure):
    _fields_ = [("ptr", c_vp), ("C", c_i32), ("W", c_i32), ("H", c_i32), ("N", c_i32),
                ("sW", c_i64), ("sH", c_i64), ("sN", c_i64)]


class GemmSeg(C.Structure):
    _fields_ = [("view", c_i32), ("dh", c_i32), ("dw", c_i32), ("wk0", c_i32), ("wtap", c_i32)]


class OutSlice(C.Structure):
    _fields_ = [("ptr", c_vp), ("out_C", c_i32), ("col0", c_i32), ("ncols", c_i32), ("accumulate", c_i32)]


class ConvGemmDesc(C.Structure):
    _fields_ = [("nviews", c_i32), ("views", ActView * MAX_VIEWS), ("nseg", c_i32), ("seg", GemmSeg * MAX_SEGS),
                ("wpack", c_vp), ("w_ntaps", c_i32), ("w_ktot", c_i32), ("ncols", c_i32),
                ("W", c_i32), ("H", c_i32), ("N", c_i32), ("epi_mode", c_i32), ("out", c_vp), ("out_C", c_i32),
                ("up_k", c_i32), ("up_cp", c_i32), ("bias", c_vp), ("stat_sum", c_vp), ("stat_sq", c_vp),
                ("stat_C", c_i32), ("accumulate", c_i32), ("nouts", c_i32), ("outs", OutSlice * MAX_VIEWS),
                ("dtype", c_i32), ("wpack_lo", c_vp),
                ("bwd_y", c_vp), ("bwd_mean", c_vp), ("bwd_rstd", c_vp), ("bwd_gamma", c_vp), ("bwd_beta", c_vp),
                ("bwd_slope", c_f32), ("stat_fold", c_i32)]


class WgradTap(C.Structure):
    _fields_ = [("a_view", c_i32), ("a_dh", c_i32), ("a_dw", c_i32), ("b_view", c_i32)]


class WgradDesc(C.Structure):
    _fields_ = [("a_nviews", c_i32), ("a_views", ActView * 4), ("b_nviews", c_i32), ("b_views", ActView * 4),
                ("ntaps", c_i32), ("taps", WgradTap * 9), ("W", c_i32), ("H", c_i32), ("N", c_i32),
                ("dw_acc", c_vp), ("n_rows", c_i32), ("ld_k", c_i32), ("k0", c_i32), ("splits", c_i32),
                ("dtype", c_i32)]


class WgradMultiDesc(C.Structure):
    _fields_ = [("nsrc", c_i32), ("x", ActView * 8), ("k0", c_i32 * 8), ("dy", ActView), ("W", c_i32), ("H", c_i32),
                ("N", c_i32), ("dw_acc", c_vp), ("n_rows", c_i32), ("ld_k", c_i32), ("splits", c_i32)]


class ConvTBwdDesc(C.Structure):
    _fields_ = [("x", ActView), ("dy", ActView * 4), ("wd", c_vp), ("wd_rows", c_i32), ("wd_ld", c_i32),
                ("dw_acc", c_vp), ("n_rows", c_i32), ("ld_k", c_i32), ("dbias", c_vp), ("dx", c_vp), ("dx_C", c_i32),
                ("accumulate", c_i32), ("Cout", c_i32)]


class ParamJob(C.Structure):
    _fields_ = [("kind", c_i32), ("i", c_i32 * 11), ("src", c_vp), ("dst0", c_vp), ("dst1", c_vp)]


JOB_COPY_F32, JOB_PACK_CONV, JOB_PACK_CONVT, JOB_UNPACK_CONV, JOB_UNPACK_CONVT, JOB_PACK_CONV_PAIR = range(6)


# name -> argtypes (return type is int unless listed in _RESTYPES)
_SIGS = {
    "mtbc_abi_version": [],
    "mtbc_device_check": [],
    "mtbc_set_mode": [c_i32],
    "mtbc_get_mode": [],
    "mtbc_query_workspace_bytes": [C.c_char_p, c_i32, c_i64, c_i32],
    "mtbc_in_stats_det": [c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp],
    "mtbc_conv_gemm_create": [C.POINTER(ConvGemmDesc), C.POINTER(c_vp)],
    "mtbc_wgrad_create": [C.POINTER(WgradDesc), C.POINTER(c_vp)],
    "mtbc_wgrad_multi_create": [C.POINTER(WgradMultiDesc), C.POINTER(c_vp)],
    "mtbc_convT_bwd_create": [C.POINTER(ConvTBwdDesc), C.POINTER(c_vp)],
    "mtbc_param_jobs_create": [C.POINTER(ParamJob), c_i32, C.POINTER(c_vp)],
    "mtbc_op_launch": [c_vp, c_vp],
    "mtbc_ops_launch": [C.POINTER(c_vp), c_i32, c_vp],
    "mtbc_op_destroy": [c_vp],
    "mtbc_op_flops": [c_vp],
    "mtbc_pack_conv_weight": [c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_i32, c_i32, c_i32, c_vp, c_i32, c_i32,
                              c_vp],
    "mtbc_pack_convT_weight": [c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_i32, c_vp, c_i32, c_i32, c_vp],
    "mtbc_unpack_conv_wgrad": [c_vp, c_i32, c_i32, c_i32, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp],
    "mtbc_unpack_convT_wgrad": [c_vp, c_i32, c_i32, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp],
    "mtbc_conv_first_fwd": [c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_i32, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp],
    "mtbc_conv_first_wgrad": [c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_i32, c_i32, c_vp, c_vp],
    "mtbc_add": [c_vp, c_vp, c_vp, c_i64, c_vp],
    "mtbc_accumulate": [c_vp, c_vp, c_i64, c_i32, c_vp],
    "mtbc_dropout_fwd": [c_vp, c_vp, c_vp, c_i64, c_f32, C.c_uint64, c_vp, c_i32, c_i32, c_vp],
    "mtbc_dropout_bwd": [c_vp, c_vp, c_vp, c_i64, c_f32, c_i32, c_vp],
    "mtbc_zero_stuff2": [c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp],
    "mtbc_bn_pool_fwd": [c_vp, c_vp, c_i32, c_i32, c_i32, c_i64, c_i32, c_f32, c_vp, c_vp, c_vp, c_vp],
    "mtbc_bn_pool_bwd": [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp],
    "mtbc_in_stats": [c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp],
    "mtbc_in_apply": [c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_i32, c_f32, c_f32, c_vp, c_vp,
                      c_vp, c_vp, c_vp],
    "mtbc_in_bwd_reduce": [c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_f32, c_vp, c_vp, c_vp],
    "mtbc_in_bwd_apply": [c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_f32, c_vp, c_vp, c_vp, c_vp,
                          c_vp, c_i32, c_vp],
    "mtbc_in_bwd": [c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32,
                    c_vp, c_vp],
    "mtbc_maxpool2_bwd": [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_i32, c_vp],
    "mtbc_upsample2_fwd": [c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp],
    "mtbc_upsample2_bwd": [c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_i32, c_vp],
    "mtbc_channel_sum": [c_vp, c_i64, c_i32, c_i32, c_vp, c_i32, c_vp],
    "mtbc_head1x1_fwd": [c_vp, c_i64, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp],
    "mtbc_head1x1_bwd": [c_vp, c_vp, c_i64, c_i32, c_i32, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp],
    "mtbc_dshead_compose": [c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_vp],
    "mtbc_dshead_fwd": [c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp],
    "mtbc_dshead_bwd_parts": [c_i32, c_i32, c_i32],
    "mtbc_dshead_bwd": [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_i32, c_vp, c_vp, c_i32, c_vp],
    "mtbc_dshead_decompose": [c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp],
    "mtbc_gap_fc_fwd": [c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_i32, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp,
                        c_vp],
    "mtbc_gap_fc_bwd": [c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_i32, c_vp, c_i32, c_vp, c_vp, c_vp, c_i32, c_vp,
                        c_vp, c_vp, c_vp, c_vp, c_vp],
    "mtbc_flat_fc_fwd": [c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_i32, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp],
    "mtbc_flat_fc_bwd": [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_i32, c_vp, c_i32, c_vp, c_vp, c_i32, c_vp,
                         c_vp, c_vp, c_vp, c_vp, c_vp],
    "mtbc_softmax_rows_fwd": [c_vp, c_i32, c_i32, c_vp, c_vp],
    "mtbc_softmax_rows_bwd": [c_vp, c_vp, c_i32, c_i32, c_vp, c_vp],
    "mtbc_dice_sums": [c_vp, c_vp, c_i32, c_i64, c_vp, c_vp],
    "mtbc_dice_finalize": [c_vp, c_i32, c_vp, c_vp],
    "mtbc_dice_bwd": [c_vp, c_vp, c_i32, c_i64, c_vp, c_vp, c_f32, c_vp, c_vp],
    "mtbc_focal_fwd": [c_vp, c_vp, c_i32, c_i32, c_f32, c_f32, c_vp, c_vp],
    "mtbc_focal_bwd": [c_vp, c_vp, c_i32, c_i32, c_f32, c_f32, c_vp, c_f32, c_vp, c_vp],
    "mtbc_multitask_loss": [c_vp, c_i32, c_i32, c_vp, c_f32, c_vp, c_vp],
    "mtbc_refine_predictions": [c_vp, c_vp, c_i32, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp],
    "mtbc_hard_dice_counts": [c_vp, c_vp, c_i64, c_vp, c_vp],
    "mtbc_confusion_counts": [c_vp, c_vp, c_i32, c_i64, c_vp, c_vp],
    "mtbc_row_hausdorff": [c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp],
    "mtbc_metrics_accumulate": [c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_vp],
    "mtbc_augment_batch": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp],
    "mtbc_adam_step": [c_vp, c_vp, c_vp, c_vp, c_i64, c_f32, c_f32, c_f32, c_f32, c_f32, c_i32, c_vp],
    "mtbc_adam_step_dev": [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_f32, c_f32, c_f32, c_f32, c_vp, c_vp],
    "mtbc_increment_i32": [c_vp, c_vp],
    "mtbc_fill_f32": [c_vp, c_i64, c_f32, c_vp],
    "mtbc_zero_bytes": [c_vp, c_i64, c_vp],
    "mtbc_copy_f32": [c_vp, c_vp, c_i64, c_vp],
    "mtbc_f32_to_bf16_nhwc": [c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_i32, c_vp],
    "mtbc_bf16_nhwc_to_f32": [c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp],
}
_RESTYPES = {"mtbc_op_destroy": None, "mtbc_op_flops": c_f64, "mtbc_query_workspace_bytes": c_i64}
MODE_ACT_FP32, MODE_DETERMINISTIC = 1, 2

EXPORTED_SYMBOLS = sorted(list(_SIGS) + ["mtbc_last_error", "mtbc_build_digest"])
ABI_VERSION = 4


class MtbcError(RuntimeError):
    pass


_lib = None


def load():
    """Load libmtbc.so (raises if it has not been built -- there is no CPU or library fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _b
    if _b.lib_digest() != _b._digest():
        # stale or missing binary: rebuild from the sources next to it when a compiler is there, else fail loudly
        try:
            _b.build()
        except Exception as e:  # noqa: BLE001 - reported as the loader's error below
            if LIB_PATH.exists():
                raise MtbcError(f"{LIB_PATH} was built from other sources and could not be rebuilt: {e}") from e
    if not LIB_PATH.exists():
        raise MtbcError(f"{LIB_PATH} not found: build it with `python -m multi_task_breast_cancer_b200.build` "
                        "(or __graft_entry__.build()); there is no fallback path")
    lib = C.CDLL(str(LIB_PATH))
    _verify(lib)
    for name, args in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = _RESTYPES.get(name, c_i32)
    lib.mtbc_last_error.argtypes = []
    lib.mtbc_last_error.restype = C.c_char_p
    _lib = lib
    return lib


def _verify(lib):
    """Refuse a stale binary: the digest compiled into the library must equal the digest of the sources on disk, and
    the ABI version must be the one these ctypes signatures were written for.  (A stamp file next to the .so could not
    tell: the .so is git-ignored, so a checkout changes the sources but not the binary.)"""
    try:
        fn = lib.mtbc_build_digest
    except AttributeError:
        raise MtbcError(f"{LIB_PATH} predates mtbc_build_digest: rebuild it (python -m multi_task_breast_cancer_b200.build)")
    fn.restype, fn.argtypes = C.c_char_p, []
    have = fn().decode()
    lib.mtbc_abi_version.restype, lib.mtbc_abi_version.argtypes = c_i32, []
    if lib.mtbc_abi_version() != ABI_VERSION:
        raise MtbcError(f"{LIB_PATH}: ABI version {lib.mtbc_abi_version()}, this binding needs {ABI_VERSION}: rebuild")
    from . import build as _b
    if (_b.CSRC / "api.cu").exists():          # sources present (always, in-tree): compare
        want = _b._digest()
        if have != want:
            raise MtbcError(f"{LIB_PATH} was built from other sources (digest {have[:12]}.. != {want[:12]}..): "
                            "rebuild it with `python -m multi_task_breast_cancer_b200.build`")


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().mtbc_last_error().decode(errors="replace")
        raise MtbcError(f"{what or 'mtbc call'} failed (status {rc}): {msg}")


def call(name: str, *args):
    """Call an int-returning entry point and raise MtbcError on a non-zero status."""
    lib = load()
    check(getattr(lib, name)(*args), name)
