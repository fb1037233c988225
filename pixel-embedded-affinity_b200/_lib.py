"""ctypes binding of csrc/libpea_b200.so (the C ABI declared in include/pea_b200.h).

There is no CPU path and no fallback: if the shared library is missing or fails to load, importing this
module raises immediately with the build command to run.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PEA_B200_LIB: development knob for A/B runs of alternative builds of the same ABI (never a CPU path)
LIB_PATH = os.environ.get("PEA_B200_LIB") or os.path.join(_HERE, "csrc", "libpea_b200.so")

c_float_p = ctypes.POINTER(ctypes.c_float)
c_int32_p = ctypes.POINTER(ctypes.c_int32)
vp = ctypes.c_void_p
ci = ctypes.c_int

MULTI_MAX_PROBLEMS, MULTI_MAX_OFFSETS = 8, 12


class LossProblem(ctypes.Structure):
    """pea_loss_problem of include/pea_b200.h"""
    _fields_ = [("emb", vp), ("partner", vp), ("target", vp), ("weight", vp), ("mask", vp), ("affs", vp), ("grad_emb", vp),
                ("loss_out", vp), ("offsets", c_int32_p), ("loss_norm", c_float_p), ("loss_coef", c_float_p),
                ("B", ctypes.c_int32), ("D", ctypes.c_int32), ("H", ctypes.c_int32), ("W", ctypes.c_int32),
                ("K", ctypes.c_int32), ("border", ctypes.c_int32), ("mask_dtype", ctypes.c_int32)]


# name -> (restype, argtypes); mirrors include/pea_b200.h one to one
SIGNATURES = {
    "pea_abi_version": (ci, []),
    "pea_strerror": (ctypes.c_char_p, [ci]),
    "pea_debug_configure": (ci, [ci, ci]),
    "pea_workspace_bytes": (ctypes.c_size_t, [ci, ci, ci, ci, ci, ci, ci]),
    "pea_affs_fwd": (ci, [vp, vp, ci, ci, ci, ci, ci, c_int32_p, ci, ci, ci, ci, ctypes.c_int64, ctypes.c_int64, vp, ci,
                          ctypes.c_int64, vp, vp, ctypes.c_size_t, vp]),
    "pea_zshard_stage_halo": (ci, [vp, ctypes.c_int64, vp, ci, ci, ci, ci, ci, ci, ci, vp, ctypes.c_uint32, vp]),
    "pea_zshard_affs_fwd": (ci, [vp, ci, ci, ci, ci, ci, ctypes.c_int64, ctypes.c_int64, vp, ci, ctypes.c_int64, vp, vp,
                                 ctypes.c_uint32, ci, ci, ci, vp, vp]),
    "pea_affs_loss_fwd_bwd": (ci, [vp, vp, ci, ci, ci, ci, ci, c_int32_p, ci, ci, ci, vp, vp, vp, ci, c_float_p,
                                   c_float_p, vp, vp, vp, vp, vp, ctypes.c_size_t, vp]),
    "pea_affs_bwd": (ci, [vp, vp, ci, ci, ci, ci, ci, c_int32_p, ci, ci, vp, vp, vp, vp, ctypes.c_size_t, vp]),
    "pea_scale_inplace": (ci, [vp, ctypes.c_int64, vp, vp]),
    "pea_consistency_flip": (ci, [vp, vp, ci, ci, ci, ci, ci, vp, vp]),
    "pea_blend_add": (ci, [vp, vp, ci, ci, ci, ci, vp, vp, ci, ci, ci, ci, ci, ci, vp]),
    "pea_blend_finish": (ci, [vp, vp, ci, ci, ci, ci, ci, ci, ci, ci, vp, vp]),
    "pea_affs_loss_multi_workspace": (ctypes.c_size_t, [ctypes.POINTER(LossProblem), ci]),
    "pea_affs_loss_multi": (ci, [ctypes.POINTER(LossProblem), ci, ci, vp, ctypes.c_size_t, vp, vp]),
    "pea_scale_multi": (ci, [ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(vp), ci, vp]),
    "pea_gt_affs2d": (ci, [vp, ci, ci, ci, ci, c_int32_p, ci, ci, ci, vp, vp, vp]),
    "pea_gt_affs3d": (ci, [vp, ci, ci, ci, ci, ci, c_int32_p, ci, ci, vp, vp]),
    "pea_gt_counts_bytes": (ctypes.c_size_t, [ci, ci]),
    "pea_gt_targets": (ci, [vp, ci, ci, ci, ci, ci, c_int32_p, ci, ci, ci, ci, ctypes.c_double, vp, vp, vp, vp,
                            ctypes.c_size_t, vp]),
    "pea_gt_pair_weights": (ci, [vp, ci, ci, ci, ci, ci, c_int32_p, ci, ci, ci, ci, ctypes.c_double, vp, vp, ctypes.c_size_t, vp]),
    "pea_affs_loss_labels_fwd_bwd": (ci, [vp, vp, ci, ci, ci, ci, ci, c_int32_p, ci, ci, ci, ci, vp, c_float_p, c_float_p, vp, vp,
                                          vp, vp, ctypes.c_size_t, vp]),
    "pea_replicate_first": (ci, [vp, ci, ci, ci, ci, ci, ctypes.c_int64, ci, vp]),
    "pea_wbr_stats_bytes": (ctypes.c_size_t, [ci]),
    "pea_weight_binary_ratio": (ci, [vp, ci, ctypes.c_int64, ctypes.c_double, vp, vp, ctypes.c_size_t, vp]),
}

ABI_VERSION = 7


class PeaError(RuntimeError):
    """A pea_* entry point returned a non-zero code (negative: rejected argument, positive: cudaError_t)."""

    def __init__(self, fn, code, msg):
        super().__init__(f"{fn} failed with code {code}: {msg}")
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found. The B200 kernels are not built; there is no CPU fallback. "
            f"Run `python -c 'import __graft_entry__ as g; g.build()'` at the repo root "
            f"(or `make -C {os.path.join(_HERE, 'csrc')}`).")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means the .so is stale: rebuild
        fn.restype = res
        fn.argtypes = args
    got = lib.pea_abi_version()
    if got != ABI_VERSION:
        raise ImportError(f"{LIB_PATH} has ABI version {got}, Python layer expects {ABI_VERSION}: rebuild")
    return lib


lib = _load()


def check(fn_name, code):
    if code != 0:
        raise PeaError(fn_name, code, lib.pea_strerror(code).decode())
