"""ctypes binding of include/wavernn_b200.h.  Fails loudly: there is no fallback path."""
import ctypes
import os

from . import build as _build

i32, i64, u64, vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64, ctypes.c_void_p

MODE = {"RAW": 0, "MOL": 1}
PRECISION = {"fp32": 0, "bf16": 1, "bf16-dense": 2}
ABI_VERSION = 3


class Config(ctypes.Structure):
    _fields_ = [(n, i32) for n in ("rnn_dims", "fc_dims", "feat_dims", "aux_dims", "n_classes", "mode", "precision")]


class Weights(ctypes.Structure):
    FIELDS = ("I_w", "I_b", "r1_wih", "r1_whh", "r1_bih", "r1_bhh", "r2_wih", "r2_whh", "r2_bih", "r2_bhh",
              "fc1_w", "fc1_b", "fc2_w", "fc2_b", "fc3_w", "fc3_b")
    KEYS = ("I.weight", "I.bias",
            "rnn1.weight_ih_l0", "rnn1.weight_hh_l0", "rnn1.bias_ih_l0", "rnn1.bias_hh_l0",
            "rnn2.weight_ih_l0", "rnn2.weight_hh_l0", "rnn2.bias_ih_l0", "rnn2.bias_hh_l0",
            "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias", "fc3.weight", "fc3.bias")
    _fields_ = [(n, vp) for n in FIELDS]


class Info(ctypes.Structure):
    _fields_ = [("ctas", i32), ("threads", i32), ("smem_bytes", i32), ("folds_per_group", i32),
                ("max_folds_per_launch", i32), ("exchanges_per_step", i32), ("sm_count", i32),
                ("launches", i64), ("epilogue_launches", i64), ("last_kernel_status", i32),
                ("last_kernel_ms", ctypes.c_float), ("kernel_kind", i32)]


# every symbol include/wavernn_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "wrnn_abi_version": (i32, []),
    "wrnn_last_error": (ctypes.c_char_p, []),
    "wrnn_create": (i32, [ctypes.POINTER(Config), i32, ctypes.POINTER(vp)]),
    "wrnn_destroy": (None, [vp]),
    "wrnn_load_weights": (i32, [vp, ctypes.POINTER(Weights)]),
    "wrnn_packed_floats": (i64, [ctypes.POINTER(Config)]),
    "wrnn_pack_weights_host": (i32, [ctypes.POINTER(Config), ctypes.POINTER(Weights), vp, i64]),
    "wrnn_wide_packed_floats": (i64, [ctypes.POINTER(Config), ctypes.POINTER(i64)]),
    "wrnn_wide_pack_host": (i32, [ctypes.POINTER(Config), ctypes.POINTER(Weights), vp, i64]),
    "wrnn_dense_layout": (i32, [ctypes.POINTER(Config), ctypes.POINTER(i64)]),
    "wrnn_dense_pack_host": (i32, [ctypes.POINTER(Config), ctypes.POINTER(Weights), vp, vp, vp]),
    "wrnn_fold_index": (i32, [i64, i64, i64, ctypes.POINTER(i64), ctypes.POINTER(i64)]),
    "wrnn_generate_folds": (i32, [vp, vp, vp, i64, vp, vp, i32, i32, vp, u64, vp, vp, vp, vp, vp]),
    "wrnn_generate_folds_frames": (i32, [vp, vp, i64, vp, i64, vp, i32, i32, vp, i32, i32, vp, u64, vp, vp, vp, vp, vp]),
    "wrnn_xfade_unfold": (i32, [vp, i32, i32, i32, i32, i32, i64, i32, vp, vp]),
    "wrnn_xfade_unfold_segment": (i32, [vp, i32, i32, i32, i32, i64, i32, i64, i64, i64, i64, vp, vp]),
    "wrnn_synchronize": (i32, [vp]),
    "wrnn_query": (i32, [vp, ctypes.POINTER(i32), ctypes.POINTER(i32)]),
    "wrnn_get_info": (i32, [vp, ctypes.POINTER(Info)]),
    "wrnn_set_profiling": (i32, [vp, i32]),
    "wrnn_get_stage_cycles": (i32, [vp, vp, i32]),
    "wrnn_measure_exchange": (i32, [vp, i32, ctypes.POINTER(ctypes.c_float)]),
    "wrnn_set_kernel": (i32, [vp, i32]),
    "wrnn_cond_blob_floats": (i64, [i32]),
    "wrnn_cond_create": (i32, [i32, vp, i64, i32, ctypes.POINTER(vp)]),
    "wrnn_cond_destroy": (None, [vp]),
    "wrnn_cond_launches": (i64, [vp]),
    "wrnn_cond_frames": (i32, [vp, vp, vp, i32, vp, vp]),
}

_LIB = None


class WaveRNNLibraryError(RuntimeError):
    pass


def lib():
    """Load (building first if the sources are newer) libwavernn_b200.so; raises if impossible."""
    global _LIB
    if _LIB is None:
        path = os.environ.get("WRNN_LIB") or _build.LIB          # WRNN_LIB: development knob for A/B runs of two builds on one box
        if path == _build.LIB and _build.needs_build():
            try:
                _build.build()
            except Exception as e:  # no nvcc on the box and no prebuilt .so
                if not os.path.exists(path):
                    raise WaveRNNLibraryError("libwavernn_b200.so is missing and could not be built: %s" % e)
        L = ctypes.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)          # AttributeError if the .so lacks a declared symbol
            fn.restype, fn.argtypes = res, args
        if L.wrnn_abi_version() != ABI_VERSION:
            raise WaveRNNLibraryError("ABI version mismatch: library %d, binding %d" % (L.wrnn_abi_version(), ABI_VERSION))
        _LIB = L
    return _LIB


def check(rc):
    if rc != 0:
        raise RuntimeError("wavernn_b200: %s (status %d)" % (lib().wrnn_last_error().decode(), rc))


def fold_index(total_len, target, overlap):
    n, pl = i64(), i64()
    check(lib().wrnn_fold_index(total_len, target, overlap, ctypes.byref(n), ctypes.byref(pl)))
    return int(n.value), int(pl.value)
