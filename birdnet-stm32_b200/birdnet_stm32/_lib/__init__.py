"""ctypes binding of `libbn_b200.so` (the C ABI declared in `include/bn_engine.h`).

The library is built in-tree by `__graft_entry__.build()` (or `make -C csrc`).
There is no fallback: if the shared object is missing, or no CUDA device is
present when an engine is created, the call raises.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbn_b200.so")

# every symbol `include/bn_engine.h` declares
EXPORTS = (
    "bn_create", "bn_destroy", "bn_query", "bn_set_option", "bn_infer_spec_f32", "bn_frontend_pcm16",
    "bn_infer_pcm16", "bn_infer_pool", "bn_pool_scores", "bn_dump_tensor", "bn_launch_count",
    "bn_profile_read", "bn_host_alloc", "bn_host_free", "bn_last_error", "bn_version",
    "bn_infer_wave_f32", "bn_infer_pool_wave_f32", "bn_frontend_wave_f32",
    # include/bn_features.h
    "bn_features_create", "bn_features_destroy", "bn_features_rows", "bn_features_pcm16",
    # include/bn_ingest.h
    "bn_ingest_create", "bn_ingest_destroy", "bn_ingest_out_len", "bn_ingest_num_chunks", "bn_ingest_filter",
    "bn_ingest_window", "bn_ingest_chunks", "bn_ingest_launch_count",
    # include/bn_metrics.h
    "bn_metrics_compute", "bn_metrics_bootstrap_ap",
    # include/bn_reader.h
    "bn_wav_probe", "bn_read_pcm16_batch", "bn_read_raw_batch",
)

BN_SAMPLE_FORMAT = {"s16": 0, "s24": 1, "s32": 2, "f32": 3, "u8": 4}

BN_POOL = {"avg": 0, "mean": 0, "average": 0, "max": 1, "lme": 2, "log_mean_exp": 2, "log_mean_exponential": 2}
BN_OPT_ROUNDING, BN_OPT_MEAN_VARIANT, BN_OPT_FORCE_GENERIC, BN_OPT_WAVE, BN_OPT_PROFILE, BN_OPT_TENSOR_CORE = 1, 2, 3, 4, 5, 6
BN_OPT_FUSION = 7
BN_OPT_HOST_WAVE = 8


class BnInfo(C.Structure):
    _fields_ = [
        ("frontend_kind", C.c_int32), ("sample_rate", C.c_int32), ("chunk_len", C.c_int32), ("n_fft", C.c_int32),
        ("hop", C.c_int32), ("spec_width", C.c_int32), ("fft_bins", C.c_int32), ("num_classes", C.c_int32),
        ("input_elems", C.c_int64), ("n_ops", C.c_int32), ("n_tensors", C.c_int32), ("device", C.c_int32),
        ("wave", C.c_int32), ("workspace_bytes", C.c_int64), ("fast_path", C.c_int32), ("reserved", C.c_int32 * 7),
    ]


class BnFeatParams(C.Structure):
    """`bn_feat_params` of include/bn_features.h"""

    _fields_ = [
        ("sample_rate", C.c_int32), ("chunk_len", C.c_int32), ("n_fft", C.c_int32), ("spec_width", C.c_int32),
        ("n_mels", C.c_int32), ("mode", C.c_int32), ("mag_scale", C.c_int32), ("n_mfcc", C.c_int32),
        ("pcen_b", C.c_float), ("reserved", C.c_int32 * 7),
    ]


class BnMetricsResult(C.Structure):
    """`bn_metrics_result` of include/bn_metrics.h"""

    _fields_ = [
        ("roc_auc_micro", C.c_double), ("map_micro", C.c_double), ("cmap", C.c_double), ("precision", C.c_double),
        ("recall", C.c_double), ("f1", C.c_double), ("n_positive", C.c_int64), ("n_cells", C.c_int64),
        ("classes_without_positives", C.c_int32), ("n_launches", C.c_int32), ("reserved", C.c_int32 * 4),
    ]


class BnReaderFile(C.Structure):
    """`bn_reader_file` of include/bn_reader.h"""

    _fields_ = [
        ("status", C.c_int32), ("n_chunks", C.c_int32), ("sample_rate", C.c_int32), ("channels", C.c_int32), ("fmt", C.c_int32),
        ("peak", C.c_float), ("n_frames", C.c_int64), ("data_offset", C.c_int64), ("container", C.c_int32), ("reserved", C.c_int32),
    ]


BN_FEAT_MODE = {"mel": 0, "log_mel": 1, "mfcc": 2}
BN_MAG_SCALE = {"none": 0, "pwl": 1, "pcen": 2, "db": 3}


class EngineError(RuntimeError):
    """A `bn_*` call returned a negative status."""

    def __init__(self, code: int, message: str):
        super().__init__(f"[bn status {code}] {message}")
        self.code = code


_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()' "
            "or make -C birdnet-stm32_b200/csrc).  There is no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    vp, i32, f32 = C.c_void_p, C.c_int, C.c_float
    L.bn_create.argtypes = [vp, C.c_size_t, i32, C.POINTER(vp)]
    L.bn_destroy.argtypes = [vp]
    L.bn_destroy.restype = None
    L.bn_query.argtypes = [vp, C.POINTER(BnInfo)]
    L.bn_set_option.argtypes = [vp, i32, i32]
    L.bn_infer_spec_f32.argtypes = [vp, vp, i32, vp, vp]
    L.bn_frontend_pcm16.argtypes = [vp, vp, vp, i32, vp, vp]
    L.bn_infer_pcm16.argtypes = [vp, vp, vp, i32, vp, vp]
    L.bn_infer_pool.argtypes = [vp, vp, vp, vp, i32, i32, f32, vp, vp]
    L.bn_pool_scores.argtypes = [vp, vp, vp, i32, i32, i32, f32, vp, vp]
    L.bn_dump_tensor.argtypes = [vp, i32, vp, C.c_size_t]
    L.bn_launch_count.argtypes = [vp]
    L.bn_launch_count.restype = C.c_int64
    L.bn_profile_read.argtypes = [vp, i32, C.c_char_p, C.c_size_t, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    L.bn_host_alloc.argtypes = [C.c_size_t]
    L.bn_host_alloc.restype = vp
    L.bn_host_free.argtypes = [vp]
    L.bn_host_free.restype = None
    L.bn_last_error.restype = C.c_char_p
    L.bn_version.restype = C.c_char_p
    L.bn_features_create.argtypes = [C.POINTER(BnFeatParams), vp, vp, i32, C.POINTER(vp)]
    L.bn_features_destroy.argtypes = [vp]
    L.bn_features_destroy.restype = None
    L.bn_features_rows.argtypes = [vp]
    L.bn_features_pcm16.argtypes = [vp, vp, vp, i32, vp, vp]
    L.bn_infer_wave_f32.argtypes = [vp, vp, vp, i32, vp, vp]
    L.bn_infer_pool_wave_f32.argtypes = [vp, vp, vp, vp, i32, i32, f32, vp, vp]
    L.bn_frontend_wave_f32.argtypes = [vp, vp, vp, i32, vp, vp]
    i64 = C.c_int64
    L.bn_ingest_create.argtypes = [i32, C.POINTER(vp)]
    L.bn_ingest_destroy.argtypes = [vp]
    L.bn_ingest_destroy.restype = None
    L.bn_ingest_out_len.argtypes = [i64, i32, i32]
    L.bn_ingest_out_len.restype = i64
    L.bn_ingest_num_chunks.argtypes = [i64, i32, i32]
    L.bn_ingest_filter.argtypes = [i32, i32, vp, i32, C.POINTER(i32), C.POINTER(i32)]
    L.bn_ingest_window.argtypes = [vp, vp, i32, i64, i32, i32, i32, i32, vp, vp, vp]
    L.bn_ingest_chunks.argtypes = [vp, vp, i32, i64, i32, i32, i32, i32, i32, vp, i32, C.POINTER(i32), vp, vp]
    L.bn_ingest_launch_count.argtypes = [vp]
    L.bn_ingest_launch_count.restype = i64
    L.bn_metrics_compute.argtypes = [vp, vp, i32, i32, i32, C.POINTER(BnMetricsResult), vp]
    L.bn_metrics_bootstrap_ap.argtypes = [vp, vp, i32, i32, i32, C.c_uint64, vp, vp, i32]
    L.bn_wav_probe.argtypes = [C.c_char_p, C.c_double, C.POINTER(BnReaderFile)]
    L.bn_read_raw_batch.argtypes = [C.POINTER(C.c_char_p), i32, C.c_double, vp, C.c_int64, i32, C.POINTER(BnReaderFile), C.POINTER(C.c_int64)]
    L.bn_read_pcm16_batch.argtypes = [C.POINTER(C.c_char_p), i32, i32, i32, i32, C.c_double, vp, i32, i32, C.POINTER(BnReaderFile),
                                      C.POINTER(i32)]
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        raise EngineError(rc, load().bn_last_error().decode("utf-8", "replace"))
