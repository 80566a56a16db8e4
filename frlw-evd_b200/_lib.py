"""ctypes binding of ``libevrep.so`` (the C ABI declared in ``include/evrep.h``).

There is no CPU fallback: if the shared library is missing, or a call fails, an
exception is raised.  The library is built in-tree by ``frlw_evd_b200.build``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libevrep.so")
HEADER_PATH = os.path.join(os.path.dirname(HERE), "include", "evrep.h")

P = c_void_p


class TafWindow(ctypes.Structure):
    """``evrep_taf_window`` (include/evrep.h)."""
    _fields_ = [("ev_begin", c_int64), ("ev_end", c_int64), ("start_time", c_int64),
                ("n_bins", c_int32), ("fresh", c_int32)]


class EvWindow(ctypes.Structure):
    """``evrep_ev_window`` (include/evrep.h)."""
    _fields_ = [("ev_begin", c_int64), ("ev_end", c_int64), ("t0", c_int64)]


class EvSegment(ctypes.Structure):
    """``evrep_ev_segment`` (include/evrep.h)."""
    _fields_ = [("ev_begin", c_int64), ("ev_end", c_int64), ("start_time", c_int64)]


class EvSpan(ctypes.Structure):
    """``evrep_ev_span`` (include/evrep.h)."""
    _fields_ = [("first_segment", c_int32), ("last_segment", c_int32), ("t0", c_int64), ("tw", c_int64)]


class CountSegment(ctypes.Structure):
    """``evrep_count_segment`` (include/evrep.h)."""
    _fields_ = [("ev_begin", c_int64), ("ev_end", c_int64)]


class CountEmit(ctypes.Structure):
    """``evrep_count_emit`` (include/evrep.h)."""
    _fields_ = [("first_segment", c_int32), ("last_segment", c_int32)]


class SaeWindow(ctypes.Structure):
    """``evrep_sae_window`` (include/evrep.h)."""
    _fields_ = [("ev_begin", c_int64), ("ev_end", c_int64), ("now", c_int64), ("t_first", c_int64), ("t_last", c_int64)]


# name -> (restype, argtypes); must list every symbol include/evrep.h declares
SIGNATURES = {
    "evrep_version": (c_int, []),
    "evrep_strerror": (c_char_p, [c_int]),
    "evrep_last_cuda_error": (c_char_p, []),
    "evrep_device_info": (c_int, [P, P, P]),
    "evrep_decode_dat": (c_int, [P, c_int64, P, P, P, P, P]),
    "evrep_soa_to_aos64": (c_int, [P, P, P, P, c_int64, P, P]),
    "evrep_count_accumulate": (c_int, [P, P, P, c_int64, c_int, c_int, P, P, P, P]),
    "evrep_count_accumulate_aos64": (c_int, [P, c_int64, c_int, c_int, c_int, P, P]),
    "evrep_count_finalize": (c_int, [P, c_int, c_int, P, c_int, P]),
    "evrep_count_image": (c_int, [P, P, P, c_int64, c_int, c_int, P, P, P, P, P]),
    "evrep_count_images_u8": (c_int, [P, P, P, c_int64, P, c_int, c_int, c_int, P, P, c_int, c_int, P, P, P, P, P]),
    "evrep_sae_u8": (c_int, [P, P, P, P, c_int64, c_int, c_int, P, P, c_int, c_int, P, P, c_float, c_float, P, c_int,
                             P, P, P, P, P]),
    "evrep_sae": (c_int, [P, P, P, P, c_int64, c_int, c_int, P, P, c_float, c_float, P, c_int, P, P, P, P, P]),
    "evrep_sae_aos64": (c_int, [P, c_int64, c_int, c_int, c_int, c_float, c_float, P, c_int, P, P, P, P, P]),
    "evrep_event_volume": (c_int, [P, P, P, P, c_int64, c_int64, c_int64, c_int, c_int, c_int, P, P, P, P]),
    "evrep_event_volume_aos64": (c_int, [P, c_int64, c_int, c_int, c_int, c_int, P, P]),
    "evrep_taf_bin_scratch_bytes": (c_int64, [c_int, c_int]),
    "evrep_taf_bin": (c_int, [P, P, P, P, c_int64, c_int64, c_double, c_int, c_int, c_int, P, P, P, P, P, P, P]),
    "evrep_taf_bin_aos64": (c_int, [P, c_int64, c_int, c_int, c_int, c_int, P, P, P, P, P]),
    "evrep_taf_stream_scratch_bytes": (c_int64, [c_int64, c_int, c_int64, c_int, c_int]),
    "evrep_taf_stream": (c_int, [P, P, P, P, c_int64, P, c_int, c_int, c_int, c_int, c_int, P, P, c_int, c_int, P, c_int,
                                 P, c_int64, P, c_int64, P, P, P]),
    "evrep_taf_stream_ordered_scratch_bytes": (c_int64, [c_int64, c_int, c_int64, c_int, c_int]),
    "evrep_taf_stream_ordered_status_offset": (c_int64, [c_int64, c_int, c_int64, c_int, c_int]),
    "evrep_taf_stream_ordered": (c_int, [P, P, P, P, c_int64, P, c_int, c_int, c_int, c_int, c_int, P, P, c_int, c_int, P, c_int,
                                         P, c_int64, P, c_int64, P, P, P]),
    "evrep_taf_stream_sliced_scratch_bytes": (c_int64, [c_int64, c_int, c_int64, c_int, c_int, c_int]),
    "evrep_taf_stream_sliced": (c_int, [P, P, P, P, c_int64, P, c_int, c_int, c_int, c_int, c_int, P, P, c_int, c_int, P,
                                         c_int, P, c_int64, P, c_int64, P, c_int64, P, P, P]),
    "evrep_stream_order_violations": (c_int, [P, P, P]),
    "evrep_events_order_check": (c_int, [P, c_int64, P, P]),
    "evrep_event_volume_stream_scratch_bytes": (c_int64, [c_int64, c_int, c_int, c_int]),
    "evrep_event_volume_stream": (c_int, [P, P, P, P, c_int64, P, c_int, c_int64, c_int, c_int, c_int, P, P, c_int, c_int,
                                          P, c_int64, P, c_int64, P]),
    "evrep_event_volume_spans_scratch_bytes": (c_int64, [c_int64, c_int, c_int, c_int, c_int, c_int]),
    "evrep_event_volume_spans": (c_int, [P, P, P, P, c_int64, P, c_int, P, c_int, c_int, c_int, c_int, P, P, c_int, c_int,
                                         P, c_int64, P, c_int64, P, c_int64, P, P, P]),
    "evrep_event_volume_u8_batch": (c_int, [P, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P, P]),
    "evrep_count_stream_scratch_bytes": (c_int64, [c_int64, c_int, c_int, c_int, c_int]),
    "evrep_count_stream": (c_int, [P, P, P, P, c_int64, P, c_int, P, c_int, c_int, c_int, P, P, c_int, c_int, P, c_int64,
                                   P, c_int64, P]),
    "evrep_count_lut_u8_batch": (c_int, [P, c_int64, c_int64, c_int, c_int, c_int, c_int, P, P, P, P]),
    "evrep_sae_stream_scratch_bytes": (c_int64, [c_int64, P, c_int, c_int, c_int]),
    "evrep_sae_stream": (c_int, [P, P, P, P, c_int64, P, c_int, c_int, c_int, P, P, c_int, c_int, P, c_int, P, c_int64,
                                 P, c_int64, P]),
    "evrep_sae_decay_u8_batch": (c_int, [P, c_int64, P, c_int64, c_int, c_int, c_int, c_int, P, P, P, c_int, P, P]),
    "evrep_timesurface_scratch_bytes": (c_int64, [c_int, c_int]),
    "evrep_timesurface": (c_int, [P, P, P, c_int64, c_int, c_int, P, P, P, P]),
    "evrep_load_samples": (c_int, [P, c_int64, c_int, c_int, c_int, c_int, P, c_int, c_int, P, P]),
    "evrep_nearest_resize": (c_int, [P, c_int, c_int, c_int, c_int, c_int, P, P, P, P]),
    "evrep_quantize_u8": (c_int, [P, c_int64, c_int, P, P]),
    "evrep_taf_leaky_u8": (c_int, [P, c_int, c_int, c_int, c_int, c_int, P, P, P, P]),
    "evrep_taf_leaky_u8_batch": (c_int, [P, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P, P]),
    "evrep_leaky_transform": (c_int, [P, c_int64, P, P]),
    "evrep_sparse_splat": (c_int, [P, c_int64, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_float, c_float, P, P]),
    "evrep_pixel_major_to_planar": (c_int, [P, c_int, c_int64, c_int, P, P]),
    "evrep_sparse_agile_shift": (c_int, [P, P, c_int64, c_int, P, P]),
    "evrep_sparse_taf": (c_int, [P, c_int64, c_int, c_int, c_int, c_int, P, P]),
    "evrep_sparse_event_frame": (c_int, [P, c_int64, c_int, c_int, c_int, P, P]),
    "evrep_sparse_to_dense": (c_int, [P, P, c_int64, c_int, c_int, c_int, c_int, P, P]),
    "evrep_compact_scratch_bytes": (c_int64, [c_int64]),
    "evrep_dense_to_sparse": (c_int, [P, P, c_int, c_int, P, P, P, P, P]),
    "evrep_event_memory_update": (c_int, [P, c_int64, P, c_int64, c_double, P, P, P, P, P]),
    "evrep_taf_online_scratch_bytes": (c_int64, [c_int, c_int, c_int]),
    "evrep_taf_online_bin": (c_int, [P, c_int64, c_double, c_double, c_double, c_int, c_int, c_int, c_int, P, P, P, P]),
    "evrep_event_queue_tensor": (c_int, [P, c_int64, c_int, c_int, c_int, c_int, P, c_int, P, P, P]),
}

_lib = None


class EvrepError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load (once) and return the shared library; raise if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise EvrepError(
                "libevrep.so is missing (%s). Build it with `python -m frlw_evd_b200.build`; "
                "there is no CPU fallback for the encoders." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        lib = load()
        msg = lib.evrep_strerror(int(rc)).decode()
        if rc == -2:
            msg += ": " + lib.evrep_last_cuda_error().decode()
        raise EvrepError("%s failed: %s (code %d)" % (what or "libevrep call", msg, rc))


def call(name: str, *args):
    """Invoke an ``int``-returning entry point and raise on a non-zero code."""
    check(getattr(load(), name)(*args), name)
