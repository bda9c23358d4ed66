"""ctypes binding of include/tob200.h (tensororder_b200/csrc/libtob200.so).

This is the stub a TensorOrder maintainer would add next to
`src/tensor_network/tensor_apis/numpy_apis.py`.  The product path has no
fallback: if the shared library is missing, importing this module raises."""
import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_double, c_float, c_int32, c_int64, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libtob200.so")

TOB_OK, TOB_E_INVALID, TOB_E_OOM, TOB_E_CUDA, TOB_E_NODEVICE = 0, 1, 2, 3, 4


class tob_plan_desc(ctypes.Structure):
    _fields_ = [
        ("n_nodes", c_int32),
        ("node_left", POINTER(c_int32)),
        ("node_right", POINTER(c_int32)),
        ("node_leaf", POINTER(c_int32)),
        ("n_leaves", c_int32),
        ("leaf_rank", POINTER(c_int32)),
        ("leaf_data_offset", POINTER(c_int64)),
        ("leaf_axis_start", POINTER(c_int32)),
        ("leaf_axis_edge", POINTER(c_int32)),
        ("n_slice_groups", c_int32),
        ("leaf_data_len", c_int64),
    ]


class tob_options(ctypes.Structure):
    _fields_ = [
        ("device", c_int32),
        ("use_graph", c_int32),
        ("kernel_policy", c_int32),
        ("hoist_invariant", c_int32),
        ("mem_limit_bytes", c_int64),
        ("use_microtree", c_int32),
        ("slice_lanes", c_int32),
        ("dag_branches", c_int32),
    ]


# every symbol include/tob200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "tob_default_options": (None, [POINTER(tob_options)]),
    "tob_plan_create": (c_int32, [POINTER(tob_plan_desc), POINTER(tob_options), POINTER(c_void_p)]),
    "tob_plan_peak_bytes": (c_int64, [c_void_p]),
    "tob_plan_num_slices": (c_uint64, [c_void_p]),
    "tob_plan_describe": (c_int64, [c_void_p, c_char_p, c_int64]),
    "tob_plan_upload": (c_int32, [c_void_p, POINTER(c_double), c_int64]),
    "tob_plan_update_leaves": (c_int32, [c_void_p, POINTER(c_double), c_int64]),
    "tob_plan_release": (c_int32, [c_void_p]),
    "tob_pool_trim": (c_int32, [c_int32]),
    "tob_plan_run": (c_int32, [c_void_p, c_uint64, c_uint64, c_uint64, POINTER(c_double)]),
    "tob_plan_run_ex": (c_int32, [c_void_p, c_uint64, c_uint64, c_uint64, c_double, c_int32, POINTER(c_double)]),
    "tob_plan_run_async": (c_int32, [c_void_p, c_uint64, c_uint64, c_uint64, c_double, c_int32, c_void_p]),
    "tob_plan_join": (c_int32, [c_void_p, c_void_p]),
    "tob_plan_wait": (c_int32, [c_void_p, POINTER(c_double)]),
    "tob_plan_last_ms": (c_double, [c_void_p]),
    "tob_plan_last_issue_ms": (c_double, [c_void_p]),
    "tob_plan_last_launches": (c_int64, [c_void_p]),
    "tob_plan_set_modulus": (c_int32, [c_void_p, c_double]),
    "tob_plan_set_gemm_timing": (c_int32, [c_void_p, c_int32]),
    "tob_plan_last_gemm": (c_int32, [c_void_p, POINTER(c_double), POINTER(c_double), POINTER(c_int64)]),
    "tob_plan_set_stream": (c_int32, [c_void_p, c_void_p]),
    "tob_plan_num_ops": (c_int64, [c_void_p]),
    "tob_plan_work": (c_int32, [c_void_p, POINTER(c_double), POINTER(c_double), POINTER(c_int64)]),
    "tob_plan_profile": (c_int32, [c_void_p, c_uint64, POINTER(c_float), c_int64, POINTER(c_double)]),
    "tob_plan_debug_run": (c_int32, [c_void_p, c_uint64, c_int64]),
    "tob_plan_debug_read": (c_int32, [c_void_p, c_int32, c_int64, c_int64, POINTER(c_double)]),
    "tob_plan_destroy": (None, [c_void_p]),
    "tob_tensordot_device": (c_int32, [c_void_p, c_int32, c_void_p, c_int32, POINTER(c_int32), POINTER(c_int32),
                                       c_int32, c_void_p, c_void_p, c_int64, c_int32, c_void_p, POINTER(c_float)]),
    "tob_tensordot_host": (c_int32, [POINTER(c_double), c_int32, POINTER(c_double), c_int32, POINTER(c_int32),
                                     POINTER(c_int32), c_int32, POINTER(c_double)]),
    "tob_permute_device": (c_int32, [c_void_p, c_void_p, c_int32, POINTER(c_int32), c_void_p, POINTER(c_float)]),
    "tob_tuning_set": (c_int32, [c_char_p, c_double]),
    "tob_tuning_get": (c_int32, [c_char_p, POINTER(c_double)]),
    "tob_gemm_time_model_us": (c_double, [c_int32, c_int32, c_int32, c_int32, c_int32, c_int32]),
    "tob_warm": (c_int32, [c_int32]),
    "tob_device_count": (c_int32, []),
    "tob_version": (c_char_p, []),
    "tob_last_error": (c_char_p, []),
}

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "tensororder_b200: %s is missing. Build it with `python -m tensororder_b200.build` "
        "(needs nvcc); there is no CPU fallback." % LIB_PATH
    )

lib = ctypes.CDLL(LIB_PATH)
for _name, (_res, _args) in SYMBOLS.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args


def last_error() -> str:
    return lib.tob_last_error().decode("utf-8", "replace")
