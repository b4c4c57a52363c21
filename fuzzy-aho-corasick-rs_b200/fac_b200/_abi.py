"""ctypes mirror of include/fac.h (struct layouts + the libfacgpu.so loader).

The product path has no CPU fallback: `load_library()` raises if the CUDA library has not
been built, and every entry point fails loudly (FAC_CUDA_ERROR) without a usable GPU.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "csrc", "libfacgpu.so")

FAC_OK, FAC_HAYSTACK_TOO_LARGE, FAC_INVALID_UTF8, FAC_CUDA_ERROR, FAC_OOM, FAC_INVALID_ARGUMENT, \
    FAC_UNSUPPORTED, FAC_IO_ERROR = range(8)
ORDER_UNSORTED, ORDER_DEFAULT, ORDER_GREEDY, ORDER_COVERAGE_WEIGHTED = range(4)
OVERLAP_KEEP, OVERLAP_NON_OVERLAPPING, OVERLAP_NON_OVERLAPPING_UNIQUE = range(3)


class fac_limits(C.Structure):
    _fields_ = [("insertions", C.c_int16), ("deletions", C.c_int16), ("substitutions", C.c_int16),
                ("swaps", C.c_int16), ("edits", C.c_int16)]


class fac_pattern(C.Structure):
    _fields_ = [("text", C.c_char_p), ("len", C.c_size_t), ("weight", C.c_float), ("has_limits", C.c_int32),
                ("limits", fac_limits), ("unique_id", C.c_int64)]


class fac_sim_pair(C.Structure):
    _fields_ = [("a", C.c_uint32), ("b", C.c_uint32), ("similarity", C.c_float)]


class fac_mapping(C.Structure):
    _fields_ = [("a", C.c_char_p), ("a_len", C.c_size_t), ("b", C.c_char_p), ("b_len", C.c_size_t),
                ("score", C.c_float)]


class fac_config(C.Structure):
    _fields_ = [("case_insensitive", C.c_int32), ("has_limits", C.c_int32), ("limits", fac_limits),
                ("has_penalties", C.c_int32), ("penalty_insertion", C.c_float), ("penalty_deletion", C.c_float),
                ("penalty_substitution", C.c_float), ("penalty_swap", C.c_float), ("beam_width", C.c_uint64),
                ("has_auto_beam", C.c_int32), ("auto_beam_budget", C.c_uint64), ("auto_beam_width", C.c_uint64),
                ("min_symbol_similarity", C.c_float), ("has_similarity", C.c_int32),
                ("similarity", C.POINTER(fac_sim_pair)), ("n_similarity", C.c_size_t),
                ("mappings", C.POINTER(fac_mapping)), ("n_mappings", C.c_size_t)]


class fac_match(C.Structure):
    _fields_ = [("start", C.c_uint64), ("end", C.c_uint64), ("pattern_index", C.c_uint32),
                ("similarity", C.c_float), ("insertions", C.c_uint8), ("deletions", C.c_uint8),
                ("substitutions", C.c_uint8), ("swaps", C.c_uint8), ("edits", C.c_uint8), ("pad_", C.c_uint8 * 3)]


class fac_window(C.Structure):
    _fields_ = [("text", C.c_void_p), ("len", C.c_size_t), ("base", C.c_uint64), ("commit", C.c_size_t)]


class fac_shard(C.Structure):
    _fields_ = [("own_begin", C.c_size_t), ("own_end", C.c_size_t), ("read_end", C.c_size_t)]


class fac_search_args(C.Structure):
    _fields_ = [("haystack", C.c_void_p), ("len", C.c_size_t), ("own_begin", C.c_size_t), ("own_end", C.c_size_t),
                ("base", C.c_uint64), ("threshold", C.c_float), ("order", C.c_int), ("overlap", C.c_int),
                ("use_prefilter", C.c_int32), ("flags", C.c_uint32)]


class fac_stream_stats(C.Structure):
    _fields_ = [("bytes_read", C.c_uint64), ("bytes_written", C.c_uint64), ("windows", C.c_uint64), ("matches", C.c_uint64),
                ("states", C.c_uint64), ("device_ms", C.c_double), ("expand_ms", C.c_double), ("kernel_launches", C.c_uint32),
                ("devices", C.c_uint32)]


FAC_HAYSTACK_ON_DEVICE, FAC_RESULT_ON_DEVICE, FAC_TEXT_IS_UNICODE, FAC_APPLY_PRESORTED = 1, 2, 4, 8

assert C.sizeof(fac_match) == 32

READ_FN = C.CFUNCTYPE(C.c_int64, C.c_void_p, C.POINTER(C.c_uint8), C.c_size_t)
WRITE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint8), C.c_size_t)
MATCH_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(fac_match))
REPLACE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(fac_match), C.c_uint64, C.POINTER(C.c_uint8), C.c_size_t,
                         C.POINTER(C.c_void_p), C.POINTER(C.c_size_t))

# Every symbol include/fac.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("fac_last_error_string", C.c_char_p, []),
    ("fac_abi_version", C.c_int, []),
    ("fac_build_source_hash", C.c_char_p, []),
    ("fac_engine_create", C.c_int, [C.POINTER(fac_config), C.POINTER(fac_pattern), C.c_size_t, C.POINTER(C.c_void_p)]),
    ("fac_engine_create_on", C.c_int, [C.c_int, C.POINTER(fac_config), C.POINTER(fac_pattern), C.c_size_t, C.POINTER(C.c_void_p)]),
    ("fac_engine_create_multi", C.c_int, [C.POINTER(C.c_int), C.c_size_t, C.POINTER(fac_config), C.POINTER(fac_pattern), C.c_size_t, C.POINTER(C.c_void_p)]),
    ("fac_engine_free", None, [C.c_void_p]),
    ("fac_engine_max_match_graphemes", C.c_size_t, [C.c_void_p]),
    ("fac_engine_prefilter_active", C.c_int, [C.c_void_p]),
    ("fac_engine_num_nodes", C.c_size_t, [C.c_void_p]),
    ("fac_engine_num_patterns", C.c_size_t, [C.c_void_p]),
    ("fac_engine_device", C.c_int, [C.c_void_p]),
    ("fac_engine_num_devices", C.c_size_t, [C.c_void_p]),
    ("fac_search", C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_float, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    ("fac_search_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_float, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    ("fac_last_haystack_graphemes", C.c_uint64, []),
    ("fac_search_shard", C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint64, C.c_float, C.c_int, C.POINTER(C.c_void_p)]),
    ("fac_matches_apply", C.c_int, [C.c_void_p, C.POINTER(fac_match), C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    ("fac_plan_shards", C.c_int, [C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.POINTER(fac_shard)]),
    ("fac_search_ex", C.c_int, [C.c_void_p, C.POINTER(fac_search_args), C.POINTER(C.c_void_p)]),
    ("fac_matches_apply_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_uint32, C.POINTER(C.c_void_p)]),
    ("fac_matches_device_data", C.c_void_p, [C.c_void_p]),
    ("fac_search_windows", C.c_int, [C.c_void_p, C.POINTER(fac_window), C.c_size_t, C.c_float, C.POINTER(C.c_void_p)]),
    ("fac_matches_data", C.POINTER(fac_match), [C.c_void_p]),
    ("fac_matches_len", C.c_size_t, [C.c_void_p]),
    ("fac_matches_states_pushed", C.c_uint64, [C.c_void_p]),
    ("fac_matches_device_ms", C.c_double, [C.c_void_p]),
    ("fac_matches_expand_ms", C.c_double, [C.c_void_p]),
    ("fac_matches_kernel_launches", C.c_uint32, [C.c_void_p]),
    ("fac_matches_free", None, [C.c_void_p]),
    ("fac_search_stream", C.c_int, [C.c_void_p, READ_FN, C.c_void_p, C.c_float, MATCH_FN, C.c_void_p, C.POINTER(C.c_uint64)]),
    ("fac_replace_stream_table", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.POINTER(C.c_void_p),
                                           C.POINTER(C.c_size_t), C.c_size_t, C.POINTER(fac_stream_stats)]),
    ("fac_search_stream_stats", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.POINTER(fac_stream_stats)]),
    ("fac_replace_stream", C.c_int, [C.c_void_p, READ_FN, C.c_void_p, WRITE_FN, C.c_void_p, C.c_float,
                                     REPLACE_FN, C.c_void_p, C.POINTER(C.c_uint64)]),
]

_lib = None
CSRC_DIR = os.path.join(os.path.dirname(_HERE), "csrc")
HEADER = os.path.join(os.path.dirname(os.path.dirname(_HERE)), "include", "fac.h")


def source_hash():
    """SHA-256 over the library's sources (csrc/*.{cu,cuh,h,cpp,inc} + include/fac.h), the stamp build() compiles in."""
    import hashlib
    h = hashlib.sha256()
    files = sorted(f for f in os.listdir(CSRC_DIR) if f.endswith((".cu", ".cuh", ".h", ".cpp", ".inc")))
    for f in [os.path.join(CSRC_DIR, f) for f in files] + [HEADER]:
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()


def load_library(path=None):
    """Load libfacgpu.so and type every exported symbol.  Raises if it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            "libfacgpu.so not found at %s: build it with `python __graft_entry__.py` "
            "(there is no CPU fallback for the search path)" % p)
    lib = C.CDLL(p)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the ABI is incomplete
        fn.restype = res
        fn.argtypes = args
    if path is None and os.path.isdir(CSRC_DIR) and os.environ.get("FAC_ALLOW_STALE_LIB") != "1":
        built, want = lib.fac_build_source_hash().decode(), source_hash()
        if built != want:
            raise RuntimeError("libfacgpu.so was built from other sources (stamp %s..., tree %s...): rebuild with "
                               "`python __graft_entry__.py`" % (built[:12], want[:12]))
    if path is None:
        _lib = lib
    return lib
