"""ctypes binding of the CPU oracle (oracle/fac_oracle.cpp) with the same Python-side backend
interface as fac_b200.GpuBackend, so one test body runs against both.  TEST CODE ONLY."""
import ctypes as C
import os
import subprocess

from fac_b200 import _abi
from fac_b200._abi import fac_config, fac_match, fac_pattern
from fac_b200.api import HaystackTooLarge, SearchError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "_build", "libfac_oracle.so")


def build_oracle():
    src = os.path.join(ORACLE_DIR, "fac_oracle.cpp")
    if (not os.path.exists(ORACLE_LIB)) or os.path.getmtime(ORACLE_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return ORACLE_LIB


class OracleBackend:
    name = "oracle"

    def __init__(self):
        lib = C.CDLL(build_oracle())
        self.lib = lib
        vp = C.c_void_p
        lib.orc_engine_create.restype = vp
        lib.orc_engine_create.argtypes = [C.POINTER(fac_config), C.POINTER(fac_pattern), C.c_size_t]
        lib.orc_engine_free.argtypes = [vp]
        lib.orc_engine_max_match_graphemes.restype = C.c_size_t
        lib.orc_engine_max_match_graphemes.argtypes = [vp]
        lib.orc_engine_num_nodes.restype = C.c_size_t
        lib.orc_engine_num_nodes.argtypes = [vp]
        lib.orc_engine_max_edits_fast.argtypes = [vp]
        lib.orc_engine_prefilter_active.argtypes = [vp]
        lib.orc_search.argtypes = [vp, C.c_char_p, C.c_size_t, C.c_float, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
        lib.orc_window_states.argtypes = [vp, C.c_char_p, C.c_size_t, C.c_float, C.POINTER(C.c_uint32), C.c_size_t]
        lib.orc_apply.argtypes = [vp, C.POINTER(fac_match), C.c_size_t, C.c_int, C.c_int, C.POINTER(vp)]
        lib.orc_search_stream.argtypes = [vp, C.c_char_p, C.c_size_t, C.c_size_t, C.c_float, C.c_int, C.POINTER(vp)]
        lib.orc_search_parallel.argtypes = [vp, C.c_void_p, C.c_size_t, C.c_float, C.c_int, C.POINTER(vp)]
        lib.orc_cut_windows.argtypes = [vp, C.c_char_p, C.c_size_t, C.c_size_t, C.POINTER(C.c_uint64), C.c_size_t]
        lib.orc_replace_stream.argtypes = [vp, C.c_char_p, C.c_size_t, C.c_size_t, C.c_float, _abi.REPLACE_FN,
                                           C.c_void_p, C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_size_t)]
        lib.orc_free_buf.argtypes = [C.POINTER(C.c_uint8)]
        lib.orc_matches_data.restype = C.POINTER(fac_match)
        lib.orc_matches_data.argtypes = [vp]
        lib.orc_matches_len.restype = C.c_size_t
        lib.orc_matches_len.argtypes = [vp]
        lib.orc_matches_states_pushed.restype = C.c_uint64
        lib.orc_matches_states_pushed.argtypes = [vp]
        lib.orc_matches_free.argtypes = [vp]
        lib.orc_grapheme_starts.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_uint64), C.c_size_t]
        lib.orc_to_lowercase.restype = C.c_size_t
        lib.orc_to_lowercase.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
        lib.orc_utf8_valid_up_to.restype = C.c_size_t
        lib.orc_utf8_valid_up_to.argtypes = [C.c_char_p, C.c_size_t]

    # ---- backend interface ----
    def create(self, cfg, pats, n, device=None):
        return C.c_void_p(self.lib.orc_engine_create(C.byref(cfg), pats, n))

    def free(self, h):
        self.lib.orc_engine_free(h)

    def max_match_graphemes(self, h):
        return self.lib.orc_engine_max_match_graphemes(h)

    def prefilter_active(self, h):
        return bool(self.lib.orc_engine_prefilter_active(h))

    def num_nodes(self, h):
        return self.lib.orc_engine_num_nodes(h)

    def _take(self, mh):
        n = self.lib.orc_matches_len(mh)
        arr = (fac_match * n)()
        if n:
            C.memmove(arr, self.lib.orc_matches_data(mh), n * C.sizeof(fac_match))
        stats = {"states_pushed": int(self.lib.orc_matches_states_pushed(mh))}
        self.lib.orc_matches_free(mh)
        return arr, stats

    def search(self, h, data, thr, order, overlap, use_prefilter):
        mh = C.c_void_p()
        st = self.lib.orc_search(h, data, len(data), thr, order, overlap, int(use_prefilter), C.byref(mh))
        if st == _abi.FAC_HAYSTACK_TOO_LARGE:
            raise HaystackTooLarge(st, "haystack too large")
        if st != 0:
            raise SearchError(st, "oracle status %d" % st)
        return self._take(mh)

    def search_parallel(self, h, ptr, n, thr, threads):
        """Raw search of n bytes at host address `ptr` over `threads` cores (ASCII input)."""
        mh = C.c_void_p()
        st = self.lib.orc_search_parallel(h, C.c_void_p(ptr), n, thr, threads, C.byref(mh))
        if st != 0:
            raise SearchError(st, "oracle parallel search status %d" % st)
        return self._take(mh)

    def apply(self, h, arr, n, order, overlap):
        mh = C.c_void_p()
        self.lib.orc_apply(h, arr, n, order, overlap, C.byref(mh))
        return self._take(mh)

    def window_states(self, h, data, thr):
        cnt = (C.c_uint32 * max(1, len(data)))()
        n = self.lib.orc_window_states(h, data, len(data), thr, cnt, len(data))
        return list(cnt[:n])

    def cut_windows(self, h, data, read_block=0):
        cap = len(data) // 1024 + 16
        tri = (C.c_uint64 * (3 * cap))()
        n = self.lib.orc_cut_windows(h, data, len(data), read_block, tri, cap)
        return [(tri[3 * i], tri[3 * i + 1], tri[3 * i + 2]) for i in range(n)]

    def search_stream_mem(self, h, data, thr, threads=1, read_block=0):
        mh = C.c_void_p()
        self.lib.orc_search_stream(h, data, len(data), read_block, thr, threads, C.byref(mh))
        return self._take(mh)

    @staticmethod
    def _read_all(reader):
        parts = []
        while True:
            b = reader.read()
            if not b:
                break
            parts.append(bytes(b))
        return b"".join(parts)

    def search_stream(self, h, reader, thr, on_match):
        data = self._read_all(reader)
        arr, _ = self.search_stream_mem(h, data, thr)
        for c in arr:
            on_match(c)
        return len(data)

    def replace_stream(self, h, reader, writer, thr, callback, read_block=0):
        data = self._read_all(reader)
        keep = []

        def rp(_u, pm, base, text, n, out_p, out_n):
            r = callback(pm.contents, C.string_at(text, n))
            if r is None:
                return 0
            b = C.create_string_buffer(r.encode("utf-8") if isinstance(r, str) else bytes(r))
            keep[:] = [b]
            out_p[0] = C.cast(b, C.c_void_p).value
            out_n[0] = len(b) - 1
            return 1

        ob = C.POINTER(C.c_uint8)()
        ol = C.c_size_t(0)
        self.lib.orc_replace_stream(h, data, len(data), read_block, thr, _abi.REPLACE_FN(rp), None, C.byref(ob),
                                    C.byref(ol))
        out = C.string_at(ob, ol.value)
        self.lib.orc_free_buf(ob)
        writer.write(out)
        return len(out)

    # ---- unicode helpers ----
    def grapheme_starts(self, data):
        out = (C.c_uint64 * max(1, len(data)))()
        n = self.lib.orc_grapheme_starts(data, len(data), out, len(data))
        return list(out[:n])

    def to_lowercase(self, data):
        buf = C.create_string_buffer(3 * len(data) + 8)
        n = self.lib.orc_to_lowercase(data, len(data), buf, len(buf))
        return buf.raw[:n]
