"""ctypes binding of the host emulator (tests/emu/fac_emu.cpp): the product's flattened automaton,
segmentation predicates and slot formulation run sequentially on the CPU.  TEST CODE ONLY."""
import ctypes as C
import os
import subprocess

from fac_b200._abi import fac_config, fac_match, fac_pattern

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_LIB = os.path.join(EMU_DIR, "_build", "libfac_emu.so")
CSRC = os.path.join(ROOT, "fuzzy-aho-corasick-rs_b200", "csrc")


def build_emu():
    srcs = [os.path.join(EMU_DIR, "fac_emu.cpp"), os.path.join(CSRC, "fac_builder.cpp")]
    deps = srcs + [os.path.join(CSRC, f) for f in ("fac_core.h", "fac_types.h", "fac_unicode.h", "fac_builder.h", "fac_succinct.h", "fac_flat.h")]
    if (not os.path.exists(EMU_LIB)) or any(os.path.getmtime(EMU_LIB) < os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(EMU_LIB), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off"] + srcs +
                              ["-o", EMU_LIB])
    return EMU_LIB


class EmuBackend:
    """Same Python-side backend interface as GpuBackend / OracleBackend (search only)."""
    name = "emu"

    def __init__(self, tile=32):
        self.lib = C.CDLL(build_emu())
        self.tile = tile
        self.lib.emu_search.argtypes = [C.POINTER(fac_config), C.POINTER(fac_pattern), C.c_size_t, C.c_char_p,
                                        C.c_size_t, C.c_float, C.c_uint32, C.POINTER(C.POINTER(fac_match)),
                                        C.POINTER(C.c_size_t), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
        self.lib.emu_free.argtypes = [C.c_void_p]
        self.lib.emu_engine_info.argtypes = [C.POINTER(fac_config), C.POINTER(fac_pattern), C.c_size_t,
                                             C.POINTER(C.c_uint64)]
        self.lib.emu_segment.argtypes = [C.POINTER(fac_config), C.POINTER(fac_pattern), C.c_size_t, C.c_char_p,
                                         C.c_size_t, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32),
                                         C.POINTER(C.c_uint32), C.c_size_t]
        self.lib.emu_search_succinct.argtypes = [C.POINTER(fac_config), C.POINTER(fac_pattern), C.c_size_t, C.c_char_p,
                                                 C.c_size_t, C.c_float, C.POINTER(C.POINTER(fac_match)),
                                                 C.POINTER(C.c_size_t), C.POINTER(C.c_uint64)]
        self.lib.emu_search_flat.argtypes = self.lib.emu_search_succinct.argtypes
        self.lib.emu_set_faithful_flat.argtypes = [C.c_int]
        self.flat = False          # True: run the general stack-machine fast path (merged records) where it applies
        self.flat_used = 0
        self._keep = {}
        self._next = 1
        self.succinct = False      # True: run the succinct-trie fast path where it applies
        self.succinct_used = 0
        self.succinct_dirty = 0

    def create(self, cfg, pats, n, device=None):
        info = (C.c_uint64 * 8)()
        st = self.lib.emu_engine_info(C.byref(cfg), pats, n, info)
        if st != 0:
            raise RuntimeError("emu build status %d" % st)
        h = C.c_void_p(self._next)
        self._next += 1
        self._keep[h.value] = (cfg, pats, n, list(info))
        return h

    def free(self, h):
        self._keep.pop(h.value, None)

    def info(self, h):
        return self._keep[h.value][3]

    def max_match_graphemes(self, h):
        return self.info(h)[1]

    def prefilter_active(self, h):
        return bool(self.info(h)[3])

    def num_nodes(self, h):
        return self.info(h)[0]

    def search(self, h, data, thr, order, overlap, use_prefilter, per_window=False):
        cfg, pats, n, _ = self._keep[h.value]
        assert order == 0 and overlap == 0 and not use_prefilter, "emulator covers the raw search only"
        out = C.POINTER(fac_match)()
        cnt = C.c_size_t(0)
        states = C.c_uint64(0)
        if self.succinct:
            info = (C.c_uint64 * 4)()
            st = self.lib.emu_search_succinct(C.byref(cfg), pats, n, data, len(data), thr, C.byref(out), C.byref(cnt), info)
            if st == 0:
                self.succinct_used += 1
                self.succinct_dirty += info[0]
                arr = (fac_match * cnt.value)()
                if cnt.value:
                    C.memmove(arr, out, cnt.value * C.sizeof(fac_match))
                self.lib.emu_free(out)
                return arr, {"states_pushed": info[1], "dirty_windows": info[0], "candidates": info[2], "keys": info[3]}
            if st != -3:
                raise RuntimeError("emu succinct status %d" % st)
        if self.flat:
            info = (C.c_uint64 * 4)()
            st = self.lib.emu_search_flat(C.byref(cfg), pats, n, data, len(data), thr, C.byref(out), C.byref(cnt), info)
            if st == 0:
                self.flat_used += 1
                arr = (fac_match * cnt.value)()
                if cnt.value:
                    C.memmove(arr, out, cnt.value * C.sizeof(fac_match))
                self.lib.emu_free(out)
                return arr, {"states_pushed": info[1], "dirty_windows": info[0], "candidates": info[2], "keys": info[3]}
            if st != -3:
                raise RuntimeError("emu flat status %d" % st)
        pw = (C.c_uint32 * max(1, len(data)))() if per_window else None
        st = self.lib.emu_search(C.byref(cfg), pats, n, data, len(data), thr, self.tile, C.byref(out), C.byref(cnt),
                                 C.byref(states), pw)
        if st != 0:
            raise RuntimeError("emu status %d" % st)
        arr = (fac_match * cnt.value)()
        if cnt.value:
            C.memmove(arr, out, cnt.value * C.sizeof(fac_match))
        self.lib.emu_free(out)
        stats = {"states_pushed": states.value}
        if per_window:
            stats["per_window"] = list(pw)
        return arr, stats
