"""Host-side mirror of the reference's public interface over the C ABI (include/fac.h).

Names, argument meaning and error behaviour follow kakserpom/fuzzy-aho-corasick-rs v0.5.0
(paths under /root/reference) so the parity tests read like the reference's own tests:

  FuzzyAhoCorasickBuilder  src/builder.rs:22-184      FuzzyLimits / FuzzyPenalties  src/structs.rs:292-420
  Pattern                  src/structs.rs:597-754     SearchOptions / Order / Overlap  src/options.rs
  FuzzyAhoCorasick.search  src/query.rs:30-38         segmented helpers  src/query.rs:46-201
  FuzzyMatches helpers     src/matches.rs:151-595     Prefiltered  src/prefilter.rs:101-156
  streaming                src/stream.rs:319-638      FuzzyReplacer  src/replacer.rs

The search itself (grapheme segmentation, frontier expansion, best-per-span dedup, ranking,
overlap resolution, window cutting) runs behind the C ABI on the GPU; what lives here is the
O(matches) string assembly the reference keeps in Rust above its search call.
"""
import ctypes as C

from . import _abi
from ._abi import (fac_config, fac_limits, fac_mapping, fac_match, fac_pattern, fac_sim_pair,
                   ORDER_UNSORTED, ORDER_DEFAULT, ORDER_GREEDY, ORDER_COVERAGE_WEIGHTED,
                   OVERLAP_KEEP, OVERLAP_NON_OVERLAPPING, OVERLAP_NON_OVERLAPPING_UNIQUE)

DEFAULT_THRESHOLD = 0.0  # src/options.rs:7


class SearchError(Exception):
    """SearchError (src/error.rs).  `graphemes` is set for HaystackTooLarge."""

    def __init__(self, status, message, graphemes=None):
        super().__init__(message)
        self.status = status
        self.graphemes = graphemes


class HaystackTooLarge(SearchError):
    pass


class Order:
    Unsorted, Default, Greedy, CoverageWeighted = (ORDER_UNSORTED, ORDER_DEFAULT, ORDER_GREEDY,
                                                   ORDER_COVERAGE_WEIGHTED)


class Overlap:
    Keep, NonOverlapping, NonOverlappingUnique = (OVERLAP_KEEP, OVERLAP_NON_OVERLAPPING,
                                                  OVERLAP_NON_OVERLAPPING_UNIQUE)


class SearchOptions:
    """src/options.rs:45-132 (chainable, value semantics)."""

    def __init__(self, threshold=DEFAULT_THRESHOLD, order=Order.Unsorted, overlap=Overlap.Keep):
        self.threshold_, self.order_, self.overlap_ = threshold, order, overlap

    @staticmethod
    def new():
        return SearchOptions()

    def _with(self, **kw):
        o = SearchOptions(self.threshold_, self.order_, self.overlap_)
        for k, v in kw.items():
            setattr(o, k, v)
        return o

    def threshold(self, t):
        return self._with(threshold_=t)

    def order(self, o):
        return self._with(order_=o)

    def overlap(self, o):
        return self._with(overlap_=o)

    def sorted(self):
        return self.order(Order.Default)

    def greedy(self):
        return self.order(Order.Greedy)

    def coverage_weighted(self):
        return self.order(Order.CoverageWeighted)

    def non_overlapping(self):
        return self.overlap(Overlap.NonOverlapping)

    def non_overlapping_unique(self):
        return self.overlap(Overlap.NonOverlappingUnique)


class FuzzyLimits:
    """src/structs.rs:292-363; None = unset."""

    def __init__(self):
        self.insertions_ = self.deletions_ = self.substitutions_ = self.swaps_ = self.edits_ = None

    @staticmethod
    def new():
        return FuzzyLimits()

    def _set(self, k, v):
        o = FuzzyLimits()
        o.__dict__.update(self.__dict__)
        setattr(o, k, v)
        return o

    def insertions(self, n):
        return self._set("insertions_", n)

    def deletions(self, n):
        return self._set("deletions_", n)

    def substitutions(self, n):
        return self._set("substitutions_", n)

    def swaps(self, n):
        return self._set("swaps_", n)

    def edits(self, n):
        return self._set("edits_", n)

    def to_c(self):
        f = lambda v: -1 if v is None else int(v)
        return fac_limits(f(self.insertions_), f(self.deletions_), f(self.substitutions_), f(self.swaps_),
                          f(self.edits_))


class FuzzyPenalties:
    """src/structs.rs:369-420.  Defaults are computed in f32 by the library."""

    def __init__(self):
        self.insertion_ = self.deletion_ = self.substitution_ = self.swap_ = None

    @staticmethod
    def default():
        return FuzzyPenalties()

    def _set(self, k, v):
        o = FuzzyPenalties()
        o.__dict__.update(self.__dict__)
        setattr(o, k, v)
        return o

    def insertion(self, p):
        return self._set("insertion_", p)

    def deletion(self, p):
        return self._set("deletion_", p)

    def substitution(self, p):
        return self._set("substitution_", p)

    def swap(self, p):
        return self._set("swap_", p)


def _f32(x):
    return C.c_float(x).value


class Pattern:
    """src/structs.rs:597-754."""

    def __init__(self, pattern, weight=1.0, limits=None, custom_unique_id=None):
        self.pattern = pattern
        self.weight_ = weight
        self.limits = limits
        self.custom_unique_id_ = custom_unique_id

    @staticmethod
    def from_(x):
        if isinstance(x, Pattern):
            return x
        if isinstance(x, str):
            return Pattern(x)
        if isinstance(x, tuple) and len(x) == 2:
            return Pattern(x[0], x[1])
        if isinstance(x, tuple) and len(x) == 3:  # (text, weight, max_edits) structs.rs:732-754
            return Pattern(x[0], x[1], FuzzyLimits().edits(x[2]))
        raise TypeError("cannot convert %r into a Pattern" % (x,))

    def weight(self, w):
        return Pattern(self.pattern, w, self.limits, self.custom_unique_id_)

    def fuzzy(self, limits):
        return Pattern(self.pattern, self.weight_, limits, self.custom_unique_id_)

    def custom_unique_id(self, uid):
        return Pattern(self.pattern, self.weight_, self.limits, uid)

    def as_str(self):
        return self.pattern

    def __len__(self):  # bytes, structs.rs:628-630
        return len(self.pattern.encode("utf-8"))

    def __repr__(self):
        return "Pattern(%r)" % self.pattern


class FuzzyMatch:
    """src/structs.rs:757-781."""
    __slots__ = ("insertions", "deletions", "substitutions", "swaps", "edits", "pattern_index", "pattern",
                 "start", "end", "similarity", "text")

    def key(self):
        return (self.start, self.end, self.pattern_index)

    def as_tuple(self):
        """(start, end, pattern, similarity bits, ins, del, sub, swap, edits): the parity tuple."""
        return (self.start, self.end, self.pattern_index, C.c_uint32.from_buffer(C.c_float(self.similarity)).value,
                self.insertions, self.deletions, self.substitutions, self.swaps, self.edits)

    def __repr__(self):
        return ("FuzzyMatch(pattern=%r, text=%r, span=(%d,%d), sim=%.4f, ins=%d del=%d sub=%d swap=%d)" %
                (self.pattern.pattern, self.text, self.start, self.end, self.similarity, self.insertions,
                 self.deletions, self.substitutions, self.swaps))


class Segment:
    """src/structs.rs:785-846: kind is 'matched' or 'unmatched'."""
    __slots__ = ("kind", "start", "end", "text", "match")

    def __init__(self, kind, start, end, text, match=None):
        self.kind, self.start, self.end, self.text, self.match = kind, start, end, text, match

    def matched(self):
        return self.match if self.kind == "matched" else None

    def as_str(self):
        return self.text


class FuzzyMatches:
    """src/structs.rs:853-889 + the helpers of src/matches.rs:151-595 (host-side string assembly)."""

    def __init__(self, haystack_bytes, inner, stats=None):
        self.haystack = haystack_bytes
        self.inner = inner
        self.stats = stats or {}

    def __iter__(self):
        return iter(self.inner)

    def __len__(self):
        return len(self.inner)

    def __getitem__(self, i):
        return self.inner[i]

    def is_empty(self):
        return not self.inner

    def tuples(self):
        return [m.as_tuple() for m in self.inner]

    def matched_spans(self):  # matches.rs:481-483
        return [(m.start, m.end) for m in self.inner]

    def matched_strings(self):  # matches.rs:509-511
        return [m.text for m in self.inner]

    def _s(self, a, b=None):
        return self.haystack[a:b].decode("utf-8")

    def replace(self, callback):  # matches.rs:165-188
        out, last = [], 0
        for m in self.inner:
            if m.start >= last:
                out.append(self._s(last, m.start))
                last = m.end
                r = callback(m)
                out.append(m.text if r is None else r)
        out.append(self._s(last))
        return "".join(out)

    def segment_iter(self):  # matches.rs:526-558
        segs, last = [], 0
        for m in self.inner:
            if m.start >= last:
                if m.start > last:
                    segs.append(Segment("unmatched", last, m.start, self._s(last, m.start)))
                last = m.end
                segs.append(Segment("matched", m.start, m.end, m.text, m))
        n = len(self.haystack)
        if last < n:
            segs.append(Segment("unmatched", last, n, self._s(last)))
        return segs

    def strip_prefix(self):  # matches.rs:192-218
        out, skipping = [], True
        for seg in self.segment_iter():
            if seg.kind == "matched":
                if skipping:
                    continue
                out.append(seg.text)
            else:
                if skipping:
                    if not _rust_trim(seg.text):
                        continue
                    skipping = False
                    out.append(_rust_trim_start(seg.text))
                else:
                    out.append(seg.text)
        return "".join(out)

    def strip_suffix(self):  # matches.rs:222-254
        buf, keep = [], 0
        for seg in self.segment_iter():
            buf.append(seg)
            if seg.kind == "unmatched" and _rust_trim(seg.text):
                keep = len(buf)
        out = []
        for i, seg in enumerate(buf[:keep]):
            if seg.kind == "unmatched" and i + 1 == keep:
                out.append(_rust_trim_end(seg.text))
            else:
                out.append(seg.text)
        return "".join(out)

    def split(self):  # matches.rs:256-267
        return [s.text for s in self.segment_iter() if s.kind == "unmatched"]

    def segment_text(self):  # matches.rs:561-594
        space = (" ", "\t")
        no_lead = (",", ".", "?", "!", ";", ":", "—", "-", "…")
        result, prev_matched = "", False
        for seg in self.segment_iter():
            if seg.kind == "matched":
                if prev_matched or (result and not result.endswith(space)):
                    result += " "
                prev_matched = True
                result += seg.text
            else:
                if prev_matched and not seg.text.startswith(no_lead):
                    result += " "
                prev_matched = False
                result += seg.text
        return result


_RUST_WS = frozenset([0x20, 0x85, 0xA0, 0x1680, 0x2028, 0x2029, 0x202F, 0x205F, 0x3000]
                     + list(range(0x09, 0x0E)) + list(range(0x2000, 0x200B)))


def _is_rust_ws(ch):
    # char::is_whitespace == Unicode White_Space
    return ord(ch) in _RUST_WS


def _rust_trim_start(s):
    i = 0
    while i < len(s) and _is_rust_ws(s[i]):
        i += 1
    return s[i:]


def _rust_trim_end(s):
    j = len(s)
    while j > 0 and _is_rust_ws(s[j - 1]):
        j -= 1
    return s[:j]


def _rust_trim(s):
    return _rust_trim_end(_rust_trim_start(s))


class StreamMatch:
    """src/stream.rs:39-60 (absolute u64 offsets, owned text)."""
    __slots__ = ("start", "end", "pattern_index", "similarity", "insertions", "deletions", "substitutions",
                 "swaps", "edits", "text")

    def as_tuple(self):
        return (self.start, self.end, self.pattern_index, C.c_uint32.from_buffer(C.c_float(self.similarity)).value,
                self.insertions, self.deletions, self.substitutions, self.swaps, self.edits)


# ---------------------------------------------------------------------------------------------
# Backend: the thin layer that talks to the C ABI.  The GPU backend is the only product backend.
# ---------------------------------------------------------------------------------------------
class _MatchesOwner:
    """Keeps a fac_matches handle alive for as long as a zero-copy view of its records exists."""

    def __init__(self, lib, mh):
        self.lib, self.mh = lib, mh

    def __del__(self):
        try:
            self.lib.fac_matches_free(self.mh)
        except Exception:
            pass


class DeviceMatches:
    """A match list that stayed in device memory (FAC_RESULT_ON_DEVICE): `ptr` / `n` describe n fac_match records
    (32 bytes each) on the engine's device; the buffer lives until this object is collected."""

    def __init__(self, lib, mh):
        self._owner = _MatchesOwner(lib, mh)
        self.n = int(lib.fac_matches_len(mh))
        self.ptr = int(lib.fac_matches_device_data(mh) or 0)

    def __len__(self):
        return self.n

    @property
    def __cuda_array_interface__(self):
        return {"shape": (self.n * 32,), "typestr": "|u1", "data": (self.ptr, False), "version": 3, "strides": None}

    def as_tensor(self, device):
        """Zero-copy torch uint8 view [n * 32] (keeps this object alive through the tensor)."""
        import torch
        if self.n == 0:
            return torch.empty(0, dtype=torch.uint8, device=device)
        t = torch.as_tensor(self, device=device)
        t._fac_owner = self
        return t


class GpuBackend:
    name = "libfacgpu"

    def __init__(self, lib=None):
        self.lib = lib or _abi.load_library()

    def _err(self, status):
        msg = self.lib.fac_last_error_string().decode("utf-8", "replace")
        if status == _abi.FAC_HAYSTACK_TOO_LARGE:
            return HaystackTooLarge(status, msg, int(self.lib.fac_last_haystack_graphemes()))
        return SearchError(status, "fac status %d: %s" % (status, msg))

    def create(self, cfg, pats, n, device=None):
        h = C.c_void_p()
        if device is None:
            st = self.lib.fac_engine_create(C.byref(cfg), pats, n, C.byref(h))
        elif isinstance(device, (list, tuple)):   # one replica per device: the stream calls deal windows over them
            devs = (C.c_int * len(device))(*device)
            st = self.lib.fac_engine_create_multi(devs, len(device), C.byref(cfg), pats, n, C.byref(h))
        else:
            st = self.lib.fac_engine_create_on(device, C.byref(cfg), pats, n, C.byref(h))
        if st != 0:
            raise self._err(st)
        return h

    def free(self, h):
        self.lib.fac_engine_free(h)

    def max_match_graphemes(self, h):
        return self.lib.fac_engine_max_match_graphemes(h)

    def prefilter_active(self, h):
        return bool(self.lib.fac_engine_prefilter_active(h))

    def num_nodes(self, h):
        return self.lib.fac_engine_num_nodes(h)

    def _take(self, mh):
        """fac_matches handle -> (ctypes array of fac_match, stats).  Large results are exposed in place
        (the array aliases the library's pinned buffer and frees the handle when it is collected), small
        ones are copied; either way the caller sees an ordinary `fac_match * n` array."""
        n = self.lib.fac_matches_len(mh)
        data = self.lib.fac_matches_data(mh)
        stats = {"states_pushed": int(self.lib.fac_matches_states_pushed(mh)),
                 "device_ms": float(self.lib.fac_matches_device_ms(mh)),
                 "expand_ms": float(self.lib.fac_matches_expand_ms(mh)),
                 "kernel_launches": int(self.lib.fac_matches_kernel_launches(mh))}
        if n >= (1 << 15):
            arr = (fac_match * n).from_address(C.cast(data, C.c_void_p).value)
            arr._owner = _MatchesOwner(self.lib, mh)
            return arr, stats
        arr = (fac_match * n)()
        if n:
            C.memmove(arr, data, n * C.sizeof(fac_match))
        self.lib.fac_matches_free(mh)
        return arr, stats

    def search(self, h, data, thr, order, overlap, use_prefilter):
        mh = C.c_void_p()
        buf = (C.c_char * len(data)).from_buffer_copy(data) if len(data) else None
        st = self.lib.fac_search(h, buf, len(data), thr, order, overlap, int(use_prefilter), C.byref(mh))
        if st != 0:
            raise self._err(st)
        return self._take(mh)

    def search_host_ptr(self, h, ptr, n, thr, order, overlap, use_prefilter):
        """fac_search on a caller-owned host buffer (e.g. pinned memory) given by address."""
        mh = C.c_void_p()
        st = self.lib.fac_search(h, C.c_void_p(ptr), n, thr, order, overlap, int(use_prefilter), C.byref(mh))
        if st != 0:
            raise self._err(st)
        return self._take(mh)

    def search_device(self, h, dptr, n, thr, order, overlap, use_prefilter):
        mh = C.c_void_p()
        st = self.lib.fac_search_device(h, C.c_void_p(dptr), n, thr, order, overlap, int(use_prefilter), C.byref(mh))
        if st != 0:
            raise self._err(st)
        return self._take(mh)

    def search_shard(self, h, data_or_ptr, n, own_begin, own_end, base, thr, on_device):
        mh = C.c_void_p()
        if on_device:
            p = C.c_void_p(data_or_ptr)
        else:
            p = (C.c_char * n).from_buffer_copy(data_or_ptr) if n else None
        st = self.lib.fac_search_shard(h, p, n, own_begin, own_end, base, thr, int(on_device), C.byref(mh))
        if st != 0:
            raise self._err(st)
        return self._take(mh)

    def _stats(self, mh):
        return {"states_pushed": int(self.lib.fac_matches_states_pushed(mh)),
                "device_ms": float(self.lib.fac_matches_device_ms(mh)),
                "expand_ms": float(self.lib.fac_matches_expand_ms(mh)),
                "kernel_launches": int(self.lib.fac_matches_kernel_launches(mh))}

    def search_ex(self, h, ptr, n, own_begin, own_end, base, thr, order, overlap, flags, use_prefilter=False):
        """fac_search_ex on a raw host / device address.  With FAC_RESULT_ON_DEVICE returns a DeviceMatches."""
        a = _abi.fac_search_args(C.c_void_p(ptr), n, own_begin, own_end, base, thr, order, overlap, int(use_prefilter), flags)
        mh = C.c_void_p()
        st = self.lib.fac_search_ex(h, C.byref(a), C.byref(mh))
        if st != 0:
            raise self._err(st)
        if flags & _abi.FAC_RESULT_ON_DEVICE:
            return DeviceMatches(self.lib, mh), self._stats(mh)
        return self._take(mh)

    def apply_device(self, h, dptr, n, order, overlap, flags):
        """fac_matches_apply_device: device-resident fac_match records in, final list out (host array, or a
        DeviceMatches with FAC_RESULT_ON_DEVICE)."""
        mh = C.c_void_p()
        st = self.lib.fac_matches_apply_device(h, C.c_void_p(dptr), n, order, overlap, flags, C.byref(mh))
        if st != 0:
            raise self._err(st)
        if flags & _abi.FAC_RESULT_ON_DEVICE:
            return DeviceMatches(self.lib, mh), self._stats(mh)
        return self._take(mh)

    def apply(self, h, arr, n, order, overlap):
        mh = C.c_void_p()
        st = self.lib.fac_matches_apply(h, arr, n, order, overlap, C.byref(mh))
        if st != 0:
            raise self._err(st)
        return self._take(mh)

    def search_windows(self, h, windows, thr):
        """windows: list of (bytes, base, commit)."""
        n = len(windows)
        arr = (_abi.fac_window * n)()
        keep = []
        for i, (data, base, commit) in enumerate(windows):
            b = C.create_string_buffer(data, len(data))
            keep.append(b)
            arr[i].text = C.cast(b, C.c_void_p)
            arr[i].len = len(data)
            arr[i].base = base
            arr[i].commit = commit
        mh = C.c_void_p()
        st = self.lib.fac_search_windows(h, arr, n, thr, C.byref(mh))
        if st != 0:
            raise self._err(st)
        return self._take(mh)

    def search_stream(self, h, reader, thr, on_match):
        err = []

        def rd(_u, buf, cap):
            try:
                b = reader.read(cap)
                if b is None:
                    return 0
                C.memmove(buf, b, len(b))
                return len(b)
            except Exception as e:  # io::Error
                err.append(e)
                return -1

        def cb(_u, pm):
            on_match(pm.contents)

        total = C.c_uint64(0)
        st = self.lib.fac_search_stream(h, _abi.READ_FN(rd), None, thr, _abi.MATCH_FN(cb), None, C.byref(total))
        if err:
            raise err[0]
        if st != 0:
            raise self._err(st)
        return total.value

    def replace_stream(self, h, reader, writer, thr, callback):
        """callback(fac_match, matched_text_bytes) -> str | None."""
        err = []
        keep = []

        def rd(_u, buf, cap):
            try:
                b = reader.read(cap)
                if b is None:
                    return 0
                C.memmove(buf, b, len(b))
                return len(b)
            except Exception as e:
                err.append(e)
                return -1

        def wr(_u, buf, n):
            try:
                writer.write(C.string_at(buf, n))
                return 0
            except Exception as e:
                err.append(e)
                return 1

        def rp(_u, pm, base, text, n, out_p, out_n):
            r = callback(pm.contents, C.string_at(text, n))
            if r is None:
                return 0
            b = C.create_string_buffer(r.encode("utf-8") if isinstance(r, str) else bytes(r))
            keep[:] = [b]  # must outlive the return
            out_p[0] = C.cast(b, C.c_void_p).value
            out_n[0] = len(b) - 1
            return 1

        total = C.c_uint64(0)
        st = self.lib.fac_replace_stream(h, _abi.READ_FN(rd), None, _abi.WRITE_FN(wr), None, thr,
                                         _abi.REPLACE_FN(rp), None, C.byref(total))
        if err:
            raise err[0]
        if st != 0:
            raise self._err(st)
        return total.value


# ---------------------------------------------------------------------------------------------
# Builder / engine
# ---------------------------------------------------------------------------------------------
class FuzzyAhoCorasickBuilder:
    """src/builder.rs:22-184."""

    def __init__(self, backend=None):
        self._backend = backend
        self._similarity = None
        self._limits = None
        self._penalties = FuzzyPenalties()
        self._ci = False
        self._beam = None
        self._auto_beam = None
        self._mappings = []
        self._min_sym = 0.0
        self._device = None

    @staticmethod
    def new(backend=None):
        return FuzzyAhoCorasickBuilder(backend)

    def similarity(self, pairs):
        """pairs: {(a, b): score} with single-character strings (Similarity::from_map)."""
        self._similarity = dict(pairs)
        return self

    def fuzzy(self, limits):
        self._limits = limits
        return self

    def penalties(self, p):
        self._penalties = p
        return self

    def case_insensitive(self, v):
        self._ci = bool(v)
        return self

    def beam_width(self, w):
        self._beam = int(w)
        return self

    def auto_beam(self, budget, width):
        self._auto_beam = (int(budget), int(width))
        return self

    def mapping(self, a, b):
        return self.mapping_scored(a, b, 1.0)

    def mapping_scored(self, a, b, score):
        self._mappings.append((a, b, score))
        return self

    def min_symbol_similarity(self, m):
        self._min_sym = m
        return self

    def device(self, index):
        """B200 extension: CUDA device to place the automaton on (fac_engine_create_on), or a list of devices
        (fac_engine_create_multi: one replica per device, stream windows are dealt over them)."""
        self._device = index
        return self

    def build(self, inputs):
        pats = [Pattern.from_(x) for x in inputs]
        backend = self._backend or GpuBackend()
        cfg = fac_config()
        cfg.case_insensitive = int(self._ci)
        if self._limits is not None:
            cfg.has_limits = 1
            cfg.limits = self._limits.to_c()
        p = self._penalties
        if any(v is not None for v in (p.insertion_, p.deletion_, p.substitution_, p.swap_)):
            # unset fields keep the f32 defaults of FuzzyPenalties::default()
            m = _f32(1.3)
            d_sub, d_ins = _f32(_f32(1.1) * m), _f32(_f32(0.4) * m)
            d_del, d_swap = _f32(_f32(0.7) * m), _f32(_f32(0.4) * m)
            cfg.has_penalties = 1
            cfg.penalty_insertion = d_ins if p.insertion_ is None else p.insertion_
            cfg.penalty_deletion = d_del if p.deletion_ is None else p.deletion_
            cfg.penalty_substitution = d_sub if p.substitution_ is None else p.substitution_
            cfg.penalty_swap = d_swap if p.swap_ is None else p.swap_
        cfg.beam_width = self._beam or 0
        if self._auto_beam is not None:
            cfg.has_auto_beam = 1
            cfg.auto_beam_budget, cfg.auto_beam_width = self._auto_beam
        cfg.min_symbol_similarity = self._min_sym
        keep = []
        if self._similarity is not None:
            items = list(self._similarity.items())
            arr = (fac_sim_pair * max(1, len(items)))()
            for i, ((a, b), s) in enumerate(items):
                arr[i] = fac_sim_pair(ord(a), ord(b), s)
            cfg.has_similarity = 1
            cfg.similarity = arr
            cfg.n_similarity = len(items)
            keep.append(arr)
        if self._mappings:
            marr = (fac_mapping * len(self._mappings))()
            for i, (a, b, s) in enumerate(self._mappings):
                ab, bb = a.encode("utf-8"), b.encode("utf-8")
                keep += [ab, bb]
                marr[i] = fac_mapping(ab, len(ab), bb, len(bb), s)
            cfg.mappings = marr
            cfg.n_mappings = len(self._mappings)
            keep.append(marr)
        parr = (fac_pattern * max(1, len(pats)))()
        for i, pt in enumerate(pats):
            tb = pt.pattern.encode("utf-8")
            keep.append(tb)
            parr[i].text = tb
            parr[i].len = len(tb)
            parr[i].weight = pt.weight_
            if pt.limits is not None:
                parr[i].has_limits = 1
                parr[i].limits = pt.limits.to_c()
            parr[i].unique_id = -1 if pt.custom_unique_id_ is None else pt.custom_unique_id_
        if self._device is not None and hasattr(backend, "lib"):
            handle = backend.create(cfg, parr, len(pats), self._device)
        else:
            handle = backend.create(cfg, parr, len(pats))
        eng = FuzzyAhoCorasick(backend, handle, pats)
        eng._has_auto_beam = self._auto_beam is not None
        return eng

    def build_replacer(self, pairs):
        pairs = list(pairs)
        engine = self.build([p for p, _ in pairs])
        return FuzzyReplacer(engine, [r for _, r in pairs])


class FuzzyAhoCorasick:
    """src/structs.rs:528-567 (handle to the device-resident automaton)."""

    def __init__(self, backend, handle, patterns):
        self._b, self._h, self._patterns = backend, handle, patterns

    def __del__(self):
        try:
            if self._h is not None:
                self._b.free(self._h)
                self._h = None
        except Exception:
            pass

    def patterns(self):
        return self._patterns

    def num_nodes(self):
        return self._b.num_nodes(self._h)

    def max_match_graphemes(self):  # stream.rs:213
        return self._b.max_match_graphemes(self._h)

    # -- result construction -------------------------------------------------------------
    def _wrap(self, data, arr, stats):
        inner = []
        for c in arr:
            m = FuzzyMatch()
            m.insertions, m.deletions, m.substitutions, m.swaps, m.edits = (c.insertions, c.deletions,
                                                                            c.substitutions, c.swaps, c.edits)
            m.pattern_index = c.pattern_index
            m.pattern = self._patterns[c.pattern_index]
            m.start, m.end, m.similarity = c.start, c.end, c.similarity
            m.text = data[c.start:c.end].decode("utf-8")
            inner.append(m)
        return FuzzyMatches(data, inner, stats)

    @staticmethod
    def _bytes(haystack):
        return haystack.encode("utf-8") if isinstance(haystack, str) else bytes(haystack)

    def _search(self, haystack, thr, order, overlap, use_prefilter=False):
        data = self._bytes(haystack)
        arr, stats = self._b.search(self._h, data, thr, order, overlap, use_prefilter)
        return self._wrap(data, arr, stats)

    # -- query.rs ------------------------------------------------------------------------
    def search(self, haystack, opts=None):
        opts = opts or SearchOptions()
        return self._search(haystack, opts.threshold_, opts.order_, opts.overlap_)

    def segmented(self, haystack, opts=None, use_prefilter=False):  # query.rs:46-64
        opts = opts or SearchOptions()
        order = Order.Default if opts.order_ == Order.Unsorted else opts.order_
        overlap = Overlap.NonOverlapping if opts.overlap_ == Overlap.Keep else opts.overlap_
        return self._search(haystack, opts.threshold_, order, overlap, use_prefilter)

    def replace(self, text, opts, callback):
        return self.segmented(text, opts).replace(callback)

    def strip_prefix(self, haystack, opts=None):
        return self.segmented(haystack, opts).strip_prefix()

    def strip_suffix(self, haystack, opts=None):
        return self.segmented(haystack, opts).strip_suffix()

    def split(self, haystack, opts=None):
        return self.segmented(haystack, opts).split()

    def segment_iter(self, haystack, opts=None):
        return self.segmented(haystack, opts).segment_iter()

    def segment_text(self, haystack, opts=None):
        return self.segmented(haystack, opts).segment_text()

    def with_prefilter(self):  # prefilter.rs:113
        return Prefiltered(self)

    # -- stream.rs -----------------------------------------------------------------------
    def _stream_match(self, c):
        m = StreamMatch()
        m.start, m.end, m.pattern_index, m.similarity = c.start, c.end, c.pattern_index, c.similarity
        m.insertions, m.deletions, m.substitutions, m.swaps, m.edits = (c.insertions, c.deletions,
                                                                        c.substitutions, c.swaps, c.edits)
        m.text = None
        return m

    def search_stream(self, reader, threshold, on_match):
        """search_stream / search_stream_parallel (stream.rs:319-429): returns bytes read."""
        return self._b.search_stream(self._h, reader, threshold, lambda c: on_match(self._stream_match(c)))

    def search_stream_parallel(self, reader, threshold, threads, on_match):
        # the device batches windows itself; `threads` is accepted for interface parity
        return self.search_stream(reader, threshold, on_match)

    def stream_matches(self, reader, threshold):
        out = []
        self.search_stream(reader, threshold, out.append)
        return iter(out)

    def replace_stream(self, reader, writer, threshold, callback):
        """replace_stream / replace_stream_parallel (stream.rs:465-638); callback(StreamMatch) -> str | None."""
        def cb(c, text):
            m = self._stream_match(c)
            m.text = text.decode("utf-8")
            return callback(m)
        return self._b.replace_stream(self._h, reader, writer, threshold, cb)

    def replace_stream_parallel(self, reader, writer, threads, threshold, callback):
        return self.replace_stream(reader, writer, threshold, callback)


class Prefiltered:
    """src/prefilter.rs:58-156."""

    def __init__(self, engine):
        self.engine = engine

    def is_active(self):
        return self.engine._b.prefilter_active(self.engine._h)

    def search(self, haystack, opts=None):
        opts = opts or SearchOptions()
        return self.engine._search(haystack, opts.threshold_, opts.order_, opts.overlap_, use_prefilter=True)


class FuzzyReplacer:
    """src/replacer.rs."""

    def __init__(self, engine, replacements):
        self._engine, self.replacements = engine, list(replacements)

    def engine(self):
        return self._engine

    def replace(self, text, opts=None):
        reps = self.replacements
        return self._engine.replace(text, opts, lambda m: reps[m.pattern_index] if m.pattern_index < len(reps) else None)

    def replace_stream(self, reader, writer, threshold):
        reps = self.replacements
        return self._engine.replace_stream(
            reader, writer, threshold, lambda m: reps[m.pattern_index] if m.pattern_index < len(reps) else None)
