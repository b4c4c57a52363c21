// fac_apply.cuh -- K4: ranking and overlap resolution on the device
// (FuzzyMatches::apply, src/matches.rs:7-149).
//
//   ranking      : the three total orders of matches.rs:24-84 as comparators for a device merge sort
//                  (window id is the leading key, so a batch of stream windows ranks in one call)
//   non_overlapping (matches.rs:86-112): the rank-ordered greedy interval selection equals the
//                  lexicographically-first maximal independent set of the conflict graph.  It is
//                  computed without a sequential sweep: matches are laid out by start offset, every
//                  undecided match inspects only its overlapping neighbours, accepts itself once all
//                  better-ranked conflicting neighbours are rejected and rejects itself as soon as one
//                  is accepted; rounds repeat until nothing is undecided (k_overlap_round).
//   non_overlapping_unique (matches.rs:116-149): the one-per-pattern-identity constraint couples all
//                  matches of a window, so it is replayed in rank order by one warp per window
//                  (k_unique_select), with the neighbour scan spread over the lanes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fac_core.h"
#include "fac_types.h"

// Order::{Unsorted, Default, Greedy, CoverageWeighted}; Unsorted is the canonical
// (start, end, pattern) order (unpinned item U2).
struct RankLess {
    int order;
    const uint32_t *pat_bytes;
    __device__ __forceinline__ bool tail(const WMatch &l, const WMatch &r) const {
        if (l.start != r.start) return l.start < r.start;
        if (l.end != r.end) return l.end < r.end;
        return l.pat < r.pat;
    }
    __device__ __forceinline__ bool operator()(const WMatch &l, const WMatch &r) const {
        if (l.win != r.win) return l.win < r.win;
        if (order == 0) return tail(l, r);
        const uint32_t ls = fac_total_order_u32(l.sim), rs = fac_total_order_u32(r.sim);
        const uint32_t lp = pat_bytes[l.pat], rp = pat_bytes[r.pat];
        if (order == 1) {  // default_sort, matches.rs:24-41
            if (ls != rs) return ls > rs;
            if (lp != rp) return lp > rp;
            const uint64_t lt = l.end - l.start, rt = r.end - r.start;
            if (lt != rt) return lt > rt;
            return tail(l, r);
        }
        if (order == 2) {  // greedy_sort, matches.rs:46-61
            if (lp != rp) return lp > rp;
            if (ls != rs) return ls > rs;
            return tail(l, r);
        }
        // coverage_weighted_sort, matches.rs:67-84: similarity * similarity * len as f32
        const uint32_t lc = fac_total_order_u32(__fmul_rn(__fmul_rn(l.sim, l.sim), (float)lp));
        const uint32_t rc = fac_total_order_u32(__fmul_rn(__fmul_rn(r.sim, r.sim), (float)rp));
        if (lc != rc) return lc > rc;
        if (ls != rs) return ls > rs;
        return tail(l, r);
    }
};

// Position order used by the overlap kernels: (window, start, rank).  Elements are ranks (indices
// into the ranked array R).
struct PosLess {
    const WMatch *R;
    __device__ __forceinline__ bool operator()(const uint32_t &a, const uint32_t &b) const {
        const WMatch &x = R[a], &y = R[b];
        if (x.win != y.win) return x.win < y.win;
        if (x.start != y.start) return x.start < y.start;
        return a < b;
    }
};

struct WinEnd {
    uint32_t win;
    uint32_t pad;
    uint64_t end;
};
struct WinEndMax {  // segmented running max of `end` (associative)
    __device__ __forceinline__ WinEnd operator()(const WinEnd &a, const WinEnd &b) const {
        if (a.win != b.win) return b;
        WinEnd r = b;
        r.end = a.end > b.end ? a.end : b.end;
        return r;
    }
};

__global__ void k_iota(uint32_t *p, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}
__global__ void k_gather_winend(const WMatch *R, const uint32_t *P, WinEnd *out, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const WMatch m = R[P[i]]; WinEnd w; w.win = m.win; w.pad = 0; w.end = m.end; out[i] = w; }
}

// The reference's acceptance test of a later-ranked match m against an accepted interval o
// (matches.rs:93-100): with `occupied` sorted and disjoint, checking the two neighbours of the
// insertion point equals this pairwise predicate (asymmetric only for empty spans).
__device__ __forceinline__ bool fac_conflict(uint64_t ms, uint64_t me, uint64_t os, uint64_t oe) {
    return (os < ms && oe > ms) || (os >= ms && os < me);
}

enum : uint8_t { OV_UNDECIDED = 0, OV_ACCEPTED = 1, OV_REJECTED = 2 };

// One round of the parallel greedy.  P: ranks in position order; pmax: running max end in position
// order; st_in / st_out: decision per position (double buffered so a round reads a consistent
// snapshot).  *pending is set when something stays undecided.
__global__ void __launch_bounds__(256) k_overlap_round(const WMatch *__restrict__ R, const uint32_t *__restrict__ P,
                                                       const WinEnd *__restrict__ pmax, const uint8_t *__restrict__ st_in,
                                                       uint8_t *__restrict__ st_out, uint32_t n, uint32_t *pending) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint8_t cur = st_in[p];
    if (cur != OV_UNDECIDED) { st_out[p] = cur; return; }
    const uint32_t my_rank = P[p];
    const WMatch m = R[my_rank];
    bool any_accepted = false, any_undecided = false;
    // neighbours to the left: start <= m.start; stop once nothing further left can reach m.start
    for (uint32_t q = p; q-- > 0;) {
        const WinEnd pm = pmax[q];
        if (pm.win != m.win) break;
        const uint32_t orank = P[q];
        const WMatch o = R[orank];
        if (!(pm.end > m.start || o.start == m.start)) break;
        if (orank < my_rank && fac_conflict(m.start, m.end, o.start, o.end)) {
            const uint8_t s = st_in[q];
            if (s == OV_ACCEPTED) { any_accepted = true; break; }
            if (s == OV_UNDECIDED) any_undecided = true;
        }
    }
    if (!any_accepted) {
        for (uint32_t q = p + 1; q < n; q++) {
            const uint32_t orank = P[q];
            const WMatch o = R[orank];
            // right neighbours start at or after m.start; equal starts with a better rank sit on the
            // left (position order breaks start ties by rank), so only o.start < m.end can conflict
            if (o.win != m.win || o.start >= m.end) break;
            if (orank < my_rank && fac_conflict(m.start, m.end, o.start, o.end)) {
                const uint8_t s = st_in[q];
                if (s == OV_ACCEPTED) { any_accepted = true; break; }
                if (s == OV_UNDECIDED) any_undecided = true;
            }
        }
    }
    uint8_t res = OV_UNDECIDED;
    if (any_accepted) res = OV_REJECTED;
    else if (!any_undecided) res = OV_ACCEPTED;
    else atomicOr(pending, 1u);
    st_out[p] = res;
}

// non_overlapping_unique: one warp per window replays the ranked list.  `win_rank_off[w]` is the
// first rank of window w in R (R is ranked with the window id as leading key); posof[rank] is the
// position of that match in P.  used[] is a bitmap over dense pattern identities, one stripe per
// warp-resident window (cleared by the kernel after use).
__global__ void __launch_bounds__(32) k_unique_select(const WMatch *__restrict__ R, const uint32_t *__restrict__ P,
                                                      const uint32_t *__restrict__ posof, const WinEnd *__restrict__ pmax,
                                                      const uint32_t *__restrict__ win_rank_off, uint32_t n_windows, uint32_t n,
                                                      const uint32_t *__restrict__ pat_uid_dense, uint32_t uid_words,
                                                      uint32_t *__restrict__ used, uint8_t *__restrict__ st) {
    const uint32_t lane = threadIdx.x;
    for (uint32_t w = blockIdx.x; w < n_windows; w += gridDim.x) {
        uint32_t *my_used = used + (size_t)blockIdx.x * uid_words;
        const uint32_t r0 = win_rank_off[w], r1 = win_rank_off[w + 1];
        for (uint32_t r = r0; r < r1; r++) {
            const WMatch m = R[r];
            const uint32_t uid = pat_uid_dense[m.pat];
            const bool was_used = (my_used[uid >> 5] >> (uid & 31u)) & 1u;
            bool conflict = false;
            const uint32_t p = posof[r];
            if (!was_used) {
                // left neighbours, strided over the lanes in chunks of 32
                bool done = false;
                for (uint32_t base = 0; !done && !conflict; base += 32) {
                    const uint32_t d = base + lane + 1;
                    bool lane_stop = true, lane_conf = false;
                    if (d <= p) {
                        const uint32_t q = p - d;
                        const WinEnd pm = pmax[q];
                        const WMatch o = R[P[q]];
                        if (pm.win == m.win && (pm.end > m.start || o.start == m.start)) {
                            lane_stop = false;
                            lane_conf = st[q] == OV_ACCEPTED && fac_conflict(m.start, m.end, o.start, o.end);
                        }
                    }
                    conflict = __any_sync(0xFFFFFFFFu, lane_conf);
                    done = __any_sync(0xFFFFFFFFu, lane_stop);
                    // lanes beyond the first stopping lane may have looked too far left; their conflicts
                    // are still genuine intervals of the same window only if every nearer lane continued
                    if (done && conflict) {
                        const uint32_t stop_mask = __ballot_sync(0xFFFFFFFFu, lane_stop);
                        const uint32_t conf_mask = __ballot_sync(0xFFFFFFFFu, lane_conf);
                        const uint32_t first_stop = __ffs(stop_mask) - 1;
                        conflict = (conf_mask & ((1u << first_stop) - 1u)) != 0u;
                    }
                }
                for (uint32_t base = 0; !conflict; base += 32) {
                    const uint32_t q = p + 1 + base + lane;
                    bool lane_stop = true, lane_conf = false;
                    if (q < n) {
                        const WMatch o = R[P[q]];
                        if (o.win == m.win && o.start < m.end) {
                            lane_stop = false;
                            lane_conf = st[q] == OV_ACCEPTED && fac_conflict(m.start, m.end, o.start, o.end);
                        }
                    }
                    const uint32_t stop_mask = __ballot_sync(0xFFFFFFFFu, lane_stop);
                    const uint32_t conf_mask = __ballot_sync(0xFFFFFFFFu, lane_conf);
                    if (stop_mask) {
                        const uint32_t first_stop = __ffs(stop_mask) - 1;
                        conflict = (conf_mask & ((1u << first_stop) - 1u)) != 0u;
                        break;
                    }
                    conflict = conf_mask != 0u;
                }
            }
            const bool accept = !was_used && !conflict;
            if (lane == 0) {
                st[p] = accept ? OV_ACCEPTED : OV_REJECTED;
                if (accept) my_used[uid >> 5] |= 1u << (uid & 31u);
            }
            __syncwarp();
        }
        for (uint32_t k = lane; k < uid_words; k += 32) my_used[k] = 0;
        __syncwarp();
    }
}

__global__ void k_posof(const uint32_t *P, uint32_t *posof, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) posof[P[i]] = i;
}
// first rank of each window in the ranked array (window id is the leading sort key)
__global__ void k_win_rank_off(const WMatch *R, uint32_t n, uint32_t *off, uint32_t n_windows) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    const uint32_t w_here = i < n ? R[i].win : n_windows;
    const uint32_t w_prev = i > 0 ? R[i - 1].win + 1 : 0;
    for (uint32_t w = w_prev; w <= w_here && w <= n_windows; w++) off[w] = i;
}
__global__ void k_accept_flags(const uint8_t *st, uint8_t *flags, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = st[i] == OV_ACCEPTED;
}
// ---- Order::Unsorted as a radix sort: key = start | (end - start) | pattern packed into 64 bits ----
// bounds[0..3] = max start, max (end - start), max pattern, max window of the list (sizes the key fields)
__global__ void __launch_bounds__(256) k_key_bounds(const WMatch *__restrict__ m, uint32_t n, unsigned long long *__restrict__ bounds) {
    unsigned long long ms = 0, ml = 0;
    uint32_t mp = 0, mw = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint4 a = reinterpret_cast<const uint4 *>(m)[2 * (size_t)i], b = reinterpret_cast<const uint4 *>(m)[2 * (size_t)i + 1];
        const unsigned long long st = ((unsigned long long)a.y << 32) | a.x, en = ((unsigned long long)a.w << 32) | a.z;
        ms = max(ms, st); ml = max(ml, en - st); mp = max(mp, b.x); mw = max(mw, b.w);
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        ms = max(ms, __shfl_xor_sync(0xFFFFFFFFu, ms, d)); ml = max(ml, __shfl_xor_sync(0xFFFFFFFFu, ml, d));
        mp = max(mp, __shfl_xor_sync(0xFFFFFFFFu, mp, d)); mw = max(mw, __shfl_xor_sync(0xFFFFFFFFu, mw, d));
    }
    if ((threadIdx.x & 31u) == 0) {
        atomicMax(&bounds[0], ms); atomicMax(&bounds[1], ml);
        atomicMax(&bounds[2], (unsigned long long)mp); atomicMax(&bounds[3], (unsigned long long)mw);
    }
}
__global__ void __launch_bounds__(256) k_make_keys(const WMatch *__restrict__ m, uint32_t n, uint32_t len_shift, uint32_t start_shift,
                                                   unsigned long long *__restrict__ keys, uint32_t *__restrict__ idx) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 a = reinterpret_cast<const uint4 *>(m)[2 * (size_t)i];
    const uint32_t pat = reinterpret_cast<const uint32_t *>(m)[8 * (size_t)i + 4];
    const unsigned long long st = ((unsigned long long)a.y << 32) | a.x, en = ((unsigned long long)a.w << 32) | a.z;
    keys[i] = (st << start_shift) | ((en - st) << len_shift) | pat;
    idx[i] = i;
}
// 32-byte records moved as two 128-bit halves
__global__ void __launch_bounds__(256) k_gather_matches16(const WMatch *__restrict__ R, const uint32_t *__restrict__ sel, uint32_t n, WMatch *__restrict__ out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t i = t >> 1, h = t & 1u;
    if (i < n) reinterpret_cast<uint4 *>(out)[2 * (size_t)i + h] = reinterpret_cast<const uint4 *>(R)[2 * (size_t)sel[i] + h];
}

// kept matches in position order -> final records
__global__ void k_gather_matches(const WMatch *R, const uint32_t *sel, uint32_t n, WMatch *out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = R[sel[i]];
}

// WMatch -> fac_match with the window's base offset; flags the matches a stream window owns
// (start < commit, src/stream.rs:278-279).
__global__ void k_finalize(const WMatch *in, uint32_t n, const FacWindow *windows, fac_match *out, uint8_t *keep) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const WMatch m = in[i];
    const FacWindow w = windows[m.win];
    fac_match o;
    o.start = m.start + w.base; o.end = m.end + w.base; o.pattern_index = m.pat; o.similarity = m.sim;
    o.insertions = m.cnt & 0xFF; o.deletions = (m.cnt >> 8) & 0xFF; o.substitutions = (m.cnt >> 16) & 0xFF; o.swaps = m.cnt >> 24;
    o.edits = (uint8_t)fac_edits_of(m.cnt);
    o.pad_[0] = o.pad_[1] = o.pad_[2] = 0;
    out[i] = o;
    keep[i] = m.start < w.commit;
}

// fac_match (absolute offsets) -> WMatch of a single window with base 0: the inverse of k_finalize, for
// fac_matches_apply_device.  *bad is set when a pattern index is out of range.
__global__ void k_unfinalize(const fac_match *in, uint32_t n, uint32_t n_patterns, WMatch *out, uint32_t *bad) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const fac_match m = in[i];
    if (m.pattern_index >= n_patterns) atomicOr(bad, 1u);
    WMatch w;
    w.start = m.start; w.end = m.end; w.pat = m.pattern_index; w.sim = m.similarity; w.win = 0;
    w.cnt = (uint32_t)m.insertions | ((uint32_t)m.deletions << 8) | ((uint32_t)m.substitutions << 16) | ((uint32_t)m.swaps << 24);
    out[i] = w;
}
