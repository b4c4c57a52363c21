// fac_bitap.cuh -- K2: the bitap pre-filter of src/prefilter.rs on the device.
//
// The scan runs over a stream of one byte per grapheme: the haystack bytes themselves for an ASCII haystack
// (Offsets::Identity, prefilter.rs:253-260) or the symbol-id stream K1 emits for a non-ASCII one (`transcode`,
// prefilter.rs:262-281: id of the folded grapheme in the filter's symbol table, 0 = other).
//
// BitapFilter::search_unsorted (prefilter.rs:304-374) scans the haystack once PER PATTERN with a
// Wu-Manber shift-AND automaton of k+1 u64 rows (bitap_windows, :410-435), pushes the candidate
// window [end-(m+k), end] for every end position whose row k has bit m-1, merges overlapping /
// touching windows (:335-342) and runs the engine on every merged slice as its own haystack.
//
//   k_bitap_scan   one warp per text sub-chunk, one LANE per pattern: the 32 lanes read the same
//                  haystack byte (one broadcast load, 16 bytes fetched per 128-bit load) and each
//                  advances its own pattern's rows held in registers; the byte -> mask table of the
//                  block's 32 patterns is staged in shared memory transposed ([byte][lane]) so the
//                  per-lane mask fetch is conflict-free.  A sub-chunk is warmed up over the
//                  preceding max(m+k) symbols: row bits only depend on the last m+k symbols, and
//                  the "deleted prefix is free" start state (1<<d)-1 is re-established by the
//                  recurrence at every position, so the state inside the sub-chunk equals the
//                  reference's sequential scan.
//                  Hits set the window's unit segments [s, e) in a coverage bitmap (atomicOr).
//   k_cov_edges    per 32-bit word of the bitmap: number of run starts / run ends
//   k_cov_emit     ordered compaction of run starts and run ends: two windows merge in the
//                  reference iff s2 <= e1, i.e. iff their unit segments form one run, so the
//                  k-th run == the k-th merged slice (gs, ge).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define BITAP_SUB 2048u      // graphemes owned by one warp
#define BITAP_WARPS 8u

struct BitapParams {
    const uint8_t *text;          // ASCII haystack bytes, or the K1 symbol-id stream of a non-ASCII haystack
    uint32_t n;                   // graphemes (== stream bytes)
    const uint64_t *bytemask;     // [P][rows]: mask of stream byte b for pattern p, prefilter.rs:210-231
    uint32_t rows;                // 128 (ASCII bytes, case folding baked in) or alphabet + 1 (symbol ids)
    const uint8_t *m;             // [P] pattern length in graphemes (1..63)
    const uint8_t *k;             // [P] edit budget of this call (k_for, prefilter.rs:285-302)
    uint32_t n_patterns;
    uint32_t warm;                // max over patterns of m + k
    uint32_t *cov;                // coverage bitmap over unit segments [g, g+1)
    unsigned long long *hits;     // statistics: number of hit positions
};

__device__ __forceinline__ void bitap_mark(uint32_t *cov, uint32_t s, uint32_t e) {  // set bits [s, e)
    uint32_t w0 = s >> 5, w1 = (e - 1u) >> 5;
    for (uint32_t w = w0; w <= w1; w++) {
        uint32_t lo = (w == w0) ? (s & 31u) : 0u, hi = (w == w1) ? ((e - 1u) & 31u) : 31u;
        const uint32_t mask = (0xFFFFFFFFu >> (31u - hi)) & (0xFFFFFFFFu << lo);
        atomicOr(&cov[w], mask);
    }
}

template <int KMAX>
__global__ void __launch_bounds__(BITAP_WARPS * 32) k_bitap_scan(const BitapParams P) {
    extern __shared__ __align__(16) uint64_t s_mask[];  // [rows][lane]
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t p = blockIdx.y * 32u + lane;
    const bool live = p < P.n_patterns;
    const uint32_t rows = P.rows;
    for (uint32_t i = threadIdx.x; i < rows * 32u; i += blockDim.x) {
        const uint32_t b = i >> 5, l = i & 31u, pp = blockIdx.y * 32u + l;
        s_mask[i] = pp < P.n_patterns ? P.bytemask[(size_t)pp * rows + b] : 0ull;
    }
    __syncthreads();
    const uint32_t m = live ? P.m[p] : 1u, k = live ? P.k[p] : 0u;
    const uint64_t match_bit = 1ull << (m - 1u);
    const uint32_t span = m + k;
    const uint64_t chunk = (uint64_t)blockIdx.x * BITAP_WARPS + warp;
    const uint64_t own0 = chunk * BITAP_SUB;
    if (own0 >= P.n) return;
    const uint32_t own1 = (uint32_t)min((uint64_t)P.n, own0 + BITAP_SUB);
    const uint32_t w0 = own0 > P.warm ? (uint32_t)own0 - P.warm : 0u;
    uint64_t r[KMAX + 1], nr[KMAX + 1];
#pragma unroll
    for (int d = 0; d <= KMAX; d++) r[d] = (1ull << d) - 1ull;  // prefilter.rs:416-418
    uint32_t n_hits = 0;
    auto step = [&](uint32_t c, uint32_t i) {
        const uint64_t bc = s_mask[min(c, rows - 1u) * 32u + lane];
        nr[0] = ((r[0] << 1) | 1ull) & bc;
        uint64_t rk = nr[0];
#pragma unroll
        for (int d = 1; d <= KMAX; d++) {
            nr[d] = ((r[d] << 1) & bc) | ((r[d - 1] | nr[d - 1]) << 1) | r[d - 1] | 1ull;  // prefilter.rs:422-428
            if ((uint32_t)d == k) rk = nr[d];
        }
#pragma unroll
        for (int d = 0; d <= KMAX; d++) r[d] = nr[d];
        if (live && i >= own0 && (rk & match_bit)) {  // prefilter.rs:430-433
            const uint32_t end = i + 1u;
            bitap_mark(P.cov, end > span ? end - span : 0u, end);
            n_hits++;
        }
    };
    // all lanes read the same haystack byte: 16 bytes per 128-bit load on the aligned body
    uint32_t i = w0;
    while (i < own1 && ((uintptr_t)(P.text + i) & 15u)) { step(P.text[i], i); i++; }
    while ((uint64_t)i + 16u <= own1) {
        const uint4 v = *reinterpret_cast<const uint4 *>(P.text + i);
        const uint32_t wd[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 16; q++) step((wd[q >> 2] >> ((q & 3) * 8)) & 0xFFu, i + q);
        i += 16u;
    }
    while (i < own1) { step(P.text[i], i); i++; }
    n_hits = __reduce_add_sync(0xFFFFFFFFu, n_hits);
    if (lane == 0 && n_hits) atomicAdd(P.hits, (unsigned long long)n_hits);
}

// run starts / ends per bitmap word; bit g = unit segment [g, g+1)
__device__ __forceinline__ void cov_edges(const uint32_t *cov, uint32_t n_words, uint32_t w, uint32_t &starts, uint32_t &ends) {
    const uint32_t v = cov[w];
    const uint32_t prev = w ? cov[w - 1] >> 31 : 0u;
    const uint32_t next = (w + 1 < n_words) ? (cov[w + 1] & 1u) : 0u;
    starts = v & ~((v << 1) | prev);
    ends = v & ~((v >> 1) | (next << 31));
}
__global__ void __launch_bounds__(256) k_cov_edges(const uint32_t *cov, uint32_t n_words, uint32_t *cnt) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    uint32_t s, e;
    cov_edges(cov, n_words, w, s, e);
    cnt[w] = __popc(s) | (__popc(e) << 16);   // both counts per word are <= 16
}
// off[w] = exclusive prefix over words of (starts | ends << 32) packed as u64
__global__ void __launch_bounds__(256) k_cov_emit(const uint32_t *cov, uint32_t n_words, const unsigned long long *off, uint32_t *gs, uint32_t *ge) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    uint32_t s, e;
    cov_edges(cov, n_words, w, s, e);
    const unsigned long long o = off[w];
    uint32_t os = (uint32_t)(o & 0xFFFFFFFFull), oe = (uint32_t)(o >> 32);
    while (s) { const uint32_t b = __ffs(s) - 1u; s &= s - 1u; gs[os++] = w * 32u + b; }
    while (e) { const uint32_t b = __ffs(e) - 1u; e &= e - 1u; ge[oe++] = w * 32u + b + 1u; }
}
struct CovCountToU64 {
    __host__ __device__ __forceinline__ unsigned long long operator()(const uint32_t &c) const {
        return (unsigned long long)(c & 0xFFFFu) | ((unsigned long long)(c >> 16) << 32);
    }
};

// Per merged slice of a non-ASCII haystack: bit 0 = the slice holds a non-ASCII byte, bit 1 = it holds a "\r\n" pair.
// The reference searches every slice as its own haystack and re-tests `is_ascii` per slice (prefilter.rs:346-350,
// search.rs:196): an all-ASCII slice uses the byte-per-grapheme storage, where CR LF are two graphemes instead of the
// one cluster of UAX #29 -- the only case in which the two storages segment ASCII text differently.
__global__ void __launch_bounds__(256) k_slice_classify(const uint8_t *__restrict__ text, const uint32_t *__restrict__ off32, const uint64_t *__restrict__ off64,
                                                        const uint32_t *__restrict__ gs, const uint32_t *__restrict__ ge, uint32_t n_slices,
                                                        uint32_t *__restrict__ flags) {
    const uint32_t sl = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (sl >= n_slices) return;
    const uint64_t b0 = off64 ? off64[gs[sl]] : (uint64_t)off32[gs[sl]], b1 = off64 ? off64[ge[sl]] : (uint64_t)off32[ge[sl]];
    uint32_t f = 0;
    for (uint64_t i = b0 + lane; i < b1; i += 32u) {
        const uint8_t c = text[i];
        if (c & 0x80u) f |= 1u;
        if (c == '\r' && i + 1 < b1 && text[i + 1] == '\n') f |= 2u;
    }
    f = __reduce_or_sync(0xFFFFFFFFu, f);
    if (lane == 0) flags[sl] = f;
}
