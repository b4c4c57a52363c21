// fac_segment.cuh -- K1: UTF-8 -> grapheme streams on the device.
//
// Replaces build_unicode_graphemes (src/search.rs:398-416: `grapheme_indices(true)` + per-grapheme
// `to_lowercase`) and the `text_chars` vector (src/search.rs:302) for non-ASCII haystacks, and the
// pre-filter's transcode (src/prefilter.rs:251-281).  Every scalar start decides its own cluster
// boundary with the look-back predicate of fac_unicode.h (no sequential state), a prefix sum turns
// the boundary marks into grapheme indices, and one thread per cluster folds it and emits
//   first[g]  first char of the folded grapheme (u32)
//   gid[g]    id of the folded grapheme in the engine's symbol table (engines with mappings)
//   off[g]    byte offset (+ sentinel off[n] = len)
//   sym[g]    pre-filter symbol id (u8, 0 = other) when requested
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fac_types.h"
#include "fac_unicode.h"

// Structural UTF-8 validation (what `str::from_utf8` accepts), one thread per byte: a lead byte
// checks its own sequence, a continuation byte checks that a lead within 3 bytes covers it.
__global__ void __launch_bounds__(256) k_validate_utf8(const uint8_t *__restrict__ s, uint64_t n, uint32_t *bad) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t b = s[i];
    bool ok = true;
    if (b < 0x80) ok = true;
    else if ((b & 0xC0) == 0x80) {
        // must be the k-th continuation (k = 1..3) of a lead that needs >= k continuations
        ok = false;
        for (uint32_t k = 1; k <= 3 && k <= i; k++) {
            const uint8_t l = s[i - k];
            if ((l & 0xC0) == 0x80) continue;
            const uint32_t need = l >= 0xF0 ? 3u : (l >= 0xE0 ? 2u : (l >= 0xC0 ? 1u : 0u));
            ok = need >= k;
            break;
        }
    } else {
        uint32_t extra;
        uint8_t lo = 0x80, hi = 0xBF;
        if (b >= 0xC2 && b <= 0xDF) extra = 1;
        else if (b == 0xE0) { extra = 2; lo = 0xA0; }
        else if (b == 0xED) { extra = 2; hi = 0x9F; }
        else if (b >= 0xE1 && b <= 0xEF) extra = 2;
        else if (b == 0xF0) { extra = 3; lo = 0x90; }
        else if (b >= 0xF1 && b <= 0xF3) extra = 3;
        else if (b == 0xF4) { extra = 3; hi = 0x8F; }
        else { extra = 0; ok = false; }
        if (ok) {
            if (n - i <= extra) ok = false;
            else {
                if (s[i + 1] < lo || s[i + 1] > hi) ok = false;
                for (uint32_t k = 2; k <= extra && ok; k++)
                    if ((s[i + k] & 0xC0) != 0x80) ok = false;
            }
        }
    }
    if (!ok) atomicOr(bad, 1u);
}

// mark[i] = 1 iff an extended grapheme cluster starts at byte i.  `seg_lo` (optional) gives, per
// 4 KiB block of the text, nothing: segmentation restarts only at `lo` (the start of the text or
// of the stream window that contains byte i).
__global__ void __launch_bounds__(256) k_seg_mark(const uint8_t *__restrict__ s, uint64_t lo, uint64_t hi, UnicodeTables U,
                                                  uint8_t *__restrict__ mark) {
    const uint64_t i = lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi) return;
    uint8_t m = 0;
    if (!fac_is_cont(s[i])) m = fac_break_before(U, s, lo, i) ? 1 : 0;
    mark[i] = m;
}

struct SegEmitParams {
    const uint8_t *s;
    uint64_t lo, hi;            // byte range [lo, hi) being segmented
    const uint8_t *mark;        // per byte
    const uint32_t *gidx;       // exclusive prefix sum of mark (per byte), relative to lo
    uint32_t g_base;            // grapheme index of the first cluster of this range
    UnicodeTables U;
    const FacSymbol *symbols;   // engine symbol table (mappings) or nullptr
    uint32_t sym_mask;
    const uint8_t *pool;
    const FacSymbol *pf_symbols;  // pre-filter symbol table or nullptr
    uint32_t pf_mask;
    const uint8_t *pf_pool;
    int fold;                   // case_insensitive
    uint32_t *first, *gid;
    uint32_t *off32;
    uint64_t *off64;
    uint8_t *pf_sym;
};

__global__ void __launch_bounds__(256) k_seg_emit(const SegEmitParams P) {
    const uint64_t i = P.lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.hi || !P.mark[i]) return;
    const uint32_t g = P.g_base + P.gidx[i - P.lo];
    uint64_t e = i + 1;
    while (e < P.hi && !P.mark[e]) e++;
    if (P.off32) P.off32[g] = (uint32_t)i;
    if (P.off64) P.off64[g] = i;
    uint32_t fc = 0, flen;
    if (P.symbols) {
        P.gid[g] = fac_symbol_lookup(P.U, P.symbols, P.sym_mask, P.pool, P.s, i, e, P.fold != 0, fc);
    } else {
        fac_grapheme_hash(P.U, P.s, i, e, P.fold != 0, fc, flen);
    }
    P.first[g] = fc;
    if (P.pf_symbols) {
        uint32_t fc2;
        P.pf_sym[g] = (uint8_t)fac_symbol_lookup(P.U, P.pf_symbols, P.pf_mask, P.pf_pool, P.s, i, e, P.fold != 0, fc2);
    }
}
