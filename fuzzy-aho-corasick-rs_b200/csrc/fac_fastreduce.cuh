// fac_fastreduce.cuh -- reduction of the FAST kernel's candidates.
//
// Candidates of the FAST kernel carry no FIFO position, so the reference's tie-break ("first among
// equal similarities", src/search.rs:705-722) cannot be replayed from them.  It does not have to
// be: per (start, end, pattern) the maximum similarity is order-independent, and when
// every candidate that attains it carries the SAME edit counts the reference's record is
// determined without knowing who came first.  Keys whose maximum is attained with two different
// count vectors (~1.5 % of result keys) mark their start window dirty; dirty windows
// are searched again by the order-faithful kernel and take ALL their results from it.
//   k_fbest_max    : slot per key (representative-index hashing), atomicMax of the similarity
//   k_fbest_minmax : among candidates at the maximum: min / max of the packed counts, min index
//   k_fbest_mark   : min != max  -> set the window's bit in the dirty bitmap
//   k_fbest_emit   : winners (min index at the maximum) of clean windows -> WMatch records
//   k_dirty_tiles  : dirty bitmap -> one-window tile descriptors for the faithful pass
#pragma once
#include "fac_kernels.cuh"

struct FBestParams {
    BestParams B;
    uint32_t *tab_sim;    // [tab_size] total-order image of the best similarity (0 = none)
    uint32_t *tab_cmin;   // [tab_size]
    uint32_t *tab_cmax;   // [tab_size]
    uint32_t *tab_first;  // [tab_size]
    uint32_t *dirty;      // bitmap over start windows, bit (sg - dirty_base)
    uint32_t dirty_base;
};

__global__ void __launch_bounds__(256) k_fbest_max(const FBestParams F) {
    const BestParams &P = F.B;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_cands) return;
    const FacCand c = P.cands[i];
    if (fac_cand_dead(P, c)) { P.cslot[i] = FAC_EMPTY; return; }
    const uint32_t mask = P.tab_size - 1u;
    uint32_t h = fac_hash3(c.sg, c.eg, c.pat) & mask;
    for (;;) {
        uint32_t r = *((volatile uint32_t *)&P.tab_rep[h]);
        if (r == FAC_EMPTY) {
            const uint32_t old = atomicCAS(&P.tab_rep[h], FAC_EMPTY, i);
            r = (old == FAC_EMPTY) ? i : old;
        }
        if (r == i) break;
        const FacCand rc = P.cands[r];
        if (rc.sg == c.sg && rc.eg == c.eg && rc.pat == c.pat) break;
        h = (h + 1u) & mask;
    }
    P.cslot[i] = h;
    atomicMax(&F.tab_sim[h], fac_total_order_u32(c.sim));
}

__global__ void __launch_bounds__(256) k_fbest_minmax(const FBestParams F) {
    const BestParams &P = F.B;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_cands) return;
    const uint32_t h = P.cslot[i];
    if (h == FAC_EMPTY) return;
    const FacCand c = P.cands[i];
    if (F.tab_sim[h] != fac_total_order_u32(c.sim)) return;
    atomicMin(&F.tab_cmin[h], c.cnt);
    atomicMax(&F.tab_cmax[h], c.cnt);
    atomicMin(&F.tab_first[h], i);
}

__global__ void __launch_bounds__(256) k_fbest_mark(const FBestParams F) {
    const BestParams &P = F.B;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_cands) return;
    const uint32_t h = P.cslot[i];
    if (h == FAC_EMPTY || F.tab_first[h] != i) return;
    if (F.tab_cmin[h] != F.tab_cmax[h]) {
        const uint32_t w = P.cands[i].sg - F.dirty_base;
        atomicOr(&F.dirty[w >> 5], 1u << (w & 31u));
    }
}

__global__ void __launch_bounds__(256) k_fbest_emit(const FBestParams F) {
    const BestParams &P = F.B;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_cands) return;
    const uint32_t h = P.cslot[i];
    if (h == FAC_EMPTY || F.tab_first[h] != i) return;
    const FacCand c = P.cands[i];
    const uint32_t w = c.sg - F.dirty_base;
    if ((F.dirty[w >> 5] >> (w & 31u)) & 1u) return;  // the whole window is taken from the faithful pass
    const uint32_t win = c.tag & 0x7FFFFFFFu;
    const FacWindow wd = P.windows[win];
    const uint64_t sb = fac_byte_offset(P.tv, c.sg) - wd.byte_begin;
    const uint64_t eb = fac_byte_offset(P.tv, c.eg) - wd.byte_begin;
    const unsigned long long o = atomicAdd(P.out_count, 1ull);
    if (o >= P.out_cap) return;
    WMatch m;
    m.start = sb; m.end = eb; m.pat = c.pat; m.sim = c.sim; m.cnt = c.cnt; m.win = win;
    P.out[o] = m;
}

__global__ void __launch_bounds__(256) k_dirty_tiles(const uint32_t *dirty, uint32_t n_words, uint32_t dirty_base, uint32_t text_end,
                                                     uint4 *tiles, uint32_t cap, unsigned long long *count) {
    const uint32_t wi = blockIdx.x * blockDim.x + threadIdx.x;
    if (wi >= n_words) return;
    uint32_t bits = dirty[wi];
    while (bits) {
        const uint32_t b = __ffs(bits) - 1;
        bits &= bits - 1;
        const unsigned long long o = atomicAdd(count, 1ull);
        if (o < cap) tiles[o] = make_uint4(dirty_base + wi * 32u + b, 1u, text_end, 0u);
    }
}
