// fac_kernels.cuh -- hand-written sm_100a kernels of the fuzzy search path.
//
//   k_scan_bytes        haystack classification (is_ascii, src/search.rs:196) in one coalesced pass
//   k_expand            K3, generic / order-faithful variant: fuzzy frontier expansion, one CTA per tile of
//                       start windows (search_unsorted_impl, src/search.rs:533-1104).  Engines inside the
//                       fast kernel's domain run k_expand_succinct (fac_succinct.cuh) instead and come here
//                       only for the windows whose tie-break needs the reference's FIFO order.
//   k_best_insert/select  best-per-(start,end,pattern) reduction of the raw candidates
//                       (the `best` map, src/search.rs:705-735 + :1111-1118)
//
// No tensor cores: nothing on this path is a dense contraction.  The work is integer / f32 scalar
// graph expansion; the B200 levers are (i) keeping the frontier, the dedup table and the text tile
// on chip or L2-resident, (ii) order-preserving ballot/popc compaction so the frontier order equals
// the reference's FIFO order (tie-breaking is observable), (iii) 128-bit state loads /
// stores, (iv) a persistent grid sized to the SM count with dynamic tile fetch.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fac_core.h"
#include "fac_types.h"
#include "fac_unicode.h"

#ifndef FAC_BLOCK
#define FAC_BLOCK 256
#endif
#define FAC_NWARPS (FAC_BLOCK / 32)
#define FAC_EMPTY 0xFFFFFFFFu

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t fac_lane() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t fac_warp() { return threadIdx.x >> 5; }

// Block-wide exclusive scan of one u32 per thread.  `buf` is a [2][FAC_NWARPS+1] shared array,
// `parity` alternates per call.  ONE barrier per call: every warp publishes its total, and after the
// barrier every warp scans the FAC_NWARPS totals itself (the barrier of the next call orders the
// reads of this call before the buffer is written again two calls later).
template <int NW>
__device__ __forceinline__ uint32_t fac_block_offsets_t(uint32_t warp_total, uint32_t (*buf)[NW + 1], uint32_t &parity, uint32_t &total) {
    if (NW == 1) { __syncwarp(); total = warp_total; return 0u; }   // single-warp CTAs need no exchange
    uint32_t *b = buf[parity & 1u];
    parity++;
    if (fac_lane() == 0) b[fac_warp()] = warp_total;
    __syncthreads();
    const uint32_t x = fac_lane() < (uint32_t)NW ? b[fac_lane()] : 0u;
    uint32_t xi = x;
#pragma unroll
    for (int d = 1; d < NW; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, xi, d);
        if (fac_lane() >= (uint32_t)d) xi += t;
    }
    total = __shfl_sync(0xFFFFFFFFu, xi, NW - 1);
    return __shfl_sync(0xFFFFFFFFu, xi - x, fac_warp());
}
template <int NW>
__device__ __forceinline__ uint32_t fac_block_scan_t(uint32_t v, uint32_t (*buf)[NW + 1], uint32_t &parity, uint32_t &total) {
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (fac_lane() >= (uint32_t)d) incl += t;
    }
    const uint32_t warp_total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    return fac_block_offsets_t<NW>(warp_total, buf, parity, total) + incl - v;
}
// Same for a 1-bit predicate: ballot + popc inside the warp (order-preserving compaction rank).
template <int NW>
__device__ __forceinline__ uint32_t fac_block_rank_t(bool p, uint32_t (*buf)[NW + 1], uint32_t &parity, uint32_t &total) {
    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, p);
    const uint32_t in_warp = __popc(bal & ((1u << fac_lane()) - 1u));
    return fac_block_offsets_t<NW>(__popc(bal), buf, parity, total) + in_warp;
}
__device__ __forceinline__ uint32_t fac_block_scan(uint32_t v, uint32_t (*buf)[FAC_NWARPS + 1], uint32_t &parity, uint32_t &total) {
    return fac_block_scan_t<FAC_NWARPS>(v, buf, parity, total);
}
__device__ __forceinline__ uint32_t fac_block_rank(bool p, uint32_t (*buf)[FAC_NWARPS + 1], uint32_t &parity, uint32_t &total) {
    return fac_block_rank_t<FAC_NWARPS>(p, buf, parity, total);
}

// ---- TMA (bulk async copy engine) 1-D global -> shared with an mbarrier -------------------------
__device__ __forceinline__ uint32_t fac_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void fac_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fac_smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fac_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fac_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fac_mbar_wait(uint64_t *bar, uint32_t phase) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(fac_smem_u32(bar)), "r"(phase) : "memory");
}
// dst / src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void fac_tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(fac_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(fac_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fac_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__global__ void k_fill_u32(uint32_t *p, uint32_t v, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = v;
}

// ---------------------------------------------------------------------------------------------
// k_scan_bytes: OR of all haystack bytes' high bits (is_ascii) -- 16 bytes per thread per step.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_scan_bytes(const uint8_t *__restrict__ s, uint64_t n, uint32_t *__restrict__ flags) {
    uint32_t acc = 0;
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t head = ((16u - ((uintptr_t)s & 15u)) & 15u);
    const uint64_t h = head < n ? head : n;
    for (uint64_t i = tid; i < h; i += nth) acc |= s[i];
    const uint4 *v = (const uint4 *)(s + h);
    const uint64_t nv = (n - h) / 16;
    for (uint64_t i = tid; i < nv; i += nth) {
        const uint4 q = v[i];
        acc |= q.x | q.y | q.z | q.w;
    }
    for (uint64_t i = h + nv * 16 + tid; i < n; i += nth) acc |= s[i];
    acc &= 0x80808080u;
    acc = __reduce_or_sync(0xFFFFFFFFu, acc);
    if (fac_lane() == 0 && acc) atomicOr(flags, 1u);
}

// ---------------------------------------------------------------------------------------------
// K3: frontier expansion
// ---------------------------------------------------------------------------------------------
struct ExpandParams {
    AutomatonView A;
    TextView tv;
    float thr, maxpen;
    // tiling: mode 0 = uniform tiles of `tile` windows over [seg_begin, seg_end), haystack end text_end;
    //         mode 1 = explicit descriptors {start, count, text_end, _} (pre-filter slices, stream windows, retries)
    uint32_t mode, seg_begin, seg_end, text_end, tile, n_tiles;
    const uint4 *tiles;
    uint32_t lookahead;  // graphemes a window may read past its start
    uint32_t pass;       // 0 main pass, 1 retry pass
    // per-CTA scratch (index = blockIdx.x)
    FacState *queue;
    uint32_t *nxt, *hslot;
    uint32_t qcap;
    uint32_t *gtab_rep, *gtab_head;
    float *gtab_min;
    uint32_t gtab_size;       // power of two
    uint32_t smem_tab_size;   // power of two (slots of the shared-memory dedup table)
    uint32_t smem_text_cap;   // graphemes the shared text tile can hold
    // outputs
    FacCand *cands;
    uint32_t cand_cap;
    unsigned long long *counters;  // [0] next tile, [1] candidates, [2] states pushed, [3] failed tiles, [4] matches, [5] dirty, [6] queued states
    uint32_t *failed_tiles;
    uint32_t failed_cap;
    uint32_t *failed_bitmap;  // one bit per tile of the main pass
    uint32_t *per_window;     // optional: queue.len() of every start window (auto-beam accounting)
    int use_tma;
};

// Text accessor over the staged shared-memory tile with a global-memory fallback.
template <bool ASCII>
struct TileText {
    const uint8_t *sb;     // ASCII: folded bytes
    const uint32_t *sf;    // unicode: first chars
    const uint32_t *sg;    // unicode: grapheme ids (engines with mappings)
    uint32_t base, len;    // tile covers graphemes [base, base+len)
    FacTextDirect G;
    __device__ __forceinline__ uint32_t first(uint32_t j) const {
        const uint32_t r = j - base;
        if (r < len) return ASCII ? (uint32_t)sb[r] : sf[r];
        return G.first(j);
    }
    __device__ __forceinline__ uint32_t gid(uint32_t j) const {
        const uint32_t r = j - base;
        if (r < len) return ASCII ? G.ascii_gid[sb[r]] : sg[r];
        return G.gid(j);
    }
};

__device__ __forceinline__ uint32_t fac_hash3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t h = a * 0x9E3779B1u;
    h = (h ^ (h >> 16)) + b * 0x85EBCA6Bu;
    h = (h ^ (h >> 13)) + c * 0xC2B2AE35u;
    h ^= h >> 16;
    h *= 0x7FEB352Du;
    h ^= h >> 15;
    return h;
}

// Append one raw candidate (a `best.entry(key)` visit, search.rs:705-735) to the global list.
__device__ __forceinline__ void fac_emit_cand(const ExpandParams &P, uint32_t sg, uint32_t eg, uint32_t pat, float sim, uint32_t cnt,
                                              uint32_t seq, uint32_t tile, uint32_t tag) {
    const unsigned long long ci = atomicAdd(&P.counters[1], 1ull);
    if (ci < P.cand_cap) {
        FacCand cd;
        cd.sg = sg; cd.eg = eg; cd.pat = pat; cd.sim = sim; cd.cnt = cnt; cd.seq = seq; cd.tile = tile; cd.tag = tag;
        uint4 *dst = reinterpret_cast<uint4 *>(&P.cands[ci]);
        dst[0] = reinterpret_cast<uint4 *>(&cd)[0];
        dst[1] = reinterpret_cast<uint4 *>(&cd)[1];
    }
}

// FAST mode: a child that has spent its whole edit budget can only follow exact transitions
// (sub / swap / ins / del all need `edits < MAX_EDITS_FAST`, search.rs:810, 937, 1003, 1043), so its
// entire future is one chain of exact steps.  Instead of materialising that chain state by state
// it is walked here, with the same per-pop checks (node ceiling :638-642, outputs :659-737, exact
// transition :776-798).  Returns the number of chain states visited.
template <class Text>
__device__ __forceinline__ uint32_t fac_walk_exhausted(const ExpandParams &P, const Text &T, uint32_t start, uint32_t text_end,
                                                       const FacState &child, uint32_t tile, uint32_t tag) {
    const AutomatonView &A = P.A;
    uint32_t node = child.node;
    uint32_t jr = (child.pos >> FAC_POS_J_SHIFT) & FAC_POS_MASK, mr = child.pos & FAC_POS_MASK;
    uint32_t steps = 0;
    for (;;) {
        steps++;
        if (fac_over_ceiling(A, node, child.pen, P.thr)) break;
        const uint32_t o1 = A.node_out_off[node + 1];
        for (uint32_t o = A.node_out_off[node]; o < o1; o++) {
            const uint32_t pat = A.out_pat[o];
            float sim;
            if (fac_eval_output(A, P.thr, pat, child.pen, child.cnt, sim)) fac_emit_cand(P, start, start + mr, pat, sim, child.cnt, 0u, tile, tag);
        }
        const uint32_t j = start + jr;
        if (j >= text_end) break;
        const uint32_t nx = A.has_mappings ? fac_lookup(A, node, T.gid(j)) : fac_lookup(A, node, T.first(j));
        if (nx == FAC_NONE) break;
        node = nx; jr++; mr = jr;
    }
    return steps;
}

template <bool ASCII, bool MAPP, bool FAST>
__global__ void __launch_bounds__(FAC_BLOCK) k_expand(const __grid_constant__ ExpandParams P) {
    extern __shared__ __align__(16) uint8_t dyn_smem[];
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ uint32_t s_scan[2][FAC_NWARPS + 1];
    __shared__ uint32_t s_tile_idx;
    __shared__ uint32_t c_node[FAC_BLOCK], c_cnt[FAC_BLOCK], c_pos[FAC_BLOCK], c_exact[FAC_BLOCK], c_flags[FAC_BLOCK];
    __shared__ float c_pen[FAC_BLOCK];
    __shared__ uint32_t s_off[FAC_BLOCK + 1];

    const AutomatonView &A = P.A;
    const uint32_t tid = threadIdx.x;
    uint32_t parity = 0;
    uint32_t mbar_phase = 0;

    // dynamic shared memory carve-up: [text tile][dedup table rep][dedup table head]
    const uint32_t text_bytes = ASCII ? ((P.smem_text_cap + 31u) & ~15u) : ((P.smem_text_cap * 4u + 31u) & ~15u);
    uint8_t *s_text_b = dyn_smem;
    uint32_t *s_text_f = (uint32_t *)dyn_smem;
    uint32_t *s_text_g = (uint32_t *)(dyn_smem + text_bytes);
    uint8_t *after_text = dyn_smem + text_bytes * ((!ASCII && MAPP) ? 2u : 1u);
    uint32_t *s_rep = (uint32_t *)after_text;
    uint32_t *s_head = s_rep + P.smem_tab_size;

    FacState *const queue = P.queue + (size_t)blockIdx.x * P.qcap;
    uint32_t *const nxt = P.nxt + (size_t)blockIdx.x * P.qcap;
    uint32_t *const hslot = P.hslot + (size_t)blockIdx.x * P.qcap;
    uint32_t *const g_rep = P.gtab_rep + (size_t)blockIdx.x * P.gtab_size;
    uint32_t *const g_head = P.gtab_head + (size_t)blockIdx.x * P.gtab_size;
    float *const g_min = P.gtab_min + (size_t)blockIdx.x * P.gtab_size;

    for (uint32_t k = tid; k < P.smem_tab_size; k += FAC_BLOCK) { s_rep[k] = FAC_EMPTY; s_head[k] = FAC_EMPTY; }
    if (tid == 0) fac_mbar_init(&s_mbar, 1);
    __syncthreads();

    for (;;) {
        if (tid == 0) s_tile_idx = (uint32_t)atomicAdd(&P.counters[0], 1ull);
        __syncthreads();
        const uint32_t t = s_tile_idx;
        if (t >= P.n_tiles) break;
        uint32_t tile_start, count, text_end;
        uint32_t win_tag = P.pass << 31;  // window id of the tile | pass bit
        if (P.mode == 0) {
            tile_start = P.seg_begin + t * P.tile;
            count = min(P.tile, P.seg_end - tile_start);
            text_end = P.text_end;
        } else {
            const uint4 d = P.tiles[t];
            tile_start = d.x; count = d.y; text_end = d.z; win_tag |= d.w;
        }

        // ---- stage the grapheme tile into shared memory (TMA bulk copy + tail by plain loads) ----
        TileText<ASCII> T;
        T.G.tv = P.tv; T.G.ascii_gid = A.ascii_gid; T.G.ci = A.ci;
        T.sb = s_text_b; T.sf = s_text_f; T.sg = s_text_g;
        {
            const uint32_t want = min(text_end - tile_start, count + P.lookahead);
            const uint32_t elem = ASCII ? 1u : 4u;
            // align the window down to 16 bytes so the bulk engine can be used
            uint32_t lead = ASCII ? (uint32_t)(((uintptr_t)(P.tv.bytes + tile_start)) & 15u) : (tile_start & 3u);
            if (lead > tile_start) lead = 0;  // cannot align below the start of the buffer
            const uint32_t base = tile_start - lead;
            const bool aligned = ASCII ? ((((uintptr_t)(P.tv.bytes + base)) & 15u) == 0u) : ((base & 3u) == 0u);
            uint32_t len = min(want + lead, P.smem_text_cap);
            T.base = base; T.len = len;
            const uint32_t bulk_elems = (P.use_tma && aligned) ? ((len * elem) & ~15u) / elem : 0u;
            if (bulk_elems) {
                if (tid == 0) {
                    fac_fence_proxy_async();
                    const uint32_t nb = bulk_elems * elem * ((!ASCII && MAPP) ? 2u : 1u);
                    fac_mbar_expect_tx(&s_mbar, nb);
                    if (ASCII) fac_tma_load_1d(s_text_b, P.tv.bytes + base, bulk_elems, &s_mbar);
                    else {
                        fac_tma_load_1d(s_text_f, P.tv.first + base, bulk_elems * 4u, &s_mbar);
                        if (MAPP) fac_tma_load_1d(s_text_g, P.tv.gid + base, bulk_elems * 4u, &s_mbar);
                    }
                }
            }
            for (uint32_t k = bulk_elems + tid; k < len; k += FAC_BLOCK) {
                if (ASCII) s_text_b[k] = P.tv.bytes[base + k];
                else { s_text_f[k] = P.tv.first[base + k]; if (MAPP) s_text_g[k] = P.tv.gid[base + k]; }
            }
            if (bulk_elems) { fac_mbar_wait(&s_mbar, mbar_phase & 1u); mbar_phase++; }
            __syncthreads();
            if (ASCII && A.ci) {  // to_ascii_lowercase in place (AsciiGraphemes::gs_first_char, grapheme.rs:110-117)
                for (uint32_t k = tid; k < len; k += FAC_BLOCK) { const uint8_t b = s_text_b[k]; if (b >= 'A' && b <= 'Z') s_text_b[k] = b + 32; }
                __syncthreads();
            }
        }

        // ---- level 0: one root state per (non-skipped) start window, in window order ----
        uint32_t qlen = 0;
        uint32_t tail_steps = 0;  // FAST: chain states walked in place instead of being queued
        bool failed = false;
        for (uint32_t w0 = 0; w0 < count; w0 += FAC_BLOCK) {
            const uint32_t w = w0 + tid;
            bool push = false;
            if (w < count) {
                const uint32_t start = tile_start + w;
                const bool has1 = start + 1 < text_end;
                push = !fac_window_skipped(A, T.first(start), has1, has1 ? T.first(start + 1) : 0u);
            }
            uint32_t total;
            const uint32_t r = fac_block_rank(push, s_scan, parity, total);
            if (push) {
                FacState S; S.node = 0; S.pen = 0.f; S.cnt = 0; S.pos = fac_make_pos(w, 0, 0);
                *reinterpret_cast<uint4 *>(&queue[qlen + r]) = *reinterpret_cast<uint4 *>(&S);
            }
            qlen += total;
        }
        __syncthreads();

        uint32_t lb = 0, le = qlen;
        while (lb < le && !failed) {
            const uint32_t C = le - lb;
            const bool use_smem_tab = !MAPP && (C * 2u <= P.smem_tab_size);
            uint32_t *const rep = use_smem_tab ? s_rep : g_rep;
            uint32_t *const head = use_smem_tab ? s_head : g_head;
            // the per-level table is sized to the level (power of two >= 2C) so its footprint stays
            // L2-resident; the persistent table of engines with mappings keeps one size per tile
            uint32_t tsize = use_smem_tab ? P.smem_tab_size : P.gtab_size;
            if (!MAPP && !use_smem_tab) {
                const uint32_t need = 2u * C;
                uint32_t p2 = 1u << (32 - __clz(need - 1u));
                if (p2 < tsize) tsize = p2;
            }
            const uint32_t tmask = tsize - 1u;

            // ---- phase A: group the level's states by dedup key (VisitedKey, search.rs:31-37) ----
            for (uint32_t i = lb + tid; i < le; i += FAC_BLOCK) {
                const uint4 q = *reinterpret_cast<const uint4 *>(&queue[i]);
                uint32_t h = fac_hash3(q.x, q.z, q.w) & tmask;
                for (;;) {
                    uint32_t r = *((volatile uint32_t *)&rep[h]);
                    if (r == FAC_EMPTY) {
                        const uint32_t old = atomicCAS(&rep[h], FAC_EMPTY, i);
                        r = (old == FAC_EMPTY) ? i : old;
                    }
                    if (r == i) break;
                    const uint4 rq = *reinterpret_cast<const uint4 *>(&queue[r]);
                    if (rq.x == q.x && rq.z == q.z && rq.w == q.w) break;
                    h = (h + 1u) & tmask;
                }
                nxt[i] = atomicExch(&head[h], i);
                hslot[i] = h;
            }
            __syncthreads();

            // ---- phase B: chunks of FAC_BLOCK states -> (state, slot) work items -> next level ----
            uint32_t nbase = le;
            for (uint32_t c0 = lb; c0 < le && !failed; c0 += FAC_BLOCK) {
                const uint32_t i = c0 + tid;
                uint32_t nslots = 0;
                if (i < le) {
                    const uint4 q = *reinterpret_cast<const uint4 *>(&queue[i]);
                    FacState S; S.node = q.x; S.pen = __uint_as_float(q.y); S.cnt = q.z; S.pos = q.w;
                    const uint32_t w = S.pos >> FAC_POS_W_SHIFT;
                    const uint32_t start = tile_start + w;
                    if (P.per_window) atomicAdd(&P.per_window[start], 1u);
                    // dedup verdict: expanded iff strictly lower than every earlier same-key state
                    // (visited.entry, search.rs:608-628)
                    const uint32_t h = hslot[i];
                    float m = MAPP ? g_min[h] : __int_as_float(0x7F800000);
                    for (uint32_t k = head[h]; k != FAC_EMPTY; k = nxt[k])
                        if (k < i) m = fminf(m, queue[k].pen);
                    const bool expanded = !(m <= S.pen);
                    if (expanded && !fac_over_ceiling(A, S.node, S.pen, P.thr)) {
                        // outputs (search.rs:659-737): raw candidates, reduced later
                        const uint32_t o1 = A.node_out_off[S.node + 1];
                        for (uint32_t o = A.node_out_off[S.node]; o < o1; o++) {
                            const uint32_t pat = A.out_pat[o];
                            float sim;
                            if (fac_eval_output(A, P.thr, pat, S.pen, S.cnt, sim))
                                fac_emit_cand(P, start, start + (S.pos & FAC_POS_MASK), pat, sim, S.cnt, i, t, win_tag);
                        }
                        FacCtx Cx;
                        fac_make_ctx(A, T, P.maxpen, start, text_end, S, Cx);
                        nslots = Cx.nslots;
                        c_node[tid] = Cx.node; c_pen[tid] = Cx.pen; c_cnt[tid] = Cx.cnt; c_pos[tid] = Cx.pos;
                        c_exact[tid] = Cx.exact; c_flags[tid] = Cx.flags;
                    }
                }
                uint32_t W;
                const uint32_t excl = fac_block_scan(nslots, s_scan, parity, W);
                s_off[tid] = excl;
                if (tid == 0) s_off[FAC_BLOCK] = W;
                __syncthreads();
                for (uint32_t k0 = 0; k0 < W; k0 += FAC_BLOCK) {
                    const uint32_t k = k0 + tid;
                    bool push = false;
                    FacState child;
                    if (k < W) {
                        // owner state = last s with s_off[s] <= k
                        uint32_t lo = 0, hi = FAC_BLOCK;
                        while (hi - lo > 1u) {
                            const uint32_t mid = (lo + hi) >> 1;
                            if (s_off[mid] <= k) lo = mid; else hi = mid;
                        }
                        FacCtx Cx;
                        Cx.node = c_node[lo]; Cx.pen = c_pen[lo]; Cx.cnt = c_cnt[lo]; Cx.pos = c_pos[lo];
                        Cx.exact = c_exact[lo]; Cx.flags = c_flags[lo]; Cx.nslots = 0;
                        const uint32_t start = tile_start + (Cx.pos >> FAC_POS_W_SHIFT);
                        push = fac_eval_slot(A, T, P.maxpen, start, text_end, Cx, k - s_off[lo], child);
                        if (FAST && push && (int)fac_edits_of(child.cnt) >= A.mef) {
                            tail_steps += fac_walk_exhausted(P, T, start, text_end, child, t, win_tag);
                            push = false;
                        }
                    }
                    uint32_t total;
                    const uint32_t r = fac_block_rank(push, s_scan, parity, total);
                    if (nbase + total > P.qcap) { failed = true; break; }
                    if (push) *reinterpret_cast<uint4 *>(&queue[nbase + r]) = *reinterpret_cast<uint4 *>(&child);
                    nbase += total;
                }
                __syncthreads();
            }

            // ---- phase C: release (per-level table) or fold minima in (persistent table) ----
            for (uint32_t i = lb + tid; i < le; i += FAC_BLOCK) {
                const uint32_t h = hslot[i];
                if (MAPP) {
                    // pens are non-negative finite floats in practice; the CAS loop is exact for any order
                    const float pen = queue[i].pen;
                    int *addr = (int *)&g_min[h];
                    int old = *addr;
                    while (__int_as_float(old) > pen) {
                        const int prev = atomicCAS(addr, old, __float_as_int(pen));
                        if (prev == old) break;
                        old = prev;
                    }
                    head[h] = FAC_EMPTY;
                } else {
                    rep[h] = FAC_EMPTY;
                    head[h] = FAC_EMPTY;
                }
            }
            __syncthreads();
            lb = le;
            le = nbase;
        }

        if (MAPP) {  // reset the persistent table for the next tile
            for (uint32_t i = tid; i < lb; i += FAC_BLOCK) {
                const uint32_t h = hslot[i];
                g_rep[h] = FAC_EMPTY; g_head[h] = FAC_EMPTY; g_min[h] = __int_as_float(0x7F800000);
            }
        }
        if (FAST && !failed) {
            tail_steps = __reduce_add_sync(0xFFFFFFFFu, tail_steps);
            if (fac_lane() == 0 && tail_steps) atomicAdd(&P.counters[2], (unsigned long long)tail_steps);
        }
        if (tid == 0) {
            if (failed) {
                const unsigned long long fi = atomicAdd(&P.counters[3], 1ull);
                if (fi < P.failed_cap) P.failed_tiles[fi] = t;
                if (P.failed_bitmap) atomicOr(&P.failed_bitmap[t >> 5], 1u << (t & 31u));
            } else {
                atomicAdd(&P.counters[2], (unsigned long long)le);
                atomicAdd(&P.counters[6], (unsigned long long)le);  // queued states only (tile sizing)
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// candidate reduction: best per (start, end, pattern) = max similarity, earliest FIFO position on
// ties (`if similarity > entry.similarity` keeps the first, search.rs:705-722)
// ---------------------------------------------------------------------------------------------
struct BestParams {
    const FacCand *cands;
    uint32_t n_cands;
    uint32_t *tab_rep;            // [tab_size] candidate index or FAC_EMPTY
    unsigned long long *tab_val;  // [tab_size]
    uint32_t tab_size;            // power of two
    uint32_t *cslot;              // [n_cands]
    // candidates of the main pass whose tile failed are superseded by the retry pass
    const uint32_t *failed_bitmap;
    uint32_t sg_end;              // candidates of start windows >= sg_end are ignored
    TextView tv;
    const FacWindow *windows;  // haystack windows of this call (a whole-haystack search has one)
    WMatch *out;
    unsigned long long *out_count;
    uint32_t out_cap;
};

__device__ __forceinline__ unsigned long long fac_cand_val(const FacCand &c) {
    return ((unsigned long long)fac_total_order_u32(c.sim) << 32) | (unsigned long long)(0xFFFFFFFFu - c.seq);
}
__device__ __forceinline__ bool fac_cand_dead(const BestParams &P, const FacCand &c) {
    if (c.sg >= P.sg_end) return true;  // windows past the auto_beam crossing are redone beamed
    if ((c.tag >> 31) != 0 || !P.failed_bitmap) return false;
    return (P.failed_bitmap[c.tile >> 5] >> (c.tile & 31u)) & 1u;
}

__global__ void __launch_bounds__(256) k_best_insert(const BestParams P) {
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
    atomicMax(&P.tab_val[h], fac_cand_val(c));
}

__global__ void __launch_bounds__(256) k_best_select(const BestParams P) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_cands) return;
    const uint32_t h = P.cslot[i];
    if (h == FAC_EMPTY) return;
    const FacCand c = P.cands[i];
    if (P.tab_val[h] != fac_cand_val(c)) return;
    // start/end bytes (search.rs:667-676); the offset table carries a sentinel, and a window's
    // matched_end never exceeds its own text_end, so one lookup serves both branches
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
