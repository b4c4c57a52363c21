// fac_stack.cuh -- K3, general stack-machine variant: fuzzy frontier expansion for the engines of the
// reference's fast monomorphisations `search_unsorted_impl<MAPPINGS, _, MAX_EDITS_FAST = 1..6>`
// (src/search.rs:204-393) that are OUTSIDE the succinct kernel's domain: multi-character mappings
// (src/search.rs:883-923), multi-byte / non-ASCII pattern graphemes (src/structs.rs:452-519), alphabets of
// more than 63 symbols, similarity tables with non-ASCII members -- the Unicode workload (cfg3).
//
// Same execution shape as k_expand_succinct (fac_succinct.cuh), with the slot formulation of fac_core.h (every
// potential child of a state is a numbered slot) over merged 16-byte node / edge records (fac_flat.h: one record
// load per state, one per child) instead of the succinct trie:
//
//   * persistent CTAs, tiles of start windows fetched with one atomicAdd per tile; the tile's grapheme
//     stream (folded first chars, + grapheme ids for engines with mappings, or the haystack bytes) staged
//     into shared memory by TMA bulk copies (cp.async.bulk + mbarrier);
//   * every WARP runs one depth-first stack machine in shared memory over a stream of start windows:
//     pop <= 32 states (one per lane; node ceiling src/search.rs:638-642, outputs :659-737, slot count),
//     flatten the (state, slot) pairs of the pop with a warp prefix sum and turn them into child states
//     32 per round (exact :776-798, substitutions :814-874, mapping transitions :883-923, swap :935-989,
//     insertion :994-1029, deletions :1035-1089), pushed by __ballot_sync / __popc compaction;
//   * a child whose node ceiling already rejects it is dropped at push time (the reference drops it when it
//     is popped: result-neutral, only the visited-state statistic changes);
//   * children that have spent the edit budget can only follow exact transitions (sub / mapping / swap /
//     ins / del all need edits < MAX_EDITS_FAST): they go to a per-warp walk queue and are walked 32 at a
//     time.
//
// Order-independent like the succinct kernel: no dedup map (result-neutral), candidates reduced by
// maximum similarity with tie detection (fac_fastreduce.cuh); tied windows and windows whose states did not fit
// the warp stack are redone by the order-faithful kernel (k_expand).  No global-memory frontier.
#pragma once
#include "fac_flat.h"
#include "fac_kernels.cuh"

#define STK_WQ_CAP 96u
#define STK_THREADS 256

struct StackParams {
    ExpandParams E;
    FlatView F;           // per-call merged records (k_flat_prepare)
    uint32_t stack_cap;   // states per warp stack
    uint32_t feed_below;  // a new root is fed while fewer than this many states are stacked
    uint32_t *dirty;      // bitmap over start windows (bit start - E.seg_begin): a state of the window did not fit
};

// per-call records: node ceiling = prune_len - prune_len_over_weight * threshold (search.rs:638-642, exact f32 ops);
// an edge record carries the ceiling of its child so that a doomed child is never pushed
__global__ void __launch_bounds__(256) k_flat_prepare_nodes(const uint4 *__restrict__ stat, const float *__restrict__ plen, const float *__restrict__ plow,
                                                            float thr, uint32_t n, uint4 *__restrict__ nrec) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint4 v = stat[i];
    v.z = __float_as_uint(__fsub_rn(plen[i], __fmul_rn(plow[i], thr)));
    nrec[i] = v;
}
__global__ void __launch_bounds__(256) k_flat_prepare_edges(const uint4 *__restrict__ stat, const uint4 *__restrict__ nrec, uint32_t n, uint4 *__restrict__ erec) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint4 v = stat[i];
    v.w = nrec[v.x & 0x7FFFFFFFu].z;
    erec[i] = v;
}

struct StkEmit {
    const ExpandParams *P;
    uint32_t tile, tag;
    __device__ __forceinline__ void operator()(uint32_t sg, uint32_t eg, uint32_t pat, float sim, uint32_t cnt) const {
        fac_emit_cand(*P, sg, eg, pat, sim, cnt, 0u, tile, tag);
    }
};

template <bool ASCII, bool MAPP>
__global__ void __launch_bounds__(STK_THREADS) k_expand_stack(const __grid_constant__ StackParams SP) {
    extern __shared__ __align__(16) uint8_t dyn_smem[];
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ uint32_t s_tile_idx, s_next_win;
    constexpr uint32_t NW = STK_THREADS / 32;
    const ExpandParams &P = SP.E;
    const AutomatonView &A = P.A;
    const FlatView F = SP.F;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;

    // carve-up: [text tile: first chars / bytes][grapheme ids][warp stacks][walk queues]
    const uint32_t text_bytes = ASCII ? ((P.smem_text_cap + 31u) & ~15u) : ((P.smem_text_cap * 4u + 31u) & ~15u);
    uint8_t *s_text_b = dyn_smem;
    uint32_t *s_text_f = (uint32_t *)dyn_smem;
    uint32_t *s_text_g = (uint32_t *)(dyn_smem + text_bytes);
    uint4 *s_stack = (uint4 *)(dyn_smem + text_bytes * ((!ASCII && MAPP) ? 2u : 1u));
    uint4 *s_wq = s_stack + (size_t)NW * SP.stack_cap;
    uint4 *const stk = s_stack + (size_t)warp * SP.stack_cap;
    uint4 *const wq = s_wq + (size_t)warp * STK_WQ_CAP;
    const uint32_t cap = SP.stack_cap;
    uint32_t mbar_phase = 0;
    uint32_t n_states = 0;

    if (tid == 0) fac_mbar_init(&s_mbar, 1);
    __syncthreads();

    for (;;) {
        if (tid == 0) { s_tile_idx = (uint32_t)atomicAdd(&P.counters[0], 1ull); s_next_win = 0; }
        __syncthreads();
        const uint32_t t = s_tile_idx;
        if (t >= P.n_tiles) break;
        uint32_t tile_start, count, text_end;
        uint32_t win_tag = 0;
        if (P.mode == 0) {
            tile_start = P.seg_begin + t * P.tile;
            count = min(P.tile, P.seg_end - tile_start);
            text_end = P.text_end;
        } else {
            const uint4 d = P.tiles[t];
            tile_start = d.x; count = d.y; text_end = d.z; win_tag = d.w;
        }

        // ---- stage the grapheme tile (TMA bulk copy of the aligned body + tail by plain loads) ----
        TileText<ASCII> T;
        T.G.tv = P.tv; T.G.ascii_gid = A.ascii_gid; T.G.ci = A.ci;
        T.sb = s_text_b; T.sf = s_text_f; T.sg = s_text_g;
        {
            const uint32_t want = min(text_end - tile_start, count + P.lookahead);
            const uint32_t elem = ASCII ? 1u : 4u;
            uint32_t lead = ASCII ? (uint32_t)(((uintptr_t)(P.tv.bytes + tile_start)) & 15u) : (tile_start & 3u);
            if (lead > tile_start) lead = 0;
            const uint32_t base = tile_start - lead;
            const bool aligned = ASCII ? ((((uintptr_t)(P.tv.bytes + base)) & 15u) == 0u) : ((base & 3u) == 0u);
            const uint32_t len = min(want + lead, P.smem_text_cap);
            T.base = base; T.len = len;
            const uint32_t bulk_elems = (P.use_tma && aligned) ? ((len * elem) & ~15u) / elem : 0u;
            if (bulk_elems && tid == 0) {
                fac_fence_proxy_async();
                fac_mbar_expect_tx(&s_mbar, bulk_elems * elem * ((!ASCII && MAPP) ? 2u : 1u));
                if (ASCII) fac_tma_load_1d(s_text_b, P.tv.bytes + base, bulk_elems, &s_mbar);
                else {
                    fac_tma_load_1d(s_text_f, P.tv.first + base, bulk_elems * 4u, &s_mbar);
                    if (MAPP) fac_tma_load_1d(s_text_g, P.tv.gid + base, bulk_elems * 4u, &s_mbar);
                }
            }
            for (uint32_t k = bulk_elems + tid; k < len; k += STK_THREADS) {
                if (ASCII) s_text_b[k] = P.tv.bytes[base + k];
                else { s_text_f[k] = P.tv.first[base + k]; if (MAPP) s_text_g[k] = P.tv.gid[base + k]; }
            }
            if (bulk_elems) { fac_mbar_wait(&s_mbar, mbar_phase & 1u); mbar_phase++; }
            __syncthreads();
            if (ASCII && A.ci) {  // to_ascii_lowercase in place (grapheme.rs:110-117)
                for (uint32_t k = tid; k < len; k += STK_THREADS) { const uint8_t b = s_text_b[k]; if (b >= 'A' && b <= 'Z') s_text_b[k] = b + 32; }
                __syncthreads();
            }
        }

        // ---- every warp: one stack machine over a stream of the tile's start windows ----
        {
            uint32_t top = 0, wn = 0;      // stack height, walk-queue length (warp-uniform)
            uint32_t b0 = 0, total = 0;    // item rounds of the current pop
            uint32_t off = 0;              // exclusive prefix of the lanes' slot counts
            bool more = true, fed = false;
            FlatCtx C;
            C.node = C.cnt = C.pos = C.eoff = C.shape = C.lists = C.nslots = 0; C.exact = C.row = FAC_NONE; C.pen = 0.f;
            const StkEmit emit{&P, t, win_tag};
            for (;;) {
                __syncwarp();
                // (1) exhausted children: exact transitions only, 32 at a time
                if (wn >= 32u || (wn && b0 >= total && top == 0 && !more)) {
                    const uint32_t n = min(wn, 32u);
                    if (lane < n) {
                        const uint4 q = wq[wn - n + lane];
                        FacState c;
                        c.node = q.x; c.pen = __uint_as_float(q.y); c.cnt = q.z; c.pos = q.w;
                        n_states += flat_walk(A, F, T, P.thr, emit, tile_start + (q.w >> FAC_POS_W_SHIFT), text_end, c);
                    }
                    wn -= n;
                    continue;
                }
                // (2) (state, slot) pairs of the last pop, 32 per round
                if (b0 < total) {
                    const uint32_t it = b0 + lane;
                    b0 += 32u;
                    uint32_t lo = 0;
#pragma unroll
                    for (int step = 16; step; step >>= 1) {
                        const uint32_t cand = lo + step;
                        const uint32_t v = __shfl_sync(0xFFFFFFFFu, off, cand & 31u);
                        if (cand < 32u && v <= it) lo = cand;
                    }
                    FlatCtx O;
                    O.node = __shfl_sync(0xFFFFFFFFu, C.node, lo);
                    O.pen = __shfl_sync(0xFFFFFFFFu, C.pen, lo);
                    O.cnt = __shfl_sync(0xFFFFFFFFu, C.cnt, lo);
                    O.pos = __shfl_sync(0xFFFFFFFFu, C.pos, lo);
                    O.exact = __shfl_sync(0xFFFFFFFFu, C.exact, lo);
                    O.eoff = __shfl_sync(0xFFFFFFFFu, C.eoff, lo);
                    O.shape = __shfl_sync(0xFFFFFFFFu, C.shape, lo);
                    O.lists = __shfl_sync(0xFFFFFFFFu, C.lists, lo);
                    O.row = __shfl_sync(0xFFFFFFFFu, C.row, lo);
                    O.nslots = 0;
                    const uint32_t r = it - __shfl_sync(0xFFFFFFFFu, off, lo);
                    FacState c;
                    c.node = 0; c.pen = 0.f; c.cnt = 0; c.pos = 0;
                    bool ok = false;
                    if (it < total) {
                        const uint32_t start = tile_start + (O.pos >> FAC_POS_W_SHIFT);
                        ok = flat_eval_slot<true>(A, F, T, P.maxpen, start, text_end, O, r, c);
                    }
                    const bool exhausted = (int)fac_edits_of(c.cnt) >= A.mef;
                    const bool to_walk = ok && exhausted, to_stack = ok && !exhausted;
                    const uint32_t bw = __ballot_sync(0xFFFFFFFFu, to_walk), bs = __ballot_sync(0xFFFFFFFFu, to_stack);
                    const uint4 cv = make_uint4(c.node, __float_as_uint(c.pen), c.cnt, c.pos);
                    if (to_walk) wq[wn + __popc(bw & lt_mask)] = cv;
                    if (to_stack) stk[top + __popc(bs & lt_mask)] = cv;
                    wn += __popc(bw); top += __popc(bs);
                    continue;
                }
                // (3) feed start windows while the stack is short (a whole warp's worth for edits(1) engines)
                if (more && !fed && top < SP.feed_below) {
                    const uint32_t nf = A.mef <= 1 ? 32u - min(top, 31u) : 1u;
                    uint32_t w0 = 0;
                    if (lane == 0) w0 = atomicAdd(&s_next_win, nf);
                    w0 = __shfl_sync(0xFFFFFFFFu, w0, 0);
                    if (w0 >= count) { more = false; continue; }
                    if (w0 + nf >= count) more = false;
                    const uint32_t w = w0 + lane;
                    bool push = lane < nf && w < count;
                    if (push) {
                        const uint32_t start = tile_start + w;
                        const bool has1 = start + 1 < text_end;
                        push = !fac_window_skipped(A, T.first(start), has1, has1 ? T.first(start + 1) : 0u);   // search.rs:535-553
                    }
                    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, push);
                    if (push) stk[top + __popc(bal & lt_mask)] = make_uint4(0u, 0u, 0u, fac_make_pos(w, 0, 0));
                    top += __popc(bal);
                    fed = bal != 0u;
                    continue;
                }
                if (top == 0) {
                    if (!more && wn == 0) break;
                    fed = false;
                    continue;
                }
                // (4) pop up to 32 states
                fed = false;
                const uint32_t navail = min(top, 32u);
                const bool has = lane < navail;
                FacState S;
                S.node = 0; S.pen = 0.f; S.cnt = 0; S.pos = 0;
                if (has) { const uint4 sv = stk[top - 1u - lane]; S.node = sv.x; S.pen = __uint_as_float(sv.y); S.cnt = sv.z; S.pos = sv.w; }
                const uint32_t start = tile_start + (S.pos >> FAC_POS_W_SHIFT);
                FlatRec nr;
                nr.x = nr.y = nr.w = 0u; nr.z = 0xFF800000u;   // -inf ceiling: not live
                if (has) { const uint4 v = reinterpret_cast<const uint4 *>(F.nrec)[S.node]; nr.x = v.x; nr.y = v.y; nr.z = v.z; nr.w = v.w; }
                const bool live = has && !(S.pen > __uint_as_float(nr.z));   // node ceiling, search.rs:638-642
                const bool last = (int)fac_edits_of(S.cnt) + 1 >= A.mef;
                C.nslots = 0; C.shape = 0; C.lists = 0; C.exact = FAC_NONE;
                if (live) flat_make_ctx<true>(A, F, T, P.maxpen, start, text_end, S, nr, C);
                // stack pushes of this state in the worst case: a state on its last edit keeps only its exact child
                const uint32_t ub = !live ? 0u : (last ? ((C.shape & FLAT_F_EXACT) ? 1u : 0u) : C.nslots);
                uint32_t n_pop = navail;
                if (__any_sync(0xFFFFFFFFu, ub > 1u)) {
                    uint32_t incl = ub;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                        if (lane >= (uint32_t)d) incl += v;
                    }
                    const uint32_t viol = __ballot_sync(0xFFFFFFFFu, has && (incl > cap - top + lane + 1u));
                    if (viol) n_pop = (uint32_t)(__ffs(viol) - 1);
                }
                if (n_pop == 0) {  // the top state alone does not fit: its window is redone by the faithful kernel
                    if (lane == 0) {
                        const uint32_t wabs = start - P.seg_begin;
                        atomicOr(&SP.dirty[wabs >> 5], 1u << (wabs & 31u));
                        atomicAdd(&P.counters[7], 1ull);
                    }
                    top -= 1u;
                    total = 0; b0 = 0;
                    continue;
                }
                const bool active = lane < n_pop && live;
                top -= n_pop;
                if (active) {
                    n_states++;
                    const uint32_t no = flat_nout(nr);
                    for (uint32_t o = 0; o < no; o++) {   // outputs, search.rs:659-737 (`edits > MAX_EDITS_FAST` never holds)
                        const uint32_t pat = A.out_pat[nr.w + o];
                        const float total = A.pat_glen[pat];
                        const float sim = __fmul_rn(__fdiv_rn(__fsub_rn(total, S.pen), total), A.pat_weight[pat]);   // search.rs:698-699
                        if (!(sim < P.thr)) emit(start, start + (S.pos & FAC_POS_MASK), pat, sim, S.cnt);
                    }
                }
                const uint32_t n_items = active ? C.nslots : 0u;
                off = n_items;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, off, d);
                    if (lane >= (uint32_t)d) off += v;
                }
                total = __shfl_sync(0xFFFFFFFFu, off, 31);
                off -= n_items;
                b0 = 0;
            }
        }
        __syncthreads();  // every warp is done with the tile before it is restaged
    }
    n_states = __reduce_add_sync(0xFFFFFFFFu, n_states);
    if (lane == 0 && n_states) atomicAdd(&P.counters[2], (unsigned long long)n_states);
}
