// fac_succinct.cuh -- K3 fast kernel: fuzzy frontier expansion over the succinct BFS-ordered trie.
//
// search_unsorted_impl (src/search.rs:418-1119) for engines without mappings whose trie edges are single ASCII
// bytes over at most 63 symbols: the fast monomorphisations <MAPPINGS=false, _, MAX_EDITS_FAST=1..6>, the generic
// MAX_EDITS_FAST=255 path with per-pattern / per-type limits (template LIM) and engines without limits
// (exact-only).  Haystacks are ASCII bytes or the K1 stream of folded first chars.  Per-state logic is in
// fac_succinct.h (shared with the CPU emulator); this file is the SIMT orchestration:
//
//   * one persistent CTA per SM, tiles of start windows fetched with one atomicAdd per tile;
//   * the tile's haystack elements (+ look-ahead) staged into shared memory by one TMA bulk copy,
//     then case-folded and translated to dense symbols in place;
//   * the first `n_smem_nodes` node records (BFS order == shallow levels first, where most visits
//     land) copied to shared memory once per CTA; deeper records are 128-bit loads that hit L2;
//   * each warp runs ONE depth-first stack machine in shared memory over a stream of start windows
//     (a state carries its window; the next root is fed when fewer than 32 states are left):
//     pop <= 32 states (one per lane), push exact/swap/insertion children by ballot/popc
//     compaction; the children that survive the last-edit dead-end filter are read off the
//     precomputed grandchild masks (two loads per state instead of one record per child),
//     flattened over the warp with a prefix sum and turned into child states 32 per round;
//   * children that exhausted the edit budget can only follow exact transitions: they are queued in
//     a per-warp shared-memory walk queue and walked 32 at a time, so the divergent chain walk runs
//     with (nearly) full warps.
//
// No global-memory frontier: DRAM traffic is the haystack once plus the emitted candidates.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fac_kernels.cuh"
#include "fac_succinct.h"

struct SuccParams {
    const uint4 *rec;        // [n_nodes] per-call node records (SuccRec)
    const uint4 *out2;       // SuccOut entries
    const float *sub_pen;    // [ROW][SUCC_SP_STRIDE]
    const uint8_t *sym_of;   // [256]
    const uint8_t *text;     // ASCII haystack bytes, or
    const uint32_t *first;   // non-ASCII haystack: first char of every folded grapheme (K1 stream); null for ASCII
    uint32_t n_nodes, n_smem_nodes;
    SuccConsts K;
    int32_t ci, wskip;
    int32_t exact_only;      // engine without FuzzyLimits: only the exact chain from the root can emit
    unsigned long long first_mask, second_mask;
    uint32_t seg_begin, seg_end, text_end, tile, n_tiles, lookahead;
    const uint4 *tiles;      // optional explicit tiles {start, count (<= tile), text_end, window id} (pre-filter slices, stream batches); null = uniform tiling
    const void *masks;       // survivor masks (SuccGMDev layout): u32 (narrow) or u64 (wide) entries
    uint32_t gm_nodes;       // nodes covered by the one-deep table
    uint32_t gm2_nodes;      // nodes covered by the two-deep table
    uint32_t gm2_off;        // entry offset of the two-deep table inside `masks`
    uint32_t n3, np2, r3, gm3_off, pm3_off, pm2_off, n4, pm4_off;   // deep tables (SuccGMDev), narrow layout only
    uint32_t stack_cap;      // states per warp stack
    uint32_t feed;           // start windows fed at once when the stack runs low (engines with more than one edit)
    uint32_t text_cap;       // elements of the shared text tile (multiple of 16)
    FacCand *cands;
    uint32_t cand_cap;
    unsigned long long *counters;  // [0] next tile, [1] candidates, [2] states visited, [7] overflowed windows
    uint32_t *dirty;         // bitmap over start windows (bit sg - seg_begin): set when a state of the window did not fit the stack
};

// per-call records: ceiling = prune_len - prune_low * thr (search.rs:638-642), exact f32 ops.
// narrow: bm is u32 [n], fc_sym = first_child | in-symbol << 27;  wide: bm is u64 [n] (bit 63 = has outputs), << 26.
template <bool W>
__global__ void __launch_bounds__(256) k_succ_prepare(const void *__restrict__ bm, const uint32_t *__restrict__ fc_sym, const float *__restrict__ plen,
                                                      const float *__restrict__ plow, const uint32_t *__restrict__ out_idx, float thr, uint32_t n,
                                                      uint4 *__restrict__ rec) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float c = __fsub_rn(plen[i], __fmul_rn(plow[i], thr));
    if (W) {
        const unsigned long long b = reinterpret_cast<const unsigned long long *>(bm)[i];
        rec[i] = make_uint4((uint32_t)b, (uint32_t)(b >> 32), fc_sym[i], __float_as_uint(c));
    } else rec[i] = make_uint4(reinterpret_cast<const uint32_t *>(bm)[i], fc_sym[i], __float_as_uint(c), out_idx[i]);
}

struct SuccRecsDev {
    const uint4 *s, *g;
    uint32_t ns;
    __device__ __forceinline__ SuccRec operator()(uint32_t n) const {
        // one generic 128-bit load through a selected pointer: no divergent branch between the shared-memory
        // copy of the shallow records and the L2-resident rest
        const uint4 *p = n < ns ? s + n : g + n;
        const uint4 v = *p;
        SuccRec r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w;
        return r;
    }
};
// Survivor-mask tables (fac_succinct.h) in ONE buffer, node index fastest:
//     gmT[y][node]  (n1 nodes)   followed at `off2` by   gm2T[y1][y2][node]  (the first n2 nodes).
// Siblings are numbered consecutively and sit next to each other on the warp stack with the same text context, so the
// lanes of a pop that hold a sibling group read consecutive words (one or two 128-byte lines instead of one per lane).
// One 4- or 8-byte load per question; the index (32-bit) is selected, the load is shared between the two tables.
template <bool W>
struct SuccGMDev {
    typedef typename SuccW<W>::M M;
    const M *buf;
    uint32_t n1, n2, off2;
    // deep tables (narrow layout only; 0 nodes otherwise): gm3T[y1][y2][y3][node] at off3, pm3T[a][b][c][node] at offp3 (n3 nodes each),
    // pm2T[a][b][node] at offp2 (np2 nodes), pm4T[a][b][c][d][node] at offp4 (n4 nodes); symbol row r3, symbols clamped to
    // r3 - 1 = "no symbol"
    uint32_t n3, np2, r3, off3, offp3, offp2, n4, offp4;
    __device__ __forceinline__ bool four(uint32_t node) const { return !W && node < n4; }
    __device__ __forceinline__ bool two_deep(uint32_t node) const { return node < n2 || node < n3; }
    __device__ __forceinline__ M row3(uint32_t node, uint32_t y1, uint32_t y2, uint32_t y3) const {
        if (W) return row2(node, y1, y2);
        const uint32_t q = r3 - 1u;
        const uint32_t i3 = off3 + ((min(y1, q) * r3 + min(y2, q)) * r3 + min(y3, q)) * n3 + node;
        const uint32_t i2 = node < n2 ? off2 + (y1 * SuccW<W>::ROW + y2) * n2 + node : y1 * n1 + node;
        return node < n1 ? __ldg(buf + (node < n3 ? i3 : i2)) : ~M(0);
    }
    __device__ __forceinline__ M deep(uint32_t node, bool q_pm, uint32_t a, uint32_t b, uint32_t c, uint32_t d) const {
        if (W) return q_pm ? ~M(0) : row2(node, a, b);
        const uint32_t q = r3 - 1u;
        const uint32_t ab = min(a, q) * r3 + min(b, q), abc = ab * r3 + min(c, q);
        const uint32_t ip = node < n4 ? offp4 + (abc * r3 + min(d, q)) * n4 + node : (node < n3 ? offp3 + abc * n3 + node : offp2 + ab * np2 + node);
        const uint32_t ig = node < n3 ? off3 + abc * n3 + node : (node < n2 ? off2 + (a * SuccW<W>::ROW + b) * n2 + node : a * n1 + node);
        const bool valid = q_pm ? node < np2 : node < n1;
        return valid ? __ldg(buf + (q_pm ? ip : ig)) : ~M(0);
    }
    __device__ __forceinline__ M pm(uint32_t node, uint32_t a, uint32_t b, uint32_t c, uint32_t d) const {
        if (W) return ~M(0);
        const uint32_t q = r3 - 1u;
        const uint32_t ab = min(a, q) * r3 + min(b, q), abc = ab * r3 + min(c, q);
        const uint32_t i4 = offp4 + (abc * r3 + min(d, q)) * n4 + node;
        const uint32_t i3 = offp3 + abc * n3 + node;
        const uint32_t i2 = offp2 + ab * np2 + node;
        return node < np2 ? __ldg(buf + (node < n4 ? i4 : (node < n3 ? i3 : i2))) : ~M(0);
    }
    __device__ __forceinline__ M row(uint32_t node, uint32_t y) const {
        return node < n1 ? __ldg(buf + (y * n1 + node)) : ~M(0);
    }
    __device__ __forceinline__ M row2(uint32_t node, uint32_t y1, uint32_t y2) const {
        const uint32_t idx = node < n2 ? off2 + (y1 * SuccW<W>::ROW + y2) * n2 + node : y1 * n1 + node;
        return node < n1 ? __ldg(buf + idx) : ~M(0);
    }
};
struct SuccTextDev {
    const uint32_t *ctxw;  // text context words of the tile (succ_ctx_pack)
    uint32_t base;
    __device__ __forceinline__ uint32_t ctx(uint32_t j) const { return ctxw[j - base]; }
    __device__ __forceinline__ uint32_t byte(uint32_t j) const { return succ_ctx_byte(ctxw[j - base]); }
    __device__ __forceinline__ uint32_t sym(uint32_t j) const { return succ_ctx_s0(ctxw[j - base]); }
};
struct SuccEmitDev {
    FacCand *cands;
    uint32_t cap;
    unsigned long long *counter;
    uint32_t tag;   // haystack-window id of the tile being searched (stream batches), 0 otherwise
    __device__ __forceinline__ void operator()(uint32_t sg, uint32_t eg, uint32_t pat, float sim, uint32_t cnt) {
        const unsigned long long ci = atomicAdd(counter, 1ull);
        if (ci < cap) {
            uint4 *dst = reinterpret_cast<uint4 *>(&cands[ci]);
            dst[0] = make_uint4(sg, eg, pat, __float_as_uint(sim));
            dst[1] = make_uint4(cnt, 0u, 0u, tag);
        }
    }
};

__device__ __forceinline__ uint32_t succ_lanemask_lt() { uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

#define SUCC_WQ_CAP 96u

// LIM = limits mode: per-pattern / per-type FuzzyLimits evaluated per state (the reference's MAX_EDITS_FAST = 255 path).
// W = wide layout: 32..63 pattern symbols, 64-bit child bitmaps and masks.
template <int NT, bool LIM, bool W>
__global__ void __launch_bounds__(NT, 1) k_expand_succinct(const __grid_constant__ SuccParams P) {
    extern __shared__ __align__(16) uint8_t dyn_smem[];
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ uint32_t s_tile_idx, s_next_win;
    constexpr int NW = NT / 32;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

    // carve-up: [node records][warp stacks][walk queues][sub_pen][text context words][raw tile][sym_of]
    uint4 *s_rec = reinterpret_cast<uint4 *>(dyn_smem);
    uint4 *s_stack = s_rec + P.n_smem_nodes;
    uint4 *s_wq = s_stack + (size_t)NW * P.stack_cap;
    float *s_subpen = reinterpret_cast<float *>(s_wq + (size_t)NW * SUCC_WQ_CAP);
    typedef typename SuccW<W>::M M;
    constexpr uint32_t NOSYM = SuccW<W>::NOSYM;
    uint32_t *s_ctx = reinterpret_cast<uint32_t *>(s_subpen + SuccW<W>::ROW * SUCC_SP_STRIDE);
    uint8_t *s_raw = reinterpret_cast<uint8_t *>(s_ctx + P.text_cap);
    uint8_t *s_symof = s_raw + (size_t)P.text_cap * (P.first ? 4u : 1u);

    for (uint32_t k = tid; k < P.n_smem_nodes; k += NT) s_rec[k] = P.rec[k];
    for (uint32_t k = tid; k < SuccW<W>::ROW * SUCC_SP_STRIDE; k += NT) s_subpen[k] = P.sub_pen[k];
    for (uint32_t k = tid; k < 256; k += NT) s_symof[k] = P.sym_of[k];
    if (tid == 0) fac_mbar_init(&s_mbar, 1);
    __syncthreads();

    const SuccConsts K = P.K;
    const SuccRecsDev R{s_rec, P.rec, P.n_smem_nodes};
    const SuccGMDev<W> G{reinterpret_cast<const M *>(P.masks), P.gm_nodes, P.gm2_nodes, P.gm2_off, P.n3, P.np2, P.r3, P.gm3_off, P.pm3_off, P.pm2_off, P.n4, P.pm4_off};
    const SuccOut *out2 = reinterpret_cast<const SuccOut *>(P.out2);
    SuccEmitDev emit{P.cands, P.cand_cap, &P.counters[1], 0u};
    uint4 *const stk = s_stack + (size_t)warp * P.stack_cap;
    uint4 *const wq = s_wq + (size_t)warp * SUCC_WQ_CAP;
    const uint32_t cap = P.stack_cap;
    const uint32_t lt_mask = succ_lanemask_lt();
    uint32_t mbar_phase = 0;
    uint32_t n_states = 0;  // per-lane count of visited states (summed at the end)

    for (;;) {
        if (tid == 0) { s_tile_idx = (uint32_t)atomicAdd(&P.counters[0], 1ull); s_next_win = 0; }
        __syncthreads();
        const uint32_t t = s_tile_idx;
        if (t >= P.n_tiles) break;
        uint32_t tile_start, count, text_end;
        if (P.tiles) { const uint4 d = P.tiles[t]; tile_start = d.x; count = d.y; text_end = d.z; emit.tag = d.w; }
        else { tile_start = P.seg_begin + t * P.tile; count = min(P.tile, P.seg_end - tile_start); text_end = P.text_end; }

        // ---- stage the tile: TMA bulk copy of the 16-byte aligned body, plain loads for the tail ----
        // elements are haystack bytes (ASCII) or u32 first chars of the K1 grapheme stream (non-ASCII haystack)
        const uint32_t esz = P.first ? 4u : 1u;
        const uint8_t *src = P.first ? reinterpret_cast<const uint8_t *>(P.first) : P.text;
        uint32_t lead = (uint32_t)(((uintptr_t)(src + (size_t)tile_start * esz)) & 15u) / esz;
        if (lead > tile_start) lead = 0;
        const uint32_t base = tile_start - lead;
        const uint32_t span = count + P.lookahead + lead;                 // positions the tile must answer
        const uint32_t avail = min(span + 3u, text_end - base);           // elements that exist (a context word looks 3 ahead)
        const bool aligned = ((((uintptr_t)(src + (size_t)base * esz)) & 15u) == 0u);
        const uint32_t bulk = aligned ? ((avail * esz) & ~15u) : 0u;      // bytes
        if (bulk && tid == 0) {
            fac_fence_proxy_async();
            fac_mbar_expect_tx(&s_mbar, bulk);
            fac_tma_load_1d(s_raw, src + (size_t)base * esz, bulk, &s_mbar);
        }
        for (uint32_t k = bulk + tid; k < avail * esz; k += NT) s_raw[k] = src[(size_t)base * esz + k];
        if (bulk) { fac_mbar_wait(&s_mbar, mbar_phase & 1u); mbar_phase++; }
        __syncthreads();
        // text context words: folded first char + the dense symbols of positions k .. k+3 (fac_succinct.h)
        for (uint32_t k = tid; k < span; k += NT) {
            uint32_t b0 = 0, sy[4];
#pragma unroll
            for (uint32_t q = 0; q < 4; q++) {
                uint32_t b = 0, sq = NOSYM;
                if (k + q < avail) {
                    if (P.first) {   // already folded by K1; non-ASCII first chars match no edge and have similarity 0
                        const uint32_t c = reinterpret_cast<const uint32_t *>(s_raw)[k + q];
                        b = c < 128u ? c : SUCC_NONASCII;
                        sq = c < 128u ? s_symof[c] : NOSYM;
                    } else {
                        b = s_raw[k + q];
                        if (P.ci && b >= 'A' && b <= 'Z') b += 32u;   // to_ascii_lowercase, grapheme.rs:110-117
                        sq = s_symof[b];
                    }
                }
                if (q == 0) b0 = b;
                sy[q] = sq;
            }
            s_ctx[k] = succ_ctx_pack(b0, sy[0], sy[1], sy[2], sy[3]);
        }
        __syncthreads();
        const SuccTextDev T{s_ctx, base};

        if (P.exact_only) {
            // no edit is ever accepted (search.rs:166-168): one LANE per start window walks the exact chain
            for (uint32_t w = tid; w < count; w += NT)
                n_states += succ_walk<false, W>(K, R, out2, T, emit, tile_start + w, text_end, 0u, R(0u), 0.f, 0u, 0u, 0u);
            __syncthreads();
            continue;
        }
        // ---- windows of the tile: every warp runs ONE stack machine over a stream of start windows ----
        // A state carries its window (position word bits 20..31), so the next window's root is fed in as soon as
        // fewer than 32 states are left: pops, item rounds and walk drains stay full across window boundaries.
        {
            uint32_t top = 0, wn = 0;      // stack height, walk-queue length (warp-uniform)
            uint32_t b0 = 0, total = 0;    // item rounds of the current pop
            uint32_t off = 0;              // exclusive prefix of the lanes' item counts
            bool more = true, fed = false; // windows left in the tile; one window fed since the last pop
            SuccCtx2<W> C;
            C.bm = C.sub_m = C.del_m = 0; C.fc = C.cnt = C.pos = C.packed = C.flags = 0; C.pen = 0.f;
            for (;;) {
                __syncwarp();
                // (1) exhausted children: exact transitions only, walked 32 at a time
                if (wn >= 32u || (wn && b0 >= total && top == 0 && !more)) {
                    const uint32_t n = min(wn, 32u);
                    if (lane < n) {
                        const uint4 q = wq[wn - n + lane];
                        n_states += succ_walk<LIM, W>(K, R, out2, T, emit, tile_start + (q.w >> 20), text_end, q.x, R(q.x), __uint_as_float(q.y), q.z,
                                                      succ_jr(q.w), succ_mr(q.w));
                    }
                    wn -= n;
                    continue;
                }
                // (2) surviving (state, child) pairs of the last pop, 32 per round
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
                    SuccCtx2<W> O;
                    O.bm = __shfl_sync(0xFFFFFFFFu, C.bm, lo);
                    O.fc = __shfl_sync(0xFFFFFFFFu, C.fc, lo);
                    O.pen = __shfl_sync(0xFFFFFFFFu, C.pen, lo);
                    O.cnt = __shfl_sync(0xFFFFFFFFu, C.cnt, lo);
                    O.pos = __shfl_sync(0xFFFFFFFFu, C.pos, lo);
                    O.packed = __shfl_sync(0xFFFFFFFFu, C.packed, lo);
                    O.flags = __shfl_sync(0xFFFFFFFFu, C.flags, lo);
                    O.sub_m = __shfl_sync(0xFFFFFFFFu, C.sub_m, lo);
                    O.del_m = __shfl_sync(0xFFFFFFFFu, C.del_m, lo);
                    const uint32_t r = it - __shfl_sync(0xFFFFFFFFu, off, lo);
                    FacState c;
                    c.node = 0; c.pen = 0.f; c.cnt = 0; c.pos = 0;
                    const bool ok = it < total && succ_item2<W>(K, s_subpen, O, r, c);
                    const bool to_walk = ok && (O.flags & SUCC_F_LAST);
                    const bool to_stack = ok && !(O.flags & SUCC_F_LAST);
                    const uint32_t bw = __ballot_sync(0xFFFFFFFFu, to_walk), bs = __ballot_sync(0xFFFFFFFFu, to_stack);
                    const uint4 cv = make_uint4(c.node, __float_as_uint(c.pen), c.cnt, c.pos);
                    if (to_walk) wq[wn + __popc(bw & lt_mask)] = cv;
                    if (to_stack) stk[top + __popc(bs & lt_mask)] = cv;
                    wn += __popc(bw); top += __popc(bs);
                    continue;
                }
                // (3) feed more start windows when the stack runs low: one root at a time when a root fans out into
                //     dozens of children, a whole warp's worth when the root is already on its last edit (edits(1))
                if (more && !fed && top < 32u) {
                    const uint32_t nf = K.mef <= 1 ? 32u - top : min(P.feed, 32u - top);
                    uint32_t w0 = 0;
                    if (lane == 0) w0 = atomicAdd(&s_next_win, nf);
                    w0 = __shfl_sync(0xFFFFFFFFu, w0, 0);
                    if (w0 >= count) { more = false; continue; }
                    if (w0 + nf >= count) more = false;   // this fetch takes the tail of the tile
                    const uint32_t w = w0 + lane;
                    bool push = lane < nf && w < count;
                    if (push && P.wskip) {  // 2-gram window skip (search.rs:535-553); result-neutral.  Only ASCII first chars take part
                        const uint32_t start = tile_start + w;
                        const uint32_t cw = T.ctx(start);
                        if (succ_ctx_byte(cw) != SUCC_NONASCII && !((P.first_mask >> succ_ctx_s0(cw)) & 1u))
                            push = !(start + 1 >= text_end || (T.byte(start + 1) != SUCC_NONASCII && !((P.second_mask >> succ_ctx_s1(cw)) & 1u)));
                    }
                    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, push);
                    if (push) stk[top + __popc(bal & lt_mask)] = make_uint4(0u, 0u, 0u, w << 20);
                    top += __popc(bal);
                    fed = bal != 0u;
                    continue;
                }
                if (top == 0) {
                    if (!more && wn == 0) break;
                    fed = false;   // nothing to pop: feed again, or fall into the final walk drain
                    continue;
                }
                // (4) pop up to 32 states
                fed = false;
                const uint32_t navail = min(top, 32u);
                const bool has = lane < navail;
                uint4 sv = make_uint4(0, 0, 0, 0);
                SuccRec rec; rec.x = rec.y = rec.z = 0; rec.w = FAC_NONE;
                if (has) { sv = stk[top - 1u - lane]; rec = R(sv.x); }
                const uint32_t start = tile_start + (sv.w >> 20);
                const float pen = __uint_as_float(sv.y);
                const bool dead = !has || pen > succ_ceil<W>(rec);   // node ceiling, search.rs:638-642
                const bool last = (int)fac_edits_of(sv.z) + 1 >= K.mef;
                // children of every live candidate first (no side effects), so that the stack check below uses the exact
                // number of pushes instead of the worst case 2 * fan-out + 3
                bool p_ex = false, p_sw = false, p_in = false;
                FacState c_ex, c_sw, c_in;
                c_ex.node = c_sw.node = c_in.node = 0; c_ex.pen = c_sw.pen = c_in.pen = 0.f;
                c_ex.cnt = c_sw.cnt = c_in.cnt = 0; c_ex.pos = c_sw.pos = c_in.pos = 0;
                C.sub_m = C.del_m = 0;
                if (!dead) {
                    succ_make_ctx2<LIM, W>(K, T, G, start, text_end, sv.x, rec, pen, sv.z, sv.w, C);
                    const uint32_t jr = succ_jr(sv.w);
                    const uint32_t cur_s = succ_ctx_s0(C.packed);
                    if (C.flags & SUCC_F_EXACT) {   // exact transition, search.rs:776-798 (unless the child provably cannot emit)
                        p_ex = true;
                        c_ex.node = succ_child<W>(rec, cur_s); c_ex.pen = pen; c_ex.cnt = sv.z; c_ex.pos = succ_repos(sv.w, jr + 1, jr + 1);
                    }
                    p_sw = succ_swap2<LIM, W>(K, R, C, c_sw);
                    p_in = succ_ins2<LIM, W>(K, C, sv.x, c_in);
                }
                uint32_t n_items = succ_popc(C.sub_m) + succ_popc(C.del_m);
                // stack pushes of this state: a state on its last edit only pushes its exact child (the rest goes to the walk queue)
                const uint32_t ub = (uint32_t)p_ex + (last ? 0u : (uint32_t)p_sw + (uint32_t)p_in + n_items);
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
                if (n_pop == 0) {  // the top state alone does not fit: its window goes to the faithful kernel, the state is dropped
                    if (lane == 0) {
                        const uint32_t wabs = start - P.seg_begin;
                        atomicOr(&P.dirty[wabs >> 5], 1u << (wabs & 31u));
                        atomicAdd(&P.counters[7], 1ull);
                    }
                    top -= 1u;
                    b0 = total = 0;
                    continue;
                }
                const bool active = lane < n_pop && !dead;
                top -= n_pop;
                __syncwarp();
                if (active) {
                    n_states++;
                    if (succ_has_out<W>(rec)) succ_outputs<LIM>(K, out2, emit, succ_out_idx<W>(K, rec, sv.x), pen, sv.z, start, start + succ_mr(sv.w));
                } else {   // not popped in this round (or dead): nothing to push
                    p_ex = p_sw = p_in = false;
                    C.sub_m = C.del_m = 0;
                    n_items = 0;
                }
                {
                    // exact children and the swap / insertion children of states with budget left go back on the stack,
                    // the exhausted swap / insertion children into the walk queue: five ballots, no branches
                    const bool lw = active && last;
                    const uint32_t b_ex = __ballot_sync(0xFFFFFFFFu, p_ex);
                    const uint32_t b_sw_w = __ballot_sync(0xFFFFFFFFu, p_sw && lw), b_in_w = __ballot_sync(0xFFFFFFFFu, p_in && lw);
                    const uint32_t b_sw_s = __ballot_sync(0xFFFFFFFFu, p_sw && !lw), b_in_s = __ballot_sync(0xFFFFFFFFu, p_in && !lw);
                    if (p_ex) stk[top + __popc(b_ex & lt_mask)] = make_uint4(c_ex.node, __float_as_uint(c_ex.pen), c_ex.cnt, c_ex.pos);
                    top += __popc(b_ex);
                    const uint4 v_sw = make_uint4(c_sw.node, __float_as_uint(c_sw.pen), c_sw.cnt, c_sw.pos);
                    const uint4 v_in = make_uint4(c_in.node, __float_as_uint(c_in.pen), c_in.cnt, c_in.pos);
                    if (p_sw && lw) wq[wn + __popc(b_sw_w & lt_mask)] = v_sw;
                    wn += __popc(b_sw_w);
                    if (p_in && lw) wq[wn + __popc(b_in_w & lt_mask)] = v_in;
                    wn += __popc(b_in_w);
                    if (b_sw_s | b_in_s) {
                        if (p_sw && !lw) stk[top + __popc(b_sw_s & lt_mask)] = v_sw;
                        top += __popc(b_sw_s);
                        if (p_in && !lw) stk[top + __popc(b_in_s & lt_mask)] = v_in;
                        top += __popc(b_in_s);
                    }
                }
                off = n_items;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, off, d);
                    if (lane >= (uint32_t)d) off += v;
                }
                total = __shfl_sync(0xFFFFFFFFu, off, 31);
                off -= n_items;  // exclusive
                b0 = 0;
            }
        }
        __syncthreads();  // every warp is done with the tile before it is restaged
    }
    n_states = __reduce_add_sync(0xFFFFFFFFu, n_states);
    if (lane == 0 && n_states) atomicAdd(&P.counters[2], (unsigned long long)n_states);
}
