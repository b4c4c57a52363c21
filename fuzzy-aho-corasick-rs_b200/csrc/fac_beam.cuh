// fac_beam.cuh -- K3b: beamed frontier expansion (src/search.rs:577-589), one CTA per start window.
//
// The reference checks the beam before EVERY pop: if more than 2*bw states are un-popped it keeps
// the bw lowest-penalty ones and truncates the queue.  The pop loop is replayed in chunks of up to
// FAC_BLOCK un-popped states: a chunk is evaluated speculatively (dedup verdicts, push counts), a
// scan gives the queue length every state would have seen at its pop, the first state that trips
// the beam splits the chunk, the states before it are committed (visited map, outputs, children in
// FIFO order) and the rest of the queue is cut to the bw lowest by (penalty total order, queue
// position), preserving queue order (DESIGN.md unpinned item U3: the reference's
// select_nth_unstable_by leaves the choice among equal penalties and the resulting permutation to
// the Rust standard library).  Same-key states can recur at any distance in a beamed queue, so the
// visited map is persistent for the window: an append-only log of (key, min penalty) entries in
// the CTA's scratch, indexed by an open-addressing table of log indices (an entry is published only
// after it is fully written, so readers never see a partial key).
//
// Scratch split of the CTA's queue region (qcap states): [0,Q) queue | [Q,2Q) cut staging |
// [2Q,4Q) visited log, with Q = qcap/4.  With bw == 0 the kernel is the exact search, one window per
// CTA (used to locate the auto_beam crossing window, src/search.rs:1096-1103).
#pragma once
#include "fac_kernels.cuh"

// BS = threads per CTA: 32 (one warp per start window: no cross-warp barriers, 32 windows in flight per SM; the
// default) or 256 (big-queue retry of windows that overflow the small per-warp scratch).
template <int BS>
__global__ void __launch_bounds__(BS) k_expand_beam(const __grid_constant__ ExpandParams P, const uint32_t bw) {
    constexpr int NWB = BS / 32;
    __shared__ uint32_t s_scan[2][NWB + 1];
    __shared__ uint32_t s_tile_idx, s_cut, s_vcount;
    __shared__ uint32_t c_node[BS], c_cnt[BS], c_pos[BS], c_exact[BS], c_flags[BS];
    __shared__ float c_pen[BS];
    __shared__ uint32_t k_node[BS], k_cnt[BS], k_pos[BS];
    __shared__ float k_pen[BS];
    __shared__ uint8_t k_exp[BS];
    __shared__ uint32_t s_off[BS + 1], s_pc[BS], s_pb[BS + 1];

    const AutomatonView &A = P.A;
    const uint32_t tid = threadIdx.x;
    uint32_t parity = 0;
    const uint32_t Q = P.qcap / 4u;
    FacState *const queue = P.queue + (size_t)blockIdx.x * P.qcap;
    FacState *const stage = queue + Q;
    FacState *const vlog = queue + 2u * Q;
    const uint32_t vcap = 2u * Q;
    uint32_t *const vslot = P.hslot + (size_t)blockIdx.x * P.qcap;  // table slot of log entry k (for the reset)
    uint32_t *const vtab = P.gtab_rep + (size_t)blockIdx.x * P.gtab_size;
    const uint32_t tmask = P.gtab_size - 1u;
    const float INF = __int_as_float(0x7F800000);

    FacTextDirect T;
    T.tv = P.tv; T.ascii_gid = A.ascii_gid; T.ci = A.ci;

    for (;;) {
        if (tid == 0) { s_tile_idx = (uint32_t)atomicAdd(&P.counters[0], 1ull); s_vcount = 0; }
        __syncthreads();
        const uint32_t t_idx = s_tile_idx;
        if (t_idx >= P.n_tiles) break;
        uint32_t start, text_end, win_tag = P.pass << 31;
        if (P.mode == 0) { start = P.seg_begin + t_idx; text_end = P.text_end; }
        else { const uint4 d = P.tiles[t_idx]; start = d.x; text_end = d.z; win_tag |= d.w; }
        const bool has1 = start + 1 < text_end;
        const bool skipped = fac_window_skipped(A, T.first(start), has1, has1 ? T.first(start + 1) : 0u);
        uint32_t h = 0, t = 0;  // un-popped states live in queue[h, t); t is the reference's queue.len()
        bool failed = false;
        if (!skipped) {
            if (tid == 0) { FacState S; S.node = 0; S.pen = 0.f; S.cnt = 0; S.pos = 0; queue[0] = S; }
            t = 1;
        }
        __syncthreads();
        while (h < t && !failed) {
            const uint32_t m = min(t - h, (uint32_t)BS);
            // ---- load the chunk ----
            FacState S; S.node = 0; S.pen = 0.f; S.cnt = 0; S.pos = 0;
            if (tid < m) {
                const uint4 q = *reinterpret_cast<const uint4 *>(&queue[h + tid]);
                S.node = q.x; S.pen = __uint_as_float(q.y); S.cnt = q.z; S.pos = q.w;
                k_node[tid] = S.node; k_cnt[tid] = S.cnt; k_pos[tid] = S.pos; k_pen[tid] = S.pen;
            }
            s_pc[tid] = 0;
            if (tid == 0) s_cut = m;
            __syncthreads();
            // ---- speculative verdicts: visited log + earlier states of this chunk (search.rs:608-628) ----
            uint32_t nslots = 0;
            bool expanded = false;
            if (tid < m) {
                float mn = INF;
                uint32_t hh = fac_hash3(S.node, S.cnt, S.pos) & tmask;
                for (;;) {
                    const uint32_t r = vtab[hh];
                    if (r == FAC_EMPTY) break;
                    const uint4 e = *reinterpret_cast<const uint4 *>(&vlog[r]);
                    if (e.x == S.node && e.z == S.cnt && e.w == S.pos) { mn = __uint_as_float(e.y); break; }
                    hh = (hh + 1u) & tmask;
                }
                for (uint32_t k = 0; k < tid; k++)
                    if (k_node[k] == S.node && k_cnt[k] == S.cnt && k_pos[k] == S.pos) mn = fminf(mn, k_pen[k]);
                expanded = !(mn <= S.pen);
                if (expanded && !fac_over_ceiling(A, S.node, S.pen, P.thr)) {
                    FacCtx Cx;
                    fac_make_ctx(A, T, P.maxpen, start, text_end, S, Cx);
                    nslots = Cx.nslots;
                    c_node[tid] = Cx.node; c_pen[tid] = Cx.pen; c_cnt[tid] = Cx.cnt; c_pos[tid] = Cx.pos;
                    c_exact[tid] = Cx.exact; c_flags[tid] = Cx.flags;
                }
            }
            k_exp[tid] = expanded ? 1 : 0;
            uint32_t W;
            const uint32_t excl = fac_block_scan_t<NWB>(nslots, s_scan, parity, W);
            s_off[tid] = excl;
            if (tid == 0) s_off[BS] = W;
            __syncthreads();
            // ---- pass 1: push count of every state ----
            for (uint32_t k0 = 0; k0 < W; k0 += BS) {
                const uint32_t k = k0 + tid;
                if (k < W) {
                    uint32_t lo = 0, hi = BS;
                    while (hi - lo > 1u) { const uint32_t mid = (lo + hi) >> 1; if (s_off[mid] <= k) lo = mid; else hi = mid; }
                    FacCtx Cx;
                    Cx.node = c_node[lo]; Cx.pen = c_pen[lo]; Cx.cnt = c_cnt[lo]; Cx.pos = c_pos[lo];
                    Cx.exact = c_exact[lo]; Cx.flags = c_flags[lo]; Cx.nslots = 0;
                    FacState child;
                    if (fac_eval_slot(A, T, P.maxpen, start, text_end, Cx, k - s_off[lo], child)) atomicAdd(&s_pc[lo], 1u);
                }
            }
            __syncthreads();
            uint32_t total_push;
            const uint32_t pb = fac_block_scan_t<NWB>(s_pc[tid], s_scan, parity, total_push);  // pushes before state tid
            s_pb[tid] = pb;
            if (tid == 0) s_pb[BS] = total_push;
            // ---- where does the beam trip?  remaining = queue.len() - q_idx at the pop of state tid ----
            if (bw && tid < m) {
                const uint32_t remaining = (t + pb) - (h + tid);
                if (remaining > 2u * bw) atomicMin(&s_cut, tid);
            }
            __syncthreads();
            const uint32_t cm = s_cut;  // states [0, cm) of the chunk are committed
            const bool cut = cm < m;
            const uint32_t commit_push = s_pb[cm];  // s_pb[m] == total_push: threads >= m push nothing
            if (t + commit_push > Q) failed = true;
            if (!failed) {
                // ---- commit: visited log and outputs of the committed prefix ----
                if (tid < cm && expanded) {
                    // only the last expanded occurrence of a key in the committed prefix carries its minimum
                    bool last = true;
                    for (uint32_t k = tid + 1; k < cm; k++)
                        if (k_exp[k] && k_node[k] == S.node && k_cnt[k] == S.cnt && k_pos[k] == S.pos) { last = false; break; }
                    if (last) {
                        uint32_t hh = fac_hash3(S.node, S.cnt, S.pos) & tmask;
                        uint32_t mine = FAC_EMPTY;
                        for (;;) {
                            uint32_t r = *((volatile uint32_t *)&vtab[hh]);
                            if (r == FAC_EMPTY) {
                                if (mine == FAC_EMPTY) {
                                    mine = atomicAdd(&s_vcount, 1u);
                                    if (mine >= vcap) break;
                                    uint4 e; e.x = S.node; e.y = __float_as_uint(S.pen); e.z = S.cnt; e.w = S.pos;
                                    *reinterpret_cast<uint4 *>(&vlog[mine]) = e;
                                    vslot[mine] = FAC_EMPTY;
                                    __threadfence_block();
                                }
                                const uint32_t old = atomicCAS(&vtab[hh], FAC_EMPTY, mine);
                                if (old == FAC_EMPTY) { vslot[mine] = hh; break; }
                                r = old;
                            }
                            const uint4 e = *reinterpret_cast<const uint4 *>(&vlog[r]);
                            if (e.x == S.node && e.z == S.cnt && e.w == S.pos) {
                                vlog[r].pen = S.pen;  // expanded => strictly below the recorded minimum
                                break;
                            }
                            hh = (hh + 1u) & tmask;
                        }
                    }
                    if (!fac_over_ceiling(A, S.node, S.pen, P.thr)) {
                        const uint32_t o1 = A.node_out_off[S.node + 1];
                        for (uint32_t o = A.node_out_off[S.node]; o < o1; o++) {
                            const uint32_t pat = A.out_pat[o];
                            float sim;
                            if (fac_eval_output(A, P.thr, pat, S.pen, S.cnt, sim)) {
                                const unsigned long long ci = atomicAdd(&P.counters[1], 1ull);
                                if (ci < P.cand_cap) {
                                    FacCand cd;
                                    cd.sg = start; cd.eg = start + (S.pos & FAC_POS_MASK); cd.pat = pat; cd.sim = sim;
                                    cd.cnt = S.cnt; cd.seq = h + tid; cd.tile = t_idx; cd.tag = win_tag;
                                    uint4 *dst = reinterpret_cast<uint4 *>(&P.cands[ci]);
                                    dst[0] = reinterpret_cast<uint4 *>(&cd)[0];
                                    dst[1] = reinterpret_cast<uint4 *>(&cd)[1];
                                }
                            }
                        }
                    }
                }
                // ---- commit: children of the committed prefix, appended in FIFO order at queue[t..] ----
                const uint32_t Wc = s_off[cm];  // work items of the committed states (s_off[BS] == W)
                uint32_t nbase = t;
                for (uint32_t k0 = 0; k0 < Wc; k0 += BS) {
                    const uint32_t k = k0 + tid;
                    bool push = false;
                    FacState child;
                    if (k < Wc) {
                        uint32_t lo = 0, hi = BS;
                        while (hi - lo > 1u) { const uint32_t mid = (lo + hi) >> 1; if (s_off[mid] <= k) lo = mid; else hi = mid; }
                        FacCtx Cx;
                        Cx.node = c_node[lo]; Cx.pen = c_pen[lo]; Cx.cnt = c_cnt[lo]; Cx.pos = c_pos[lo];
                        Cx.exact = c_exact[lo]; Cx.flags = c_flags[lo]; Cx.nslots = 0;
                        push = fac_eval_slot(A, T, P.maxpen, start, text_end, Cx, k - s_off[lo], child);
                    }
                    uint32_t total;
                    const uint32_t r = fac_block_rank_t<NWB>(push, s_scan, parity, total);
                    if (push) *reinterpret_cast<uint4 *>(&queue[nbase + r]) = *reinterpret_cast<uint4 *>(&child);
                    nbase += total;
                }
                __syncthreads();
                if (!cut) { h += m; t = nbase; }
                else {
                    // ---- beam cut over queue[hc, nbase): keep the bw lowest (pen, position), in queue order ----
                    const uint32_t hc = h + cm, ncut = nbase - hc;
                    uint32_t kept = 0;
                    for (uint32_t e0 = 0; e0 < ncut; e0 += BS) {
                        const uint32_t e = e0 + tid;
                        bool keep = false;
                        uint4 me = make_uint4(0, 0, 0, 0);
                        if (e < ncut) me = *reinterpret_cast<const uint4 *>(&queue[hc + e]);
                        const uint32_t mykey = fac_total_order_u32(__uint_as_float(me.y));
                        // rank among the un-popped states by (penalty total order, queue position); the keys of
                        // 256 states at a time are staged in shared memory (s_pc is free until the next chunk)
                        uint32_t rank = 0;
                        for (uint32_t f0 = 0; f0 < ncut; f0 += BS) {
                            __syncthreads();
                            s_pc[tid] = (f0 + tid < ncut) ? fac_total_order_u32(queue[hc + f0 + tid].pen) : 0xFFFFFFFFu;
                            __syncthreads();
                            const uint32_t nf = min((uint32_t)BS, ncut - f0);
                            for (uint32_t f = 0; f < nf; f++) {
                                const uint32_t ok = s_pc[f];
                                rank += (ok < mykey || (ok == mykey && f0 + f < e)) ? 1u : 0u;
                            }
                        }
                        keep = e < ncut && rank < bw;
                        uint32_t total;
                        const uint32_t r = fac_block_rank_t<NWB>(keep, s_scan, parity, total);
                        if (keep) *reinterpret_cast<uint4 *>(&stage[kept + r]) = me;
                        kept += total;
                    }
                    __syncthreads();
                    for (uint32_t e = tid; e < kept; e += BS)
                        *reinterpret_cast<uint4 *>(&queue[hc + e]) = *reinterpret_cast<const uint4 *>(&stage[e]);
                    h = hc; t = hc + kept;  // queue.truncate(q_idx + bw)
                }
            }
            __syncthreads();
        }
        // ---- reset the visited table; account the window ----
        const uint32_t vc = min(s_vcount, vcap);
        for (uint32_t k = tid; k < vc; k += BS) { const uint32_t hh = vslot[k]; if (hh != FAC_EMPTY) vtab[hh] = FAC_EMPTY; }
        if (s_vcount > vcap) failed = true;
        if (tid == 0) {
            if (failed) {
                const unsigned long long fi = atomicAdd(&P.counters[3], 1ull);
                if (fi < P.failed_cap) P.failed_tiles[fi] = t_idx;
                if (P.failed_bitmap) atomicOr(&P.failed_bitmap[t_idx >> 5], 1u << (t_idx & 31u));
            } else {
                atomicAdd(&P.counters[2], (unsigned long long)t);
                if (P.per_window) P.per_window[start] = t;
            }
        }
        __syncthreads();
    }
}
