// fac_beam2.cuh -- K3b, shared-memory variant: beamed frontier expansion (src/search.rs:577-589), one WARP per start
// window with the whole window state on chip.
//
// Same semantics as k_expand_beam (fac_beam.cuh): the reference pops one state at a time, checks the beam before EVERY
// pop (`queue.len() - q_idx > 2 * bw` -> keep the bw lowest penalties, truncate), dedups through the visited map
// (search.rs:608-628) and pushes children in a fixed order; here a chunk of up to 32 un-popped states is evaluated
// speculatively, a scan gives the queue length every state would have seen at its pop, the first state that trips the
// beam splits the chunk, the states before it are committed and the rest of the queue is cut.  What changed is where
// the state lives and how much is read per state:
//
//   * FIFO queue = a 512-entry ring in shared memory (the un-popped part never exceeds 2*bw + one state's pushes);
//   * visited map = a 512- or 1024-slot open-addressing table in shared memory (64-bit packed key + f32 penalty);
//   * children are evaluated ONCE: they are written behind the tail speculatively, per-parent push counts come from
//     shared-memory atomics, and a cut simply truncates the tail to the committed prefix;
//   * the cut is a 4-pass radix select on the total-order image of the penalty (shared-memory histogram) followed by
//     an order-preserving in-place compaction: K lowest by (penalty, queue position), queue order kept (DESIGN.md,
//     unpinned item U3: the reference's select_nth_unstable_by leaves the choice among equal penalties to std);
//   * per-state arithmetic over the merged 16-byte node / edge records of fac_flat.h (one load per state, one per
//     child) with the short output-children lists / survivor masks for states on their last edit.
//
// Domain: engines with MAX_EDITS_FAST 1..6 (the packed key holds 3 bits per edit count) and 2*bw + 2*fan-out <= ring.
// Windows that overflow the ring or the visited table are reported as failed tiles and redone by k_expand_beam<256>.
#pragma once
#include "fac_flat.h"
#include "fac_kernels.cuh"

#define BM2_QCAP 512u
// visited slots per window (template VCAP: 512 or 1024); a window that records more than 3/4 of them is handed over
#define BM2_TEXT 64u      // graphemes of the window staged in shared memory (first chars + grapheme ids)
#define BM2_SMEM_PER_WARP(VCAP) (BM2_QCAP * 16u + (VCAP) * 12u + 256u * 4u + 32u * 4u + BM2_TEXT * 8u)

// The window's own graphemes from shared memory (already case folded / translated to grapheme ids), anything beyond from
// the global streams.
struct WarpText {
    const uint32_t *sf, *sg;
    uint32_t base, len;
    FacTextDirect G;
    __device__ __forceinline__ uint32_t first(uint32_t j) const { const uint32_t r = j - base; return r < len ? sf[r] : G.first(j); }
    __device__ __forceinline__ uint32_t gid(uint32_t j) const { const uint32_t r = j - base; return r < len ? sg[r] : G.gid(j); }
};

struct Beam2Params {
    ExpandParams E;
    FlatView F;
};

__device__ __forceinline__ unsigned long long bm2_key(uint32_t node, uint32_t cnt, uint32_t pos) {
    const uint32_t c = (cnt & 7u) | (((cnt >> 8) & 7u) << 3) | (((cnt >> 16) & 7u) << 6) | (((cnt >> 24) & 7u) << 9);
    return 0x8000000000000000ull | ((unsigned long long)node << 32) | ((unsigned long long)(pos & 0xFFFFFu) << 12) | c;
}
__device__ __forceinline__ uint32_t bm2_hash(unsigned long long k) {
    k ^= k >> 31; k *= 0x9E3779B97F4A7C15ull; k ^= k >> 29;
    return (uint32_t)k;
}

template <uint32_t BM2_VCAP>
__global__ void __launch_bounds__(256) k_beam_warp(const __grid_constant__ Beam2Params BP, const uint32_t bw) {
    constexpr uint32_t BM2_VMAX = BM2_VCAP / 4u * 3u;
    extern __shared__ __align__(16) uint8_t dyn_smem[];   // per warp: [ring][visited keys][visited penalties][histogram][push counts]
    const ExpandParams &P = BP.E;
    const AutomatonView &A = P.A;
    const FlatView F = BP.F;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint8_t *const mine_smem = dyn_smem + (size_t)warp * BM2_SMEM_PER_WARP(BM2_VCAP);
    uint4 *const Q = reinterpret_cast<uint4 *>(mine_smem);
    unsigned long long *const vkey = reinterpret_cast<unsigned long long *>(mine_smem + BM2_QCAP * 16u);
    float *const vpen = reinterpret_cast<float *>(mine_smem + BM2_QCAP * 16u + BM2_VCAP * 8u);
    uint32_t *const hist = reinterpret_cast<uint32_t *>(mine_smem + BM2_QCAP * 16u + BM2_VCAP * 12u);
    uint32_t *const pc = hist + 256;
    uint32_t *const s_first = pc + 32, *const s_gid = s_first + BM2_TEXT;
    constexpr uint32_t QM = BM2_QCAP - 1u, VM = BM2_VCAP - 1u;
    const float INF = __int_as_float(0x7F800000);

    WarpText T;
    T.G.tv = P.tv; T.G.ascii_gid = A.ascii_gid; T.G.ci = A.ci;
    T.sf = s_first; T.sg = s_gid; T.base = 0; T.len = 0;

    for (;;) {
        uint32_t t_idx = 0;
        if (lane == 0) t_idx = (uint32_t)atomicAdd(&P.counters[0], 1ull);
        t_idx = __shfl_sync(0xFFFFFFFFu, t_idx, 0);
        if (t_idx >= P.n_tiles) break;
        uint32_t start, text_end, win_tag = P.pass << 31;
        if (P.mode == 0) { start = P.seg_begin + t_idx; text_end = P.text_end; }
        else { const uint4 d = P.tiles[t_idx]; start = d.x; text_end = d.z; win_tag |= d.w; }
        for (uint32_t k = lane; k < BM2_VCAP; k += 32u) vkey[k] = 0ull;
        {   // stage the window's graphemes: every later text access of the window is a shared-memory load
            const uint32_t len = min(BM2_TEXT, text_end - start);
            __syncwarp();
            for (uint32_t k = lane; k < len; k += 32u) { s_first[k] = T.G.first(start + k); if (A.has_mappings) s_gid[k] = T.G.gid(start + k); }
            T.base = start; T.len = len;
            __syncwarp();
        }
        const bool has1 = start + 1 < text_end;
        const bool skipped = fac_window_skipped(A, T.first(start), has1, has1 ? T.first(start + 1) : 0u);
        uint32_t head = 0, tail = 0, vcount = 0, mcap = 32u;   // un-popped states live in ring positions [head, tail); tail = queue.len()
        bool failed = false;
        if (!skipped) { if (lane == 0) Q[0] = make_uint4(0u, 0u, 0u, 0u); tail = 1; }
        __syncwarp();

        while (head < tail && !failed) {
            const uint32_t m = min(tail - head, mcap);
            const bool has = lane < m;
            FacState S;
            S.node = 0; S.pen = 0.f; S.cnt = 0; S.pos = 0;
            if (has) { const uint4 q = Q[(head + lane) & QM]; S.node = q.x; S.pen = __uint_as_float(q.y); S.cnt = q.z; S.pos = q.w; }
            const unsigned long long key = has ? bm2_key(S.node, S.cnt, S.pos) : (unsigned long long)lane;   // dummies are distinct
            pc[lane] = 0u;
            // ---- speculative verdict: visited map + earlier same-key states of this chunk (search.rs:608-628) ----
            float mn = INF;
            if (has) {
                uint32_t h = bm2_hash(key) & VM;
                for (;;) {
                    const unsigned long long k = vkey[h];
                    if (k == 0ull) break;
                    if (k == key) { mn = vpen[h]; break; }
                    h = (h + 1u) & VM;
                }
            }
            const uint32_t same_all = __match_any_sync(0xFFFFFFFFu, key);
            uint32_t same = same_all & lt_mask;
            while (__any_sync(0xFFFFFFFFu, same != 0u)) {
                const uint32_t src = same ? (31u - (uint32_t)__clz(same)) : lane;
                const float v = __shfl_sync(0xFFFFFFFFu, S.pen, src);
                if (same) { mn = fminf(mn, v); same &= ~(1u << src); }
            }
            const bool expanded = has && !(mn <= S.pen);
            FlatRec nr;
            nr.x = nr.y = nr.w = 0u; nr.z = 0xFF800000u;
            if (expanded) { const uint4 v = reinterpret_cast<const uint4 *>(F.nrec)[S.node]; nr.x = v.x; nr.y = v.y; nr.z = v.z; nr.w = v.w; }
            const bool live = expanded && !(S.pen > __uint_as_float(nr.z));   // node ceiling, search.rs:638-642
            FlatCtx C;
            C.node = C.cnt = C.pos = C.eoff = C.shape = C.lists = C.nslots = 0; C.exact = C.row = FAC_NONE; C.pen = 0.f;
            if (live) flat_make_ctx<false>(A, F, T, P.maxpen, start, text_end, S, nr, C);
            uint32_t off = C.nslots;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, off, d);
                if (lane >= (uint32_t)d) off += v;
            }
            const uint32_t total = __shfl_sync(0xFFFFFFFFu, off, 31);
            off -= C.nslots;
            __syncwarp();
            // ---- children, evaluated once and written behind the tail in FIFO order (owner, slot) ----
            uint32_t base = tail;
            bool overflow = false;
            for (uint32_t b0 = 0; b0 < total; b0 += 32u) {
                const uint32_t it = b0 + lane;
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
                const bool ok = it < total && flat_eval_slot<false>(A, F, T, P.maxpen, start, text_end, O, r, c);
                const uint32_t bal = __ballot_sync(0xFFFFFFFFu, ok);
                if (base + __popc(bal) - head > BM2_QCAP) { overflow = true; break; }
                if (ok) {
                    Q[(base + __popc(bal & lt_mask)) & QM] = make_uint4(c.node, __float_as_uint(c.pen), c.cnt, c.pos);
                    atomicAdd(&pc[lo], 1u);
                }
                base += __popc(bal);
            }
            __syncwarp();
            if (overflow) {   // the speculative children do not fit the ring: smaller chunk, or hand the window over
                if (mcap > 1u) { mcap >>= 1; continue; }
                failed = true;
                break;
            }
            // ---- where does the beam trip?  remaining = queue.len() - q_idx at the pop of state `lane` ----
            const uint32_t mypc = pc[lane];
            uint32_t pb = mypc;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, pb, d);
                if (lane >= (uint32_t)d) pb += v;
            }
            pb -= mypc;   // pushes before this state
            uint32_t cm = m;
            if (bw) {
                const uint32_t viol = __ballot_sync(0xFFFFFFFFu, has && (tail + pb) - (head + lane) > 2u * bw);
                if (viol) cm = (uint32_t)__ffs(viol) - 1u;
            }
            const uint32_t commit_push = cm == m ? base - tail : __shfl_sync(0xFFFFFFFFu, pb, cm & 31u);
            // ---- commit the prefix [0, cm): visited map (the last expanded occurrence of a key carries its minimum) ----
            {
                const bool committed = lane < cm && expanded;
                const uint32_t cmask = __ballot_sync(0xFFFFFFFFu, committed);
                const uint32_t later = same_all & ~lt_mask & ~(1u << lane) & cmask;
                bool fresh = false;
                if (committed && !later) {
                    uint32_t h = bm2_hash(key) & VM;
                    for (;;) {
                        const unsigned long long k = vkey[h];
                        if (k == key) { vpen[h] = S.pen; break; }      // expanded => strictly below the recorded minimum
                        if (k == 0ull) {
                            const unsigned long long old = atomicCAS(&vkey[h], 0ull, key);
                            if (old == 0ull || old == key) { vpen[h] = S.pen; fresh = old == 0ull; break; }
                        }
                        h = (h + 1u) & VM;
                    }
                }
                vcount += __popc(__ballot_sync(0xFFFFFFFFu, fresh));
                if (vcount > BM2_VMAX) failed = true;
            }
            // ---- commit: outputs of the prefix (search.rs:659-737) ----
            if (lane < cm && live) {
                const uint32_t no = flat_nout(nr);
                for (uint32_t o = 0; o < no; o++) {
                    const uint32_t pat = A.out_pat[nr.w + o];
                    float sim;
                    if (fac_eval_output(A, P.thr, pat, S.pen, S.cnt, sim))
                        fac_emit_cand(P, start, start + (S.pos & FAC_POS_MASK), pat, sim, S.cnt, head + lane, t_idx, win_tag);
                }
            }
            __syncwarp();
            if (cm == m) { head += m; tail = base; mcap = 32u; continue; }
            // ---- beam cut over the un-popped states [hc, tail): keep the bw lowest (pen, position), in queue order ----
            tail += commit_push;
            const uint32_t hc = head + cm, n = tail - hc;
            uint32_t prefix = 0, need = bw;
#pragma unroll 1
            for (int pass = 0; pass < 4; pass++) {
                const uint32_t shift = 24u - 8u * (uint32_t)pass;
                for (uint32_t k = lane; k < 256u; k += 32u) hist[k] = 0u;
                __syncwarp();
                for (uint32_t e = lane; e < n; e += 32u) {
                    const uint32_t k = fac_total_order_u32(__uint_as_float(Q[(hc + e) & QM].y));
                    if (pass == 0 || (k >> (shift + 8u)) == (prefix >> (shift + 8u))) atomicAdd(&hist[(k >> shift) & 255u], 1u);
                }
                __syncwarp();
                uint32_t bins[8], mine = 0;
#pragma unroll
                for (int b = 0; b < 8; b++) { bins[b] = hist[lane * 8u + b]; mine += bins[b]; }
                uint32_t incl = mine;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                    if (lane >= (uint32_t)d) incl += v;
                }
                const uint32_t hitl = (uint32_t)__ffs(__ballot_sync(0xFFFFFFFFu, incl >= need)) - 1u;   // n > bw >= need: some lane reaches it
                uint32_t digit = 0, below = incl - mine;
                if (lane == hitl) {
#pragma unroll
                    for (int b = 0; b < 8; b++) {
                        if (below + bins[b] >= need) { digit = lane * 8u + (uint32_t)b; break; }
                        below += bins[b];
                    }
                }
                digit = __shfl_sync(0xFFFFFFFFu, digit, hitl);
                below = __shfl_sync(0xFFFFFFFFu, below, hitl);
                need -= below;
                prefix |= digit << shift;
                __syncwarp();
            }
            // prefix = the bw-th smallest key; `need` of the entries equal to it are kept (the earliest ones)
            uint32_t kept = 0, eq_seen = 0;
            for (uint32_t e0 = 0; e0 < n; e0 += 32u) {
                const uint32_t e = e0 + lane;
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (e < n) v = Q[(hc + e) & QM];
                const uint32_t k = fac_total_order_u32(__uint_as_float(v.y));
                const bool eq = e < n && k == prefix;
                const uint32_t beq = __ballot_sync(0xFFFFFFFFu, eq);
                const bool keep = e < n && (k < prefix || (eq && eq_seen + __popc(beq & lt_mask) < need));
                const uint32_t bk = __ballot_sync(0xFFFFFFFFu, keep);
                __syncwarp();
                if (keep) Q[(hc + kept + __popc(bk & lt_mask)) & QM] = v;
                kept += __popc(bk);
                eq_seen += __popc(beq);
                __syncwarp();
            }
            head = hc; tail = hc + kept; mcap = 32u;   // queue.truncate(q_idx + bw)
        }
        if (lane == 0) {
            if (failed) {
                const unsigned long long fi = atomicAdd(&P.counters[3], 1ull);
                if (fi < P.failed_cap) P.failed_tiles[fi] = t_idx;
                if (P.failed_bitmap) atomicOr(&P.failed_bitmap[t_idx >> 5], 1u << (t_idx & 31u));
            } else {
                atomicAdd(&P.counters[2], (unsigned long long)tail);
                if (P.per_window) P.per_window[start] = tail;
            }
        }
        __syncwarp();
    }
}
