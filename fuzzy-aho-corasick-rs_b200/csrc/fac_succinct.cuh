// fac_succinct.cuh -- K3 fast kernel: fuzzy frontier expansion over the succinct BFS-ordered trie.
//
// search_unsorted_impl<MAPPINGS=false, WINDOW_SKIP, MAX_EDITS_FAST=1..6> (src/search.rs:418-1119)
// for ASCII haystacks and single-byte pattern alphabets of at most 31 symbols.  Per-state logic is
// in fac_succinct.h (shared with the CPU emulator); this file is the SIMT orchestration:
//
//   * one persistent CTA per SM, tiles of start windows fetched with one atomicAdd per tile;
//   * the tile's haystack bytes (+ look-ahead) staged into shared memory by one TMA bulk copy,
//     then case-folded and translated to dense symbols in place;
//   * the first `n_smem_nodes` node records (BFS order == shallow levels first, where most visits
//     land) copied to shared memory once per CTA; deeper records come through L1/L2 as 128-bit loads;
//   * each warp owns one start window at a time and runs a depth-first stack machine in shared
//     memory: pop <= 32 states (one per lane), push exact/swap/insertion children by ballot/popc
//     compaction, flatten the (state, child edge) pairs of the popped states with a warp prefix sum
//     and evaluate substitution + deletion through 32 edges per round; children that exhausted the
//     edit budget are walked in place (exact transitions only).
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
    const float *sub_pen;    // [32 * 128]
    const uint8_t *sym_of;   // [256]
    const uint8_t *text;     // ASCII haystack bytes
    uint32_t n_nodes, n_smem_nodes;
    SuccConsts K;
    int32_t ci, wskip;
    uint32_t first_mask, second_mask;
    uint32_t seg_begin, seg_end, text_end, tile, n_tiles, lookahead;
    uint32_t stack_cap;      // states per warp stack
    uint32_t text_cap;       // bytes of the shared text tile (multiple of 16)
    FacCand *cands;
    uint32_t cand_cap;
    unsigned long long *counters;  // [0] next tile, [1] candidates, [2] states visited, [7] overflowed windows
    uint32_t *dirty;         // bitmap over start windows (bit sg - seg_begin): set when a window's stack overflowed
};

// per-call records: ceiling = prune_len - prune_low * thr (search.rs:638-642), exact f32 ops
__global__ void __launch_bounds__(256) k_succ_prepare(const uint32_t *__restrict__ bm, const uint32_t *__restrict__ fc_sym, const float *__restrict__ plen,
                                                      const float *__restrict__ plow, const uint32_t *__restrict__ out_idx, float thr, uint32_t n,
                                                      uint4 *__restrict__ rec) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float c = __fsub_rn(plen[i], __fmul_rn(plow[i], thr));
    rec[i] = make_uint4(bm[i], fc_sym[i], __float_as_uint(c), out_idx[i]);
}

struct SuccRecsDev {
    const uint4 *s, *g;
    uint32_t ns;
    __device__ __forceinline__ SuccRec operator()(uint32_t n) const {
        const uint4 v = n < ns ? s[n] : __ldg(&g[n]);
        SuccRec r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w;
        return r;
    }
};
struct SuccTextDev {
    const uint8_t *sb, *ss;  // folded bytes / symbols of the tile
    uint32_t base;
    __device__ __forceinline__ uint32_t byte(uint32_t j) const { return sb[j - base]; }
    __device__ __forceinline__ uint32_t sym(uint32_t j) const { return ss[j - base]; }
};
struct SuccEmitDev {
    FacCand *cands;
    uint32_t cap;
    unsigned long long *counter;
    __device__ __forceinline__ void operator()(uint32_t sg, uint32_t eg, uint32_t pat, float sim, uint32_t cnt) {
        const unsigned long long ci = atomicAdd(counter, 1ull);
        if (ci < cap) {
            uint4 *dst = reinterpret_cast<uint4 *>(&cands[ci]);
            dst[0] = make_uint4(sg, eg, pat, __float_as_uint(sim));
            dst[1] = make_uint4(cnt, 0u, 0u, 0u);
        }
    }
};

__device__ __forceinline__ uint32_t succ_lanemask_lt() { uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

__device__ __forceinline__ void succ_warp_push(uint4 *stk, uint32_t &top, bool p, const FacState &c) {
    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, p);
    if (p) stk[top + __popc(bal & succ_lanemask_lt())] = make_uint4(c.node, __float_as_uint(c.pen), c.cnt, c.pos);
    top += __popc(bal);
}

template <int NT>
__global__ void __launch_bounds__(NT, 1) k_expand_succinct(const __grid_constant__ SuccParams P) {
    extern __shared__ __align__(16) uint8_t dyn_smem[];
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ uint32_t s_tile_idx, s_next_win;
    constexpr int NW = NT / 32;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

    // carve-up: [node records][warp stacks][sub_pen][raw tile][folded bytes][symbols][sym_of]
    uint4 *s_rec = reinterpret_cast<uint4 *>(dyn_smem);
    uint4 *s_stack = s_rec + P.n_smem_nodes;
    float *s_subpen = reinterpret_cast<float *>(s_stack + (size_t)NW * P.stack_cap);
    uint8_t *s_raw = reinterpret_cast<uint8_t *>(s_subpen + 32 * 128);
    uint8_t *s_byte = s_raw + P.text_cap;
    uint8_t *s_sym = s_byte + P.text_cap;
    uint8_t *s_symof = s_sym + P.text_cap;

    for (uint32_t k = tid; k < P.n_smem_nodes; k += NT) s_rec[k] = P.rec[k];
    for (uint32_t k = tid; k < 32 * 128; k += NT) s_subpen[k] = P.sub_pen[k];
    for (uint32_t k = tid; k < 256; k += NT) s_symof[k] = P.sym_of[k];
    if (tid == 0) fac_mbar_init(&s_mbar, 1);
    __syncthreads();

    const SuccConsts K = P.K;
    const SuccRecsDev R{s_rec, P.rec, P.n_smem_nodes};
    const SuccOut *out2 = reinterpret_cast<const SuccOut *>(P.out2);
    SuccEmitDev emit{P.cands, P.cand_cap, &P.counters[1]};
    uint4 *const stk = s_stack + (size_t)warp * P.stack_cap;
    const uint32_t cap = P.stack_cap;
    uint32_t mbar_phase = 0;
    uint32_t n_states = 0;  // per-lane count of visited states (summed at the end)

    for (;;) {
        if (tid == 0) { s_tile_idx = (uint32_t)atomicAdd(&P.counters[0], 1ull); s_next_win = 0; }
        __syncthreads();
        const uint32_t t = s_tile_idx;
        if (t >= P.n_tiles) break;
        const uint32_t tile_start = P.seg_begin + t * P.tile;
        const uint32_t count = min(P.tile, P.seg_end - tile_start);
        const uint32_t text_end = P.text_end;

        // ---- stage the tile: TMA bulk copy of the 16-byte aligned body, plain loads for the tail ----
        uint32_t lead = (uint32_t)(((uintptr_t)(P.text + tile_start)) & 15u);
        if (lead > tile_start) lead = 0;
        const uint32_t base = tile_start - lead;
        const uint32_t span = count + P.lookahead + lead;                 // positions the tile must answer
        const uint32_t avail = min(span, text_end - base);                // bytes that exist
        const bool aligned = ((((uintptr_t)(P.text + base)) & 15u) == 0u);
        const uint32_t bulk = aligned ? (avail & ~15u) : 0u;
        if (bulk && tid == 0) {
            fac_fence_proxy_async();
            fac_mbar_expect_tx(&s_mbar, bulk);
            fac_tma_load_1d(s_raw, P.text + base, bulk, &s_mbar);
        }
        for (uint32_t k = bulk + tid; k < avail; k += NT) s_raw[k] = P.text[base + k];
        if (bulk) { fac_mbar_wait(&s_mbar, mbar_phase & 1u); mbar_phase++; }
        __syncthreads();
        for (uint32_t k = tid; k < span; k += NT) {
            uint32_t b = 0, s = SUCC_NOSYM;
            if (k < avail) {
                b = s_raw[k];
                if (P.ci && b >= 'A' && b <= 'Z') b += 32u;   // to_ascii_lowercase, grapheme.rs:110-117
                s = s_symof[b];
            }
            s_byte[k] = (uint8_t)b; s_sym[k] = (uint8_t)s;
        }
        __syncthreads();
        const SuccTextDev T{s_byte, s_sym, base};

        // ---- windows of the tile, one per warp at a time ----
        for (;;) {
            uint32_t w = 0;
            if (lane == 0) w = atomicAdd(&s_next_win, 1u);
            w = __shfl_sync(0xFFFFFFFFu, w, 0);
            if (w >= count) break;
            const uint32_t start = tile_start + w;
            if (P.wskip) {  // 2-gram window skip (search.rs:535-553); result-neutral
                if (!((P.first_mask >> T.sym(start)) & 1u)) {
                    if (start + 1 >= text_end) continue;
                    if (!((P.second_mask >> T.sym(start + 1)) & 1u)) continue;
                }
            }
            uint32_t top = 1;
            if (lane == 0) stk[0] = make_uint4(0u, 0u, 0u, 0u);
            while (top) {
                __syncwarp();
                const uint32_t navail = min(top, 32u);
                const bool has = lane < navail;
                uint4 sv = make_uint4(0, 0, 0, 0);
                SuccRec rec; rec.x = rec.y = rec.z = 0; rec.w = FAC_NONE;
                if (has) { sv = stk[top - 1u - lane]; rec = R(sv.x); }
                const float pen = __uint_as_float(sv.y);
                const bool dead = !has || pen > __uint_as_float(rec.z);   // node ceiling, search.rs:638-642
                const bool last = (int)fac_edits_of(sv.z) + 1 >= K.mef;
                const uint32_t deg = __popc(rec.x);
                // worst-case pushes of this state: exact only when its edit-children are exhausted
                const uint32_t ub = dead ? 0u : (last ? 1u : 2u * deg + 3u);
                uint32_t incl = ub;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                    if (lane >= (uint32_t)d) incl += v;
                }
                const uint32_t viol = __ballot_sync(0xFFFFFFFFu, has && (incl > cap - top + lane + 1u));
                const uint32_t n_pop = viol ? (uint32_t)(__ffs(viol) - 1) : navail;
                if (n_pop == 0) {  // the top state alone does not fit: give the window to the faithful kernel
                    if (lane == 0) {
                        atomicOr(&P.dirty[(start - P.seg_begin) >> 5], 1u << ((start - P.seg_begin) & 31u));
                        atomicAdd(&P.counters[7], 1ull);
                    }
                    break;
                }
                const bool active = lane < n_pop && !dead;
                top -= n_pop;
                __syncwarp();

                SuccCtx C;
                C.fc = 0; C.pen = 0.f; C.cnt = 0; C.pos = 0; C.packed = 0xFF000000u; C.flags = 0;
                uint32_t n_items = 0;
                bool p_ex = false, p_sw = false, p_in = false;
                FacState c_ex, c_sw, c_in;
                SuccRec r_sw; r_sw.x = r_sw.y = r_sw.z = 0; r_sw.w = FAC_NONE;
                if (active) {
                    n_states++;
                    if (rec.w != FAC_NONE) succ_outputs(K, out2, emit, rec.w, pen, sv.z, start, start + (sv.w & 1023u));
                    succ_make_ctx(K, T, start, text_end, rec, pen, sv.z, sv.w, C);
                    const uint32_t jr = sv.w >> 10;
                    if ((C.packed >> 24) != 0xFFu) {
                        p_ex = true;
                        c_ex.node = (rec.y & SUCC_FC_MASK) + (C.packed >> 24); c_ex.pen = pen; c_ex.cnt = sv.z; c_ex.pos = succ_make_pos(jr + 1, jr + 1);
                    }
                    p_sw = succ_swap(K, R, rec, C, r_sw, c_sw);
                    p_in = succ_ins(K, rec, C, sv.x, c_in);
                    if ((C.flags & (SUCC_F_IN_TEXT | SUCC_F_DEL)) != 0) n_items = deg;
                    if (last) {  // edit-children are exhausted: walk them in place
                        if (p_sw) n_states += succ_walk(K, R, out2, T, emit, start, text_end, r_sw, c_sw.pen, c_sw.cnt, c_sw.pos >> 10, c_sw.pos & 1023u);
                        if (p_in) n_states += succ_walk(K, R, out2, T, emit, start, text_end, rec, c_in.pen, c_in.cnt, c_in.pos >> 10, c_in.pos & 1023u);
                        p_sw = p_in = false;
                    }
                }
                succ_warp_push(stk, top, p_ex, c_ex);
                if (__any_sync(0xFFFFFFFFu, p_sw)) succ_warp_push(stk, top, p_sw, c_sw);
                if (__any_sync(0xFFFFFFFFu, p_in)) succ_warp_push(stk, top, p_in, c_in);

                // ---- (state, child edge) pairs, 32 per round ----
                uint32_t off = n_items;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, off, d);
                    if (lane >= (uint32_t)d) off += v;
                }
                const uint32_t total = __shfl_sync(0xFFFFFFFFu, off, 31);
                off -= n_items;  // exclusive
                for (uint32_t b0 = 0; b0 < total; b0 += 32u) {
                    const uint32_t it = b0 + lane;
                    const bool valid = it < total;
                    uint32_t lo = 0;
#pragma unroll
                    for (int step = 16; step; step >>= 1) {
                        const uint32_t cand = lo + step;
                        const uint32_t v = __shfl_sync(0xFFFFFFFFu, off, cand & 31u);
                        if (cand < 32u && v <= it) lo = cand;
                    }
                    SuccCtx O;
                    O.fc = __shfl_sync(0xFFFFFFFFu, C.fc, lo);
                    O.pen = __shfl_sync(0xFFFFFFFFu, C.pen, lo);
                    O.cnt = __shfl_sync(0xFFFFFFFFu, C.cnt, lo);
                    O.pos = __shfl_sync(0xFFFFFFFFu, C.pos, lo);
                    O.packed = __shfl_sync(0xFFFFFFFFu, C.packed, lo);
                    O.flags = __shfl_sync(0xFFFFFFFFu, C.flags, lo);
                    const uint32_t k = it - __shfl_sync(0xFFFFFFFFu, off, lo);
                    bool p_sub = false, p_del = false;
                    FacState c_sub, c_del;
                    if (valid) {
                        const SuccRec crec = R((O.fc & SUCC_FC_MASK) + k);
                        p_sub = succ_sub(K, s_subpen, O, k, crec, c_sub);
                        p_del = succ_del(K, O, k, crec, c_del);
                        if (O.flags & SUCC_F_LAST) {
                            if (p_sub) n_states += succ_walk(K, R, out2, T, emit, start, text_end, crec, c_sub.pen, c_sub.cnt, c_sub.pos >> 10, c_sub.pos & 1023u);
                            if (p_del) n_states += succ_walk(K, R, out2, T, emit, start, text_end, crec, c_del.pen, c_del.cnt, c_del.pos >> 10, c_del.pos & 1023u);
                            p_sub = p_del = false;
                        }
                    }
                    if (__any_sync(0xFFFFFFFFu, p_sub)) succ_warp_push(stk, top, p_sub, c_sub);
                    if (__any_sync(0xFFFFFFFFu, p_del)) succ_warp_push(stk, top, p_del, c_del);
                }
            }
        }
        __syncthreads();  // every warp is done with the tile before it is restaged
    }
    n_states = __reduce_add_sync(0xFFFFFFFFu, n_states);
    if (lane == 0 && n_states) atomicAdd(&P.counters[2], (unsigned long long)n_states);
}
