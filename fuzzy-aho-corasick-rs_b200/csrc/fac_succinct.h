// fac_succinct.h -- per-state logic of the succinct-trie expansion (K3 fast kernel), shared by the
// CUDA kernel (fac_succinct.cuh) and the CPU-side emulator the tests use.
//
// Applies to engines the reference dispatches to its fast monomorphisations
// `search_unsorted_impl<MAPPINGS=false, _, MAX_EDITS_FAST=1..6>` (src/search.rs:204-393) when, in
// addition, every trie edge is one ASCII byte and the pattern alphabet has at most 31 symbols (narrow
// layout, W = false) or at most 63 symbols (wide layout, W = true: 64-bit child bitmaps).
//
// Layout.  Nodes are renumbered in BFS order with the children of a node contiguous and sorted by
// symbol, so the whole node is one 16-byte record
//     narrow: x = child bitmap over the dense symbol alphabet (bit 31 never set)
//             y = first_child | (symbol of the edge that leads INTO this node) << 27
//             z = f32 bits of the node ceiling  prune_len - prune_len_over_weight * threshold  (search.rs:638-642)
//             w = index of the node's first output entry, or FAC_NONE
//     wide:   x, y = 64-bit child bitmap (bit 63 = "node has outputs"; the output index lives in a side array)
//             z = first_child | in-symbol << 26,   w = ceiling
// and  child(n, sym) = first_child + popc(x & below(sym)).  There is no edge array and no hash
// table: the exact transition (Node::find_transition_char_no_mappings, src/structs.rs:512-519) is a
// bit test + popcount, and the last-edit dead-end filter (search.rs:839-847, 1005-1007, 1057-1063:
// "child has an output or a single-byte edge equal to the next text char") is answered for all
// children of a node at once by precomputed masks (below).
//
// The expansion is order-independent: candidates are reduced by maximum similarity and ties
// between different edit-count vectors are detected and redone by the order-faithful kernel
// (fac_fastreduce.cuh), so the reference's FIFO order does not have to be kept here.  The dedup
// map (search.rs:608-628) is result-neutral and is not applied.
#pragma once
#include "fac_core.h"
#include "fac_types.h"

#define SUCC_MAX_NODES 0x03FFFFFFu
// sub_pen is [ROW][SUCC_SP_STRIDE]: column b < 128 = text first char b, column 128 = any non-ASCII text first char
// (similarity 0: only offered when the engine has no similarity entry with a non-ASCII member)
#define SUCC_SP_STRIDE 129u
#define SUCC_NONASCII 128u

#if defined(__CUDA_ARCH__)
#define FAC_POPC(x) __popc(x)
#define FAC_POPC64(x) __popcll(x)
#define FAC_AS_FLOAT(u) __uint_as_float(u)
#else
#define FAC_POPC(x) ((uint32_t)__builtin_popcount(x))
#define FAC_POPC64(x) ((uint32_t)__builtin_popcountll(x))
static inline float fac_as_float_host(uint32_t u) { union { uint32_t u; float f; } v; v.u = u; return v.f; }
#define FAC_AS_FLOAT(u) fac_as_float_host(u)
#endif

struct SuccRec { uint32_t x, y, z, w; };   // same bytes as uint4
struct SuccOut { uint32_t pat_last; uint32_t glen_bits; uint32_t weight_bits; uint32_t pad; };  // pat | last<<31; pad = limits index

// Layout traits: mask type, "no symbol" code (its bit is never an edge), table row length.
template <bool W> struct SuccW;
template <> struct SuccW<false> { typedef uint32_t M; enum : uint32_t { NOSYM = 31u, ROW = 32u, FC_MASK = 0x07FFFFFFu, SYM_SHIFT = 27u }; };
template <> struct SuccW<true> { typedef uint64_t M; enum : uint32_t { NOSYM = 63u, ROW = 64u, FC_MASK = 0x03FFFFFFu, SYM_SHIFT = 26u }; };

FAC_HD uint32_t succ_popc(uint32_t m) { return FAC_POPC(m); }
FAC_HD uint32_t succ_popc(uint64_t m) { return FAC_POPC64(m); }

template <bool W> FAC_HD typename SuccW<W>::M succ_bm(const SuccRec &r);
template <> FAC_HD uint32_t succ_bm<false>(const SuccRec &r) { return r.x; }
template <> FAC_HD uint64_t succ_bm<true>(const SuccRec &r) { return ((uint64_t)(r.y & 0x7FFFFFFFu) << 32) | r.x; }
template <bool W> FAC_HD uint32_t succ_fc(const SuccRec &r) { return (W ? r.z : r.y) & SuccW<W>::FC_MASK; }
template <bool W> FAC_HD float succ_ceil(const SuccRec &r) { return FAC_AS_FLOAT(W ? r.w : r.z); }
template <bool W> FAC_HD bool succ_has_out(const SuccRec &r) { return W ? (r.y >> 31) != 0u : r.w != FAC_NONE; }

struct SuccConsts {
    float thr, maxpen, pen_ins, pen_del, pen_swap;
    int32_t mef;             // edit budget of the fast monomorphisations (1..6); in limits mode the largest
                             // total edit count any FuzzyLimits of the engine admits
    // limits mode (LIM = true): the reference's generic MAX_EDITS_FAST = 255 path (per-pattern / per-type limits)
    const FacLimits *lim;    // [L] index 0 = global limits (valid iff has_global)
    const uint32_t *node_lim;  // [N] BFS order: limits index of the pattern that created the node, or FAC_NONE
    int32_t has_global;
    const uint32_t *out_idx;   // [N] wide layout: first output entry of the node (the narrow record carries it)
};
template <bool W> FAC_HD uint32_t succ_out_idx(const SuccConsts &K, const SuccRec &r, uint32_t node) { return W ? K.out_idx[node] : r.w; }

// `limits.or(self.limits.as_ref())` (src/search.rs:93, 109, 125, 140, 160)
FAC_HD bool succ_pick_limits(const SuccConsts &K, uint32_t idx, FacLimits &L) {
    if (idx != FAC_NONE) { L = K.lim[idx]; return true; }
    if (K.has_global) { L = K.lim[0]; return true; }
    return false;
}

template <class M> FAC_HD M succ_below(uint32_t sym) { return (M(1) << sym) - M(1); }
template <bool W> FAC_HD uint32_t succ_child(const SuccRec &r, uint32_t sym) {
    typedef typename SuccW<W>::M M;
    return succ_fc<W>(r) + succ_popc((M)(succ_bm<W>(r) & succ_below<M>(sym)));
}
template <bool W> FAC_HD bool succ_has_edge(const SuccRec &r, uint32_t sym) { return (succ_bm<W>(r) >> sym) & 1u; }
// State position word: window-in-tile << 20 | (j - start) << 10 | (matched_end - start).  The window field lets one
// warp keep states of several start windows on its stack (the kernel feeds the next window before the current
// one has drained); the emulator always uses window 0.
FAC_HD uint32_t succ_jr(uint32_t pos) { return (pos >> 10) & 1023u; }
FAC_HD uint32_t succ_mr(uint32_t pos) { return pos & 1023u; }
FAC_HD uint32_t succ_repos(uint32_t pos, uint32_t jr, uint32_t mr) { return (pos & 0xFFF00000u) | (jr << 10) | mr; }

// Text context word of position j (one 32-bit load answers everything a state needs to know about the text):
//     folded first char of grapheme j (0..127, SUCC_NONASCII = any non-ASCII char) | sym(j) << 8 | sym(j+1) << 14 |
//     sym(j+2) << 20 | sym(j+3) << 26          (6-bit dense symbols; NOSYM of the layout when there is no such symbol)
// Text contract: positions at or beyond text_end read as byte 0 / NOSYM (up to start + look-ahead).
FAC_HD uint32_t succ_ctx_pack(uint32_t byte, uint32_t s0, uint32_t s1, uint32_t s2, uint32_t s3) {
    return byte | (s0 << 8) | (s1 << 14) | (s2 << 20) | (s3 << 26);
}
FAC_HD uint32_t succ_ctx_byte(uint32_t p) { return p & 0xFFu; }
FAC_HD uint32_t succ_ctx_s0(uint32_t p) { return (p >> 8) & 63u; }
FAC_HD uint32_t succ_ctx_s1(uint32_t p) { return (p >> 14) & 63u; }
FAC_HD uint32_t succ_ctx_s2(uint32_t p) { return (p >> 20) & 63u; }
FAC_HD uint32_t succ_ctx_s3(uint32_t p) { return p >> 26; }

// Outputs of one node visit (search.rs:659-737): fast-path limit check is `edits > MAX_EDITS_FAST`,
// never true here because no state exceeds the budget.
// In limits mode every pattern is checked against its own (or the global) limits: within_limits,
// src/search.rs:151-169 (`pad` of the output entry is the pattern's limits index).
template <bool LIM, class Emit>
FAC_HD void succ_outputs(const SuccConsts &K, const SuccOut *out2, Emit &emit, uint32_t idx, float pen, uint32_t cnt, uint32_t sg, uint32_t eg) {
    for (;;) {
        const SuccOut o = out2[idx++];
        if (LIM) {
            FacLimits L;
            bool ok;
            if (succ_pick_limits(K, o.pad, L))
                ok = fac_none_or_le(L.edits, (int)fac_edits_of(cnt)) && fac_none_or_le(L.ins, (int)(cnt & 0xFF)) &&
                     fac_none_or_le(L.del, (int)((cnt >> 8) & 0xFF)) && fac_none_or_le(L.sub, (int)((cnt >> 16) & 0xFF)) &&
                     fac_none_or_le(L.swp, (int)(cnt >> 24));
            else ok = cnt == 0;
            if (!ok) { if (o.pat_last >> 31) break; continue; }
        }
        const float total = FAC_AS_FLOAT(o.glen_bits);
        const float sim = FAC_MUL(FAC_DIV(FAC_SUB(total, pen), total), FAC_AS_FLOAT(o.weight_bits));  // search.rs:698-699
        if (!(sim < K.thr)) emit(sg, eg, o.pat_last & 0x7FFFFFFFu, sim, cnt);
        if (o.pat_last >> 31) break;
    }
}

// A state that has spent its whole edit budget follows exact transitions only (sub / swap / ins /
// del all need edits < MAX_EDITS_FAST, search.rs:810, 937, 1003, 1043): walk the chain with the
// per-pop checks (ceiling :638-642, outputs :659-737, exact :776-798).  Returns the nodes visited.
template <bool LIM, bool W, class Recs, class Text, class Emit>
FAC_HD uint32_t succ_walk(const SuccConsts &K, const Recs &R, const SuccOut *out2, const Text &T, Emit &emit, uint32_t start, uint32_t text_end,
                          uint32_t node, SuccRec rec, float pen, uint32_t cnt, uint32_t jr, uint32_t mr) {
    uint32_t steps = 0;
    for (;;) {
        steps++;
        if (pen > succ_ceil<W>(rec)) break;
        if (succ_has_out<W>(rec)) succ_outputs<LIM>(K, out2, emit, succ_out_idx<W>(K, rec, node), pen, cnt, start, start + mr);
        const uint32_t j = start + jr;
        if (j >= text_end) break;
        const uint32_t s = T.sym(j);
        if (!succ_has_edge<W>(rec, s)) break;
        node = succ_child<W>(rec, s);
        rec = R(node);
        jr++; mr = jr;
    }
    return steps;
}

enum : uint32_t { SUCC_F_IN_TEXT = 1u, SUCC_F_LAST = 2u, SUCC_F_DEL = 4u, SUCC_F_HAS_NXT = 8u, SUCC_F_INS = 16u,
                  SUCC_F_SWAP_MASK = 32u,   // the masks do not rule the swap child out
                  SUCC_F_INS_MASK = 64u,    // the masks do not rule the insertion child out
                  SUCC_F_SWAP_SURE = 128u,  // ... and already prove that an exhausted swap child can reach an output
                  SUCC_F_EXACT = 256u       // exact child exists and is not ruled out by the productivity masks
};

// ---- survivor masks -------------------------------------------------------------------------------
// For a state on its last edit the dead-end filter keeps a child c iff
//     c has an output  ||  c has a single-byte edge for the look-ahead symbol
// (substitution: look-ahead = text[j+1], search.rs:839-847; deletion: text[j], :1057-1063).  Per node
// and symbol the builder precomputes  gm[node][y] = { symbols s : child(node, s) has edge y }  and
// gm[node][NOSYM] = { s : child(node, s) has an output }  ("grandchild masks", symbol space), so the
// children that survive are known from two loads instead of one record load per child.
// States that are not on their last edit keep every child.  `GM(node, y)` returns the row entry.
template <bool W>
struct SuccCtx2 {
    typename SuccW<W>::M bm, sub_m, del_m;
    uint32_t fc;       // first child
    float pen;
    uint32_t cnt, pos;
    uint32_t packed;   // text context word of the state's position (succ_ctx_pack)
    uint32_t flags;
};

//
// Two-deep refinement for the first gm2_nodes nodes:  gm2[node][y1][y2] = { s : c = child(node, s) has edge y1 and
// g = child(c, y1) has an output or an edge y2 } and gm2[node][y1][NOSYM] = { s : ... g has an output };
// gm2[node][NOSYM][*] = 0.  A child outside  outm | gm2[node][y1][y2]  walks c -> g and stops there without visiting
// an output node, i.e. it cannot emit a candidate: dropping it is result-neutral (only the visited-state statistic
// changes).  The mask accessor `G` offers
//     G.row(node, y)        gm[node][y]                                    (~0 = "unknown" beyond the table)
//     G.row2(node, y1, y2)  gm2[node][y1][y2], or gm[node][y1] for nodes beyond the two-deep table
// The same two tables answer two more questions without touching a node record:
//   * swap (search.rs:935-989: node -text[j+1]-> x -text[j]-> n2): bit text[j+1] of gm[node][text[j]] says whether x has
//     the edge text[j]; bit text[j+1] of gm2[node][text[j]][text[j+2]] additionally says that n2 has an output or the edge
//     text[j+2], i.e. that an exhausted swap child can still emit;
//   * insertion on the last edit (search.rs:994-1029: same node, j+1): the child walks node -text[j+1]-> g -text[j+2]-> h;
//     bit text[j+1] of  outm | gm2[node][text[j+2]][text[j+3]]  says that g has an output, or the edge text[j+2] to a
//     node with an output or the edge text[j+3].  Without that (and without an output at the node itself) the
//     insertion child cannot emit.
// Deep tables (narrow layout, built by build_deep_tables in fac_builder.cpp over the compact symbol row n_syms + 1):
//   G.row3(node, y1, y2, y3)  one level deeper than row2 for the first n3 nodes: d = child(node, s) has edge y1 and the exact walk
//                             from g = child(d, y1) over y2, y3 visits an output or is still alive after y3; row2 beyond the table.
//                             Used exactly where row2 was: substitution (t1,t2,t3), deletion (t0,t1,t2), swap (t0,t2,t3).
//   G.deep(node, q, a, b, c, d)  = q ? G.pm(node, a, b, c, d) : G.row3(node, a, b, c)   (one load, table selected by index)
//   G.pm(node, a, b, c, d)    productivity mask: bit s = a state on its LAST edit at child(node, s) whose position reads a, b, c, d can
//                             still emit (itself, through its exact chain, or through any exhausted edit child; four symbols
//                             for the first n4 nodes -- G.four(node) --, three for the first n3, two for the first np2, ~0 beyond).  It filters (i) the substitution / deletion
//                             children of a state whose children will be on their last edit and (ii) the exact child of a state
//                             that is on its last edit.  A filtered state can emit nothing, now or later: result-neutral.
// Limits mode: the per-state permissions follow within_limits_*_ahead of the limits of the pattern that created
// the node (src/search.rs:66-148); with neither pattern nor global limits only a fresh state may substitute
// (:143-145).  The dead-end filter does not exist on that path, but dropping exhausted children whose exact walk
// cannot reach an output stays result-neutral, so the same masks are applied once `last` (no limits of the engine
// admit another edit after this one).
template <bool LIM, bool W, class Text, class GM>
FAC_HD void succ_make_ctx2(const SuccConsts &K, const Text &T, const GM &G, uint32_t start, uint32_t text_end, uint32_t node,
                           const SuccRec &rec, float pen, uint32_t cnt, uint32_t pos, SuccCtx2<W> &C) {
    typedef typename SuccW<W>::M M;
    const uint32_t NOSYM = SuccW<W>::NOSYM;
    const uint32_t jr = succ_jr(pos), mr = succ_mr(pos);
    const uint32_t j = start + jr;
    const int edits = (int)fac_edits_of(cnt);
    const bool last = edits + 1 >= K.mef;
    const bool child_last = edits + 2 >= K.mef;   // the edit children of this state are on their last edit (or exhausted)
    const bool in_text = j < text_end, has_nxt = j + 1 < text_end;
    const uint32_t p = T.ctx(j);
    const uint32_t cur_s = succ_ctx_s0(p), nxt_s = succ_ctx_s1(p), nxt2_s = succ_ctx_s2(p), nxt3_s = succ_ctx_s3(p);
    const float remaining = FAC_SUB(K.maxpen, pen);
    bool del_ok = K.pen_del <= remaining;  // search.rs:1035
    bool sub_ok = true, ins_ok = true;
    if (LIM) {
        FacLimits L;
        if (succ_pick_limits(K, K.node_lim[node], L)) {
            const bool e_ok = fac_none_or_lt(L.edits, edits);
            sub_ok = e_ok && fac_none_or_lt(L.sub, (int)((cnt >> 16) & 0xFF));
            ins_ok = e_ok && fac_none_or_lt(L.ins, (int)(cnt & 0xFF));
            del_ok = del_ok && e_ok && fac_none_or_lt(L.del, (int)((cnt >> 8) & 0xFF));
        } else { sub_ok = edits == 0 && ((cnt >> 16) & 0xFF) == 0; ins_ok = false; del_ok = false; }
    }
    const M bm = succ_bm<W>(rec);
    const bool has_nxt_edge = (bm >> nxt_s) & 1u;   // NOSYM is never an edge
    const bool has_out = succ_has_out<W>(rec);
    // everything the swap / insertion children need besides the masks (search.rs:935-937, 994-1003); the mask rows are
    // only fetched for the states that get this far (the loads are independent and issued together)
    const bool sw_pre = in_text && has_nxt && has_nxt_edge && cur_s != NOSYM && K.pen_swap <= remaining;
    const bool ins_pre = in_text && ins_ok && !(mr == 0 && jr == 0) && K.pen_ins <= remaining;
    const bool ins_need = ins_pre && last && !has_out;    // the masks decide
    M outm = 0, m_sub = ~M(0), m_del = ~M(0), m_sw = 0, m_ins = 0;  // states not on their last edit keep every child
    const bool has_cur_edge = (bm >> cur_s) & 1u;   // exact transition, search.rs:776-798
    M m_ex = ~M(0);
    if (last || child_last) {
        // one code path for both kinds of state: a state on its last edit reads survivor rows (row3), a state whose edit
        // children will be on their last edit reads productivity rows (pm) -- G.deep selects the table by index
        const uint32_t nxt4_s = succ_ctx_s3(T.ctx(j + 1));
        if (last) outm = G.row(node, NOSYM);
        m_sub = G.deep(node, !last, nxt_s, nxt2_s, nxt3_s, nxt4_s);
        m_del = G.deep(node, !last, cur_s, nxt_s, nxt2_s, nxt3_s);
        if (last && has_cur_edge) m_ex = G.pm(node, nxt_s, nxt2_s, nxt3_s, nxt4_s);   // the exact child stays on its last edit
    }
    m_sw = sw_pre ? (last ? G.row3(node, cur_s, nxt2_s, nxt3_s) : G.row(node, cur_s)) : M(0);
    m_ins = (ins_need && has_nxt_edge) ? G.row2(node, nxt2_s, nxt3_s) : M(0);
    const bool swap_m = (m_sw >> nxt_s) & 1u;
    const bool ins_m = ins_pre && (!ins_need || (((outm | m_ins) >> nxt_s) & (M)has_nxt_edge & 1u));
    const uint32_t flags = (last ? SUCC_F_LAST : 0u) | (in_text ? SUCC_F_IN_TEXT : 0u) | (has_nxt ? SUCC_F_HAS_NXT : 0u) |
                           (del_ok ? SUCC_F_DEL : 0u) | (ins_ok ? SUCC_F_INS : 0u) | (swap_m ? SUCC_F_SWAP_MASK : 0u) |
                           (ins_m ? SUCC_F_INS_MASK : 0u) | ((last && G.two_deep(node)) ? SUCC_F_SWAP_SURE : 0u) |
                           ((has_cur_edge && ((m_ex >> cur_s) & 1u)) ? SUCC_F_EXACT : 0u);
    C.bm = bm;
    C.sub_m = (in_text && sub_ok) ? (bm & (outm | m_sub) & ~(M(1) << cur_s)) : M(0);
    C.del_m = del_ok ? (bm & (outm | m_del)) : M(0);
    C.fc = succ_fc<W>(rec); C.pen = pen; C.cnt = cnt; C.pos = pos;
    C.packed = p;
    C.flags = flags;
}

// position of the n-th (0-based) set bit of m; m must have more than n bits set
FAC_HD uint32_t succ_nth_bit(uint32_t m, uint32_t n) {
    uint32_t pos = 0;
    uint32_t c = FAC_POPC(m & 0xFFFFu);
    if (n >= c) { n -= c; pos = 16; m >>= 16; }
    c = FAC_POPC(m & 0xFFu);
    if (n >= c) { n -= c; pos += 8; m >>= 8; }
    c = FAC_POPC(m & 0xFu);
    if (n >= c) { n -= c; pos += 4; m >>= 4; }
    c = FAC_POPC(m & 0x3u);
    if (n >= c) { n -= c; pos += 2; m >>= 2; }
    if (n >= (m & 1u)) pos += 1;
    return pos;
}
FAC_HD uint32_t succ_nth_bit(uint64_t m, uint32_t n) {
    const uint32_t lo = (uint32_t)m, c = FAC_POPC(lo);
    return n >= c ? 32u + succ_nth_bit((uint32_t)(m >> 32), n - c) : succ_nth_bit(lo, n);
}

// Item r of a state's survivors: r < popc(sub_m) is a substitution, the rest are deletions.
// Returns false when the substitution penalty exceeds the remaining budget (search.rs:829-834).
// Written branch-free so substitution and deletion lanes of a warp stay converged.
template <bool W>
FAC_HD bool succ_item2(const SuccConsts &K, const float *sub_pen, const SuccCtx2<W> &C, uint32_t r, FacState &out) {
    typedef typename SuccW<W>::M M;
    const uint32_t ns = succ_popc(C.sub_m);
    const uint32_t jr = succ_jr(C.pos);
    const bool is_sub = r < ns;
    const uint32_t s = succ_nth_bit(is_sub ? C.sub_m : C.del_m, is_sub ? r : r - ns);
    // +inf in the table when similarity < min_symbol_similarity; deletions read a valid slot and ignore it
    const float tp = sub_pen[s * SUCC_SP_STRIDE + succ_ctx_byte(C.packed)];
    const float pp = is_sub ? tp : K.pen_del;
    out.node = C.fc + succ_popc((M)(C.bm & succ_below<M>(s)));
    out.pen = FAC_ADD(C.pen, pp);
    out.cnt = C.cnt + (is_sub ? 0x10000u : 0x100u);
    out.pos = is_sub ? succ_repos(C.pos, jr + 1, jr + 1) : C.pos;
    return !(is_sub && pp > FAC_SUB(K.maxpen, C.pen));
}

// Swap (search.rs:935-989: node -text[j+1]-> x -text[j]-> n2, matched_start unchanged) and insertion
// (search.rs:994-1029: forbidden before anything is consumed, matched_end unchanged, dead-end filter on the
// current node when this is the last edit).
template <bool LIM, bool W, class Recs>
FAC_HD bool succ_swap2(const SuccConsts &K, const Recs &R, const SuccCtx2<W> &C, FacState &out) {
    typedef typename SuccW<W>::M M;
    // SUCC_F_SWAP_MASK: in the text with a next grapheme, penalty affordable, edge text[j+1] present and the mask row does
    // not rule out the second edge -- almost every state ends here, before a record is touched
    if (!(C.flags & SUCC_F_SWAP_MASK)) return false;
    const uint32_t cur_s = succ_ctx_s0(C.packed), nxt_s = succ_ctx_s1(C.packed);
    const SuccRec rx = R(C.fc + succ_popc((M)(C.bm & succ_below<M>(nxt_s))));
    if (!succ_has_edge<W>(rx, cur_s)) return false;   // only reachable when the mask row was not available ("unknown")
    const uint32_t jr = succ_jr(C.pos);
    out.node = succ_child<W>(rx, cur_s); out.pen = FAC_ADD(C.pen, K.pen_swap); out.cnt = C.cnt + 0x1000000u; out.pos = succ_repos(C.pos, jr + 2, jr + 2);
    if (LIM) {  // within_limits_swap_ahead of the TARGET node's limits (search.rs:119-130, 962-975)
        FacLimits L;
        if (!succ_pick_limits(K, K.node_lim[out.node], L)) return false;
        if (!(fac_none_or_lt(L.edits, (int)fac_edits_of(C.cnt)) && fac_none_or_lt(L.swp, (int)(C.cnt >> 24)))) return false;
    }
    if ((C.flags & (SUCC_F_LAST | SUCC_F_SWAP_SURE)) == SUCC_F_LAST) {
        // beyond the two-deep table: the exhausted swap child visits n2 at j+2 and then only follows text[j+2]; without
        // an output at n2 and without that edge it cannot emit -- dropping it is result-neutral
        const SuccRec r2 = R(out.node);
        if (!succ_has_out<W>(r2) && !succ_has_edge<W>(r2, succ_ctx_s2(C.packed))) return false;
    }
    return true;
}
template <bool LIM, bool W>
FAC_HD bool succ_ins2(const SuccConsts &K, const SuccCtx2<W> &C, uint32_t node, FacState &out) {
    // SUCC_F_INS_MASK: insertion allowed here (search.rs:994-1003) and, on the last edit, not ruled out by the masks,
    // which subsume the reference's dead-end filter ("node has an output or the edge text[j+1]", :1005-1007) and look
    // two nodes further along the exact walk of the exhausted child
    if (!(C.flags & SUCC_F_INS_MASK)) return false;
    const uint32_t jr = succ_jr(C.pos), mr = succ_mr(C.pos);
    out.node = node; out.pen = FAC_ADD(C.pen, K.pen_ins); out.cnt = C.cnt + 1u; out.pos = succ_repos(C.pos, jr + 1, mr);
    return true;
}
