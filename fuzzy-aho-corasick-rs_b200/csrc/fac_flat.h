// fac_flat.h -- per-state logic of the general stack-machine kernel (fac_stack.cuh) over MERGED 16-byte records,
// shared with the CPU-side emulator of the tests.
//
// Same arithmetic as fac_core.h (fac_make_ctx / fac_eval_slot follow src/search.rs:576-1089 operator for operator),
// restricted to what the order-independent fast path needs -- engines of the reference's fast monomorphisations
// `search_unsorted_impl<MAPPINGS, _, MAX_EDITS_FAST = 1..6>` (src/search.rs:204-393), states that still have edit budget
// (exhausted states are walked, flat_walk) -- and laid out so that a state costs one record load and a child one more:
//
//   nrec[node] = { first edge, degree | outputs << 12 | mapping transitions << 22,
//                  f32 bits of the node ceiling  prune_len - prune_len_over_weight * threshold  (search.rs:638-642),
//                  first output entry }                                                     (per call: the ceiling depends on the threshold)
//   erec[edge] = { Edge::next | (child has outputs) << 31, Edge::first_char, exact-match key of the edge
//                  (first char, or grapheme id for engines with mappings), ceiling of the child }   (build order, builder.rs:336-342)
//
// The generic kernels read the same facts from seven separate arrays (CSR offsets x2, prune coefficients x2, edge_next,
// edge_char, edge_sym): dependent L2 round trips that bound those kernels.  Everything else (hash table for wide
// nodes, ASCII edge bitmaps, outputs, similarity, mapping CSR) is read through the AutomatonView as before.
#pragma once
#include "fac_core.h"
#include "fac_types.h"

#define FLAT_MAX_DEG 4095u
#define FLAT_MAX_OUT 1023u
#define FLAT_MAX_MAPS 1023u

struct FlatRec { uint32_t x, y, z, w; };   // same bytes as uint4

struct FlatView {
    const FlatRec *nrec;   // [N]
    const FlatRec *erec;   // [E]
};

FAC_HD uint32_t flat_deg(const FlatRec &n) { return n.y & 0xFFFu; }
FAC_HD uint32_t flat_nout(const FlatRec &n) { return (n.y >> 12) & 0x3FFu; }
FAC_HD uint32_t flat_nmaps(const FlatRec &n) { return n.y >> 22; }
#if defined(__CUDA_ARCH__)
#define FLAT_AS_FLOAT(u) __uint_as_float(u)
#else
static inline float flat_as_float_host(uint32_t u) { union { uint32_t u; float f; } v; v.u = u; return v.f; }
#define FLAT_AS_FLOAT(u) flat_as_float_host(u)
#endif

// Exact transition (Node::find_transition*, src/structs.rs:452-519): same rule as fac_find_edge -- narrow nodes scan
// their (contiguous) edge records, wide nodes probe the open-addressing table.
FAC_HD uint32_t flat_lookup(const AutomatonView &A, const FlatView &F, uint32_t node, const FlatRec &nr, uint32_t key) {
    const uint32_t deg = flat_deg(nr);
    if (deg <= FAC_SCAN_DEG) {
        for (uint32_t e = 0; e < deg; e++) {
            const FlatRec er = F.erec[nr.x + e];
            if (er.z == key) return er.x & 0x7FFFFFFFu;
        }
        return FAC_NONE;
    }
    return fac_lookup_hash(A, node, key);
}

enum : uint32_t { FLAT_F_IN_TEXT = 1u, FLAT_F_SWAP = 4u, FLAT_F_INS = 8u, FLAT_F_DEL = 16u, FLAT_F_LAST = 32u, FLAT_F_HAS_NXT = 64u, FLAT_F_EXACT = 128u };

struct FlatCtx {
    uint32_t node;
    float pen;
    uint32_t cnt, pos;
    uint32_t exact;   // exact_next or FAC_NONE
    uint32_t eoff;
    uint32_t shape;   // flags | degree << 8 | mapping transitions << 20
    uint32_t nslots;
};
FAC_HD uint32_t flat_ctx_deg(const FlatCtx &C) { return (C.shape >> 8) & 0xFFFu; }
FAC_HD uint32_t flat_ctx_nmaps(const FlatCtx &C) { return C.shape >> 20; }

// Guards that do not depend on the edge + the slot count of a state that still has edit budget (edits < MAX_EDITS_FAST):
// slots = [exact] [substitution per edge] [mapping transitions] [swap] [insertion] [deletion per edge].
template <class Text>
FAC_HD void flat_make_ctx(const AutomatonView &A, const FlatView &F, const Text &T, float maxpen, uint32_t start, uint32_t text_end, const FacState &S,
                          const FlatRec &nr, FlatCtx &C) {
    C.node = S.node; C.pen = S.pen; C.cnt = S.cnt; C.pos = S.pos;
    const uint32_t jr = (S.pos >> FAC_POS_J_SHIFT) & FAC_POS_MASK, mr = S.pos & FAC_POS_MASK;
    const uint32_t j = start + jr;
    const int edits = (int)fac_edits_of(S.cnt);
    const float remaining = FAC_SUB(maxpen, S.pen);          // search.rs:648
    const bool is_last = edits + 1 >= A.mef;                 // search.rs:742
    const bool in_text = j < text_end;
    const uint32_t deg = flat_deg(nr), nmaps = A.has_mappings ? flat_nmaps(nr) : 0u;
    uint32_t flags = is_last ? FLAT_F_LAST : 0u, nslots = 0, exact = FAC_NONE;
    if (in_text) {
        flags |= FLAT_F_IN_TEXT;
        const bool has_nxt = is_last && (j + 1 < text_end);  // search.rs:758-765 (edits < MAX_EDITS_FAST holds for every state here)
        if (has_nxt) flags |= FLAT_F_HAS_NXT;
        exact = flat_lookup(A, F, S.node, nr, A.has_mappings ? T.gid(j) : T.first(j));   // search.rs:776-780
        if (exact != FAC_NONE) { flags |= FLAT_F_EXACT; nslots += 1; }
        nslots += deg + nmaps;                               // substitutions + mapping transitions (search.rs:803-811: edits < MEF)
        if (j + 1 < text_end && A.pen_swap <= remaining) { flags |= FLAT_F_SWAP; nslots += 1; }   // search.rs:935-938
        bool ins_ok = (mr != 0 || jr != 0) && A.pen_ins <= remaining;                            // search.rs:994-1003
        if (ins_ok && is_last && flat_nout(nr) == 0) {       // dead-end filter on the node itself, search.rs:1005-1007
            if (!has_nxt || !fac_has_byte_edge(A, S.node, T.first(j + 1))) ins_ok = false;
        }
        if (ins_ok) { flags |= FLAT_F_INS; nslots += 1; }
    }
    if (A.pen_del <= remaining) { flags |= FLAT_F_DEL; nslots += deg; }   // search.rs:1035-1045
    C.exact = exact; C.eoff = nr.x; C.shape = flags | (deg << 8) | (nmaps << 20); C.nslots = nslots;
}

// Decide slot `slot` of a state; on success `out` is the pushed child.  A child whose own node ceiling already rejects
// it (it would be dropped when popped, search.rs:638-642) is not produced: result-neutral.
template <class Text>
FAC_HD bool flat_eval_slot(const AutomatonView &A, const FlatView &F, const Text &T, float maxpen, uint32_t start, uint32_t text_end, const FlatCtx &C,
                           uint32_t slot, FacState &out) {
    const uint32_t w = C.pos >> FAC_POS_W_SHIFT;
    const uint32_t jr = (C.pos >> FAC_POS_J_SHIFT) & FAC_POS_MASK, mr = C.pos & FAC_POS_MASK;
    const uint32_t j = start + jr;
    const uint32_t deg = flat_ctx_deg(C);
    const bool is_last = (C.shape & FLAT_F_LAST) != 0, has_nxt = (C.shape & FLAT_F_HAS_NXT) != 0;
    uint32_t s = slot;
    if (C.shape & FLAT_F_EXACT) {
        if (s == 0) {  // search.rs:781-798
            out.node = C.exact; out.pen = C.pen; out.cnt = C.cnt; out.pos = fac_make_pos(w, jr + 1, jr + 1);
            return true;
        }
        s -= 1;
    }
    if (C.shape & FLAT_F_IN_TEXT) {
        if (s < deg) {  // substitution over edge s, search.rs:814-874
            const FlatRec er = F.erec[C.eoff + s];
            const uint32_t nx = er.x & 0x7FFFFFFFu;
            if (nx == C.exact) return false;
            const uint32_t cur = T.first(j);
            const float sm = fac_similarity(A, er.y, cur);
            if (sm < A.min_sym) return false;
            const float pp = FAC_MUL(A.pen_sub, FAC_SUB(1.0f, sm));
            if (pp > FAC_SUB(maxpen, C.pen)) return false;
            if (is_last && !(er.x >> 31) && (!has_nxt || !fac_has_byte_edge(A, nx, T.first(j + 1)))) return false;
            out.node = nx; out.pen = FAC_ADD(C.pen, pp); out.cnt = C.cnt + 0x10000u; out.pos = fac_make_pos(w, jr + 1, jr + 1);
            return !(out.pen > FLAT_AS_FLOAT(er.w));
        }
        s -= deg;
        const uint32_t nmaps = flat_ctx_nmaps(C);
        if (s < nmaps) {  // mapping transition, search.rs:883-923
            const uint32_t m = A.node_map_off[C.node] + s;
            const uint32_t h0 = A.map_hay_off[m], hlen = A.map_hay_off[m + 1] - h0;
            if ((uint64_t)j + hlen > text_end) return false;
            for (uint32_t k = 0; k < hlen; k++)
                if (T.gid(j + k) != A.map_hay_gid[h0 + k]) return false;
            const float np = FAC_ADD(C.pen, A.map_pen[m]);
            if (np > maxpen) return false;
            out.node = A.map_next[m]; out.pen = np; out.cnt = C.cnt + 0x10000u; out.pos = fac_make_pos(w, jr + hlen, jr + hlen);
            return true;
        }
        s -= nmaps;
    }
    if (C.shape & FLAT_F_SWAP) {
        if (s == 0) {  // search.rs:941-988
            const uint32_t k1 = A.has_mappings ? T.gid(j + 1) : T.first(j + 1), k0 = A.has_mappings ? T.gid(j) : T.first(j);
            const uint32_t x = flat_lookup(A, F, C.node, F.nrec[C.node], k1);
            if (x == FAC_NONE) return false;
            const uint32_t n2 = flat_lookup(A, F, x, F.nrec[x], k0);
            if (n2 == FAC_NONE) return false;
            out.node = n2; out.pen = FAC_ADD(C.pen, A.pen_swap); out.cnt = C.cnt + 0x1000000u; out.pos = fac_make_pos(w, jr + 2, jr + 2);
            return true;
        }
        s -= 1;
    }
    if (C.shape & FLAT_F_INS) {
        if (s == 0) {  // search.rs:1018-1028
            out.node = C.node; out.pen = FAC_ADD(C.pen, A.pen_ins); out.cnt = C.cnt + 1u; out.pos = fac_make_pos(w, jr + 1, mr);
            return true;
        }
        s -= 1;
    }
    {   // deletion over edge s, search.rs:1055-1088
        const FlatRec er = F.erec[C.eoff + s];
        const uint32_t nx = er.x & 0x7FFFFFFFu;
        if (is_last && !(er.x >> 31) && (!(C.shape & FLAT_F_IN_TEXT) || !fac_has_byte_edge(A, nx, T.first(j)))) return false;
        out.node = nx; out.pen = FAC_ADD(C.pen, A.pen_del); out.cnt = C.cnt + 0x100u; out.pos = fac_make_pos(w, jr, mr);
        return !(out.pen > FLAT_AS_FLOAT(er.w));
    }
}

// A child that has spent the whole edit budget follows exact transitions only: ceiling (search.rs:638-642), outputs
// (:659-737; `edits > MAX_EDITS_FAST` never holds here), exact step (:776-798).  Returns the nodes visited.
template <class Text, class Emit>
FAC_HD uint32_t flat_walk(const AutomatonView &A, const FlatView &F, const Text &T, float thr, Emit &emit, uint32_t start, uint32_t text_end,
                          const FacState &child) {
    uint32_t node = child.node;
    uint32_t jr = (child.pos >> FAC_POS_J_SHIFT) & FAC_POS_MASK, mr = child.pos & FAC_POS_MASK;
    uint32_t steps = 0;
    for (;;) {
        steps++;
        const FlatRec nr = F.nrec[node];
        if (child.pen > FLAT_AS_FLOAT(nr.z)) break;
        const uint32_t no = flat_nout(nr);
        for (uint32_t o = 0; o < no; o++) {
            const uint32_t pat = A.out_pat[nr.w + o];
            const float total = A.pat_glen[pat];
            const float sim = FAC_MUL(FAC_DIV(FAC_SUB(total, child.pen), total), A.pat_weight[pat]);   // search.rs:698-699
            if (!(sim < thr)) emit(start, start + mr, pat, sim, child.cnt);
        }
        const uint32_t j = start + jr;
        if (j >= text_end) break;
        const uint32_t nx = flat_lookup(A, F, node, nr, A.has_mappings ? T.gid(j) : T.first(j));
        if (nx == FAC_NONE) break;
        node = nx; jr++; mr = jr;
    }
    return steps;
}
