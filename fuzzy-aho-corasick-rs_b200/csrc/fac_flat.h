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
#define FLAT_MAX_MAPS 255u

struct FlatRec { uint32_t x, y, z, w; };   // same bytes as uint4

struct FlatView {
    const FlatRec *nrec;     // [N]
    const FlatRec *erec;     // [E]
    // per node, in build order: the (node-relative) edges whose child has outputs.  A state on its last edit keeps a
    // child only if it has an output or a single-ASCII-byte edge for the look-ahead char (src/search.rs:839-847,
    // 1057-1063, src/structs.rs:471-475): when that char is not ASCII (or there is none) only these edges can pass,
    // so the substitution / deletion slots of such a state are this short list instead of all edges.
    const uint32_t *ooff;    // [N+1]
    const uint32_t *olist;
    // ... and when the look-ahead char IS ASCII: for branching nodes (3..64 edges) one 64-bit survivor mask per ASCII
    // char c over the node-relative edge index (== build order): bit e set iff the child of edge e has an output or a
    // single-byte edge c.  gm_row[node] = row of the node or FAC_NONE; gm[row * 128 + c].
    const uint32_t *gm_row;  // [N]
    const unsigned long long *gm;
    // Root productivity masks (engines with mappings, i.e. text compared by grapheme id; build_flat_root_pm in
    // fac_builder.cpp): pm_root[(a * pm_g + b) * pm_words + w], bit e of the row = a state on its LAST edit at the
    // child of root edge e whose position reads the grapheme ids a, b can still emit something -- through its own
    // outputs, its exact child, a mapping transition, or an exhausted substitution / deletion / insertion / swap child
    // that has an output or an exact edge for what it reads next.  Everything else the root would push for an
    // edits(2) engine (two slots per root edge) can emit nothing, now or later: dropping it is result-neutral.
    // null = no table (engines without mappings, edit budgets other than 2..6, roots wider than 1024 edges).
    const unsigned long long *pm_root;
    uint32_t pm_g;           // row length = number of grapheme ids + 1 (id 0 = unknown grapheme / end of text)
    uint32_t pm_words;       // 64-bit words per row = ceil(root degree / 64)
    uint32_t pm_k;           // symbols per row: 2 = rows (a, b), 3 = rows (a, b, c3)
    // Second-level table: px_bits[px_row[x] * ceil(g * g / 64)] bit (b * g + c3) = a state on its last edit at node x reading
    // b, c3 can still emit; rows exist for the nodes two levels below the root.  A last-edit state pushes its exact child,
    // and a state whose children are on their last edit pushes a substitution / deletion child, only when the child's bit
    // is set (result-neutral for the same reason).  null = no table.
    const uint32_t *px_row;  // [N] or null
    const unsigned long long *px_bits;
};
#define FLAT_ROW_ROOT 0xFFFFFFFEu   // FlatCtx::row of a root state whose substitution / deletion slots run over pm_root rows
#define FLAT_SUBF 0x40000000u   // substitution slots run over the output-children list
#define FLAT_DELF 0x80000000u   // deletion slots run over the output-children list
#define FLAT_SUBM 0x10000000u   // substitution slots run over the set bits of the survivor mask of the look-ahead char
#define FLAT_DELM 0x20000000u   // deletion slots likewise

#if defined(__CUDA_ARCH__)
#define FLAT_POPC64(x) ((uint32_t)__popcll(x))
#else
#define FLAT_POPC64(x) ((uint32_t)__builtin_popcountll(x))
#endif
// position of the n-th (0-based) set bit of m; m must have more than n bits set
FAC_HD uint32_t flat_nth_bit64(unsigned long long m, uint32_t n) {
    uint32_t pos = 0;
    uint32_t c = FLAT_POPC64(m & 0xFFFFFFFFull);
    if (n >= c) { n -= c; pos = 32; m >>= 32; }
    c = FLAT_POPC64(m & 0xFFFFull);
    if (n >= c) { n -= c; pos += 16; m >>= 16; }
    c = FLAT_POPC64(m & 0xFFull);
    if (n >= c) { n -= c; pos += 8; m >>= 8; }
    c = FLAT_POPC64(m & 0xFull);
    if (n >= c) { n -= c; pos += 4; m >>= 4; }
    c = FLAT_POPC64(m & 0x3ull);
    if (n >= c) { n -= c; pos += 2; m >>= 2; }
    if (n >= (uint32_t)(m & 1ull)) pos += 1;
    return pos;
}

// pm_root row of the state at position j reading the grapheme ids of j and j + 1 (0 beyond the text)
template <class Text>
FAC_HD const unsigned long long *flat_pm_row(const FlatView &F, const Text &T, uint32_t j, uint32_t text_end) {
    const uint32_t a = j < text_end ? T.gid(j) : 0u, b = j + 1u < text_end ? T.gid(j + 1u) : 0u;
    size_t r = (size_t)a * F.pm_g + b;
    if (F.pm_k == 3u) r = r * F.pm_g + (j + 2u < text_end ? T.gid(j + 2u) : 0u);
    return F.pm_root + r * F.pm_words;
}
// productivity of a last-edit state at node x sitting at position p (it reads the grapheme ids of p and p + 1)
template <class Text>
FAC_HD bool flat_px_alive(const FlatView &F, const Text &T, uint32_t x, uint32_t p, uint32_t text_end) {
    if (F.px_row == nullptr) return true;
    const uint32_t r = F.px_row[x];
    if (r == FAC_NONE) return true;
    const uint32_t b = p < text_end ? T.gid(p) : 0u, c3 = p + 1u < text_end ? T.gid(p + 1u) : 0u;
    const size_t bitpos = (size_t)b * F.pm_g + c3, row_words = ((size_t)F.pm_g * F.pm_g + 63u) / 64u;
    return (F.px_bits[(size_t)r * row_words + (bitpos >> 6)] >> (bitpos & 63u)) & 1ull;
}

FAC_HD uint32_t flat_pm_count(const unsigned long long *row, uint32_t words) {
    uint32_t n = 0;
    for (uint32_t w = 0; w < words; w++) n += FLAT_POPC64(row[w]);
    return n;
}
// position of the k-th (0-based) set bit of a multi-word row; the row must have more than k bits set
FAC_HD uint32_t flat_pm_nth(const unsigned long long *row, uint32_t words, uint32_t k) {
    for (uint32_t w = 0; w + 1u < words; w++) {
        const unsigned long long m = row[w];
        const uint32_t c = FLAT_POPC64(m);
        if (k < c) return w * 64u + flat_nth_bit64(m, k);
        k -= c;
    }
    return (words - 1u) * 64u + flat_nth_bit64(row[words - 1u], k);
}

FAC_HD uint32_t flat_deg(const FlatRec &n) { return n.y & 0xFFFu; }
FAC_HD uint32_t flat_nout(const FlatRec &n) { return (n.y >> 12) & 0x3FFu; }
FAC_HD uint32_t flat_nmaps(const FlatRec &n) { return n.y >> 22; }   // <= FLAT_MAX_MAPS
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

enum : uint32_t { FLAT_F_IN_TEXT = 1u, FLAT_F_SUB = 2u, FLAT_F_SWAP = 4u, FLAT_F_INS = 8u, FLAT_F_DEL = 16u, FLAT_F_LAST = 32u, FLAT_F_HAS_NXT = 64u, FLAT_F_EXACT = 128u };

struct FlatCtx {
    uint32_t node;
    float pen;
    uint32_t cnt, pos;
    uint32_t exact;   // exact_next or FAC_NONE
    uint32_t eoff;
    uint32_t shape;   // flags | degree << 8 | mapping transitions << 20 | list / mask modes (bits 28..31)
    uint32_t lists;   // substitution slots | deletion slots << 16 (after the last-edit filter)
    uint32_t row;     // survivor-mask row of the node (FLAT_SUBM / FLAT_DELM)
    uint32_t nslots;
};
FAC_HD uint32_t flat_ctx_deg(const FlatCtx &C) { return (C.shape >> 8) & 0xFFFu; }
FAC_HD uint32_t flat_ctx_nmaps(const FlatCtx &C) { return (C.shape >> 20) & 0xFFu; }

// Guards that do not depend on the edge + the slot count of a state:
// slots = [exact] [substitution per edge] [mapping transitions] [swap] [insertion] [deletion per edge].
// FAST = the order-independent kernel: every popped state still has edit budget (exhausted states are walked).
// !FAST = the order-faithful beamed kernel: exhausted states are popped like any other (edits < MAX_EDITS_FAST is tested).
template <bool FAST, class Text>
FAC_HD void flat_make_ctx(const AutomatonView &A, const FlatView &F, const Text &T, float maxpen, uint32_t start, uint32_t text_end, const FacState &S,
                          const FlatRec &nr, FlatCtx &C) {
    C.node = S.node; C.pen = S.pen; C.cnt = S.cnt; C.pos = S.pos;
    const uint32_t jr = (S.pos >> FAC_POS_J_SHIFT) & FAC_POS_MASK, mr = S.pos & FAC_POS_MASK;
    const uint32_t j = start + jr;
    const int edits = (int)fac_edits_of(S.cnt);
    const float remaining = FAC_SUB(maxpen, S.pen);          // search.rs:648
    const bool is_last = edits + 1 >= A.mef;                 // search.rs:742
    const bool can_edit = FAST || edits < A.mef;             // search.rs:810, 937, 1003, 1043
    const bool in_text = j < text_end;
    const uint32_t deg = flat_deg(nr), nmaps = A.has_mappings ? flat_nmaps(nr) : 0u;
    uint32_t flags = is_last ? FLAT_F_LAST : 0u, nslots = 0, exact = FAC_NONE, n_sub = 0, n_del = 0;
    // a state on its last edit keeps a child only if it has an output or a single-ASCII-byte edge for the look-ahead
    // char: branching nodes enumerate just those children (output-children list / survivor mask of the char)
    const bool filt = is_last && deg > 2u;
    // the root of an engine whose first-level states are on their last edit: only the productive children (pm_root)
    const bool root_pm = FAST && S.node == 0u && F.pm_root != nullptr && !is_last && edits + 2 >= A.mef;
    const uint32_t row = filt ? F.gm_row[S.node] : (root_pm ? FLAT_ROW_ROOT : FAC_NONE);
    const uint32_t n_out_children = filt ? F.ooff[S.node + 1] - F.ooff[S.node] : 0u;
    if (in_text) {
        flags |= FLAT_F_IN_TEXT;
        const bool has_nxt = is_last && can_edit && (j + 1 < text_end);  // search.rs:758-765
        if (has_nxt) flags |= FLAT_F_HAS_NXT;
        exact = flat_lookup(A, F, S.node, nr, A.has_mappings ? T.gid(j) : T.first(j));   // search.rs:776-780
        if (FAST && is_last && exact != FAC_NONE && !flat_px_alive(F, T, exact, j + 1u, text_end)) exact = FAC_NONE;   // the child (at j + 1, still on its last edit) cannot emit
        if (exact != FAC_NONE) { flags |= FLAT_F_EXACT; nslots += 1; }
        if (can_edit) {   // substitutions + mapping transitions, search.rs:803-811
            flags |= FLAT_F_SUB;
            const uint32_t c1 = has_nxt ? T.first(j + 1) : 0xFFFFFFFFu;
            if (root_pm) { flags |= FLAT_SUBM; n_sub = flat_pm_count(flat_pm_row(F, T, j + 1u, text_end), F.pm_words); }   // children sit at j + 1
            else if (filt && c1 >= 128u) { flags |= FLAT_SUBF; n_sub = n_out_children; }
            else if (filt && row != FAC_NONE) { flags |= FLAT_SUBM; n_sub = FLAT_POPC64(F.gm[(size_t)row * 128u + c1]); }
            else n_sub = deg;
            nslots += n_sub + nmaps;
        }
        if (can_edit && j + 1 < text_end && A.pen_swap <= remaining) { flags |= FLAT_F_SWAP; nslots += 1; }   // search.rs:935-938
        bool ins_ok = can_edit && (mr != 0 || jr != 0) && A.pen_ins <= remaining;                            // search.rs:994-1003
        if (ins_ok && is_last && flat_nout(nr) == 0) {       // dead-end filter on the node itself, search.rs:1005-1007
            if (!has_nxt || !fac_has_byte_edge(A, S.node, T.first(j + 1))) ins_ok = false;
        }
        if (ins_ok) { flags |= FLAT_F_INS; nslots += 1; }
    }
    if (can_edit && A.pen_del <= remaining) {   // search.rs:1035-1045
        flags |= FLAT_F_DEL;
        const uint32_t c0 = in_text ? T.first(j) : 0xFFFFFFFFu;
        if (root_pm) { flags |= FLAT_DELM; n_del = flat_pm_count(flat_pm_row(F, T, j, text_end), F.pm_words); }   // children stay at j
        else if (filt && c0 >= 128u) { flags |= FLAT_DELF; n_del = n_out_children; }
        else if (filt && row != FAC_NONE) { flags |= FLAT_DELM; n_del = FLAT_POPC64(F.gm[(size_t)row * 128u + c0]); }
        else n_del = deg;
        nslots += n_del;
    }
    C.exact = exact; C.eoff = nr.x; C.shape = flags | (deg << 8) | (nmaps << 20); C.lists = n_sub | (n_del << 16); C.row = row; C.nslots = nslots;
}

// Decide slot `slot` of a state; on success `out` is the pushed child.  FAST: a child whose own node ceiling already
// rejects it (it would be dropped when popped, search.rs:638-642) is not produced -- result-neutral, but it changes
// queue.len(), so the order-faithful kernel (!FAST) pushes it like the reference does.  Substitution and deletion slots
// (the bulk) share one code path so that the lanes of a warp that evaluate them stay converged.
template <bool FAST, class Text>
FAC_HD bool flat_eval_slot(const AutomatonView &A, const FlatView &F, const Text &T, float maxpen, uint32_t start, uint32_t text_end, const FlatCtx &C,
                           uint32_t slot, FacState &out) {
    const uint32_t w = C.pos >> FAC_POS_W_SHIFT;
    const uint32_t jr = (C.pos >> FAC_POS_J_SHIFT) & FAC_POS_MASK, mr = C.pos & FAC_POS_MASK;
    const uint32_t j = start + jr;
    const bool is_last = (C.shape & FLAT_F_LAST) != 0, has_nxt = (C.shape & FLAT_F_HAS_NXT) != 0, in_text = (C.shape & FLAT_F_IN_TEXT) != 0;
    const uint32_t n_sub = C.lists & 0xFFFFu, nmaps = (C.shape & FLAT_F_SUB) ? flat_ctx_nmaps(C) : 0u;
    // slot boundaries
    const uint32_t b_sub = (C.shape & FLAT_F_EXACT) ? 1u : 0u;
    const uint32_t b_map = b_sub + n_sub, b_swap = b_map + nmaps;
    const uint32_t b_ins = b_swap + ((C.shape & FLAT_F_SWAP) ? 1u : 0u);
    const uint32_t b_del = b_ins + ((C.shape & FLAT_F_INS) ? 1u : 0u);
    const bool is_sub = slot >= b_sub && slot < b_map;
    if (is_sub || slot >= b_del) {   // substitution (search.rs:814-874) or deletion (search.rs:1055-1088) over one edge
        const uint32_t k = is_sub ? slot - b_sub : slot - b_del;
        const uint32_t mode = is_sub ? C.shape : (C.shape >> 1);    // FLAT_DELF / FLAT_DELM sit one bit above FLAT_SUBF / FLAT_SUBM
        const uint32_t look_j = is_sub ? j + 1u : j;                // look-ahead position of the dead-end filter
        const bool look_ok = is_sub ? has_nxt : in_text;
        uint32_t e = k;
        if (mode & FLAT_SUBF) e = F.olist[F.ooff[C.node] + k];
        else if (mode & FLAT_SUBM)
            e = C.row == FLAT_ROW_ROOT ? flat_pm_nth(flat_pm_row(F, T, look_j, text_end), F.pm_words, k)
                                       : flat_nth_bit64(F.gm[(size_t)C.row * 128u + T.first(look_j)], k);
        const FlatRec er = F.erec[C.eoff + e];
        const uint32_t nx = er.x & 0x7FFFFFFFu;
        if (is_sub && nx == C.exact) return false;
        float pp = A.pen_del;
        if (is_sub) {
            const float sm = fac_similarity(A, er.y, T.first(j));
            if (sm < A.min_sym) return false;
            pp = FAC_MUL(A.pen_sub, FAC_SUB(1.0f, sm));
            if (pp > FAC_SUB(maxpen, C.pen)) return false;
        }
        if (is_last && !(er.x >> 31) && (!look_ok || !fac_has_byte_edge(A, nx, T.first(look_j)))) return false;
        // a child that will be on its last edit and provably cannot emit (substitution child at j + 1, deletion child at j)
        if (FAST && !is_last && (int)fac_edits_of(C.cnt) + 2 >= A.mef && !flat_px_alive(F, T, nx, look_j, text_end)) return false;
        out.node = nx; out.pen = FAC_ADD(C.pen, pp); out.cnt = C.cnt + (is_sub ? 0x10000u : 0x100u);
        out.pos = is_sub ? fac_make_pos(w, jr + 1, jr + 1) : fac_make_pos(w, jr, mr);
        return !FAST || !(out.pen > FLAT_AS_FLOAT(er.w));
    }
    if (slot < b_sub) {  // exact transition, search.rs:781-798
        out.node = C.exact; out.pen = C.pen; out.cnt = C.cnt; out.pos = fac_make_pos(w, jr + 1, jr + 1);
        return true;
    }
    if (slot < b_swap) {  // mapping transition, search.rs:883-923
        const uint32_t m = A.node_map_off[C.node] + (slot - b_map);
        const uint32_t h0 = A.map_hay_off[m], hlen = A.map_hay_off[m + 1] - h0;
        if ((uint64_t)j + hlen > text_end) return false;
        for (uint32_t k = 0; k < hlen; k++)
            if (T.gid(j + k) != A.map_hay_gid[h0 + k]) return false;
        const float np = FAC_ADD(C.pen, A.map_pen[m]);
        if (np > maxpen) return false;
        out.node = A.map_next[m]; out.pen = np; out.cnt = C.cnt + 0x10000u; out.pos = fac_make_pos(w, jr + hlen, jr + hlen);
        return true;
    }
    if (slot < b_ins) {  // swap, search.rs:941-988
        const uint32_t k1 = A.has_mappings ? T.gid(j + 1) : T.first(j + 1), k0 = A.has_mappings ? T.gid(j) : T.first(j);
        const uint32_t x = flat_lookup(A, F, C.node, F.nrec[C.node], k1);
        if (x == FAC_NONE) return false;
        const uint32_t n2 = flat_lookup(A, F, x, F.nrec[x], k0);
        if (n2 == FAC_NONE) return false;
        out.node = n2; out.pen = FAC_ADD(C.pen, A.pen_swap); out.cnt = C.cnt + 0x1000000u; out.pos = fac_make_pos(w, jr + 2, jr + 2);
        return true;
    }
    // insertion, search.rs:1018-1028
    out.node = C.node; out.pen = FAC_ADD(C.pen, A.pen_ins); out.cnt = C.cnt + 1u; out.pos = fac_make_pos(w, jr + 1, mr);
    return true;
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
