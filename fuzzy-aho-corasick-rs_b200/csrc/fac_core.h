// fac_core.h -- the per-state arithmetic of the fuzzy frontier expansion, written once for the
// CUDA kernels (and compiled for the host by the CPU-side emulator the tests use to check the
// flattened automaton + slot logic without a GPU).
//
// Follows src/search.rs:576-1089 of the reference operator for operator; all arithmetic is IEEE
// f32 with one rounding per operation (no FMA contraction: explicit *_rn intrinsics on the device).
//
// Formulation.  The reference pops one State at a time and pushes its children in a fixed order:
//   exact (search.rs:776-798) | substitutions over all edges (:814-874) | mapping transitions
//   (:883-923) | swap (:935-989) | insertion (:994-1029) | deletions over all edges (:1035-1089).
// Here every potential child is a numbered *slot* of its parent; `fac_make_ctx` computes what is
// common to all slots of a state (the cheap per-state guards and the slot count) and
// `fac_eval_slot` decides one slot independently, so a frontier level becomes a flat list of
// (state, slot) work items whose order equals the reference's FIFO push order.
#pragma once
#include "fac_types.h"

#if defined(__CUDA_ARCH__)
#define FAC_ADD(a, b) __fadd_rn((a), (b))
#define FAC_SUB(a, b) __fsub_rn((a), (b))
#define FAC_MUL(a, b) __fmul_rn((a), (b))
#define FAC_DIV(a, b) __fdiv_rn((a), (b))
#else
#define FAC_ADD(a, b) ((a) + (b))
#define FAC_SUB(a, b) ((a) - (b))
#define FAC_MUL(a, b) ((a) * (b))
#define FAC_DIV(a, b) ((a) / (b))
#endif

// Text accessor reading straight from the TextView arrays (used by the host emulator and by the
// kernels for positions outside the staged shared-memory tile).
struct FacTextDirect {
    TextView tv;
    const uint32_t *ascii_gid;
    int ci;
    FAC_HD uint32_t first(uint32_t j) const {
        if (tv.ascii) { const uint32_t b = tv.bytes[j]; return (ci && b >= 'A' && b <= 'Z') ? b + 32u : b; }
        return tv.first[j];
    }
    FAC_HD uint32_t gid(uint32_t j) const {
        if (tv.ascii) return ascii_gid[first(j)];
        return tv.gid[j];
    }
};

// Byte offset of grapheme g (gs_byte_offset, src/grapheme.rs:58-60, 95-97); g == n maps to the
// haystack length (src/search.rs:672-676).
FAC_HD uint64_t fac_byte_offset(const TextView &tv, uint32_t g) {
    if (tv.ascii) return g;
    return tv.off64 ? tv.off64[g] : (uint64_t)tv.off32[g];
}

enum : uint32_t {
    FAC_F_IN_TEXT = 1u,
    FAC_F_SUB = 2u,
    FAC_F_SWAP = 4u,
    FAC_F_INS = 8u,
    FAC_F_DEL = 16u,
    FAC_F_LAST = 32u,
    FAC_F_HAS_NXT = 64u,
    FAC_F_EXACT = 128u,
};

struct FacCtx {
    uint32_t node;
    float pen;
    uint32_t cnt;
    uint32_t pos;
    uint32_t exact;  // exact_next or FAC_NONE
    uint32_t flags;
    uint32_t nslots;
};

FAC_HD uint32_t fac_hash2(uint32_t a, uint32_t b) {
    uint32_t h = a * 0x9E3779B1u ^ (b * 0x85EBCA6Bu + 0x27D4EB2Fu);
    h ^= h >> 15;
    h *= 0x2C1B3C6Du;
    h ^= h >> 13;
    return h;
}

// Exact transition (Node::find_transition*, src/structs.rs:452-519).  `sym` is the text first char
// for engines without mappings (first edge in build order with that first char wins), else the
// grapheme id of the folded haystack grapheme (full-string identity).  Low-degree nodes (the bulk
// of a trie) are resolved by scanning their edge symbols, which are contiguous; high-degree nodes
// go through the open-addressing table.
#define FAC_SCAN_DEG 4u
FAC_HD uint32_t fac_lookup_hash(const AutomatonView &A, uint32_t node, uint32_t sym) {
    uint32_t h = fac_hash2(node, sym) & A.trans_mask;
    for (;;) {
        const FacTrans t = A.trans[h];
        if (t.node == FAC_NONE) return FAC_NONE;
        if (t.node == node && t.sym == sym) return t.next;
        h = (h + 1) & A.trans_mask;
    }
}
FAC_HD uint32_t fac_find_edge(const AutomatonView &A, uint32_t node, uint32_t eoff, uint32_t deg, uint32_t sym) {
    if (deg <= FAC_SCAN_DEG) {
        for (uint32_t e = 0; e < deg; e++)
            if (A.edge_sym[eoff + e] == sym) return A.edge_next[eoff + e] & 0x7FFFFFFFu;
        return FAC_NONE;
    }
    return fac_lookup_hash(A, node, sym);
}
FAC_HD uint32_t fac_lookup(const AutomatonView &A, uint32_t node, uint32_t sym) {
    const uint32_t eoff = A.node_edge_off[node];
    return fac_find_edge(A, node, eoff, A.node_edge_off[node + 1] - eoff, sym);
}

// Node::has_matching_edge_char (src/structs.rs:471-475): a single-ASCII-byte edge equal to `ch`.
FAC_HD bool fac_has_byte_edge(const AutomatonView &A, uint32_t node, uint32_t ch) {
    return ch < 128u && ((A.node_bitmap[node * 4u + (ch >> 5)] >> (ch & 31u)) & 1u);
}

// get_similarity (src/search.rs:76-82) + Similarity::get (src/structs.rs:82-92); ordered pair
// (pattern side, text side).
FAC_HD float fac_similarity(const AutomatonView &A, uint32_t a, uint32_t b) {
    if (a == b) return 1.0f;
    if (a < 128u && b < 128u) return A.sim_ascii[a * 128u + b];
    const uint64_t key = ((uint64_t)a << 32) | b;
    uint32_t lo = 0, hi = A.n_sim;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        const uint64_t k = A.sim_keys[mid];
        if (k == key) return A.sim_vals[mid];
        if (k < key) lo = mid + 1; else hi = mid;
    }
    return 0.0f;
}

// `limits.or(self.limits.as_ref())` (src/search.rs:93, 109, 125, 140, 160)
FAC_HD bool fac_pick_limits(const AutomatonView &A, uint32_t idx, FacLimits &L) {
    if (idx != FAC_NONE) { L = A.lim[idx]; return true; }
    if (A.has_global_limits) { L = A.lim[0]; return true; }
    return false;
}
FAC_HD bool fac_none_or_lt(int mx, int v) { return mx < 0 || v < mx; }
FAC_HD bool fac_none_or_le(int mx, int v) { return mx < 0 || v <= mx; }

FAC_HD uint32_t fac_edits_of(uint32_t cnt) { return (cnt & 0xFFu) + ((cnt >> 8) & 0xFFu) + ((cnt >> 16) & 0xFFu) + (cnt >> 24); }
FAC_HD uint32_t fac_make_pos(uint32_t w, uint32_t jr, uint32_t mr) { return (w << FAC_POS_W_SHIFT) | (jr << FAC_POS_J_SHIFT) | mr; }

// 2-gram window skip for 1-edit engines (src/search.rs:535-553). `c0`,`c1` are text_chars[start],
// text_chars[start+1]; `has1` = start+1 < text_len.  Returns true when the window is skipped.
FAC_HD bool fac_window_skipped(const AutomatonView &A, uint32_t c0, bool has1, uint32_t c1) {
    if (!A.wskip) return false;
    if (c0 < 128u && ((A.ws_first[c0 >> 5] >> (c0 & 31u)) & 1u) == 0u) {
        if (!has1) return true;
        if (c1 < 128u && ((A.ws_second[c1 >> 5] >> (c1 & 31u)) & 1u) == 0u) return true;
    }
    return false;
}

// Node ceiling (src/search.rs:638-642): true when the state must be dropped.
FAC_HD bool fac_over_ceiling(const AutomatonView &A, uint32_t node, float pen, float thr) {
    return pen > FAC_SUB(A.node_prune_len[node], FAC_MUL(A.node_prune_low[node], thr));
}

// Per-state context: guards that do not depend on the edge, and the slot count.
//   Text: object with first(j) / gid(j) returning the folded first char / grapheme id at grapheme j.
template <class Text>
FAC_HD void fac_make_ctx(const AutomatonView &A, const Text &T, float maxpen, uint32_t start, uint32_t text_end,
                         const FacState &S, FacCtx &C) {
    C.node = S.node; C.pen = S.pen; C.cnt = S.cnt; C.pos = S.pos;
    const uint32_t jr = (S.pos >> FAC_POS_J_SHIFT) & FAC_POS_MASK, mr = S.pos & FAC_POS_MASK;
    const uint32_t j = start + jr;
    const int edits = (int)fac_edits_of(S.cnt);
    const int mef = A.mef;
    const bool fast = mef != 255;
    const float remaining = FAC_SUB(maxpen, S.pen);  // search.rs:648
    const bool is_last = fast && (edits + 1 >= mef);  // search.rs:742
    const bool in_text = j < text_end;
    const uint32_t deg = A.node_edge_off[S.node + 1] - A.node_edge_off[S.node];
    uint32_t flags = 0, nslots = 0;
    uint32_t exact = FAC_NONE;
    FacLimits L;
    bool haveL = false;
    if (!fast) haveL = fac_pick_limits(A, A.node_lim[S.node], L);  // search.rs:653-657
    if (is_last) flags |= FAC_F_LAST;
    if (in_text) {
        flags |= FAC_F_IN_TEXT;
        const bool has_nxt = is_last && (edits < mef) && (j + 1 < text_end);  // search.rs:758-765
        if (has_nxt) flags |= FAC_F_HAS_NXT;
        const uint32_t cur = T.first(j);
        exact = A.has_mappings ? fac_lookup(A, S.node, T.gid(j)) : fac_lookup(A, S.node, cur);  // search.rs:776-780
        if (exact != FAC_NONE) { flags |= FAC_F_EXACT; nslots += 1; }
        bool sub_ok;  // search.rs:803-811
        if (fast) sub_ok = edits < mef;
        else sub_ok = haveL ? (fac_none_or_lt(L.edits, edits) && fac_none_or_lt(L.sub, (int)((S.cnt >> 16) & 0xFF)))
                            : (edits == 0 && ((S.cnt >> 16) & 0xFF) == 0);
        if (sub_ok) {
            flags |= FAC_F_SUB;
            nslots += deg;
            if (A.has_mappings) nslots += A.node_map_off[S.node + 1] - A.node_map_off[S.node];
        }
        // swap pre-conditions (search.rs:935-938); the two lookups happen in the slot
        if (j + 1 < text_end && A.pen_swap <= remaining && (!fast || edits < mef)) { flags |= FAC_F_SWAP; nslots += 1; }
        // insertion (search.rs:994-1008) is decided here completely
        bool ins_ok = (mr != 0 || jr != 0) && A.pen_ins <= remaining;
        if (ins_ok) {
            if (fast) ins_ok = edits < mef;
            else ins_ok = haveL ? (fac_none_or_lt(L.edits, edits) && fac_none_or_lt(L.ins, (int)(S.cnt & 0xFF))) : false;
        }
        if (ins_ok && is_last && A.node_out_off[S.node + 1] == A.node_out_off[S.node]) {
            if (!has_nxt || !fac_has_byte_edge(A, S.node, T.first(j + 1))) ins_ok = false;
        }
        if (ins_ok) { flags |= FAC_F_INS; nslots += 1; }
    }
    bool del_ok = A.pen_del <= remaining;  // search.rs:1035-1045
    if (del_ok) {
        if (fast) del_ok = edits < mef;
        else del_ok = haveL ? (fac_none_or_lt(L.edits, edits) && fac_none_or_lt(L.del, (int)((S.cnt >> 8) & 0xFF))) : false;
    }
    if (del_ok) { flags |= FAC_F_DEL; nslots += deg; }
    C.exact = exact; C.flags = flags; C.nslots = nslots;
}

// Decide slot `slot` of a state; on success fill `out` with the pushed child.
template <class Text>
FAC_HD bool fac_eval_slot(const AutomatonView &A, const Text &T, float maxpen, uint32_t start, uint32_t text_end,
                          const FacCtx &C, uint32_t slot, FacState &out) {
    const uint32_t w = C.pos >> FAC_POS_W_SHIFT;
    const uint32_t jr = (C.pos >> FAC_POS_J_SHIFT) & FAC_POS_MASK, mr = C.pos & FAC_POS_MASK;
    const uint32_t j = start + jr;
    const uint32_t eoff = A.node_edge_off[C.node];
    const uint32_t deg = A.node_edge_off[C.node + 1] - eoff;
    const bool is_last = (C.flags & FAC_F_LAST) != 0;
    const bool has_nxt = (C.flags & FAC_F_HAS_NXT) != 0;
    uint32_t s = slot;
    if (C.flags & FAC_F_EXACT) {
        if (s == 0) {  // search.rs:781-798
            out.node = C.exact; out.pen = C.pen; out.cnt = C.cnt; out.pos = fac_make_pos(w, jr + 1, jr + 1);
            return true;
        }
        s -= 1;
    }
    if (C.flags & FAC_F_SUB) {
        if (s < deg) {  // search.rs:814-874
            const uint32_t en = A.edge_next[eoff + s];
            const uint32_t nx = en & 0x7FFFFFFFu;
            if (nx == C.exact) return false;
            const uint32_t cur = T.first(j);
            const float sm = fac_similarity(A, A.edge_char[eoff + s], cur);
            if (sm < A.min_sym) return false;
            const float pp = FAC_MUL(A.pen_sub, FAC_SUB(1.0f, sm));
            if (pp > FAC_SUB(maxpen, C.pen)) return false;
            if (is_last) {
                if (!(en >> 31) && (!has_nxt || !fac_has_byte_edge(A, nx, T.first(j + 1)))) return false;
            }
            out.node = nx; out.pen = FAC_ADD(C.pen, pp); out.cnt = C.cnt + 0x10000u; out.pos = fac_make_pos(w, jr + 1, jr + 1);
            return true;
        }
        s -= deg;
        if (A.has_mappings) {  // search.rs:883-923
            const uint32_t moff = A.node_map_off[C.node];
            const uint32_t nmap = A.node_map_off[C.node + 1] - moff;
            if (s < nmap) {
                const uint32_t m = moff + s;
                const uint32_t h0 = A.map_hay_off[m], hlen = A.map_hay_off[m + 1] - h0;
                if ((uint64_t)j + hlen > text_end) return false;
                for (uint32_t k = 0; k < hlen; k++)
                    if (T.gid(j + k) != A.map_hay_gid[h0 + k]) return false;
                const float np = FAC_ADD(C.pen, A.map_pen[m]);
                if (np > maxpen) return false;
                out.node = A.map_next[m]; out.pen = np; out.cnt = C.cnt + 0x10000u; out.pos = fac_make_pos(w, jr + hlen, jr + hlen);
                return true;
            }
            s -= nmap;
        }
    }
    if (C.flags & FAC_F_SWAP) {
        if (s == 0) {  // search.rs:941-988
            uint32_t x, n2 = FAC_NONE;
            if (A.has_mappings) {
                x = fac_lookup(A, C.node, T.gid(j + 1));
                if (x != FAC_NONE) n2 = fac_lookup(A, x, T.gid(j));
            } else {
                x = fac_lookup(A, C.node, T.first(j + 1));
                if (x != FAC_NONE) n2 = fac_lookup(A, x, T.first(j));
            }
            if (n2 == FAC_NONE) return false;
            if (A.mef == 255) {  // within_limits_swap_ahead(get_node_limits(node2), ..), search.rs:119-130
                FacLimits L;
                if (!fac_pick_limits(A, A.node_lim[n2], L)) return false;
                const int edits = (int)fac_edits_of(C.cnt);
                if (!(fac_none_or_lt(L.edits, edits) && fac_none_or_lt(L.swp, (int)(C.cnt >> 24)))) return false;
            }
            out.node = n2; out.pen = FAC_ADD(C.pen, A.pen_swap); out.cnt = C.cnt + 0x1000000u; out.pos = fac_make_pos(w, jr + 2, jr + 2);
            return true;
        }
        s -= 1;
    }
    if (C.flags & FAC_F_INS) {
        if (s == 0) {  // search.rs:1018-1028
            out.node = C.node; out.pen = FAC_ADD(C.pen, A.pen_ins); out.cnt = C.cnt + 1u; out.pos = fac_make_pos(w, jr + 1, mr);
            return true;
        }
        s -= 1;
    }
    // deletion over edge s (search.rs:1055-1088)
    {
        const uint32_t en = A.edge_next[eoff + s];
        const uint32_t nx = en & 0x7FFFFFFFu;
        if (is_last) {
            const bool has_co = (C.flags & FAC_F_IN_TEXT) != 0;
            if (!(en >> 31) && (!has_co || !fac_has_byte_edge(A, nx, T.first(j)))) return false;
        }
        out.node = nx; out.pen = FAC_ADD(C.pen, A.pen_del); out.cnt = C.cnt + 0x100u; out.pos = fac_make_pos(w, jr, mr);
        return true;
    }
}

// Output check for one pattern of the node's output list (src/search.rs:678-703).
// Returns true and the similarity when the candidate reaches `best`.
FAC_HD bool fac_eval_output(const AutomatonView &A, float thr, uint32_t pat, float pen, uint32_t cnt, float &sim_out) {
    const int edits = (int)fac_edits_of(cnt);
    if (A.mef != 255) {
        if (edits > A.mef) return false;
    } else {
        FacLimits L;
        if (fac_pick_limits(A, A.pat_lim[pat], L)) {  // within_limits, search.rs:151-169
            if (!(fac_none_or_le(L.edits, edits) && fac_none_or_le(L.ins, (int)(cnt & 0xFF)) &&
                  fac_none_or_le(L.del, (int)((cnt >> 8) & 0xFF)) && fac_none_or_le(L.sub, (int)((cnt >> 16) & 0xFF)) &&
                  fac_none_or_le(L.swp, (int)(cnt >> 24))))
                return false;
        } else if (cnt != 0) return false;
    }
    const float total = A.pat_glen[pat];
    const float sim = FAC_MUL(FAC_DIV(FAC_SUB(total, pen), total), A.pat_weight[pat]);  // search.rs:698-699
    if (sim < thr) return false;
    sim_out = sim;
    return true;
}

// f32::total_cmp key (monotone unsigned image of the float bits).
FAC_HD uint32_t fac_total_order_u32(float f) {
    union { float f; uint32_t u; } v;
    v.f = f;
    return (v.u & 0x80000000u) ? ~v.u : (v.u | 0x80000000u);
}
