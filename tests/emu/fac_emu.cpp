// fac_emu.cpp -- TEST CODE.  Host-side, sequential emulation of the GPU search pipeline built from
// the SAME product headers the kernels use (csrc/fac_core.h, fac_unicode.h, fac_builder.*): the
// flattened automaton, the per-scalar segmentation predicates, the slot formulation of the
// frontier expansion (level-synchronous, tile of windows at a time, order-preserving) and the
// candidate -> best-per-span reduction.  It lets the CPU-only test-suite check all of that against
// the oracle before any GPU time is spent; it is never loaded by the product path.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "../../fuzzy-aho-corasick-rs_b200/csrc/fac_builder.h"
#include "../../fuzzy-aho-corasick-rs_b200/csrc/fac_core.h"
#include "../../fuzzy-aho-corasick-rs_b200/csrc/fac_flat.h"
#include "../../fuzzy-aho-corasick-rs_b200/csrc/fac_succinct.h"
#include "../../fuzzy-aho-corasick-rs_b200/csrc/fac_unicode.h"

using namespace fac;

struct EmuText {
    std::vector<uint32_t> first, gid;
    std::vector<uint64_t> off;
    TextView tv;
};

static void segment_host(const HostAutomaton &A, const uint8_t *s, size_t len, EmuText &T) {
    const UnicodeTables &U = host_unicode_tables();
    bool ascii = true;
    for (size_t i = 0; i < len; i++) if (s[i] >= 0x80) { ascii = false; break; }
    memset(&T.tv, 0, sizeof(T.tv));
    T.tv.n_bytes = len;
    T.tv.ascii = ascii;
    if (ascii) { T.tv.bytes = s; T.tv.n = (uint32_t)len; return; }
    // K1 logic: every scalar start decides its own boundary, then per grapheme fold + id
    std::vector<uint64_t> starts;
    for (size_t i = 0; i < len; i++)
        if (!fac_is_cont(s[i]) && fac_break_before(U, s, 0, i)) starts.push_back(i);
    const size_t n = starts.size();
    T.first.resize(n); T.gid.assign(n, 0); T.off.resize(n + 1);
    for (size_t g = 0; g < n; g++) {
        const uint64_t b = starts[g], e = g + 1 < n ? starts[g + 1] : len;
        T.off[g] = b;
        uint32_t fc = 0;
        if (A.has_mappings) {
            T.gid[g] = fac_symbol_lookup(U, (const FacSymbol *)A.symbols.data(), (uint32_t)A.symbols.size() - 1, A.symbol_pool.data(), s, b, e, A.ci, fc);
        } else {
            uint32_t flen;
            fac_grapheme_hash(U, s, b, e, A.ci, fc, flen);
        }
        T.first[g] = fc;
    }
    T.off[n] = len;
    T.tv.first = T.first.data(); T.tv.gid = T.gid.data(); T.tv.off64 = T.off.data(); T.tv.n = (uint32_t)n;
}

static int g_faithful_flat = 0;   // 1: the order-faithful emulation evaluates states with flat_make_ctx<false> / flat_eval_slot<false>

extern "C" {

void emu_set_faithful_flat(int on) { g_faithful_flat = on; }

// segmentation only: returns grapheme count; fills starts/first (cap entries)
int emu_segment(const fac_config *cfg, const fac_pattern *pats, size_t np, const uint8_t *hay, size_t len, uint64_t *starts,
                uint32_t *first, uint32_t *gid, size_t cap) {
    HostAutomaton A; std::string err;
    if (build_automaton(cfg, pats, np, A, err) != FAC_OK) return -1;
    EmuText T; segment_host(A, hay, len, T);
    if (T.tv.ascii) return -2;
    for (size_t i = 0; i < T.first.size() && i < cap; i++) { starts[i] = T.off[i]; first[i] = T.first[i]; gid[i] = T.gid[i]; }
    return (int)T.first.size();
}

int emu_engine_info(const fac_config *cfg, const fac_pattern *pats, size_t np, uint64_t *info /*[8]*/) {
    HostAutomaton A; std::string err;
    fac_status st = build_automaton(cfg, pats, np, A, err);
    if (st != FAC_OK) return (int)st;
    info[0] = A.n_nodes(); info[1] = A.max_match_graphemes; info[2] = (uint64_t)A.max_edits_fast_raw; info[3] = A.bitap.active;
    info[4] = A.has_mappings; info[5] = A.wskip; info[6] = A.edge_char.size(); info[7] = A.out_pat.size();
    return 0;
}

int emu_search(const fac_config *cfg, const fac_pattern *pats, size_t np, const uint8_t *hay, size_t len, float thr, uint32_t tile,
               fac_match **out, size_t *n_out, uint64_t *states_out, uint32_t *per_window /* may be null, cap = graphemes */) {
    HostAutomaton HA; std::string err;
    fac_status st = build_automaton(cfg, pats, np, HA, err);
    if (st != FAC_OK) { fprintf(stderr, "emu build failed: %s\n", err.c_str()); return (int)st; }
    const AutomatonView A = HA.host_view();
    EmuText ET; segment_host(HA, hay, len, ET);
    const TextView &tv = ET.tv;
    FacTextDirect T{tv, A.ascii_gid, A.ci};
    const uint32_t n = tv.n;
    std::vector<FacCand> cands;
    uint64_t states = 0;
    if (tile == 0) tile = 32;
    const float maxpen = n ? FAC_SUB(A.node_prune_len[0], FAC_MUL(A.node_prune_low[0], thr)) : 0.f;  // search.rs:487
    typedef std::tuple<uint32_t, uint32_t, uint32_t> Key;
    // the merged-record helpers in their order-faithful mode (what k_beam_warp runs): same pushes in the same order
    const bool use_flat = g_faithful_flat && HA.flat_ok && HA.mef != 255;
    std::vector<FlatRec> nrec, erec;
    if (use_flat) {
        nrec.resize(HA.n_nodes()); erec.resize(HA.edge_char.size());
        for (uint32_t i = 0; i < HA.n_nodes(); i++) {
            union { float f; uint32_t u; } c;
            c.f = FAC_SUB(HA.node_prune_len[i], FAC_MUL(HA.node_prune_low[i], thr));
            nrec[i] = FlatRec{HA.flat_nrec[i * 4], HA.flat_nrec[i * 4 + 1], c.u, HA.flat_nrec[i * 4 + 3]};
        }
        for (size_t e = 0; e < erec.size(); e++)
            erec[e] = FlatRec{HA.flat_erec[e * 4], HA.flat_erec[e * 4 + 1], HA.flat_erec[e * 4 + 2], nrec[HA.flat_erec[e * 4] & 0x7FFFFFFFu].z};
    }
    const FlatView F{nrec.data(), erec.data(), HA.flat_ooff.data(), HA.flat_olist.data(), HA.flat_gm_row.data(), (const unsigned long long *)HA.flat_gm.data(),
                     HA.flat_pm_g ? (const unsigned long long *)HA.flat_pm.data() : nullptr, HA.flat_pm_g, HA.flat_pm_words, HA.flat_pm_k,
                     HA.flat_px_row.empty() ? nullptr : HA.flat_px_row.data(), (const unsigned long long *)HA.flat_px.data()};
    for (uint32_t tile_base = 0; tile_base < n; tile_base += tile) {
        const uint32_t cnt = std::min(tile, n - tile_base);
        const uint32_t text_end = n;
        std::vector<FacState> queue;
        std::vector<uint32_t> wcount(cnt, 0);
        for (uint32_t w = 0; w < cnt; w++) {
            const uint32_t start = tile_base + w;
            const bool has1 = start + 1 < text_end;
            if (fac_window_skipped(A, T.first(start), has1, has1 ? T.first(start + 1) : 0)) continue;
            queue.push_back(FacState{0, 0.f, 0, fac_make_pos(w, 0, 0)});
        }
        std::map<Key, float> persistent;  // engines with mappings: same key may recur on later levels
        size_t lb = 0;
        while (lb < queue.size()) {
            const size_t le = queue.size();
            std::map<Key, float> level_min;
            for (size_t i = lb; i < le; i++) {
                const FacState S = queue[i];
                wcount[S.pos >> FAC_POS_W_SHIFT]++;
                const Key key(S.node, S.cnt, S.pos);
                std::map<Key, float> &vis = A.has_mappings ? persistent : level_min;
                auto it = vis.find(key);
                if (it != vis.end()) { if (it->second <= S.pen) continue; it->second = S.pen; }
                else vis.emplace(key, S.pen);
                if (fac_over_ceiling(A, S.node, S.pen, thr)) continue;
                const uint32_t w = S.pos >> FAC_POS_W_SHIFT, mr = S.pos & FAC_POS_MASK;
                const uint32_t start = tile_base + w;
                for (uint32_t o = A.node_out_off[S.node]; o < A.node_out_off[S.node + 1]; o++) {
                    float sim;
                    if (fac_eval_output(A, thr, A.out_pat[o], S.pen, S.cnt, sim))
                        cands.push_back(FacCand{start, start + mr, A.out_pat[o], sim, S.cnt, (uint32_t)i, 0, 0});
                }
                FacCtx C;
                FlatCtx FC;
                if (use_flat) { flat_make_ctx<false>(A, F, T, maxpen, start, text_end, S, nrec[S.node], FC); C.nslots = FC.nslots; }
                else fac_make_ctx(A, T, maxpen, start, text_end, S, C);
                for (uint32_t s = 0; s < C.nslots; s++) {
                    FacState child;
                    if (use_flat ? flat_eval_slot<false>(A, F, T, maxpen, start, text_end, FC, s, child) : fac_eval_slot(A, T, maxpen, start, text_end, C, s, child)) {
                        const uint32_t jr = (child.pos >> FAC_POS_J_SHIFT) & FAC_POS_MASK;
                        if (jr >= FAC_MAX_SPAN) { fprintf(stderr, "emu: span overflow\n"); return 100; }
                        queue.push_back(child);
                    }
                }
            }
            lb = le;
        }
        states += queue.size();
        if (per_window) for (uint32_t w = 0; w < cnt; w++) per_window[tile_base + w] = wcount[w];
    }
    // candidate reduction: best per (start, end, pattern): max similarity, first in FIFO order on ties
    std::map<Key, FacCand> best;
    for (const FacCand &c : cands) {
        const Key k(c.sg, c.eg, c.pat);
        auto it = best.find(k);
        if (it == best.end()) best.emplace(k, c);
        else if (c.sim > it->second.sim || (c.sim == it->second.sim && c.seq < it->second.seq)) it->second = c;
    }
    std::vector<fac_match> res;
    for (auto &kv : best) {
        const FacCand &c = kv.second;
        fac_match m; memset(&m, 0, sizeof(m));
        m.start = c.sg < n ? fac_byte_offset(tv, c.sg) : 0;
        m.end = fac_byte_offset(tv, c.eg);
        m.pattern_index = c.pat; m.similarity = c.sim;
        m.insertions = c.cnt & 0xFF; m.deletions = (c.cnt >> 8) & 0xFF; m.substitutions = (c.cnt >> 16) & 0xFF; m.swaps = c.cnt >> 24;
        m.edits = (uint8_t)fac_edits_of(c.cnt);
        res.push_back(m);
    }
    std::sort(res.begin(), res.end(), [](const fac_match &a, const fac_match &b) {
        if (a.start != b.start) return a.start < b.start;
        if (a.end != b.end) return a.end < b.end;
        return a.pattern_index < b.pattern_index;
    });
    *n_out = res.size();
    *out = (fac_match *)malloc(sizeof(fac_match) * (res.size() ? res.size() : 1));
    if (!res.empty()) memcpy(*out, res.data(), sizeof(fac_match) * res.size());
    if (states_out) *states_out = states;
    return 0;
}

}  // extern "C"

// ---- succinct-trie fast path (csrc/fac_succinct.h) run sequentially: same per-state helpers as the
// kernel, a plain LIFO stack instead of the warp stack machine, the order-independent reduction with
// tie detection, and the faithful emulation above for tied ("dirty") windows. ----
struct EmuRecs { const SuccRec *r; SuccRec operator()(uint32_t n) const { return r[n]; } };
struct EmuSText {   // first chars / symbols of the haystack graphemes; positions past the end read as (0, NOSYM) like the kernel's padded tile
    const uint8_t *b; const uint32_t *first; const uint8_t *symof; bool ci; uint32_t n; uint32_t nosym;
    uint32_t byte(uint32_t j) const {
        if (j >= n) return 0;
        if (first) return first[j] < 128u ? first[j] : SUCC_NONASCII;   // K1 stream: already folded
        const uint32_t c = b[j];
        return (ci && c >= 'A' && c <= 'Z') ? c + 32u : c;
    }
    uint32_t sym(uint32_t j) const { if (j >= n) return nosym; const uint32_t c = byte(j); return c < 128u ? symof[c] : nosym; }
    uint32_t ctx(uint32_t j) const { return succ_ctx_pack(byte(j), sym(j), sym(j + 1), sym(j + 2), sym(j + 3)); }
};
template <bool W>
struct EmuGM {   // survivor-mask tables with the kernel's accessor interface (SuccGMDev)
    typedef typename SuccW<W>::M M;
    const uint64_t *gm, *gm2; uint32_t gm_nodes, gm2_nodes;
    const uint32_t *gm3, *pm3, *pm2, *pm4; uint32_t n3, np2, r3, n4;   // deep tables (narrow layout), host layout [node][..]
    bool four(uint32_t node) const { return !W && node < n4; }
    bool two_deep(uint32_t node) const { return node < gm2_nodes || node < n3; }
    uint32_t cl(uint32_t y) const { return std::min(y, r3 - 1u); }
    M row3(uint32_t node, uint32_t y1, uint32_t y2, uint32_t y3) const {
        if (W || node >= n3 || node >= gm_nodes) return row2(node, y1, y2);
        return (M)gm3[(((size_t)node * r3 + cl(y1)) * r3 + cl(y2)) * r3 + cl(y3)];
    }
    M deep(uint32_t node, bool q_pm, uint32_t a, uint32_t b, uint32_t c, uint32_t d) const { return q_pm ? pm(node, a, b, c, d) : row3(node, a, b, c); }
    M pm(uint32_t node, uint32_t a, uint32_t b, uint32_t c, uint32_t d) const {
        if (W || node >= np2) return ~M(0);
        if (node < n4) return (M)pm4[((((size_t)node * r3 + cl(a)) * r3 + cl(b)) * r3 + cl(c)) * r3 + cl(d)];
        if (node < n3) return (M)pm3[(((size_t)node * r3 + cl(a)) * r3 + cl(b)) * r3 + cl(c)];
        return (M)pm2[((size_t)node * r3 + cl(a)) * r3 + cl(b)];
    }
    M row(uint32_t node, uint32_t y) const { return node < gm_nodes ? (M)gm[(size_t)node * SuccW<W>::ROW + y] : ~M(0); }
    M row2(uint32_t node, uint32_t y1, uint32_t y2) const {
        if (node >= gm_nodes) return ~M(0);
        return node < gm2_nodes ? (M)gm2[((size_t)node * SuccW<W>::ROW + y1) * SuccW<W>::ROW + y2] : (M)gm[(size_t)node * SuccW<W>::ROW + y1];
    }
};
struct EmuEmit {
    std::vector<FacCand> *v;
    void operator()(uint32_t sg, uint32_t eg, uint32_t pat, float sim, uint32_t cnt) { v->push_back(FacCand{sg, eg, pat, sim, cnt, 0, 0, 0}); }
};

template <bool W, bool LIMM>
static void emu_succ_expand(const HostAutomaton &HA, const EmuText &ET, const uint8_t *hay, float thr, std::vector<FacCand> &cands, uint64_t &states) {
    typedef typename SuccW<W>::M M;
    const HostSuccinct &S = HA.succ;
    const uint32_t N = (uint32_t)S.bm.size(), n = ET.tv.n, text_end = n, NOSYM = SuccW<W>::NOSYM;
    std::vector<SuccRec> recs(N);
    for (uint32_t i = 0; i < N; i++) {
        union { float f; uint32_t u; } c;
        c.f = FAC_SUB(S.prune_len[i], FAC_MUL(S.prune_low[i], thr));
        if (W) recs[i] = SuccRec{(uint32_t)S.bm[i], (uint32_t)(S.bm[i] >> 32) | (S.out_idx[i] != FAC_NONE ? 0x80000000u : 0u),
                                 S.fc[i] | ((uint32_t)S.insym[i] << SuccW<W>::SYM_SHIFT), c.u};
        else recs[i] = SuccRec{(uint32_t)S.bm[i], S.fc[i] | ((uint32_t)S.insym[i] << SuccW<W>::SYM_SHIFT), c.u, S.out_idx[i]};
    }
    SuccConsts K;
    K.thr = thr; K.maxpen = succ_ceil<W>(recs[0]); K.pen_ins = HA.pen_ins; K.pen_del = HA.pen_del; K.pen_swap = HA.pen_swap;
    K.mef = S.limits_mode ? (int32_t)S.edit_bound : HA.mef;
    K.lim = HA.lim.data(); K.node_lim = S.node_lim.data(); K.has_global = HA.has_global_limits; K.out_idx = S.out_idx.data();
    const EmuRecs R{recs.data()};
    const EmuGM<W> G{S.gmask.data(), S.gmask2.data(), S.gm_nodes, S.gm2_nodes, S.gmask3.data(), S.pmask3.data(), S.pmask2.data(), S.pmask4.data(), W ? 0u : S.n3, W ? 0u : S.np2, S.r3, W ? 0u : S.n4};
    const EmuSText T{hay, ET.tv.ascii ? nullptr : ET.first.data(), S.sym_of, HA.ci, n, NOSYM};
    const SuccOut *out2 = (const SuccOut *)S.out2.data();
    EmuEmit emit{&cands};
    for (uint32_t start = 0; start < n; start++) {
        if (HA.wskip) {
            if (T.byte(start) != SUCC_NONASCII && !((S.first_mask >> T.sym(start)) & 1u)) {
                if (start + 1 >= n) continue;
                if (T.byte(start + 1) != SUCC_NONASCII && !((S.second_mask >> T.sym(start + 1)) & 1u)) continue;
            }
        }
        if (S.exact_only) {  // engine without FuzzyLimits: only the exact chain from the root can emit
            states += succ_walk<false, W>(K, R, out2, T, emit, start, text_end, 0u, R(0u), 0.f, 0u, 0u, 0u);
            continue;
        }
        std::vector<FacState> stack;
        stack.push_back(FacState{0, 0.f, 0, 0});
        while (!stack.empty()) {
            const FacState s = stack.back();
            stack.pop_back();
            states++;
            const SuccRec rec = R(s.node);
            if (s.pen > succ_ceil<W>(rec)) continue;
            if (succ_has_out<W>(rec)) succ_outputs<LIMM>(K, out2, emit, succ_out_idx<W>(K, rec, s.node), s.pen, s.cnt, start, start + succ_mr(s.pos));
            SuccCtx2<W> C;
            succ_make_ctx2<LIMM, W>(K, T, G, start, text_end, s.node, rec, s.pen, s.cnt, s.pos, C);
            const bool last = (C.flags & SUCC_F_LAST) != 0;
            const uint32_t jr = succ_jr(s.pos);
            auto child = [&](const FacState &c) {
                if (last) states += succ_walk<LIMM, W>(K, R, out2, T, emit, start, text_end, c.node, R(c.node), c.pen, c.cnt, succ_jr(c.pos), succ_mr(c.pos));
                else stack.push_back(c);
            };
            const uint32_t cur_s = succ_ctx_s0(C.packed);
            if (C.flags & SUCC_F_EXACT) stack.push_back(FacState{succ_child<W>(rec, cur_s), s.pen, s.cnt, succ_repos(s.pos, jr + 1, jr + 1)});
            FacState c;
            if (succ_swap2<LIMM, W>(K, R, C, c)) child(c);
            if (succ_ins2<LIMM, W>(K, C, s.node, c)) child(c);
            const uint32_t n_items = succ_popc(C.sub_m) + succ_popc(C.del_m);
            for (uint32_t r = 0; r < n_items; r++)
                if (succ_item2<W>(K, S.sub_pen.data(), C, r, c)) child(c);
        }
    }
    (void)M(0);
}

extern "C" {

// returns 0 ok, -3 engine/haystack outside the fast kernel's domain.  info[0] = dirty windows, info[1] = states visited
int emu_search_succinct(const fac_config *cfg, const fac_pattern *pats, size_t np, const uint8_t *hay, size_t len, float thr,
                        fac_match **out, size_t *n_out, uint64_t *info) {
    HostAutomaton HA; std::string err;
    fac_status st = build_automaton(cfg, pats, np, HA, err);
    if (st != FAC_OK) return (int)st;
    const HostSuccinct &S = HA.succ;
    if (!S.ok || HA.beam_width != 0 || HA.has_auto_beam) return -3;
    EmuText ET; segment_host(HA, hay, len, ET);
    if (!ET.tv.ascii && !S.unicode_text_ok) return -3;
    const uint32_t n = ET.tv.n;
    std::vector<FacCand> cands;
    uint64_t states = 0;
    if (S.wide) { if (S.limits_mode) emu_succ_expand<true, true>(HA, ET, hay, thr, cands, states); else emu_succ_expand<true, false>(HA, ET, hay, thr, cands, states); }
    else { if (S.limits_mode) emu_succ_expand<false, true>(HA, ET, hay, thr, cands, states); else emu_succ_expand<false, false>(HA, ET, hay, thr, cands, states); }
    typedef std::tuple<uint32_t, uint32_t, uint32_t> Key;
    struct Best { float sim; uint32_t cmin, cmax; };
    std::map<Key, Best> best;
    for (const FacCand &c : cands) {
        const Key k(c.sg, c.eg, c.pat);
        auto it = best.find(k);
        if (it == best.end()) best.emplace(k, Best{c.sim, c.cnt, c.cnt});
        else if (c.sim > it->second.sim) it->second = Best{c.sim, c.cnt, c.cnt};
        else if (c.sim == it->second.sim) { it->second.cmin = std::min(it->second.cmin, c.cnt); it->second.cmax = std::max(it->second.cmax, c.cnt); }
    }
    std::vector<uint8_t> dirty(n + 1, 0);
    uint64_t n_dirty = 0;
    for (auto &kv : best) if (kv.second.cmin != kv.second.cmax && !dirty[std::get<0>(kv.first)]) { dirty[std::get<0>(kv.first)] = 1; n_dirty++; }
    std::vector<fac_match> res;
    for (auto &kv : best) {
        if (dirty[std::get<0>(kv.first)]) continue;
        fac_match m; memset(&m, 0, sizeof(m));
        const uint32_t cnt = kv.second.cmin;
        m.start = fac_byte_offset(ET.tv, std::get<0>(kv.first)); m.end = fac_byte_offset(ET.tv, std::get<1>(kv.first));
        m.pattern_index = std::get<2>(kv.first); m.similarity = kv.second.sim;
        m.insertions = cnt & 0xFF; m.deletions = (cnt >> 8) & 0xFF; m.substitutions = (cnt >> 16) & 0xFF; m.swaps = cnt >> 24;
        m.edits = (uint8_t)fac_edits_of(cnt);
        res.push_back(m);
    }
    if (n_dirty) {
        fac_match *fm = nullptr; size_t fn = 0; uint64_t fs = 0;
        const int rc = emu_search(cfg, pats, np, hay, len, thr, 16, &fm, &fn, &fs, nullptr);
        if (rc != 0) return rc;
        std::vector<uint8_t> dirty_byte(len + 1, 0);
        for (uint32_t g = 0; g < n; g++) if (dirty[g]) dirty_byte[fac_byte_offset(ET.tv, g)] = 1;
        for (size_t i = 0; i < fn; i++) if (dirty_byte[fm[i].start]) res.push_back(fm[i]);
        free(fm);
    }
    std::sort(res.begin(), res.end(), [](const fac_match &a, const fac_match &b) {
        if (a.start != b.start) return a.start < b.start;
        if (a.end != b.end) return a.end < b.end;
        return a.pattern_index < b.pattern_index;
    });
    *n_out = res.size();
    *out = (fac_match *)malloc(sizeof(fac_match) * (res.size() ? res.size() : 1));
    if (!res.empty()) memcpy(*out, res.data(), sizeof(fac_match) * res.size());
    if (info) { info[0] = n_dirty; info[1] = states; info[2] = cands.size(); info[3] = best.size(); }
    return 0;
}

// General stack-machine fast path (csrc/fac_flat.h + fac_stack.cuh) run sequentially: merged records, a plain LIFO
// stack, exhausted children walked, the order-independent reduction with tie detection, faithful emulation for tied
// windows.  returns 0 ok, -3 engine outside the kernel's domain.  info[0] = dirty windows, info[1] = states visited
int emu_search_flat(const fac_config *cfg, const fac_pattern *pats, size_t np, const uint8_t *hay, size_t len, float thr,
                    fac_match **out, size_t *n_out, uint64_t *info) {
    HostAutomaton HA; std::string err;
    fac_status st = build_automaton(cfg, pats, np, HA, err);
    if (st != FAC_OK) return (int)st;
    if (!HA.flat_ok || HA.mef == 255 || HA.beam_width != 0 || HA.has_auto_beam) return -3;
    const AutomatonView A = HA.host_view();
    EmuText ET; segment_host(HA, hay, len, ET);
    const TextView &tv = ET.tv;
    FacTextDirect T{tv, A.ascii_gid, A.ci};
    const uint32_t n = tv.n, N = HA.n_nodes();
    std::vector<FlatRec> nrec(N), erec(HA.edge_char.size());
    for (uint32_t i = 0; i < N; i++) {
        union { float f; uint32_t u; } c;
        c.f = FAC_SUB(HA.node_prune_len[i], FAC_MUL(HA.node_prune_low[i], thr));
        nrec[i] = FlatRec{HA.flat_nrec[i * 4], HA.flat_nrec[i * 4 + 1], c.u, HA.flat_nrec[i * 4 + 3]};
    }
    for (size_t e = 0; e < erec.size(); e++)
        erec[e] = FlatRec{HA.flat_erec[e * 4], HA.flat_erec[e * 4 + 1], HA.flat_erec[e * 4 + 2], nrec[HA.flat_erec[e * 4] & 0x7FFFFFFFu].z};
    const FlatView F{nrec.data(), erec.data(), HA.flat_ooff.data(), HA.flat_olist.data(), HA.flat_gm_row.data(), (const unsigned long long *)HA.flat_gm.data(),
                     HA.flat_pm_g ? (const unsigned long long *)HA.flat_pm.data() : nullptr, HA.flat_pm_g, HA.flat_pm_words, HA.flat_pm_k,
                     HA.flat_px_row.empty() ? nullptr : HA.flat_px_row.data(), (const unsigned long long *)HA.flat_px.data()};
    std::vector<FacCand> cands;
    EmuEmit emit{&cands};
    uint64_t states = 0;
    const float maxpen = n ? FAC_SUB(A.node_prune_len[0], FAC_MUL(A.node_prune_low[0], thr)) : 0.f;
    for (uint32_t start = 0; start < n; start++) {
        const bool has1 = start + 1 < n;
        if (fac_window_skipped(A, T.first(start), has1, has1 ? T.first(start + 1) : 0)) continue;
        std::vector<FacState> stack;
        stack.push_back(FacState{0, 0.f, 0, 0});
        while (!stack.empty()) {
            const FacState s = stack.back();
            stack.pop_back();
            const FlatRec nr = nrec[s.node];
            if (s.pen > FLAT_AS_FLOAT(nr.z)) continue;
            states++;
            for (uint32_t o = 0; o < flat_nout(nr); o++) {
                const uint32_t pat = A.out_pat[nr.w + o];
                const float total = A.pat_glen[pat];
                const float sim = FAC_MUL(FAC_DIV(FAC_SUB(total, s.pen), total), A.pat_weight[pat]);
                if (!(sim < thr)) emit(start, start + (s.pos & FAC_POS_MASK), pat, sim, s.cnt);
            }
            FlatCtx C;
            flat_make_ctx<true>(A, F, T, maxpen, start, n, s, nr, C);
            for (uint32_t k = 0; k < C.nslots; k++) {
                FacState c;
                if (!flat_eval_slot<true>(A, F, T, maxpen, start, n, C, k, c)) continue;
                if ((int)fac_edits_of(c.cnt) >= A.mef) states += flat_walk(A, F, T, thr, emit, start, n, c);
                else stack.push_back(c);
            }
        }
    }
    typedef std::tuple<uint32_t, uint32_t, uint32_t> Key;
    struct Best { float sim; uint32_t cmin, cmax; };
    std::map<Key, Best> best;
    for (const FacCand &c : cands) {
        const Key k(c.sg, c.eg, c.pat);
        auto it = best.find(k);
        if (it == best.end()) best.emplace(k, Best{c.sim, c.cnt, c.cnt});
        else if (c.sim > it->second.sim) it->second = Best{c.sim, c.cnt, c.cnt};
        else if (c.sim == it->second.sim) { it->second.cmin = std::min(it->second.cmin, c.cnt); it->second.cmax = std::max(it->second.cmax, c.cnt); }
    }
    std::vector<uint8_t> dirty(n + 1, 0);
    uint64_t n_dirty = 0;
    for (auto &kv : best) if (kv.second.cmin != kv.second.cmax && !dirty[std::get<0>(kv.first)]) { dirty[std::get<0>(kv.first)] = 1; n_dirty++; }
    std::vector<fac_match> res;
    for (auto &kv : best) {
        if (dirty[std::get<0>(kv.first)]) continue;
        fac_match m; memset(&m, 0, sizeof(m));
        const uint32_t cnt = kv.second.cmin;
        m.start = fac_byte_offset(tv, std::get<0>(kv.first)); m.end = fac_byte_offset(tv, std::get<1>(kv.first));
        m.pattern_index = std::get<2>(kv.first); m.similarity = kv.second.sim;
        m.insertions = cnt & 0xFF; m.deletions = (cnt >> 8) & 0xFF; m.substitutions = (cnt >> 16) & 0xFF; m.swaps = cnt >> 24;
        m.edits = (uint8_t)fac_edits_of(cnt);
        res.push_back(m);
    }
    if (n_dirty) {
        fac_match *fm = nullptr; size_t fn = 0; uint64_t fs = 0;
        const int rc = emu_search(cfg, pats, np, hay, len, thr, 16, &fm, &fn, &fs, nullptr);
        if (rc != 0) return rc;
        std::vector<uint8_t> dirty_byte(len + 1, 0);
        for (uint32_t g = 0; g < n; g++) if (dirty[g]) dirty_byte[fac_byte_offset(tv, g)] = 1;
        for (size_t i = 0; i < fn; i++) if (dirty_byte[fm[i].start]) res.push_back(fm[i]);
        free(fm);
    }
    std::sort(res.begin(), res.end(), [](const fac_match &a, const fac_match &b) {
        if (a.start != b.start) return a.start < b.start;
        if (a.end != b.end) return a.end < b.end;
        return a.pattern_index < b.pattern_index;
    });
    *n_out = res.size();
    *out = (fac_match *)malloc(sizeof(fac_match) * (res.size() ? res.size() : 1));
    if (!res.empty()) memcpy(*out, res.data(), sizeof(fac_match) * res.size());
    if (info) { info[0] = n_dirty; info[1] = states; info[2] = cands.size(); info[3] = best.size(); }
    return 0;
}

void emu_free(void *p) { free(p); }

}  // extern "C"

// ---- definition checks of the builder's productivity / survivor tables (TEST CODE) ----
// The builder constructs the tables by bit-parallel / pattern enumeration; here every cell is recomputed by evaluating the
// definition written in the header of build_deep_tables / build_flat_pm directly (plain recursion), on engines small
// enough that all cells can be visited.  info[0] = cells compared, info[1] = mismatches, info[2..] = table sizes.
namespace {
struct SuccDef {
    const HostSuccinct &S; uint32_t R;
    bool out(uint32_t n) const { return S.out_idx[n] != FAC_NONE; }
    bool edge(uint32_t n, uint32_t y) const { return y < S.n_syms && ((S.bm[n] >> y) & 1u); }
    uint32_t kid(uint32_t n, uint32_t y) const { return S.fc[n] + (uint32_t)__builtin_popcountll(S.bm[n] & ((1ull << y) - 1ull)); }
    bool W(uint32_t d, const uint32_t *sy, int k) const {
        if (out(d)) return true;
        if (k == 0) return true;
        return edge(d, sy[0]) && W(kid(d, sy[0]), sy + 1, k - 1);
    }
    bool L(uint32_t c, const uint32_t *sy, int k) const {
        if (out(c) || k == 0) return true;
        const uint32_t a = sy[0];
        if (edge(c, a) && L(kid(c, a), sy + 1, k - 1)) return true;                       // exact child
        for (uint32_t s = 0; s < S.n_syms; s++) {
            if (!edge(c, s)) continue;
            const uint32_t d = kid(c, s);
            if (s != a && W(d, sy + 1, k - 1)) return true;                                // substitution child
            if (W(d, sy, k)) return true;                                                  // deletion child
        }
        if (W(c, sy + 1, k - 1)) return true;                                              // insertion child
        for (uint32_t b = 0; b < S.n_syms; b++) {                                          // swap child: c -b-> x -a-> n2
            if (!edge(c, b) || (k >= 2 && sy[1] != b)) continue;
            const uint32_t x = kid(c, b);
            if (edge(x, a) && (k < 2 || W(kid(x, a), sy + 2, k - 2))) return true;
        }
        return false;
    }
};
}  // namespace

extern "C" int emu_check_deep_tables(const fac_config *cfg, const fac_pattern *pats, size_t np, uint64_t *info /*[8]*/) {
    HostAutomaton HA; std::string err;
    fac_status st = build_automaton(cfg, pats, np, HA, err);
    if (st != FAC_OK) return (int)st;
    const HostSuccinct &S = HA.succ;
    if (!S.ok || S.wide || S.exact_only || S.r3 == 0) return -3;
    const uint32_t R = S.r3;
    const SuccDef D{S, R};
    uint64_t cells = 0, bad = 0;
    auto children_mask = [&](uint32_t p, const std::function<bool(uint32_t)> &f) {
        uint32_t m = 0;
        for (uint32_t s = 0; s < S.n_syms; s++) if (D.edge(p, s) && f(D.kid(p, s))) m |= 1u << s;
        return m;
    };
    uint32_t sy[4];
    for (uint32_t p = 0; p < S.np2; p++) {
        const int k = p < S.n4 ? 4 : (p < S.n3 ? 3 : 2);
        size_t total = 1; for (int i = 0; i < k; i++) total *= R;
        for (size_t idx = 0; idx < total; idx++) {
            size_t t = idx; for (int i = k - 1; i >= 0; i--) { sy[i] = (uint32_t)(t % R); t /= R; }
            const uint32_t want = children_mask(p, [&](uint32_t c) { return D.L(c, sy, k); });
            const uint32_t got = k == 4 ? S.pmask4[(size_t)p * total + idx] : (k == 3 ? S.pmask3[(size_t)p * total + idx] : S.pmask2[(size_t)p * total + idx]);
            cells++; if (want != got) bad++;
        }
        if (p < S.n3 && p < S.n4) {   // nodes of the 4-symbol table also own 3-symbol rows
            size_t t3 = (size_t)R * R * R;
            for (size_t idx = 0; idx < t3; idx++) {
                size_t t = idx; for (int i = 2; i >= 0; i--) { sy[i] = (uint32_t)(t % R); t /= R; }
                const uint32_t want = children_mask(p, [&](uint32_t c) { return D.L(c, sy, 3); });
                cells++; if (want != S.pmask3[(size_t)p * t3 + idx]) bad++;
            }
        }
    }
    for (uint32_t p = 0; p < S.n3; p++) {   // gmask3[p][y1][y2][y3] bit s: d = child(p, s) has edge y1 and W(child(d, y1); y2, y3)
        const size_t t3 = (size_t)R * R * R;
        for (size_t idx = 0; idx < t3; idx++) {
            size_t t = idx; for (int i = 2; i >= 0; i--) { sy[i] = (uint32_t)(t % R); t /= R; }
            const uint32_t want = children_mask(p, [&](uint32_t d) { return D.edge(d, sy[0]) && D.W(D.kid(d, sy[0]), sy + 1, 2); });
            cells++; if (want != S.gmask3[(size_t)p * t3 + idx]) bad++;
        }
    }
    info[0] = cells; info[1] = bad; info[2] = S.n3; info[3] = S.np2; info[4] = S.n4; info[5] = R; info[6] = S.bm.size();
    return 0;
}

namespace {
struct FlatDef {
    const HostAutomaton &A; uint32_t G;
    bool out(uint32_t n) const { return A.node_out_off[n + 1] != A.node_out_off[n]; }
    bool maps(uint32_t n) const { return A.node_map_off[n + 1] != A.node_map_off[n]; }
    uint32_t next(uint32_t n, uint32_t key) const {
        for (uint32_t e = A.node_edge_off[n]; e < A.node_edge_off[n + 1]; e++) if (A.edge_sym[e] == key) return A.edge_next[e] & 0x7FFFFFFFu;
        return FAC_NONE;
    }
    template <class F> bool any_child(uint32_t n, F f) const {
        for (uint32_t e = A.node_edge_off[n]; e < A.node_edge_off[n + 1]; e++) if (f(A.edge_next[e] & 0x7FFFFFFFu)) return true;
        return false;
    }
    bool W(uint32_t d, const uint32_t *sy, int k) const {
        if (out(d) || k == 0) return true;
        const uint32_t g = sy[0] ? next(d, sy[0]) : FAC_NONE;
        return g != FAC_NONE && W(g, sy + 1, k - 1);
    }
    // a last-edit state at x whose exact child is not expanded further (the builder's last_state)
    bool Lx(uint32_t x, const uint32_t *sy, int k) const {
        if (out(x) || maps(x) || k == 0) return true;
        if (sy[0] && next(x, sy[0]) != FAC_NONE) return true;                                                   // exact child
        if (any_child(x, [&](uint32_t d) { return W(d, sy + 1, k - 1) || W(d, sy, k); })) return true;           // substitution / deletion
        if (W(x, sy + 1, k - 1)) return true;                                                                   // insertion
        for (uint32_t e = A.node_edge_off[x]; e < A.node_edge_off[x + 1]; e++) {                                // swap: x -sy[1]-> y -sy[0]-> n2
            if (k >= 2 && A.edge_sym[e] != sy[1]) continue;
            const uint32_t y = A.edge_next[e] & 0x7FFFFFFFu;
            const uint32_t n2 = sy[0] ? next(y, sy[0]) : FAC_NONE;
            if (n2 != FAC_NONE && (k < 2 || W(n2, sy + 2, k - 2))) return true;
        }
        return false;
    }
    // a root child: same, with the exact child expanded one level
    bool Lc(uint32_t c, const uint32_t *sy, int k) const {
        if (out(c) || maps(c)) return true;
        const uint32_t x = sy[0] ? next(c, sy[0]) : FAC_NONE;
        if (x != FAC_NONE && Lx(x, sy + 1, k - 1)) return true;
        if (any_child(c, [&](uint32_t d) { return W(d, sy + 1, k - 1) || W(d, sy, k); })) return true;
        if (W(c, sy + 1, k - 1)) return true;
        const uint32_t x1 = sy[1] ? next(c, sy[1]) : FAC_NONE;
        if (x1 != FAC_NONE) { const uint32_t n2 = sy[0] ? next(x1, sy[0]) : FAC_NONE; if (n2 != FAC_NONE && W(n2, sy + 2, k - 2)) return true; }
        return false;
    }
};
}  // namespace

extern "C" int emu_check_flat_pm(const fac_config *cfg, const fac_pattern *pats, size_t np, uint64_t *info /*[8]*/) {
    HostAutomaton A; std::string err;
    fac_status st = build_automaton(cfg, pats, np, A, err);
    if (st != FAC_OK) return (int)st;
    if (A.flat_pm_g == 0) return -3;
    const uint32_t G = A.flat_pm_g, K = A.flat_pm_k, words = A.flat_pm_words;
    const FlatDef D{A, G};
    const uint32_t e0 = A.node_edge_off[0], deg0 = A.node_edge_off[1] - e0;
    uint64_t cells = 0, bad = 0;
    uint32_t sy[3];
    size_t total = 1; for (uint32_t i = 0; i < K; i++) total *= G;
    for (size_t idx = 0; idx < total; idx++) {
        size_t t = idx; for (int i = (int)K - 1; i >= 0; i--) { sy[i] = (uint32_t)(t % G); t /= G; }
        for (uint32_t e = 0; e < deg0; e++) {
            const bool want = D.Lc(A.edge_next[e0 + e] & 0x7FFFFFFFu, sy, (int)K);
            const bool got = (A.flat_pm[idx * words + (e >> 6)] >> (e & 63)) & 1ull;
            cells++; if (want != got) bad++;
        }
    }
    uint64_t rows = 0;
    if (!A.flat_px_row.empty()) {
        const size_t row_words = ((size_t)G * G + 63) / 64;
        for (uint32_t x = 0; x < A.n_nodes(); x++) {
            const uint32_t r = A.flat_px_row[x];
            if (r == FAC_NONE) continue;
            rows++;
            for (uint32_t b = 0; b < G; b++)
                for (uint32_t c3 = 0; c3 < G; c3++) {
                    const uint32_t s2[2] = {b, c3};
                    const bool want = D.Lx(x, s2, 2);
                    const size_t bp = (size_t)b * G + c3;
                    const bool got = (A.flat_px[r * row_words + (bp >> 6)] >> (bp & 63)) & 1ull;
                    cells++; if (want != got) bad++;
                }
        }
    }
    info[0] = cells; info[1] = bad; info[2] = G; info[3] = K; info[4] = deg0; info[5] = rows;
    return 0;
}
