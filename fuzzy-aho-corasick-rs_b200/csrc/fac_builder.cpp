// fac_builder.cpp -- host-side construction of the flattened automaton (see fac_builder.h).
#include "fac_builder.h"
#include "fac_succinct.h"
#include <limits>
#include <functional>
#include <cstdlib>
#define FAC_POPC_HOST(x) __builtin_popcount(x)
#include "fac_core.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <deque>
#include <map>
#include <unordered_map>

#define FAC_TABLE_QUAL static const
#include "unicode_tables.h"

namespace fac {

const UnicodeTables &host_unicode_tables() {
    static const UnicodeTables U = {FAC_GCB_STAGE1, FAC_GCB_STAGE2, FAC_LOWER_KEYS, FAC_LOWER_VALS, FAC_LOWER_N};
    return U;
}

size_t utf8_valid_up_to(const uint8_t *s, size_t n) {
    size_t i = 0;
    while (i < n) {
        const uint8_t b = s[i];
        if (b < 0x80) { i++; continue; }
        size_t extra;
        uint8_t lo = 0x80, hi = 0xBF;  // allowed range of the first continuation byte
        if (b >= 0xC2 && b <= 0xDF) extra = 1;
        else if (b == 0xE0) { extra = 2; lo = 0xA0; }
        else if (b == 0xED) { extra = 2; hi = 0x9F; }
        else if (b >= 0xE1 && b <= 0xEF) extra = 2;
        else if (b == 0xF0) { extra = 3; lo = 0x90; }
        else if (b >= 0xF1 && b <= 0xF3) extra = 3;
        else if (b == 0xF4) { extra = 3; hi = 0x8F; }
        else return i;
        if (n - i <= extra) return i;
        if (s[i + 1] < lo || s[i + 1] > hi) return i;
        for (size_t k = 2; k <= extra; k++)
            if ((s[i + k] & 0xC0) != 0x80) return i;
        i += extra + 1;
    }
    return i;
}

static void grapheme_bounds(const std::string &s, std::vector<size_t> &starts) {
    const UnicodeTables &U = host_unicode_tables();
    const uint8_t *p = (const uint8_t *)s.data();
    starts.clear();
    for (size_t i = 0; i < s.size(); i++) {
        if (fac_is_cont(p[i])) continue;
        if (fac_break_before(U, p, 0, i)) starts.push_back(i);
    }
}

size_t count_graphemes(const std::string &s) {
    std::vector<size_t> st;
    grapheme_bounds(s, st);
    return st.size();
}

static std::string lower_grapheme(const uint8_t *p, size_t b, size_t e) {
    const UnicodeTables &U = host_unicode_tables();
    std::string out;
    uint64_t i = b;
    while (i < e) {
        uint32_t cp = fac_decode(p, i), lo[2];
        int n = fac_lower(U, cp, lo);
        for (int k = 0; k < n; k++) {
            uint8_t buf[4];
            int nb = fac_encode(lo[k], buf);
            out.append((const char *)buf, nb);
        }
    }
    return out;
}

std::vector<std::string> fold_graphemes(const std::string &s, bool ci) {
    std::vector<size_t> st;
    grapheme_bounds(s, st);
    std::vector<std::string> out;
    out.reserve(st.size());
    for (size_t k = 0; k < st.size(); k++) {
        const size_t b = st[k], e = k + 1 < st.size() ? st[k + 1] : s.size();
        out.push_back(ci ? lower_grapheme((const uint8_t *)s.data(), b, e) : s.substr(b, e - b));
    }
    return out;
}

static uint32_t first_scalar(const std::string &g) {
    if (g.empty()) return 0;
    uint64_t i = 0;
    return fac_decode((const uint8_t *)g.data(), i);
}

static uint64_t fnv1a(const std::string &g) {
    uint64_t h = 0xCBF29CE484222325ull;
    for (unsigned char c : g) { h ^= c; h *= 0x100000001B3ull; }
    return h;
}

static FacLimits limits_from_c(const fac_limits &c, bool finalize) {
    FacLimits l;
    memset(&l, 0, sizeof(l));
    l.ins = c.insertions; l.del = c.deletions; l.sub = c.substitutions; l.swp = c.swaps; l.edits = c.edits;
    auto norm = [](int16_t v) -> int16_t { return v < 0 ? (int16_t)-1 : (v > 255 ? (int16_t)255 : v); };
    l.ins = norm(l.ins); l.del = norm(l.del); l.sub = norm(l.sub); l.swp = norm(l.swp); l.edits = norm(l.edits);
    if (finalize && l.edits < 0) {  // FuzzyLimits::finalize, src/structs.rs:319-335
        if (l.ins < 0) l.ins = 0;
        if (l.del < 0) l.del = 0;
        if (l.sub < 0) l.sub = 0;
        if (l.swp < 0) l.swp = 0;
    }
    return l;
}

// Open-addressing symbol table keyed by FNV-1a of the folded grapheme bytes.
static void build_symbol_table(const std::vector<std::string> &syms, std::vector<HostSymbol> &tab, std::vector<uint8_t> &pool) {
    size_t cap = 16;
    while (cap < syms.size() * 2 + 2) cap <<= 1;
    tab.assign(cap, HostSymbol{0, 0, 0, 0, 0});
    pool.clear();
    for (size_t i = 0; i < syms.size(); i++) {
        const uint64_t h = fnv1a(syms[i]);
        size_t slot = (size_t)(h & (cap - 1));
        while (tab[slot].gid != 0) slot = (slot + 1) & (cap - 1);
        tab[slot] = HostSymbol{h, (uint32_t)(i + 1), (uint32_t)pool.size(), (uint32_t)syms[i].size(), 0};
        pool.insert(pool.end(), syms[i].begin(), syms[i].end());
    }
    pool.resize(pool.size() + 8, 0);
}

namespace {
struct TmpNode {
    std::vector<std::pair<std::string, uint32_t>> order;  // edges in first-insertion order
    std::unordered_map<std::string, uint32_t> trans;
    std::vector<uint32_t> output;
    uint32_t fail = 0;
    int64_t pattern_index = -1;
};
}  // namespace

static void build_default_similarity(std::map<std::pair<uint32_t, uint32_t>, float> &m) {  // src/builder.rs:492-526
    const std::string vowels = "aeiou";
    for (char a : vowels)
        for (char b : vowels)
            if (a != b) m[{(uint32_t)a, (uint32_t)b}] = 0.6f;
    for (char a = 'a'; a <= 'z'; a++)
        for (char b = 'a'; b <= 'z'; b++)
            if (a != b && vowels.find(a) == std::string::npos && vowels.find(b) == std::string::npos) m[{(uint32_t)a, (uint32_t)b}] = 0.4f;
    auto both = [&](char a, char b, float v) { m[{(uint32_t)a, (uint32_t)b}] = v; m[{(uint32_t)b, (uint32_t)a}] = v; };
    both('o', '0', 0.6f); both('l', '1', 0.7f); both('i', '1', 0.6f); both('s', '5', 0.5f);
}

static bool k_from_limits(const FacLimits &l, size_t &k) {  // src/prefilter.rs:388-405
    if (l.edits >= 0) { k = (l.swp == 0) ? (size_t)l.edits : 2 * (size_t)l.edits; return true; }
    if (l.ins < 0 || l.del < 0 || l.sub < 0 || l.swp < 0) return false;
    k = (size_t)l.ins + (size_t)l.del + (size_t)l.sub + 2 * (size_t)l.swp;
    return true;
}

// Deep tables of the succinct fast kernel (narrow layout; semantics in fac_succinct.h).  Both are statements about what
// the reference's expansion (src/search.rs:418-1119) can still EMIT from a state, given the next k text symbols; a symbol
// index of n_syms means "no alphabet symbol here" (foreign char or end of text), symbols beyond the known ones are unknown
// and every test on them is answered "maybe".  Penalties, ceilings and per-pattern limits only remove states, so ignoring
// them keeps the tables conservative.
//
//   W(d; y1..yk)     an exhausted state (exact transitions only, search.rs:810, 937, 1003, 1043) visiting node d with the
//                    next symbols y1..yk can emit:  out(d) || (k == 0 ? true : edge(d, y1) && W(child(d, y1); y2..yk))
//   L(c; t0..tk-1)   a state on its LAST edit at node c whose own position reads t0, then t1 .. can emit: out(c), or its exact
//                    child (L(child(c, t0); t1..)), a substitution child (W(d; t1..), d != child(c, t0)), a deletion child
//                    (W(d; t0..)), the insertion child (W(c; t1..)) or the swap child (c -t1-> x -t0-> n2, W(n2; t2..))
//
//   gmask3[c][y1][y2][y3]   bit s:  d = child(c, s) has edge y1 and W(child(d, y1); y2, y3)
//   pmask4[p][a][b][c][d]   bit s:  L(child(p, s); a, b, c, d)        (first n4 nodes)
//   pmask3[p][a][b][c]      bit s:  L(child(p, s); a, b, c)           (first n3 nodes)
//   pmask2[p][a][b]         bit s:  L(child(p, s); a, b)              (first np2 nodes)
static void build_deep_tables(HostSuccinct &S) {
    const uint32_t N = (uint32_t)S.bm.size();
    const uint32_t R = S.n_syms + 1u;   // <= 32 (narrow layout)
    const size_t RP[5] = {1, R, (size_t)R * R, (size_t)R * R * R, (size_t)R * R * R * R};
    // node counts: default = what fits the per-table budget (sized on the measured cfg2 throughput), overridable, never above
    // 256 MB a table.
    // Small tries get proportionally smaller tables (64 KB of table per trie node, at least 1 MB): creating a 30-node engine
    // must not allocate and upload 200 MB.
    const size_t scaled = std::max<size_t>((size_t)1 << 20, (size_t)N * (64u << 10));
    auto envn = [&](const char *name, size_t entries_per_node, size_t budget) {
        const char *ev = getenv(name);
        const size_t cap = ((size_t)256 << 20) / (entries_per_node * 4);
        return std::min(cap, ev && *ev ? (size_t)atoll(ev) : std::min(budget, scaled) / (entries_per_node * 4));
    };
    S.r3 = R;
    S.n3 = (uint32_t)std::min<size_t>(std::min<uint32_t>(N, S.gm_nodes), envn("FAC_GM3_NODES", RP[3], (size_t)48 << 20));
    S.np2 = (uint32_t)std::min<size_t>(N, envn("FAC_PM2_NODES", RP[2], (size_t)32 << 20));
    S.n4 = (uint32_t)std::min<size_t>(S.n3, envn("FAC_PM4_NODES", RP[4], (size_t)64 << 20));
    S.np2 = std::max(S.np2, S.n3);
    S.gmask3.assign(std::max<size_t>((size_t)S.n3 * RP[3], 1), 0);
    S.pmask3.assign(std::max<size_t>((size_t)S.n3 * RP[3], 1), 0);
    S.pmask2.assign(std::max<size_t>((size_t)S.np2 * RP[2], 1), 0);
    S.pmask4.assign(std::max<size_t>((size_t)S.n4 * RP[4], 1), 0);
    if (S.np2 == 0) return;
    // A k-symbol map (k >= 1) is R^(k-1) words; bit y of word (t0 .. t_{k-2}) is the value at (t0 .. t_{k-2}, y).
    typedef std::vector<uint32_t> Words;
    const uint32_t ALL = R >= 32u ? 0xFFFFFFFFu : ((1u << R) - 1u);
    auto out = [&](uint32_t n) { return S.out_idx[n] != FAC_NONE; };
    auto has = [&](uint32_t n, uint32_t y) { return y < S.n_syms && ((S.bm[n] >> y) & 1u) != 0; };
    auto kid = [&](uint32_t n, uint32_t y) { return S.fc[n] + (uint32_t)__builtin_popcountll(S.bm[n] & ((1ull << y) - 1ull)); };
    auto for_children = [&](uint32_t h, const std::function<void(uint32_t, uint32_t)> &fn) {  // fn(sym, child)
        uint64_t bmv = S.bm[h];
        uint32_t k = 0;
        while (bmv) { const uint32_t sy = (uint32_t)__builtin_ctzll(bmv); bmv &= bmv - 1; fn(sy, S.fc[h] + k++); }
    };
    auto or_into = [](uint32_t *dst, const uint32_t *src, size_t n) { for (size_t i = 0; i < n; i++) dst[i] |= src[i]; };
    std::unordered_map<uint64_t, Words> memo_w, memo_l;   // key = node << 3 | k, kept for k <= 3
    std::function<Words(uint32_t, int)> Wf, Lf;
    auto W_or = [&](uint32_t d, int k, uint32_t *dst) {   // dst |= W(d; k symbols), k >= 1
        if (k <= 3) {
            const uint64_t key = ((uint64_t)d << 3) | (uint64_t)k;
            auto it = memo_w.find(key);
            if (it == memo_w.end()) { Words v = Wf(d, k); it = memo_w.emplace(key, std::move(v)).first; }
            or_into(dst, it->second.data(), RP[k - 1]);
        } else { const Words v = Wf(d, k); or_into(dst, v.data(), RP[k - 1]); }
    };
    Wf = [&](uint32_t d, int k) -> Words {
        Words v(RP[k - 1], out(d) ? ALL : 0u);
        if (out(d)) return v;
        if (k == 1) { v[0] = (uint32_t)S.bm[d]; return v; }
        for_children(d, [&](uint32_t y1, uint32_t g) { W_or(g, k - 1, &v[(size_t)y1 * RP[k - 2]]); });
        return v;
    };
    auto L_or = [&](uint32_t c, int k, uint32_t *dst) {
        if (k <= 3) {
            const uint64_t key = ((uint64_t)c << 3) | (uint64_t)k;
            auto it = memo_l.find(key);
            if (it == memo_l.end()) { Words v = Lf(c, k); it = memo_l.emplace(key, std::move(v)).first; }
            or_into(dst, it->second.data(), RP[k - 1]);
        } else { const Words v = Lf(c, k); or_into(dst, v.data(), RP[k - 1]); }
    };
    Lf = [&](uint32_t c, int k) -> Words {
        // k == 1: a node without an output has children, and then the exact child (or any substitution child) keeps it alive
        Words v(RP[k - 1], (out(c) || k == 1) ? ALL : 0u);
        if (out(c) || k == 1) return v;
        const size_t bs = RP[k - 2];   // words of one t0-block
        Words sub_all(bs, 0), ins(bs, 0);
        for_children(c, [&](uint32_t, uint32_t d) { W_or(d, k, v.data()); W_or(d, k - 1, sub_all.data()); });   // deletion children
        W_or(c, k - 1, ins.data());                                                                            // insertion child
        for (uint32_t a = 0; a < R; a++) {
            uint32_t *blk = &v[(size_t)a * bs];
            if (has(c, a)) {
                L_or(kid(c, a), k - 1, blk);                                                                   // exact child
                for_children(c, [&](uint32_t s, uint32_t d) { if (s != a) W_or(d, k - 1, blk); });             // substitution children
            } else or_into(blk, sub_all.data(), bs);
            or_into(blk, ins.data(), bs);
        }
        for_children(c, [&](uint32_t b, uint32_t x) {                                                          // swap child
            for_children(x, [&](uint32_t a, uint32_t n2) {
                if (k == 2) v[a] |= 1u << b;
                else W_or(n2, k - 2, &v[((size_t)a * R + b) * RP[k - 3]]);
            });
        });
        return v;
    };
    auto scatter = [&](const uint32_t *w, size_t nw, uint32_t s, uint32_t *dst) {   // dst[word * R + bit] |= 1 << s
        for (size_t i = 0; i < nw; i++) {
            uint32_t m = w[i] & ALL;
            while (m) { const uint32_t y = (uint32_t)__builtin_ctz(m); m &= m - 1; dst[i * R + y] |= 1u << s; }
        }
    };
    auto fold = [&](uint32_t p, int k, uint32_t *dst) {   // dst[R^k] bit s = L(child(p, s); k symbols)
        for_children(p, [&](uint32_t s, uint32_t c) {
            Words tmp(RP[k - 1], 0);
            L_or(c, k, tmp.data());
            scatter(tmp.data(), RP[k - 1], s, dst);
        });
    };
    for (uint32_t p = 0; p < S.n4; p++) fold(p, 4, &S.pmask4[(size_t)p * RP[4]]);
    for (uint32_t p = 0; p < S.n3; p++) {
        fold(p, 3, &S.pmask3[(size_t)p * RP[3]]);
        uint32_t *gm = &S.gmask3[(size_t)p * RP[3]];
        for_children(p, [&](uint32_t s, uint32_t d) {
            for_children(d, [&](uint32_t y1, uint32_t g) {
                Words w(RP[1], 0);
                W_or(g, 2, w.data());
                scatter(w.data(), RP[1], s, gm + (size_t)y1 * RP[2]);
            });
        });
    }
    for (uint32_t p = S.n3; p < S.np2; p++) fold(p, 2, &S.pmask2[(size_t)p * RP[2]]);
}

// Productivity tables of the general stack-machine kernel (fac_flat.h: FlatView::pm_root, px_row / px_bits) for engines with
// mappings (text compared by grapheme id; id 0 = unknown grapheme / end of text) whose first-level states are on their last
// edit or later (edit budgets 2..6; the tables are only consulted for states whose children are on their last edit).
//
//   W(d; y1..yk)   an exhausted state (exact transitions only, search.rs:810, 937, 1003, 1043) at node d reading y1..yk can
//                  emit: out(d) | (k == 0 ? true : edge(d, y1) & W(child; y2..yk))
//   L(c; a, b, c3) a state on its LAST edit at c reading a b c3 can emit: out(c) | c has mapping transitions (search.rs:883-923,
//                  counted as "can emit") | its exact child (L one symbol later; true once the symbols run out) | a substitution
//                  child W(d; b, c3) | a deletion child W(d; a, b, c3) | the insertion child W(c; b, c3) | the swap child
//                  (c -b-> x -a-> n2, W(n2; c3))
//   pm_root[(a, b[, c3])]  bit e:  L(child of root edge e; a, b[, c3])      (3 symbols when the table fits 64 MB, else 2)
//   px_bits[row(x)][b][c3] bit  :  L(x; b, c3) for the nodes x two levels below the root -- the exact child of a last-edit
//                                  state is only pushed when its bit is set
// Penalties, ceilings, similarity and the reference's own dead-end filter only remove states, so leaving them out keeps the
// tables conservative: a state they rule out can emit nothing, now or later.  The true cells are enumerated from the trie as
// patterns with wildcards (a fixed prefix of the symbols a walk has consumed), never by evaluating L per cell.
static void build_flat_pm(HostAutomaton &A, size_t G) {
    A.flat_pm.clear(); A.flat_pm_g = 0; A.flat_pm_words = 0; A.flat_pm_k = 0;
    A.flat_px_row.clear(); A.flat_px.clear();
    const uint32_t N = A.n_nodes();
    const uint32_t e0 = A.node_edge_off[0], deg0 = A.node_edge_off[1] - e0;
    const size_t words = (deg0 + 63) / 64;
    const char *ev = getenv("FAC_FLAT_ROOT_PM");
    int mode = ev && *ev ? atoi(ev) : 3;   // 0 = off, 2 = two-symbol root table only, anything else = three symbols + second-level table
    if (mode != 0 && mode != 2) mode = 3;
    if (!(A.flat_ok && A.has_mappings && A.mef >= 2 && A.mef <= 6 && deg0 >= 2 && deg0 <= 1024 && G * G * words <= ((size_t)8 << 20) && mode != 0)) return;
    const int K = (mode >= 3 && G * G * G * words <= ((size_t)8 << 20)) ? 3 : 2;
    const uint32_t WILD = 0xFFFFFFFFu;
    struct Sym3 { uint32_t v[3]; };
    struct PSet {
        size_t G; bool all = false;
        std::vector<uint8_t> A1, B1, C1, AB, BC, AC;
        std::vector<Sym3> cells;
        explicit PSet(size_t g) : G(g), A1(g), B1(g), C1(g), AB(g * g), BC(g * g), AC(g * g) {}
        void clear() { all = false; for (auto *v : {&A1, &B1, &C1, &AB, &BC, &AC}) std::fill(v->begin(), v->end(), 0); cells.clear(); }
        void emit(const Sym3 &f) {
            const bool a = f.v[0] != WILD, b = f.v[1] != WILD, c = f.v[2] != WILD;
            if (!a && !b && !c) all = true;
            else if (a && !b && !c) A1[f.v[0]] = 1;
            else if (!a && b && !c) B1[f.v[1]] = 1;
            else if (!a && !b && c) C1[f.v[2]] = 1;
            else if (a && b && !c) AB[f.v[0] * G + f.v[1]] = 1;
            else if (!a && b && c) BC[f.v[1] * G + f.v[2]] = 1;
            else if (a && !b && c) AC[f.v[0] * G + f.v[2]] = 1;
            else cells.push_back(f);
        }
        bool base(size_t a, size_t b) const { return all || A1[a] || B1[b] || AB[a * G + b]; }
        bool at(size_t a, size_t b, size_t c) const { return base(a, b) || C1[c] || BC[b * G + c] || AC[a * G + c]; }   // + cells
    };
    auto n_out = [&](uint32_t n) { return A.node_out_off[n + 1] - A.node_out_off[n]; };
    auto n_maps = [&](uint32_t n) { return A.node_map_off[n + 1] - A.node_map_off[n]; };
    auto for_edges = [&](uint32_t n, const std::function<void(uint32_t, uint32_t)> &fn) {   // fn(grapheme id, child); ids beyond the table never occur
        for (uint32_t e = A.node_edge_off[n]; e < A.node_edge_off[n + 1]; e++)
            if (A.edge_sym[e] < G) fn(A.edge_sym[e], A.edge_next[e] & 0x7FFFFFFFu);
    };
    // W(d; symbols idx .. last-1): emits one pattern per way the walk can stay alive
    std::function<void(PSet &, uint32_t, int, int, Sym3)> walk = [&](PSet &P, uint32_t d, int idx, int last, Sym3 fix) {
        if (n_out(d) != 0 || idx >= last) { P.emit(fix); return; }
        for_edges(d, [&](uint32_t k, uint32_t g) { Sym3 f = fix; f.v[idx] = k; walk(P, g, idx + 1, last, f); });
    };
    // L(x) for a state reading symbols idx, idx+1 (idx + 1 may be beyond `last`): the caller has fixed the earlier symbols
    auto last_state = [&](PSet &P, uint32_t x, int idx, int last, Sym3 fix) {
        if (n_out(x) != 0 || n_maps(x) != 0 || idx >= last) { P.emit(fix); return; }
        for_edges(x, [&](uint32_t k, uint32_t y) {
            Sym3 f = fix; f.v[idx] = k;
            P.emit(f);                                                  // exact child: alive (its own L is not expanded further)
            walk(P, y, idx + 1, last, fix);                             // substitution child y at the next position
            walk(P, y, idx, last, fix);                                 // deletion child y at this position
        });
        walk(P, x, idx + 1, last, fix);                                 // insertion child (x has no output here)
        for_edges(x, [&](uint32_t k1, uint32_t y) {                     // swap child: x -sym[idx+1]-> y -sym[idx]-> n2, then idx + 2
            for_edges(y, [&](uint32_t k0, uint32_t n2) {
                Sym3 f = fix; f.v[idx] = k0;
                if (idx + 1 < last) { f.v[idx + 1] = k1; walk(P, n2, idx + 2, last, f); }
                else P.emit(f);
            });
        });
    };
    A.flat_pm_g = (uint32_t)G; A.flat_pm_words = (uint32_t)words; A.flat_pm_k = (uint32_t)K;
    const size_t rows = K == 3 ? G * G * G : G * G;
    A.flat_pm.assign(rows * words, 0);
    PSet P(G);
    const Sym3 none{{WILD, WILD, WILD}};
    // the patterns of all root children, transposed into masks over the root edges (one bit per child), so that the table
    // is filled with a handful of word ORs per cell instead of one test per cell and child
    std::vector<uint64_t> m_all(words, 0), m_a(G * words, 0), m_b(G * words, 0), m_c(G * words, 0), m_ab(G * G * words, 0), m_bc(G * G * words, 0),
        m_ac(G * G * words, 0);
    for (uint32_t e = 0; e < deg0; e++) {
        const uint32_t c = A.edge_next[e0 + e] & 0x7FFFFFFFu;
        P.clear();
        // L(c; a, b, c3): like last_state, but the exact child x is expanded one level (L(x; b, c3))
        if (n_out(c) != 0 || n_maps(c) != 0) P.all = true;
        else {
            for_edges(c, [&](uint32_t ka, uint32_t x) {
                Sym3 f = none; f.v[0] = ka;
                last_state(P, x, 1, K, f);                             // exact child
                walk(P, x, 1, K, none);                                // substitution child (any child; the exact one is covered above)
                walk(P, x, 0, K, none);                                // deletion child
            });
            walk(P, c, 1, K, none);                                    // insertion child
            for_edges(c, [&](uint32_t kb, uint32_t x) {                // swap child
                for_edges(x, [&](uint32_t ka, uint32_t n2) { Sym3 f = none; f.v[0] = ka; f.v[1] = kb; walk(P, n2, 2, K, f); });
            });
        }
        const uint64_t bit = 1ull << (e & 63);
        const size_t w = e >> 6;
        if (P.all) { m_all[w] |= bit; continue; }
        for (size_t k = 0; k < G; k++) {
            if (P.A1[k]) m_a[k * words + w] |= bit;
            if (P.B1[k]) m_b[k * words + w] |= bit;
            if (P.C1[k]) m_c[k * words + w] |= bit;
        }
        for (size_t k = 0; k < G * G; k++) {
            if (P.AB[k]) m_ab[k * words + w] |= bit;
            if (P.BC[k]) m_bc[k * words + w] |= bit;
            if (P.AC[k]) m_ac[k * words + w] |= bit;
        }
        for (const Sym3 &f : P.cells) A.flat_pm[(((size_t)f.v[0] * G + f.v[1]) * G + f.v[2]) * words + w] |= bit;   // only with K == 3
    }
    for (size_t a = 0; a < G; a++)
        for (size_t b = 0; b < G; b++)
            for (size_t w = 0; w < words; w++) {
                const uint64_t base = m_all[w] | m_a[a * words + w] | m_b[b * words + w] | m_ab[(a * G + b) * words + w];
                if (K == 2) { A.flat_pm[(a * G + b) * words + w] |= base; continue; }
                for (size_t c3 = 0; c3 < G; c3++)
                    A.flat_pm[((a * G + b) * G + c3) * words + w] |= base | m_c[c3 * words + w] | m_bc[(b * G + c3) * words + w] | m_ac[(a * G + c3) * words + w];
            }
    // exact-child filter: L(x; b, c3) for the nodes two levels below the root
    if (mode >= 3) {
        std::vector<uint32_t> xs;
        for (uint32_t e = 0; e < deg0; e++) {
            const uint32_t c = A.edge_next[e0 + e] & 0x7FFFFFFFu;
            for (uint32_t e2 = A.node_edge_off[c]; e2 < A.node_edge_off[c + 1]; e2++) xs.push_back(A.edge_next[e2] & 0x7FFFFFFFu);
        }
        const size_t row_words = (G * G + 63) / 64;
        if (!xs.empty() && xs.size() * row_words <= ((size_t)4 << 20)) {
            A.flat_px_row.assign(N, FAC_NONE);
            A.flat_px.assign(xs.size() * row_words, 0);
            for (size_t r = 0; r < xs.size(); r++) {
                const uint32_t x = xs[r];
                A.flat_px_row[x] = (uint32_t)r;
                P.clear();
                last_state(P, x, 1, 3, none);   // reads b (index 1) and c3 (index 2)
                uint64_t *row = &A.flat_px[r * row_words];
                for (size_t b = 0; b < G; b++)
                    for (size_t c3 = 0; c3 < G; c3++)
                        if (P.all || P.B1[b] || P.C1[c3] || P.BC[b * G + c3]) row[(b * G + c3) >> 6] |= 1ull << ((b * G + c3) & 63);
            }
        }
    }
}

fac_status build_automaton(const fac_config *cfg, const fac_pattern *pats, size_t np, HostAutomaton &A, std::string &err) {
    if (!cfg || (np && !pats)) { err = "null config or patterns"; return FAC_INVALID_ARGUMENT; }
    A.ci = cfg->case_insensitive != 0;
    {   // FuzzyPenalties::default(), src/structs.rs:381-393 -- f32 products
        const float m = 1.3f;
        volatile float s = 1.1f * m, i = 0.4f * m, d = 0.7f * m, w = 0.4f * m;
        A.pen_sub = s; A.pen_ins = i; A.pen_del = d; A.pen_swap = w;
    }
    if (cfg->has_penalties) {
        A.pen_sub = cfg->penalty_substitution; A.pen_ins = cfg->penalty_insertion;
        A.pen_del = cfg->penalty_deletion; A.pen_swap = cfg->penalty_swap;
    }
    A.min_sym = cfg->min_symbol_similarity;
    A.beam_width = cfg->beam_width;
    A.has_auto_beam = cfg->has_auto_beam != 0;
    A.ab_budget = cfg->auto_beam_budget; A.ab_width = cfg->auto_beam_width;

    // ---- patterns ----
    A.patterns.clear();
    for (size_t i = 0; i < np; i++) {
        HostPattern p;
        if (pats[i].len && !pats[i].text) { err = "pattern with null text"; return FAC_INVALID_ARGUMENT; }
        p.text.assign(pats[i].text ? pats[i].text : "", pats[i].len);
        if (utf8_valid_up_to((const uint8_t *)p.text.data(), p.text.size()) != p.text.size()) {
            err = "pattern " + std::to_string(i) + " is not valid UTF-8";
            return FAC_INVALID_UTF8;
        }
        p.glen = (uint32_t)count_graphemes(p.text);
        p.byte_len = (uint32_t)p.text.size();
        p.weight = pats[i].weight;
        p.has_limits = pats[i].has_limits != 0;
        if (p.has_limits) p.limits = limits_from_c(pats[i].limits, true);  // Pattern::fuzzy, src/structs.rs:647-650
        p.unique_id = pats[i].unique_id;
        A.patterns.push_back(std::move(p));
    }

    // ---- trie (src/builder.rs:195-237) ----
    std::vector<TmpNode> nodes(1);
    for (size_t i = 0; i < A.patterns.size(); i++) {
        const std::vector<std::string> word = fold_graphemes(A.patterns[i].text, A.ci);
        uint32_t cur = 0;
        for (const std::string &g : word) {
            uint32_t next;
            auto it = nodes[cur].trans.find(g);
            if (it != nodes[cur].trans.end()) next = it->second;
            else {
                next = (uint32_t)nodes.size();
                nodes[cur].trans.emplace(g, next);
                nodes[cur].order.emplace_back(g, next);
                nodes.emplace_back();
            }
            if (nodes[next].pattern_index < 0) nodes[next].pattern_index = (int64_t)i;
            cur = next;
        }
        nodes[cur].output.push_back((uint32_t)i);
    }
    if (nodes.size() >= 0x7FFFFFFFull) { err = "too many trie nodes"; return FAC_UNSUPPORTED; }

    // ---- fail links + output merge (src/builder.rs:240-276) ----
    {
        std::deque<uint32_t> q;
        for (auto &kv : nodes[0].order) { nodes[kv.second].fail = 0; q.push_back(kv.second); }
        while (!q.empty()) {
            const uint32_t cur = q.front();
            q.pop_front();
            for (auto &kv : nodes[cur].order) {
                const uint32_t next = kv.second;
                uint32_t f = nodes[cur].fail;
                while (f != 0 && !nodes[f].trans.count(kv.first)) f = nodes[f].fail;
                uint32_t fallback = 0;
                auto it = nodes[f].trans.find(kv.first);
                if (it != nodes[f].trans.end()) fallback = it->second;
                nodes[next].fail = fallback;
                if (fallback != next) {
                    const std::vector<uint32_t> fo = nodes[fallback].output;
                    for (uint32_t e : fo)
                        if (std::find(nodes[next].output.begin(), nodes[next].output.end(), e) == nodes[next].output.end())
                            nodes[next].output.push_back(e);
                }
                q.push_back(next);
            }
        }
    }

    // ---- effective limits (src/builder.rs:289-329) ----
    A.lim.clear();
    FacLimits global;
    memset(&global, 0, sizeof(global));
    global.ins = global.del = global.sub = global.swp = global.edits = -1;
    A.has_global_limits = false;
    if (cfg->has_limits) { A.has_global_limits = true; global = limits_from_c(cfg->limits, true); }
    else {
        bool any = false;
        for (auto &p : A.patterns)
            if (p.has_limits) {
                any = true;
                auto up = [](int16_t &acc, int16_t v) { if (v >= 0) acc = std::max<int16_t>(std::max<int16_t>(acc, 0), v); };
                up(global.edits, p.limits.edits); up(global.ins, p.limits.ins); up(global.del, p.limits.del);
                up(global.sub, p.limits.sub); up(global.swp, p.limits.swp);
            }
        A.has_global_limits = any;
    }
    A.lim.push_back(global);
    A.has_pattern_limits = false;
    A.pat_lim.assign(A.patterns.size(), FAC_NONE);
    for (size_t i = 0; i < A.patterns.size(); i++)
        if (A.patterns[i].has_limits) {
            A.has_pattern_limits = true;
            A.pat_lim[i] = (uint32_t)A.lim.size();
            A.lim.push_back(A.patterns[i].limits);
        }
    // max_edits_fast (src/builder.rs:451-468) and its dispatch (src/search.rs:205-247)
    if (A.has_pattern_limits) A.max_edits_fast_raw = 255;
    else if (!A.has_global_limits) A.max_edits_fast_raw = 0;
    else if (global.edits >= 0 && global.ins < 0 && global.del < 0 && global.sub < 0 && global.swp < 0) A.max_edits_fast_raw = global.edits;
    else A.max_edits_fast_raw = 255;
    A.mef = (A.max_edits_fast_raw >= 1 && A.max_edits_fast_raw <= 6) ? A.max_edits_fast_raw : 255;

    // ---- grapheme ids + mappings (src/builder.rs:390-442) ----
    struct Directed { std::vector<std::string> pat, hay; float pen; };
    std::vector<Directed> directed;
    for (size_t k = 0; k < cfg->n_mappings; k++) {
        const fac_mapping &mp = cfg->mappings[k];
        const std::string a(mp.a ? mp.a : "", mp.a_len), b(mp.b ? mp.b : "", mp.b_len);
        if (utf8_valid_up_to((const uint8_t *)a.data(), a.size()) != a.size() ||
            utf8_valid_up_to((const uint8_t *)b.data(), b.size()) != b.size()) { err = "mapping is not valid UTF-8"; return FAC_INVALID_UTF8; }
        const std::vector<std::string> ga = fold_graphemes(a, A.ci), gb = fold_graphemes(b, A.ci);
        if (ga.empty() || gb.empty() || ga == gb) continue;
        const float pen = A.pen_sub * (1.0f - mp.score);
        directed.push_back({ga, gb, pen});
        directed.push_back({gb, ga, pen});
    }
    const size_t N = nodes.size();
    std::vector<std::vector<std::pair<uint32_t, uint32_t>>> node_maps(N);  // (directed idx, reached node)
    A.has_mappings = false;
    A.max_map_hay = 1;
    {
        size_t mmh = 0;
        for (size_t start = 0; start < N && !directed.empty(); start++)
            for (size_t d = 0; d < directed.size(); d++) {
                uint32_t cur = (uint32_t)start;
                bool ok = true;
                for (const std::string &g : directed[d].pat) {
                    auto it = nodes[cur].trans.find(g);
                    if (it == nodes[cur].trans.end()) { ok = false; break; }
                    cur = it->second;
                }
                if (ok) { node_maps[start].emplace_back((uint32_t)d, cur); A.has_mappings = true; mmh = std::max(mmh, directed[d].hay.size()); }
            }
        A.max_map_hay = (uint32_t)std::max<size_t>(mmh, 1);
    }
    std::unordered_map<std::string, uint32_t> gid_of;
    std::vector<std::string> gid_syms;
    auto intern = [&](const std::string &g) -> uint32_t {
        auto it = gid_of.find(g);
        if (it != gid_of.end()) return it->second;
        gid_syms.push_back(g);
        gid_of.emplace(g, (uint32_t)gid_syms.size());
        return (uint32_t)gid_syms.size();
    };
    if (A.has_mappings) {
        for (auto &nd : nodes) for (auto &kv : nd.order) intern(kv.first);
        for (auto &d : directed) for (auto &g : d.hay) intern(g);
    }

    // ---- flatten nodes / edges ----
    A.node_edge_off.assign(N + 1, 0); A.node_out_off.assign(N + 1, 0); A.node_map_off.assign(N + 1, 0);
    A.node_bitmap.assign(N * 4, 0); A.node_lim.assign(N, FAC_NONE);
    A.edge_char.clear(); A.edge_next.clear(); A.edge_sym.clear(); A.out_pat.clear();
    A.map_hay_off.assign(1, 0); A.map_hay_gid.clear(); A.map_next.clear(); A.map_pen.clear();
    size_t n_trans = 0;
    for (size_t i = 0; i < N; i++) {
        A.node_edge_off[i] = (uint32_t)A.edge_char.size();
        A.node_out_off[i] = (uint32_t)A.out_pat.size();
        A.node_map_off[i] = (uint32_t)A.map_next.size();
        for (auto &kv : nodes[i].order) {  // Edge::new(first char, next, g.len()==1), src/builder.rs:336-342
            const uint32_t fc = first_scalar(kv.first);
            A.edge_char.push_back(fc);
            A.edge_next.push_back(kv.second | (nodes[kv.second].output.empty() ? 0u : 0x80000000u));
            A.edge_sym.push_back(A.has_mappings ? gid_of[kv.first] : fc);
            if (kv.first.size() == 1 && fc < 128) A.node_bitmap[i * 4 + (fc >> 5)] |= 1u << (fc & 31);
            n_trans++;
        }
        for (uint32_t p : nodes[i].output) A.out_pat.push_back(p);
        for (auto &mp : node_maps[i]) {
            const Directed &d = directed[mp.first];
            for (auto &g : d.hay) A.map_hay_gid.push_back(gid_of[g]);
            A.map_hay_off.push_back((uint32_t)A.map_hay_gid.size());
            A.map_next.push_back(mp.second);
            A.map_pen.push_back(d.pen);
        }
        if (nodes[i].pattern_index >= 0) A.node_lim[i] = A.pat_lim[(size_t)nodes[i].pattern_index];  // get_node_limits, src/search.rs:67-71
    }
    A.node_edge_off[N] = (uint32_t)A.edge_char.size();
    A.node_out_off[N] = (uint32_t)A.out_pat.size();
    A.node_map_off[N] = (uint32_t)A.map_next.size();

    // ---- exact-transition table ----
    {
        size_t cap = 64;
        while (cap < n_trans * 2 + 2) cap <<= 1;
        A.trans.assign(cap, FacTrans{FAC_NONE, 0, 0, 0});
        const uint32_t mask = (uint32_t)(cap - 1);
        for (size_t i = 0; i < N; i++)
            for (auto &kv : nodes[i].order) {
                const uint32_t sym = A.has_mappings ? gid_of[kv.first] : first_scalar(kv.first);
                uint32_t h = fac_hash2((uint32_t)i, sym) & mask;
                bool dup = false;
                while (A.trans[h].node != FAC_NONE) {
                    if (A.trans[h].node == (uint32_t)i && A.trans[h].sym == sym) { dup = true; break; }  // first edge in build order wins
                    h = (h + 1) & mask;
                }
                if (!dup) A.trans[h] = FacTrans{(uint32_t)i, sym, kv.second, 0};
            }
    }

    // ---- prune coefficients (src/builder.rs:348-381) ----
    {
        std::vector<size_t> rl(N, 0);
        std::vector<float> rw(N, 0.f);
        for (size_t i = 0; i < N; i++)
            for (uint32_t p : nodes[i].output) { rl[i] = std::max<size_t>(rl[i], A.patterns[p].glen); rw[i] = std::max(rw[i], A.patterns[p].weight); }
        for (size_t i = N; i-- > 0;)  // children have higher indices than parents: one reverse pass
            for (auto &kv : nodes[i].order) { rl[i] = std::max(rl[i], rl[kv.second]); rw[i] = std::max(rw[i], rw[kv.second]); }
        A.node_prune_len.resize(N); A.node_prune_low.resize(N);
        for (size_t i = 0; i < N; i++) {
            const float len = (float)rl[i];
            A.node_prune_len[i] = len;
            A.node_prune_low[i] = len / rw[i];
        }
    }

    // ---- patterns / similarity ----
    A.pat_glen.resize(A.patterns.size()); A.pat_weight.resize(A.patterns.size());
    A.pat_bytes.resize(A.patterns.size()); A.pat_uid.resize(A.patterns.size());
    for (size_t i = 0; i < A.patterns.size(); i++) {
        A.pat_glen[i] = (float)A.patterns[i].glen;
        A.pat_weight[i] = A.patterns[i].weight;
        A.pat_bytes[i] = A.patterns[i].byte_len;
        // UniqueId derives Ord with Automatic < Custom (src/structs.rs:586-592)
        A.pat_uid[i] = A.patterns[i].unique_id >= 0 ? ((int64_t)1 << 62) | A.patterns[i].unique_id : (int64_t)i;
    }
    std::map<std::pair<uint32_t, uint32_t>, float> simmap;
    if (cfg->has_similarity) for (size_t i = 0; i < cfg->n_similarity; i++) simmap[{cfg->similarity[i].a, cfg->similarity[i].b}] = cfg->similarity[i].similarity;
    else build_default_similarity(simmap);
    A.sim_ascii.assign(128 * 128, 0.f);
    for (int i = 0; i < 128; i++) A.sim_ascii[i * 128 + i] = 1.f;
    A.sim_keys.clear(); A.sim_vals.clear();
    float max_off_diag = 0.f;  // Similarity::max_off_diagonal, src/structs.rs:61-76
    for (auto &kv : simmap) {
        const uint32_t a = kv.first.first, b = kv.first.second;
        if (a < 128 && b < 128) A.sim_ascii[a * 128 + b] = kv.second;
        else { A.sim_keys.push_back(((uint64_t)a << 32) | b); A.sim_vals.push_back(kv.second); }  // std::map order == sorted keys
        if (a != b && !(a < 128 && b < 128)) max_off_diag = std::max(max_off_diag, kv.second);
    }
    for (int i = 0; i < 128; i++) for (int j = 0; j < 128; j++) if (i != j) max_off_diag = std::max(max_off_diag, A.sim_ascii[i * 128 + j]);

    // ---- grapheme-id tables ----
    A.ascii_gid.assign(128, 0);
    if (A.has_mappings) {
        build_symbol_table(gid_syms, A.symbols, A.symbol_pool);
        for (int b = 0; b < 128; b++) {
            auto it = gid_of.find(std::string(1, (char)b));
            if (it != gid_of.end()) A.ascii_gid[b] = it->second;
        }
    } else { A.symbols.assign(1, HostSymbol{0, 0, 0, 0, 0}); A.symbol_pool.assign(8, 0); }

    // ---- window skip bitmaps (src/search.rs:504-521) ----
    A.wskip = false;
    if (A.mef == 1 && !A.has_mappings && nodes[0].output.empty()) {
        bool child_output = false;
        uint32_t first[4], second[4] = {0, 0, 0, 0};
        for (int k = 0; k < 4; k++) first[k] = A.node_bitmap[k];
        for (auto &kv : nodes[0].order) {
            for (int k = 0; k < 4; k++) { second[k] |= A.node_bitmap[kv.second * 4 + k]; first[k] |= A.node_bitmap[kv.second * 4 + k]; }
            if (!nodes[kv.second].output.empty()) child_output = true;
        }
        if (!child_output) { A.wskip = true; for (int k = 0; k < 4; k++) { A.ws_first[k] = first[k]; A.ws_second[k] = second[k]; } }
    }

    // ---- succinct BFS-ordered trie for the fast kernel (fac_succinct.h) ----
    {
        HostSuccinct &S = A.succ;
        S = HostSuccinct();
        S.exact_only = !A.has_global_limits && !A.has_pattern_limits;
        S.limits_mode = A.mef == 255 && !S.exact_only;
        bool ok = !A.has_mappings && N <= SUCC_MAX_NODES;
        if (S.limits_mode) {
            // bound on the total edits of a state: the largest edits(e) budget plus, for limit sets without a total,
            // the per-type caps (a state may arrive at such a node having spent edits of other types elsewhere)
            uint32_t max_e = 0, cap[4] = {0, 0, 0, 0};
            for (size_t li = A.has_global_limits ? 0 : 1; li < A.lim.size(); li++) {
                const FacLimits &L = A.lim[li];
                if (L.edits >= 0) max_e = std::max<uint32_t>(max_e, (uint32_t)L.edits);
                else {
                    const int16_t t[4] = {L.ins, L.del, L.sub, L.swp};
                    for (int k = 0; k < 4; k++) cap[k] = std::max<uint32_t>(cap[k], t[k] < 0 ? 255u : (uint32_t)t[k]);
                }
            }
            S.edit_bound = max_e + cap[0] + cap[1] + cap[2] + cap[3];
            if (S.edit_bound == 0 || S.edit_bound > 6) ok = false;   // deeper budgets stay on the generic kernel
        }
        std::vector<int> sym_of_char(128, -1);
        std::vector<uint32_t> chars;
        for (size_t i = 0; i < N && ok; i++)
            for (auto &kv : nodes[i].order) {
                if (kv.first.size() != 1 || (uint8_t)kv.first[0] >= 128) { ok = false; break; }
                chars.push_back((uint8_t)kv.first[0]);
            }
        if (ok) {
            std::sort(chars.begin(), chars.end());
            chars.erase(std::unique(chars.begin(), chars.end()), chars.end());
            const char *ev = getenv("FAC_SUCC_FORCE_WIDE");
            S.wide = chars.size() > 31 || (ev && *ev == '1');
            if (chars.size() > 63) ok = false;
        }
        if (ok) {
            const uint32_t NOSYM = S.wide ? 63u : 31u, ROW = NOSYM + 1u;
            memset(S.sym_of, (int)NOSYM, sizeof(S.sym_of));
            S.n_syms = (uint32_t)chars.size();
            for (size_t k = 0; k < chars.size(); k++) { sym_of_char[chars[k]] = (int)k; S.sym_of[chars[k]] = (uint8_t)k; }
            // BFS numbering: children of a node contiguous, sorted by symbol
            S.old_of.assign(1, 0);
            S.bm.assign(N, 0); S.fc.assign(N, 0); S.insym.assign(N, 0);
            for (size_t h = 0; h < S.old_of.size(); h++) {
                const uint32_t old = S.old_of[h];
                std::vector<std::pair<uint32_t, uint32_t>> ch;  // (sym, old child)
                for (auto &kv : nodes[old].order) ch.emplace_back((uint32_t)sym_of_char[(uint8_t)kv.first[0]], kv.second);
                std::sort(ch.begin(), ch.end());
                S.fc[h] = (uint32_t)S.old_of.size();
                for (auto &c : ch) {
                    S.bm[h] |= 1ull << c.first;
                    S.insym[S.old_of.size()] = (uint8_t)c.first;
                    S.old_of.push_back(c.second);
                }
            }
            S.prune_len.resize(N); S.prune_low.resize(N); S.out_idx.assign(N, FAC_NONE); S.node_lim.assign(N, FAC_NONE);
            for (size_t h = 0; h < N; h++) {
                const uint32_t old = S.old_of[h];
                S.prune_len[h] = A.node_prune_len[old]; S.prune_low[h] = A.node_prune_low[old];
                S.node_lim[h] = A.node_lim[old];
                const auto &o = nodes[old].output;
                if (!o.empty()) {
                    S.out_idx[h] = (uint32_t)(S.out2.size() / 4);
                    for (size_t k = 0; k < o.size(); k++) {
                        union { float f; uint32_t u; } g, w;
                        g.f = A.pat_glen[o[k]]; w.f = A.pat_weight[o[k]];
                        S.out2.push_back(o[k] | (k + 1 == o.size() ? 0x80000000u : 0u));
                        S.out2.push_back(g.u); S.out2.push_back(w.u); S.out2.push_back(A.pat_lim[o[k]]);
                    }
                }
            }
            if (S.out2.empty()) S.out2.assign(4, 0x80000000u);
            const float inf = std::numeric_limits<float>::infinity();
            S.sub_pen.assign((size_t)ROW * SUCC_SP_STRIDE, inf);
            for (size_t k = 0; k < chars.size(); k++)
                for (uint32_t b = 0; b < 128; b++) {
                    const float sm = chars[k] == b ? 1.0f : A.sim_ascii[chars[k] * 128 + b];  // get_similarity, search.rs:76-82
                    if (sm < A.min_sym) continue;                                              // search.rs:822-825
                    volatile float one_minus = 1.0f - sm;
                    volatile float pp = A.pen_sub * one_minus;                                 // search.rs:829
                    S.sub_pen[k * SUCC_SP_STRIDE + b] = pp;
                }
            // non-ASCII text first chars: similarity 0 with every (ASCII) pattern char unless the similarity map says otherwise
            S.unicode_text_ok = A.sim_keys.empty();
            if (!(0.0f < A.min_sym)) {
                volatile float pp0 = A.pen_sub * (1.0f - 0.0f);
                for (size_t k = 0; k < chars.size(); k++) S.sub_pen[k * SUCC_SP_STRIDE + SUCC_NONASCII] = pp0;
            }
            S.first_mask = S.bm[0];
            for (uint32_t c = 0; c < (uint32_t)__builtin_popcountll(S.bm[0]); c++) S.second_mask |= S.bm[S.fc[0] + c];
            S.first_mask |= S.second_mask;
            auto for_children = [&](uint32_t h, const std::function<void(uint32_t, uint32_t)> &fn) {  // fn(sym, child)
                uint64_t bmv = S.bm[h];
                uint32_t k = 0;
                while (bmv) { const uint32_t sy = (uint32_t)__builtin_ctzll(bmv); bmv &= bmv - 1; fn(sy, S.fc[h] + k++); }
            };
            {   // grandchild masks: gm[n][y] = symbols s with child(n, s) having edge y; gm[n][NOSYM] = children with an output
                const char *ev = getenv("FAC_GM_NODES");
                const size_t lim = ev && *ev ? (size_t)atoll(ev) : ((size_t)4 << 20);   // every node of any realistic trie
                S.gm_nodes = (uint32_t)std::min<size_t>(N, lim);
                S.gmask.assign((size_t)std::max<uint32_t>(S.gm_nodes, 1) * ROW, 0);
                for (uint32_t h = 0; h < S.gm_nodes; h++)
                    for_children(h, [&](uint32_t sy, uint32_t c) {
                        for_children(c, [&](uint32_t y, uint32_t) { S.gmask[(size_t)h * ROW + y] |= 1ull << sy; });
                        if (S.out_idx[c] != FAC_NONE) S.gmask[(size_t)h * ROW + NOSYM] |= 1ull << sy;
                    });
            }
            {   // two-deep masks for the shallow nodes (where almost all last-edit expansions happen)
                const char *ev = getenv("FAC_GM2_NODES");
                const size_t lim = ev && *ev ? (size_t)atoll(ev) : (S.wide ? 1024 : 2048);
                S.gm2_nodes = (uint32_t)std::min<size_t>(std::min<size_t>(N, lim), S.gm_nodes);
                S.gmask2.assign((size_t)std::max<uint32_t>(S.gm2_nodes, 1) * ROW * ROW, 0);
                for (uint32_t h = 0; h < S.gm2_nodes; h++)
                    for_children(h, [&](uint32_t sy, uint32_t c) {
                        for_children(c, [&](uint32_t y1, uint32_t g) {
                            uint64_t *row = &S.gmask2[((size_t)h * ROW + y1) * ROW];
                            if (S.out_idx[g] != FAC_NONE) { for (uint32_t y2 = 0; y2 < ROW; y2++) row[y2] |= 1ull << sy; }
                            else for_children(g, [&](uint32_t y2, uint32_t) { row[y2] |= 1ull << sy; });
                        });
                    });
            }
            if (!S.wide && !S.exact_only) build_deep_tables(S);
            S.ok = true;
        }
    }

    // ---- merged records for the general stack-machine kernel (fac_flat.h) ----
    {
        A.flat_ok = true;
        A.flat_nrec.assign(N * 4, 0); A.flat_erec.assign(A.edge_char.size() * 4, 0);
        for (size_t i = 0; i < N && A.flat_ok; i++) {
            const uint32_t deg = A.node_edge_off[i + 1] - A.node_edge_off[i], no = A.node_out_off[i + 1] - A.node_out_off[i];
            const uint32_t nm = A.node_map_off[i + 1] - A.node_map_off[i];
            if (deg > 4095u || no > 1023u || nm > 255u) { A.flat_ok = false; break; }
            A.flat_nrec[i * 4 + 0] = A.node_edge_off[i]; A.flat_nrec[i * 4 + 1] = deg | (no << 12) | (nm << 22);
            A.flat_nrec[i * 4 + 3] = A.node_out_off[i];
        }
        A.flat_ooff.assign(N + 1, 0); A.flat_olist.clear();
        for (size_t i = 0; i < N; i++) {
            A.flat_ooff[i] = (uint32_t)A.flat_olist.size();
            for (uint32_t e = A.node_edge_off[i]; e < A.node_edge_off[i + 1]; e++)
                if (A.edge_next[e] >> 31) A.flat_olist.push_back(e - A.node_edge_off[i]);
        }
        A.flat_ooff[N] = (uint32_t)A.flat_olist.size();
        A.flat_gm_row.assign(N, FAC_NONE); A.flat_gm.clear();
        for (size_t i = 0; i < N; i++) {
            const uint32_t e0 = A.node_edge_off[i], deg = A.node_edge_off[i + 1] - e0;
            if (deg < 3 || deg > 64 || A.flat_gm.size() / 128 >= (1u << 20)) continue;
            A.flat_gm_row[i] = (uint32_t)(A.flat_gm.size() / 128);
            const size_t base = A.flat_gm.size();
            A.flat_gm.resize(base + 128, 0);
            for (uint32_t e = 0; e < deg; e++) {
                const uint32_t child = A.edge_next[e0 + e] & 0x7FFFFFFFu;
                const bool has_out = (A.edge_next[e0 + e] >> 31) != 0;
                for (uint32_t c = 0; c < 128; c++)
                    if (has_out || ((A.node_bitmap[child * 4 + (c >> 5)] >> (c & 31)) & 1u)) A.flat_gm[base + c] |= 1ull << e;
            }
        }
        if (A.flat_gm.empty()) A.flat_gm.assign(128, 0);
        build_flat_pm(A, gid_syms.size() + 1);
        if (A.flat_pm.empty()) A.flat_pm.assign(1, 0);
        for (size_t e = 0; e < A.edge_char.size(); e++) {
            A.flat_erec[e * 4 + 0] = A.edge_next[e]; A.flat_erec[e * 4 + 1] = A.edge_char[e]; A.flat_erec[e * 4 + 2] = A.edge_sym[e];
        }
    }

    // ---- max_match_graphemes (src/stream.rs:213-253) ----
    {
        size_t max_pattern = 0, max_edits = 0;
        auto edits_of = [](const FacLimits &l) -> size_t {
            if (l.edits >= 0) return (size_t)l.edits;
            return (size_t)std::max<int>(l.ins, 0) + (size_t)std::max<int>(l.del, 0) + (size_t)std::max<int>(l.sub, 0) + (size_t)std::max<int>(l.swp, 0);
        };
        for (auto &p : A.patterns) {
            max_pattern = std::max<size_t>(max_pattern, p.glen);
            size_t e = 0;
            if (p.has_limits) e = edits_of(p.limits);
            else if (A.has_global_limits) e = edits_of(global);
            max_edits = std::max(max_edits, e);
        }
        A.max_match_graphemes = max_pattern + max_edits * A.max_map_hay;
    }
    if (A.max_match_graphemes + 3 > FAC_MAX_SPAN) {
        err = "max_match_graphemes() = " + std::to_string(A.max_match_graphemes) + " exceeds the device state layout (" + std::to_string(FAC_MAX_SPAN - 3) + ")";
        return FAC_UNSUPPORTED;
    }

    // ---- bitap pre-filter model (BitapFilter::build, src/prefilter.rs:161-245) ----
    HostBitap &B = A.bitap;
    B = HostBitap();
    do {
        if (A.has_mappings || A.patterns.empty()) break;
        const float p_sub_min = A.pen_sub * (1.0f - max_off_diag);
        const float mults[4] = {1.0f / A.pen_ins, 1.0f / A.pen_del, 1.0f / p_sub_min, 2.0f / A.pen_swap};
        bool bad = false;
        float mult = 0.f;
        for (float m : mults) { if (!std::isfinite(m) || m <= 0.0f) bad = true; mult = std::max(mult, m); }
        if (bad) break;
        std::unordered_map<std::string, uint32_t> ids;
        std::vector<std::string> id_syms;
        std::vector<std::vector<uint32_t>> pat_ids;
        bool ok = true;
        for (auto &p : A.patterns) {
            const std::vector<std::string> gs = fold_graphemes(p.text, A.ci);
            if (gs.empty() || gs.size() > 63) { ok = false; break; }
            std::vector<uint32_t> v;
            for (auto &g : gs) {
                auto it = ids.find(g);
                uint32_t id;
                if (it == ids.end()) { id = (uint32_t)ids.size() + 1; ids.emplace(g, id); id_syms.push_back(g); } else id = it->second;
                if (id > 255) { ok = false; break; }
                v.push_back(id);
            }
            if (!ok) break;
            pat_ids.push_back(v);
            B.m.push_back((uint32_t)gs.size());
            B.weight.push_back(p.weight);
            size_t kl = 0;
            bool has = false;
            if (p.has_limits) has = k_from_limits(p.limits, kl);
            else if (A.has_global_limits) has = k_from_limits(global, kl);
            B.k_limit.push_back(has ? (int32_t)std::min<size_t>(kl, 1u << 20) : -1);
        }
        if (!ok) { B = HostBitap(); break; }
        B.edit_cost_mult = mult;
        B.alphabet = (uint32_t)ids.size();
        for (int b = 0; b < 128; b++) {
            std::string f(1, (char)((A.ci && b >= 'A' && b <= 'Z') ? b + 32 : b));
            auto it = ids.find(f);
            if (it != ids.end()) B.ascii_id[b] = (uint8_t)it->second;
        }
        B.masks.assign(pat_ids.size() * (size_t)(B.alphabet + 1), 0);
        for (size_t i = 0; i < pat_ids.size(); i++)
            for (size_t k = 0; k < pat_ids[i].size(); k++) B.masks[i * (B.alphabet + 1) + pat_ids[i][k]] |= 1ull << k;
        build_symbol_table(id_syms, B.symbols, B.symbol_pool);
        B.active = true;
    } while (0);
    if (!B.active) { B.symbols.assign(1, HostSymbol{0, 0, 0, 0, 0}); B.symbol_pool.assign(8, 0); }
    return FAC_OK;
}

AutomatonView HostAutomaton::host_view() const {
    AutomatonView V;
    memset(&V, 0, sizeof(V));
    V.n_nodes = n_nodes(); V.n_edges = (uint32_t)edge_char.size(); V.n_patterns = (uint32_t)patterns.size(); V.n_outputs = (uint32_t)out_pat.size();
    V.node_edge_off = node_edge_off.data(); V.node_prune_len = node_prune_len.data(); V.node_prune_low = node_prune_low.data();
    V.node_out_off = node_out_off.data(); V.node_bitmap = node_bitmap.data(); V.node_lim = node_lim.data(); V.node_map_off = node_map_off.data();
    V.edge_char = edge_char.data(); V.edge_next = edge_next.data(); V.edge_sym = edge_sym.data();
    V.trans = trans.data(); V.trans_mask = (uint32_t)trans.size() - 1;
    V.out_pat = out_pat.data(); V.pat_glen = pat_glen.data(); V.pat_weight = pat_weight.data(); V.pat_lim = pat_lim.data(); V.lim = lim.data();
    V.sim_ascii = sim_ascii.data(); V.sim_keys = sim_keys.data(); V.sim_vals = sim_vals.data(); V.n_sim = (uint32_t)sim_keys.size();
    V.map_hay_off = map_hay_off.data(); V.map_hay_gid = map_hay_gid.data(); V.map_next = map_next.data(); V.map_pen = map_pen.data();
    V.ascii_gid = ascii_gid.data();
    V.pen_sub = pen_sub; V.pen_ins = pen_ins; V.pen_del = pen_del; V.pen_swap = pen_swap; V.min_sym = min_sym;
    V.mef = mef; V.has_mappings = has_mappings; V.has_pattern_limits = has_pattern_limits; V.has_global_limits = has_global_limits; V.ci = ci;
    V.wskip = wskip;
    for (int k = 0; k < 4; k++) { V.ws_first[k] = ws_first[k]; V.ws_second[k] = ws_second[k]; }
    return V;
}

}  // namespace fac
