// =====================================================================================
// fac_oracle.cpp -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference algorithm.
//
// This file is the parity oracle for the CUDA search path.  It restates, operator for
// operator, the algorithm of kakserpom/fuzzy-aho-corasick-rs v0.5.0 (paths below are under
// /root/reference).  It is NOT part of the product: only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load it.  The product library
// (libfacgpu.so) never links or calls anything in oracle/.
//
// The reference is Rust and cannot be compiled in this image (no cargo/rustc), so the
// oracle is pinned by the reference's own known-answer tests (src/tests.rs, doc examples),
// transcribed in tests/test_reference_kat.py and tests/test_oracle_pins.py.  Items that no reference test pins are called
// out below as "parity unpinned" and listed in DESIGN.md:
//   (U1) edge order of a node = first-insertion order here; the reference uses hashbrown
//        iteration order under its FxHasher (src/builder.rs:331-342).
//   (U2) Order::Unsorted output order = ascending (start, end, pattern) here; the
//        reference returns FxHashMap iteration order (src/search.rs:1105-1118).
//   (U3) beam selection keeps the K lowest by (penalty, queue position) and preserves queue
//        order; the reference uses select_nth_unstable_by (src/search.rs:584).
//   (U4) final `sort_unstable_by_key(start)` of non_overlapping (src/matches.rs:111) is
//        stable here (ties only arise with empty spans).
//   (U5) grapheme segmentation / lowercase tables come from the python `regex` Unicode
//        database (tools/gen_unicode_tables.py), standing in for unicode-segmentation 1.13
//        and Rust std `to_lowercase`.
//
// All floating point is IEEE f32, one rounding per operation, no FMA contraction
// (compile with -ffp-contract=off), matching Rust semantics.
// =====================================================================================
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <map>
#include <set>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../include/fac.h"
#include "unicode_tables.h"

namespace orc {

// ------------------------------------------------------------------------------------
// Unicode helpers (stand-ins for unicode-segmentation and std::to_lowercase, see U5)
// ------------------------------------------------------------------------------------
static inline uint32_t decode_utf8(const uint8_t *s, size_t n, size_t &i) {
    uint8_t b = s[i];
    if (b < 0x80) { i += 1; return b; }
    if (b < 0xE0) { uint32_t c = ((b & 0x1F) << 6) | (s[i + 1] & 0x3F); i += 2; return c; }
    if (b < 0xF0) { uint32_t c = ((b & 0x0F) << 12) | ((s[i + 1] & 0x3F) << 6) | (s[i + 2] & 0x3F); i += 3; return c; }
    uint32_t c = ((b & 0x07) << 18) | ((s[i + 1] & 0x3F) << 12) | ((s[i + 2] & 0x3F) << 6) | (s[i + 3] & 0x3F);
    i += 4;
    return c;
}

static inline void encode_utf8(uint32_t c, std::string &out) {
    if (c < 0x80) out.push_back((char)c);
    else if (c < 0x800) { out.push_back((char)(0xC0 | (c >> 6))); out.push_back((char)(0x80 | (c & 0x3F))); }
    else if (c < 0x10000) { out.push_back((char)(0xE0 | (c >> 12))); out.push_back((char)(0x80 | ((c >> 6) & 0x3F))); out.push_back((char)(0x80 | (c & 0x3F))); }
    else { out.push_back((char)(0xF0 | (c >> 18))); out.push_back((char)(0x80 | ((c >> 12) & 0x3F))); out.push_back((char)(0x80 | ((c >> 6) & 0x3F))); out.push_back((char)(0x80 | (c & 0x3F))); }
}

// Strict UTF-8 validation (what Rust's `str::from_utf8` accepts); returns valid_up_to.
static size_t utf8_valid_up_to(const uint8_t *s, size_t n) {
    size_t i = 0;
    while (i < n) {
        uint8_t b = s[i];
        if (b < 0x80) { i++; continue; }
        size_t need; uint32_t minv;
        if (b >= 0xC2 && b <= 0xDF) { need = 1; minv = 0x80; }
        else if (b >= 0xE0 && b <= 0xEF) { need = 2; minv = 0x800; }
        else if (b >= 0xF0 && b <= 0xF4) { need = 3; minv = 0x10000; }
        else return i;
        if (i + need >= n) return i;  // truncated sequence
        uint32_t c = b & (0x3F >> need);
        for (size_t k = 1; k <= need; k++) {
            uint8_t cb = s[i + k];
            if ((cb & 0xC0) != 0x80) return i;
            c = (c << 6) | (cb & 0x3F);
        }
        if (c < minv || c > 0x10FFFF || (c >= 0xD800 && c <= 0xDFFF)) return i;
        i += need + 1;
    }
    return i;
}

static inline uint8_t gcb_byte(uint32_t cp) {
    if (cp >= 0x110000) return 0;
    return FAC_GCB_STAGE2[(size_t)FAC_GCB_STAGE1[cp >> 8] * 256 + (cp & 0xFF)];
}

// UAX #29 extended grapheme cluster boundaries (GB1-GB999 incl. GB9c, GB11, GB12/13);
// stands in for `graphemes(true)` / `grapheme_indices(true)`.
// Returns the byte offset of each cluster start.
static void grapheme_starts(const uint8_t *s, size_t n, std::vector<size_t> &starts) {
    starts.clear();
    size_t i = 0;
    int pc = -1;
    unsigned ri_run = 0;  // consecutive RI immediately before the current scalar
    int ep_state = 0;     // 1: ExtPict Extend*   2: ExtPict Extend* ZWJ
    int incb_state = 0;   // 1: Consonant [Extend]*   2: ... with a Linker seen
    while (i < n) {
        size_t at = i;
        uint32_t cp = decode_utf8(s, n, i);
        uint8_t cb = gcb_byte(cp);
        int c = cb & 0xF;
        bool extpict = (cb & FAC_GCB_EXTPICT) != 0;
        int incb = (cb >> FAC_GCB_INCB_SHIFT) & 3;
        bool brk;
        if (pc < 0) brk = true;
        else if (pc == FAC_GCB_CR && c == FAC_GCB_LF) brk = false;                                    // GB3
        else if (pc == FAC_GCB_CONTROL || pc == FAC_GCB_CR || pc == FAC_GCB_LF) brk = true;           // GB4
        else if (c == FAC_GCB_CONTROL || c == FAC_GCB_CR || c == FAC_GCB_LF) brk = true;              // GB5
        else if (pc == FAC_GCB_L && (c == FAC_GCB_L || c == FAC_GCB_V || c == FAC_GCB_LV || c == FAC_GCB_LVT)) brk = false;  // GB6
        else if ((pc == FAC_GCB_LV || pc == FAC_GCB_V) && (c == FAC_GCB_V || c == FAC_GCB_T)) brk = false;                  // GB7
        else if ((pc == FAC_GCB_LVT || pc == FAC_GCB_T) && c == FAC_GCB_T) brk = false;               // GB8
        else if (c == FAC_GCB_EXTEND || c == FAC_GCB_ZWJ) brk = false;                                // GB9
        else if (c == FAC_GCB_SPACINGMARK) brk = false;                                               // GB9a
        else if (pc == FAC_GCB_PREPEND) brk = false;                                                  // GB9b
        else if (incb_state == 2 && incb == 1) brk = false;                                           // GB9c
        else if (ep_state == 2 && extpict) brk = false;                                               // GB11
        else if (pc == FAC_GCB_RI && c == FAC_GCB_RI && (ri_run & 1)) brk = false;                    // GB12/13
        else brk = true;                                                                              // GB999
        if (brk) starts.push_back(at);
        // state updates
        ri_run = (c == FAC_GCB_RI) ? ri_run + 1 : 0;
        if (extpict) ep_state = 1;
        else if (c == FAC_GCB_EXTEND && ep_state == 1) ep_state = 1;
        else if (c == FAC_GCB_ZWJ && ep_state == 1) ep_state = 2;
        else ep_state = 0;
        if (incb == 1) incb_state = 1;
        else if (incb == 3 && incb_state >= 1) incb_state = 2;
        else if (incb == 2 && incb_state >= 1) { /* keep */ }
        else incb_state = 0;
        pc = c;
    }
}

static inline void lower_cp(uint32_t cp, std::string &out) {
    if (cp < 0x80) { out.push_back((char)((cp >= 'A' && cp <= 'Z') ? cp + 32 : cp)); return; }
    if (cp == 0x130) { encode_utf8(0x69, out); encode_utf8(0x307, out); return; }
    const uint32_t *lo = std::lower_bound(FAC_LOWER_KEYS, FAC_LOWER_KEYS + FAC_LOWER_N, cp);
    if (lo != FAC_LOWER_KEYS + FAC_LOWER_N && *lo == cp) encode_utf8(FAC_LOWER_VALS[lo - FAC_LOWER_KEYS], out);
    else encode_utf8(cp, out);
}

// `str::to_lowercase` applied to ONE grapheme (builder.rs:199, search.rs:409).  The
// Final_Sigma context rule never fires inside a single cluster, so this is a per-scalar map.
static std::string to_lowercase(const uint8_t *s, size_t n) {
    std::string out;
    size_t i = 0;
    while (i < n) { uint32_t cp = decode_utf8(s, n, i); lower_cp(cp, out); }
    return out;
}

static inline uint32_t first_char(const std::string &g) {
    if (g.empty()) return 0;
    size_t i = 0;
    return decode_utf8((const uint8_t *)g.data(), g.size(), i);
}

// `graphemes(true)` of `s`, optionally folded per grapheme.
static std::vector<std::string> fold_graphemes(const std::string &s, bool ci) {
    std::vector<size_t> st;
    grapheme_starts((const uint8_t *)s.data(), s.size(), st);
    std::vector<std::string> out;
    for (size_t k = 0; k < st.size(); k++) {
        size_t b = st[k], e = (k + 1 < st.size()) ? st[k + 1] : s.size();
        if (ci) out.push_back(to_lowercase((const uint8_t *)s.data() + b, e - b));
        else out.push_back(s.substr(b, e - b));
    }
    return out;
}

// ------------------------------------------------------------------------------------
// Data model (src/structs.rs)
// ------------------------------------------------------------------------------------
struct Limits {  // FuzzyLimits, structs.rs:292-299; -1 == None
    int ins = -1, del = -1, sub = -1, swp = -1, edits = -1;
};
static Limits finalize(Limits l) {  // structs.rs:319-335
    if (l.edits < 0) {
        if (l.ins < 0) l.ins = 0;
        if (l.del < 0) l.del = 0;
        if (l.sub < 0) l.sub = 0;
        if (l.swp < 0) l.swp = 0;
    }
    return l;
}
static Limits from_c(const fac_limits &c) {
    Limits l; l.ins = c.insertions; l.del = c.deletions; l.sub = c.substitutions; l.swp = c.swaps; l.edits = c.edits; return l;
}

struct Pattern {  // structs.rs:597-610
    std::string text;
    size_t glen = 0;  // graphemes of the ORIGINAL text (structs.rs:664)
    float weight = 1.f;
    bool has_limits = false;
    Limits limits;
    int64_t uid = -1;
};

struct Edge {  // structs.rs:185-196
    uint32_t first_char;
    uint32_t next;
    bool single_byte;
};
struct MapTrans {  // structs.rs:234-242
    std::vector<std::string> hay;
    uint32_t next;
    float pen;
};
struct Node {  // structs.rs:248-281 (fail / weight are never read by search: omitted weight, kept fail)
    std::vector<Edge> edges;
    std::vector<uint32_t> output;
    float prune_len = 0.f, prune_low = 0.f;
    uint32_t fail = 0;
    int64_t pattern_index = -1;
    std::unordered_map<std::string, uint32_t> trans;
    std::vector<std::pair<std::string, uint32_t>> order;  // insertion order (U1)
};

struct Similarity {  // structs.rs:9-93
    std::map<std::pair<uint32_t, uint32_t>, float> map;
    std::vector<float> ascii;  // 128x128
    Similarity() : ascii(128 * 128, 0.f) {}
    void finish() {
        std::fill(ascii.begin(), ascii.end(), 0.f);
        for (int i = 0; i < 128; i++) ascii[i * 128 + i] = 1.f;
        for (auto &kv : map)
            if (kv.first.first < 128 && kv.first.second < 128) ascii[kv.first.first * 128 + kv.first.second] = kv.second;
    }
    float get(uint32_t a, uint32_t b) const {  // structs.rs:82-92
        if (a < 128 && b < 128) return ascii[a * 128 + b];
        auto it = map.find({a, b});
        return it == map.end() ? 0.f : it->second;
    }
    float max_off_diagonal() const {  // structs.rs:61-76
        float mx = 0.f;
        for (int i = 0; i < 128; i++)
            for (int j = 0; j < 128; j++)
                if (i != j) mx = std::max(mx, ascii[i * 128 + j]);
        for (auto &kv : map)
            if (kv.first.first != kv.first.second) mx = std::max(mx, kv.second);
        return mx;
    }
};

static Similarity default_similarity() {  // builder.rs:492-526
    Similarity s;
    const char *vowels = "aeiou";
    auto is_vowel = [&](char c) { return strchr(vowels, c) != nullptr; };
    for (const char *a = vowels; *a; a++)
        for (const char *b = vowels; *b; b++)
            if (*a != *b) s.map[{(uint32_t)*a, (uint32_t)*b}] = 0.6f;
    for (char a = 'a'; a <= 'z'; a++)
        for (char b = 'a'; b <= 'z'; b++)
            if (!is_vowel(a) && !is_vowel(b) && a != b) s.map[{(uint32_t)a, (uint32_t)b}] = 0.4f;
    auto put = [&](char a, char b, float v) { s.map[{(uint32_t)a, (uint32_t)b}] = v; s.map[{(uint32_t)b, (uint32_t)a}] = v; };
    put('o', '0', 0.6f); put('l', '1', 0.7f); put('i', '1', 0.6f); put('s', '5', 0.5f);
    s.finish();
    return s;
}

struct Match {  // FuzzyMatch, structs.rs:757-781
    uint8_t ins, del, sub, swp, edits;
    size_t pattern_index;
    size_t start, end;
    float similarity;
};

struct Engine {  // FuzzyAhoCorasick, structs.rs:528-567
    std::vector<Node> nodes;
    std::vector<Pattern> patterns;
    Similarity similarity;
    bool has_limits = false;
    Limits limits;
    float pen_sub, pen_ins, pen_del, pen_swap;
    bool ci = false;
    bool has_pattern_limits = false;
    int max_edits_fast = 0;
    std::unordered_map<uint32_t, std::vector<MapTrans>> mappings;
    bool has_beam = false; size_t beam_width = 0;
    bool has_auto_beam = false; size_t ab_budget = 0, ab_width = 0;
    float min_symbol_similarity = 0.f;
    // ---- bitap pre-filter (prefilter.rs:66-93), built lazily by build_prefilter()
    bool pf_built = false, pf_active = false;
    std::unordered_map<std::string, uint32_t> pf_symbol_ids;
    uint8_t pf_ascii_id[128];
    struct BitapPattern { size_t m; float weight; std::vector<uint64_t> mask; bool has_k_limit; size_t k_limit; };
    std::vector<BitapPattern> pf_patterns;
    float pf_edit_cost_mult = 0.f;
    mutable uint64_t last_states_pushed = 0;
};

// ------------------------------------------------------------------------------------
// Builder (src/builder.rs:181-484)
// ------------------------------------------------------------------------------------
static Engine *build(const fac_config *cfg, const fac_pattern *pats, size_t np) {
    Engine *E = new Engine();
    E->ci = cfg->case_insensitive != 0;
    // FuzzyPenalties::default, structs.rs:381-393 (f32 products)
    const float m = 1.3f;
    E->pen_sub = 1.1f * m; E->pen_ins = 0.4f * m; E->pen_del = 0.7f * m; E->pen_swap = 0.4f * m;
    if (cfg->has_penalties) {
        E->pen_sub = cfg->penalty_substitution; E->pen_ins = cfg->penalty_insertion;
        E->pen_del = cfg->penalty_deletion; E->pen_swap = cfg->penalty_swap;
    }
    if (cfg->has_similarity) {
        for (size_t i = 0; i < cfg->n_similarity; i++) E->similarity.map[{cfg->similarity[i].a, cfg->similarity[i].b}] = cfg->similarity[i].similarity;
        E->similarity.finish();
    } else E->similarity = default_similarity();
    E->has_beam = cfg->beam_width != 0; E->beam_width = (size_t)cfg->beam_width;
    E->has_auto_beam = cfg->has_auto_beam != 0; E->ab_budget = (size_t)cfg->auto_beam_budget; E->ab_width = (size_t)cfg->auto_beam_width;
    E->min_symbol_similarity = cfg->min_symbol_similarity;

    for (size_t i = 0; i < np; i++) {
        Pattern p;
        p.text.assign(pats[i].text, pats[i].len);
        std::vector<size_t> st;
        grapheme_starts((const uint8_t *)p.text.data(), p.text.size(), st);
        p.glen = st.size();
        p.weight = pats[i].weight;
        p.has_limits = pats[i].has_limits != 0;
        if (p.has_limits) p.limits = finalize(from_c(pats[i].limits));  // Pattern::fuzzy, structs.rs:647-650
        p.uid = pats[i].unique_id;
        E->patterns.push_back(p);
    }

    auto &nodes = E->nodes;
    nodes.emplace_back();
    // trie insert, builder.rs:195-237
    for (size_t i = 0; i < E->patterns.size(); i++) {
        std::vector<std::string> word = fold_graphemes(E->patterns[i].text, E->ci);
        size_t cur = 0;
        for (auto &g : word) {
            size_t next;
            auto it = nodes[cur].trans.find(g);
            if (it != nodes[cur].trans.end()) next = it->second;
            else {
                next = nodes.size();
                nodes[cur].trans[g] = (uint32_t)next;
                nodes[cur].order.push_back({g, (uint32_t)next});
                nodes.emplace_back();
            }
            if (nodes[next].pattern_index < 0) nodes[next].pattern_index = (int64_t)i;  // builder.rs:227
            cur = next;
        }
        nodes[cur].output.push_back((uint32_t)i);  // builder.rs:235
    }
    // fail links + output merge, builder.rs:240-276
    {
        std::deque<uint32_t> q;
        for (auto &kv : nodes[0].order) { nodes[kv.second].fail = 0; q.push_back(kv.second); }
        while (!q.empty()) {
            uint32_t cur = q.front(); q.pop_front();
            for (auto &kv : nodes[cur].order) {
                const std::string &g = kv.first; uint32_t next = kv.second;
                uint32_t f = nodes[cur].fail;
                while (f != 0 && !nodes[f].trans.count(g)) f = nodes[f].fail;
                uint32_t fallback = 0;
                auto it = nodes[f].trans.find(g);
                if (it != nodes[f].trans.end()) fallback = it->second;
                // NOTE reference quirk kept: for a depth-1 child f==0 and the lookup finds the
                // child itself only when cur==0, which never happens here (cur is never root).
                nodes[next].fail = fallback;
                std::vector<uint32_t> fo = nodes[fallback].output;
                for (uint32_t e : fo)
                    if (std::find(nodes[next].output.begin(), nodes[next].output.end(), e) == nodes[next].output.end())
                        nodes[next].output.push_back(e);
                q.push_back(next);
            }
        }
    }
    // effective limits, builder.rs:289-329
    if (cfg->has_limits) { E->has_limits = true; E->limits = finalize(from_c(cfg->limits)); }
    else {
        Limits mx; bool any = false;
        for (auto &p : E->patterns)
            if (p.has_limits) {
                any = true;
                if (p.limits.edits >= 0) mx.edits = std::max(std::max(mx.edits, 0), p.limits.edits);
                if (p.limits.ins >= 0) mx.ins = std::max(std::max(mx.ins, 0), p.limits.ins);
                if (p.limits.del >= 0) mx.del = std::max(std::max(mx.del, 0), p.limits.del);
                if (p.limits.sub >= 0) mx.sub = std::max(std::max(mx.sub, 0), p.limits.sub);
                if (p.limits.swp >= 0) mx.swp = std::max(std::max(mx.swp, 0), p.limits.swp);
            }
        if (any) { E->has_limits = true; E->limits = mx; }
    }
    // edges, builder.rs:336-342 (order: U1)
    for (auto &nd : nodes)
        for (auto &kv : nd.order) nd.edges.push_back(Edge{first_char(kv.first), kv.second, kv.first.size() == 1});
    // prune coefficients, builder.rs:348-381
    {
        std::vector<size_t> rl(nodes.size(), 0);
        std::vector<float> rw(nodes.size(), 0.f);
        for (size_t i = 0; i < nodes.size(); i++)
            for (uint32_t p : nodes[i].output) { rl[i] = std::max(rl[i], E->patterns[p].glen); rw[i] = std::max(rw[i], E->patterns[p].weight); }
        bool changed = true;
        while (changed) {
            changed = false;
            for (size_t i = nodes.size(); i-- > 0;) {
                size_t bl = rl[i]; float bw = rw[i];
                for (auto &kv : nodes[i].order) { bl = std::max(bl, rl[kv.second]); bw = std::max(bw, rw[kv.second]); }
                if (bl > rl[i] || bw > rw[i]) { rl[i] = bl; rw[i] = bw; changed = true; }
            }
        }
        for (size_t i = 0; i < nodes.size(); i++) {
            float len = (float)rl[i];
            nodes[i].prune_len = len;
            nodes[i].prune_low = len / rw[i];
        }
    }
    // mapping transitions, builder.rs:390-442
    if (cfg->n_mappings) {
        struct Dir { std::vector<std::string> pat, hay; float pen; };
        std::vector<Dir> directed;
        for (size_t k = 0; k < cfg->n_mappings; k++) {
            auto ga = fold_graphemes(std::string(cfg->mappings[k].a, cfg->mappings[k].a_len), E->ci);
            auto gb = fold_graphemes(std::string(cfg->mappings[k].b, cfg->mappings[k].b_len), E->ci);
            if (ga.empty() || gb.empty() || ga == gb) continue;
            float pen = E->pen_sub * (1.0f - cfg->mappings[k].score);
            directed.push_back({ga, gb, pen});
            directed.push_back({gb, ga, pen});
        }
        for (size_t start = 0; start < nodes.size(); start++) {
            std::vector<MapTrans> mts;
            for (auto &d : directed) {
                size_t cur = start; bool ok = true;
                for (auto &g : d.pat) {
                    auto it = nodes[cur].trans.find(g);
                    if (it == nodes[cur].trans.end()) { ok = false; break; }
                    cur = it->second;
                }
                if (ok) mts.push_back(MapTrans{d.hay, (uint32_t)cur, d.pen});
            }
            if (!mts.empty()) E->mappings[(uint32_t)start] = mts;
        }
    }
    E->has_pattern_limits = false;
    for (auto &p : E->patterns) if (p.has_limits) E->has_pattern_limits = true;
    // max_edits_fast, builder.rs:451-468
    if (E->has_pattern_limits) E->max_edits_fast = 255;
    else if (!E->has_limits) E->max_edits_fast = 0;
    else if (E->limits.edits >= 0 && E->limits.ins < 0 && E->limits.del < 0 && E->limits.sub < 0 && E->limits.swp < 0) E->max_edits_fast = E->limits.edits;
    else E->max_edits_fast = 255;
    return E;
}

// stream.rs:213-253
static size_t max_match_graphemes(const Engine &E) {
    size_t max_pattern = 0;
    for (auto &p : E.patterns) max_pattern = std::max(max_pattern, p.glen);
    size_t mmh = 0; bool anym = false;
    for (auto &kv : E.mappings) for (auto &mt : kv.second) { mmh = std::max(mmh, mt.hay.size()); anym = true; }
    if (!anym) mmh = 1;
    mmh = std::max<size_t>(mmh, 1);
    auto edits_of = [](const Limits &l) -> size_t {
        if (l.edits >= 0) return (size_t)l.edits;
        return (size_t)std::max(l.ins, 0) + (size_t)std::max(l.del, 0) + (size_t)std::max(l.sub, 0) + (size_t)std::max(l.swp, 0);
    };
    size_t max_edits = 0;
    for (auto &p : E.patterns) {
        size_t e = 0;
        if (p.has_limits) e = edits_of(p.limits);
        else if (E.has_limits) e = edits_of(E.limits);
        max_edits = std::max(max_edits, e);
    }
    return max_pattern + max_edits * mmh;
}

// ------------------------------------------------------------------------------------
// Search (src/search.rs)
// ------------------------------------------------------------------------------------
struct Hay {  // GraphemeStorage, grapheme.rs:33-125
    const uint8_t *bytes; size_t len;
    bool ascii;
    bool ci;
    std::vector<size_t> off;          // unicode only
    std::vector<std::string> folded;  // unicode only (folded grapheme text)
    std::vector<uint32_t> first;      // text_chars (search.rs:203, 302)
    size_t n() const { return first.size(); }
    size_t byte_offset(size_t i) const { return ascii ? i : off[i]; }
};

static void make_hay(const Engine &E, const uint8_t *s, size_t len, Hay &H) {
    H.bytes = s; H.len = len; H.ci = E.ci;
    H.ascii = true;
    for (size_t i = 0; i < len; i++) if (s[i] >= 0x80) { H.ascii = false; break; }
    if (H.ascii) {  // AsciiGraphemes, grapheme.rs:76-125
        H.first.resize(len);
        for (size_t i = 0; i < len; i++) { uint8_t b = s[i]; H.first[i] = (E.ci && b >= 'A' && b <= 'Z') ? b + 32 : b; }
    } else {  // build_unicode_graphemes, search.rs:398-416
        grapheme_starts(s, len, H.off);
        size_t n = H.off.size();
        H.folded.resize(n); H.first.resize(n);
        for (size_t k = 0; k < n; k++) {
            size_t b = H.off[k], e = (k + 1 < n) ? H.off[k + 1] : len;
            if (E.ci) H.folded[k] = to_lowercase(s + b, e - b);
            else H.folded[k].assign((const char *)s + b, e - b);
            H.first[k] = first_char(H.folded[k]);
        }
    }
}

struct State { uint32_t node, j, ms, me; float pen; uint8_t edits; uint32_t cnt; };
struct VKey {
    uint32_t node, j, ms, me, cnt;
    bool operator==(const VKey &o) const { return node == o.node && j == o.j && ms == o.ms && me == o.me && cnt == o.cnt; }
};
struct VKeyHash {
    size_t operator()(const VKey &k) const {
        uint64_t h = 0;
        auto add = [&](uint64_t v) { h = ((h << 5) | (h >> 59)) ^ v; h *= 0x517cc1b727220a95ULL; };
        add((uint64_t)k.node | ((uint64_t)k.j << 32)); add((uint64_t)k.ms | ((uint64_t)k.me << 32)); add(k.cnt);
        return (size_t)(h ^ (h >> 29));
    }
};
// The reference's `visited` is an FxHashMap pre-sized and cleared per start window (search.rs:456-480, 608-628).
// Open addressing with an epoch stamp per slot gives the same map semantics with an O(1) clear, so that the
// CPU baseline timed by bench.py is not handicapped by node-based std::unordered_map allocation.
struct FlatVisited {
    struct Slot { VKey k; float pen; uint32_t epoch; };
    std::vector<Slot> tab;
    size_t mask = 0, count = 0;
    uint32_t epoch = 1;
    FlatVisited() { tab.assign(1024, Slot{VKey{0, 0, 0, 0, 0}, 0.f, 0}); mask = 1023; }
    void clear() {
        count = 0;
        if (++epoch == 0) { for (auto &sl : tab) sl.epoch = 0; epoch = 1; }
    }
    void grow() {
        std::vector<Slot> old;
        old.swap(tab);
        tab.assign(old.size() * 2, Slot{VKey{0, 0, 0, 0, 0}, 0.f, 0});
        mask = tab.size() - 1;
        for (const Slot &sl : old)
            if (sl.epoch == epoch) {
                size_t h = VKeyHash()(sl.k) & mask;
                while (tab[h].epoch == epoch) h = (h + 1) & mask;
                tab[h] = sl;
            }
    }
    // the stored minimum of `k` if present, else inserts (k, pen) and returns nullptr
    float *find_or_insert(const VKey &k, float pen) {
        if ((count + 1) * 2 > tab.size()) grow();
        size_t h = VKeyHash()(k) & mask;
        for (;;) {
            Slot &sl = tab[h];
            if (sl.epoch != epoch) { sl.k = k; sl.pen = pen; sl.epoch = epoch; count++; return nullptr; }
            if (sl.k == k) return &sl.pen;
            h = (h + 1) & mask;
        }
    }
};
struct BKey { size_t s, e, p; bool operator<(const BKey &o) const { return s != o.s ? s < o.s : (e != o.e ? e < o.e : p < o.p); } };

static inline int32_t total_key(float f) {  // f32::total_cmp
    int32_t b; memcpy(&b, &f, 4);
    b ^= (int32_t)(((uint32_t)(b >> 31)) >> 1);
    return b;
}

// limits helpers, search.rs:84-169.  `nl` = node/pattern limits (may be null), falls back to global.
static inline const Limits *pick(const Engine &E, const Limits *nl) { return nl ? nl : (E.has_limits ? &E.limits : nullptr); }
static inline bool none_or_lt(int mx, int v) { return mx < 0 || v < mx; }
static inline bool none_or_le(int mx, int v) { return mx < 0 || v <= mx; }

static inline const Limits *node_limits(const Engine &E, uint32_t node) {  // search.rs:67-71
    int64_t pi = E.nodes[node].pattern_index;
    if (pi < 0) return nullptr;
    const Pattern &p = E.patterns[(size_t)pi];
    return p.has_limits ? &p.limits : nullptr;
}

// First edge whose first_char == ch (structs.rs:512-519)
static inline int64_t find_no_map(const Node &nd, uint32_t ch) {
    for (auto &e : nd.edges) if (e.first_char == ch) return e.next;
    return -1;
}
// structs.rs:499-506
static inline int64_t find_char(const Node &nd, uint32_t ch) {
    for (auto &e : nd.edges) if (e.first_char == ch && e.single_byte) return e.next;
    return -1;
}
// gs_find_transition, grapheme.rs:72-74 / 120-124 + structs.rs:452-464
static inline int64_t find_full(const Hay &H, const Node &nd, size_t idx, uint32_t ch) {
    if (H.ascii) return find_char(nd, ch);
    const std::string &g = H.folded[idx];
    if (g.size() == 1) return find_char(nd, (uint8_t)g[0]);
    auto it = nd.trans.find(g);
    return it == nd.trans.end() ? -1 : (int64_t)it->second;
}
static inline bool has_matching_edge_char(const Node &nd, uint32_t ch) {  // structs.rs:471-475
    for (auto &e : nd.edges) if (e.first_char == ch && e.single_byte) return true;
    return false;
}
static inline bool hay_text_eq(const Hay &H, size_t idx, const std::string &g) {
    if (H.ascii) { return g.size() == 1 && (uint8_t)g[0] == (uint8_t)H.first[idx]; }
    return H.folded[idx] == g;
}

// search_unsorted_impl, search.rs:418-1119.  Appends raw best-per-span matches (sorted by key, U2).
// `per_window_states` (optional) receives queue.len() of every start window.
static void search_raw(const Engine &E, const uint8_t *s, size_t len, float thr, std::vector<Match> &out,
                       uint64_t *states_pushed, std::vector<uint32_t> *per_window_states = nullptr) {
    Hay H; make_hay(E, s, len, H);
    const size_t n = H.n();
    if (states_pushed) *states_pushed = 0;
    if (n == 0) return;
    const bool MAPP = !E.mappings.empty();
    const int mef = E.max_edits_fast;
    const int MEF = (mef >= 1 && mef <= 6) ? mef : 255;  // search.rs:205-247
    const bool WS = (MEF == 1);
    const uint32_t text_len = (uint32_t)n;
    const std::vector<uint32_t> &tc = H.first;
    std::map<BKey, Match> best;
    std::vector<State> queue;
    FlatVisited visited;
    const Node &root = E.nodes[0];
    const float max_pen = root.prune_len - root.prune_low * thr;  // search.rs:487
    const float min_sym = E.min_symbol_similarity;
    // window skip, search.rs:504-521
    bool wskip = false; unsigned __int128 first_bits = 0, second_bits = 0;
    if (WS && !MAPP && root.output.empty()) {
        auto bits = [](const Node &nd) { unsigned __int128 b = 0; for (auto &e : nd.edges) if (e.single_byte && e.first_char < 128) b |= ((unsigned __int128)1) << e.first_char; return b; };
        first_bits = bits(root);
        bool child_output = false;
        for (auto &e : root.edges) {
            const Node &c = E.nodes[e.next];
            auto cb = bits(c);
            second_bits |= cb; first_bits |= cb;
            if (!c.output.empty()) child_output = true;
        }
        wskip = !child_output;
    }
    bool eff_beam = E.has_beam; size_t bw = E.beam_width;
    size_t states_expanded = 0;
    uint64_t total_pushed = 0;
    if (per_window_states) per_window_states->assign(n, 0);

    for (size_t start_us = 0; start_us < n; start_us++) {
        if (wskip) {  // search.rs:535-553
            uint32_t ch = tc[start_us];
            if (ch < 128 && ((first_bits >> ch) & 1) == 0) {
                size_t ni = start_us + 1;
                if (ni >= n) continue;
                uint32_t nc = tc[ni];
                if (nc < 128 && ((second_bits >> nc) & 1) == 0) continue;
            }
        }
        queue.clear(); visited.clear();
        const uint32_t start = (uint32_t)start_us;
        queue.push_back(State{0, start, start, start, 0.f, 0, 0});
        size_t q_idx = 0;
        while (q_idx < queue.size()) {
            if (eff_beam) {  // search.rs:578-589 (selection order: U3)
                size_t remaining = queue.size() - q_idx;
                if (remaining > bw * 2) {
                    std::vector<std::pair<int32_t, size_t>> keys;
                    keys.reserve(remaining);
                    for (size_t k = q_idx; k < queue.size(); k++) keys.push_back({total_key(queue[k].pen), k});
                    std::nth_element(keys.begin(), keys.begin() + (bw - 1), keys.end());
                    keys.resize(bw);
                    std::sort(keys.begin(), keys.end(), [](auto &a, auto &b) { return a.second < b.second; });
                    std::vector<State> kept; kept.reserve(bw);
                    for (auto &kk : keys) kept.push_back(queue[kk.second]);
                    queue.resize(q_idx);
                    queue.insert(queue.end(), kept.begin(), kept.end());
                }
            }
            const State S = queue[q_idx++];
            const uint32_t node = S.node, j = S.j, ms = S.ms, me = S.me, cnt = S.cnt;
            const float pen = S.pen; const int edits = S.edits;
            VKey key{node, j, ms, me, cnt};  // search.rs:608-628
            if (float *seen = visited.find_or_insert(key, pen)) { if (*seen <= pen) continue; *seen = pen; }
            const Node &nd = E.nodes[node];
            if (pen > nd.prune_len - nd.prune_low * thr) continue;  // search.rs:638-642
            const float remaining = max_pen - pen;                    // search.rs:648
            const Limits *nl = E.has_pattern_limits ? node_limits(E, node) : nullptr;  // search.rs:653-657
            const int c_ins = cnt & 0xFF, c_del = (cnt >> 8) & 0xFF, c_sub = (cnt >> 16) & 0xFF, c_swp = (cnt >> 24) & 0xFF;
            if (!nd.output.empty()) {  // search.rs:659-737
                size_t sb = (ms < n) ? H.byte_offset(ms) : 0;
                size_t eb = (me < n) ? H.byte_offset(me) : len;
                for (uint32_t pi : nd.output) {
                    const Pattern &P = E.patterns[pi];
                    if (MEF != 255) { if (edits > MEF) continue; }
                    else {
                        const Limits *L = pick(E, P.has_limits ? &P.limits : nullptr);  // search.rs:151-169
                        bool ok;
                        if (L) ok = none_or_le(L->edits, edits) && none_or_le(L->ins, c_ins) && none_or_le(L->del, c_del) && none_or_le(L->sub, c_sub) && none_or_le(L->swp, c_swp);
                        else ok = edits == 0 && c_ins == 0 && c_del == 0 && c_sub == 0 && c_swp == 0;
                        if (!ok) continue;
                    }
                    float total = (float)P.glen;
                    float q = (total - pen) / total;
                    float sim = q * P.weight;  // search.rs:698-699
                    if (sim < thr) continue;
                    BKey bk{sb, eb, pi};
                    auto bit = best.find(bk);
                    Match mm{(uint8_t)c_ins, (uint8_t)c_del, (uint8_t)c_sub, (uint8_t)c_swp, (uint8_t)edits, pi, sb, eb, sim};
                    if (bit == best.end()) best.emplace(bk, mm);
                    else if (sim > bit->second.similarity) bit->second = mm;
                }
            }
            const bool is_last = (MEF != 255) && (edits + 1 >= MEF);  // search.rs:742
            const uint32_t cur = (j < text_len) ? tc[j] : 0;
            if (j < text_len) {
                bool has_nxt = is_last && (MEF == 255 || edits < MEF) && (j + 1 < text_len);  // search.rs:758-765
                uint32_t nxt = has_nxt ? tc[j + 1] : 0;
                uint32_t msn = (me == ms) ? j : ms;  // search.rs:766-770
                int64_t ex = MAPP ? find_full(H, nd, j, cur) : find_no_map(nd, cur);  // search.rs:776-780
                if (ex >= 0) queue.push_back(State{(uint32_t)ex, j + 1, msn, j + 1, pen, (uint8_t)edits, cnt});
                bool sub_ok;
                if (MEF == 255) {  // within_limits_subst, search.rs:134-146
                    const Limits *L = pick(E, nl);
                    sub_ok = L ? (none_or_lt(L->edits, edits) && none_or_lt(L->sub, (int)((cnt >> 16) & 0xFF))) : (edits == 0 && ((cnt >> 16) & 0xFF) == 0);
                } else sub_ok = edits < MEF;
                if (sub_ok) {
                    for (auto &e : nd.edges) {  // search.rs:814-874
                        if (ex >= 0 && (int64_t)e.next == ex) continue;
                        float sm = (e.first_char == cur) ? 1.0f : E.similarity.get(e.first_char, cur);
                        if (sm < min_sym) continue;
                        float pp = E.pen_sub * (1.0f - sm);
                        if (pp > remaining) continue;
                        if (is_last) {
                            const Node &c = E.nodes[e.next];
                            if (c.output.empty() && (!has_nxt || !has_matching_edge_char(c, nxt))) continue;
                        }
                        queue.push_back(State{e.next, j + 1, msn, j + 1, pen + pp, (uint8_t)(edits + 1), cnt + 0x10000u});
                    }
                    if (MAPP) {  // search.rs:883-923
                        auto mit = E.mappings.find(node);
                        if (mit != E.mappings.end())
                            for (auto &mt : mit->second) {
                                uint32_t hlen = (uint32_t)mt.hay.size();
                                if ((uint64_t)j + hlen > text_len) continue;
                                bool okm = true;
                                for (uint32_t k = 0; k < hlen; k++) if (!hay_text_eq(H, j + k, mt.hay[k])) { okm = false; break; }
                                if (!okm) continue;
                                float np = pen + mt.pen;
                                if (np > max_pen) continue;
                                queue.push_back(State{mt.next, j + hlen, msn, j + hlen, np, (uint8_t)(edits + 1), cnt + 0x10000u});
                            }
                    }
                }
                // swap, search.rs:935-989
                if (j + 1 < text_len && E.pen_swap <= remaining && (MEF == 255 || edits < MEF)) {
                    uint32_t nc = has_nxt ? nxt : tc[j + 1];
                    int64_t x = MAPP ? find_full(H, nd, j + 1, nc) : find_no_map(nd, nc);
                    int64_t n2 = -1;
                    if (x >= 0) n2 = MAPP ? find_full(H, E.nodes[(size_t)x], j, cur) : find_no_map(E.nodes[(size_t)x], cur);
                    if (n2 >= 0) {
                        bool ok = true;
                        if (MEF == 255) {  // within_limits_swap_ahead(get_node_limits(node2)), search.rs:119-130
                            const Limits *L = pick(E, node_limits(E, (uint32_t)n2));
                            ok = L ? (none_or_lt(L->edits, edits) && none_or_lt(L->swp, (int)(cnt >> 24))) : false;
                        }
                        if (ok) queue.push_back(State{(uint32_t)n2, j + 2, ms, j + 2, pen + E.pen_swap, (uint8_t)(edits + 1), cnt + 0x1000000u});
                    }
                }
                // insertion, search.rs:994-1029
                {
                    bool ok = (ms != me || ms != j) && E.pen_ins <= remaining;
                    if (ok) {
                        if (MEF == 255) {
                            const Limits *L = pick(E, nl);
                            ok = L ? (none_or_lt(L->edits, edits) && none_or_lt(L->ins, (int)(cnt & 0xFF))) : false;
                        } else ok = edits < MEF;
                    }
                    if (ok && is_last && nd.output.empty() && (!has_nxt || !has_matching_edge_char(nd, nxt))) ok = false;
                    if (ok) queue.push_back(State{node, j + 1, ms, me, pen + E.pen_ins, (uint8_t)(edits + 1), cnt + 1u});
                }
            }
            // deletion, search.rs:1035-1089
            {
                bool ok = E.pen_del <= remaining;
                if (ok) {
                    if (MEF == 255) {
                        const Limits *L = pick(E, nl);
                        ok = L ? (none_or_lt(L->edits, edits) && none_or_lt(L->del, (int)((cnt >> 8) & 0xFF))) : false;
                    } else ok = edits < MEF;
                }
                if (ok) {
                    bool has_co = is_last && j < text_len;
                    for (auto &e : nd.edges) {
                        if (is_last) {
                            const Node &c = E.nodes[e.next];
                            if (c.output.empty() && (!has_co || !has_matching_edge_char(c, cur))) continue;
                        }
                        queue.push_back(State{e.next, j, ms, me, pen + E.pen_del, (uint8_t)(edits + 1), cnt + 0x100u});
                    }
                }
            }
        }
        total_pushed += queue.size();
        if (per_window_states) (*per_window_states)[start_us] = (uint32_t)queue.size();
        if (E.has_auto_beam && !eff_beam) {  // search.rs:1096-1103
            states_expanded += queue.size();
            if (states_expanded > E.ab_budget) { eff_beam = true; bw = E.ab_width; }
        }
    }
    for (auto &kv : best) out.push_back(kv.second);
    if (states_pushed) *states_pushed = total_pushed;
}

// ------------------------------------------------------------------------------------
// Ranking + overlap (src/matches.rs:7-149)
// ------------------------------------------------------------------------------------
static void apply(const Engine &E, std::vector<Match> &v, int order, int overlap) {
    auto plen = [&](const Match &m) { return E.patterns[m.pattern_index].text.size(); };
    auto tail = [](const Match &l, const Match &r, bool &res) {
        if (l.start != r.start) { res = l.start < r.start; return true; }
        if (l.end != r.end) { res = l.end < r.end; return true; }
        if (l.pattern_index != r.pattern_index) { res = l.pattern_index < r.pattern_index; return true; }
        return false;
    };
    if (order == FAC_ORDER_DEFAULT) {  // matches.rs:24-41
        std::sort(v.begin(), v.end(), [&](const Match &l, const Match &r) {
            int32_t a = total_key(l.similarity), b = total_key(r.similarity);
            if (a != b) return a > b;
            if (plen(l) != plen(r)) return plen(l) > plen(r);
            size_t tl = l.end - l.start, tr = r.end - r.start;
            if (tl != tr) return tl > tr;
            bool res; if (tail(l, r, res)) return res; return false;
        });
    } else if (order == FAC_ORDER_GREEDY) {  // matches.rs:46-61
        std::sort(v.begin(), v.end(), [&](const Match &l, const Match &r) {
            if (plen(l) != plen(r)) return plen(l) > plen(r);
            int32_t a = total_key(l.similarity), b = total_key(r.similarity);
            if (a != b) return a > b;
            bool res; if (tail(l, r, res)) return res; return false;
        });
    } else if (order == FAC_ORDER_COVERAGE_WEIGHTED) {  // matches.rs:67-84
        std::sort(v.begin(), v.end(), [&](const Match &l, const Match &r) {
            float ls = l.similarity * l.similarity * (float)plen(l);
            float rs = r.similarity * r.similarity * (float)plen(r);
            int32_t a = total_key(ls), b = total_key(rs);
            if (a != b) return a > b;
            a = total_key(l.similarity); b = total_key(r.similarity);
            if (a != b) return a > b;
            bool res; if (tail(l, r, res)) return res; return false;
        });
    }
    if (overlap == FAC_OVERLAP_KEEP) return;
    const bool unique = overlap == FAC_OVERLAP_NON_OVERLAPPING_UNIQUE;
    std::set<std::pair<int, size_t>> used;  // UniqueId: Automatic < Custom (structs.rs:586-592)
    std::vector<std::pair<size_t, size_t>> occupied;
    std::vector<Match> kept;
    for (auto &m : v) {  // matches.rs:86-149
        std::pair<int, size_t> uid;
        if (unique) {
            const Pattern &P = E.patterns[m.pattern_index];
            uid = P.uid >= 0 ? std::make_pair(1, (size_t)P.uid) : std::make_pair(0, m.pattern_index);
            if (used.count(uid)) continue;
        }
        // binary_search_by(|(s,_)| s.cmp(&m.start)).unwrap_or_else(|p| p): on Ok any matching index
        // may be returned; with equal starts only possible for empty spans the choice is U4-adjacent.
        size_t pos;
        {   // lower bound on start (equal starts only arise with empty spans)
            size_t l = 0, r = occupied.size();
            while (l < r) { size_t mid = l + (r - l) / 2; if (occupied[mid].first < m.start) l = mid + 1; else r = mid; }
            pos = l;
        }
        bool prev_ok = pos == 0 || occupied[pos - 1].second <= m.start;
        bool next_ok = pos == occupied.size() || occupied[pos].first >= m.end;
        if (prev_ok && next_ok) {
            if (unique) used.insert(uid);
            occupied.insert(occupied.begin() + pos, {m.start, m.end});
            kept.push_back(m);
        }
    }
    std::stable_sort(kept.begin(), kept.end(), [](const Match &a, const Match &b) { return a.start < b.start; });  // U4
    v.swap(kept);
}

// ------------------------------------------------------------------------------------
// Bitap pre-filter (src/prefilter.rs)
// ------------------------------------------------------------------------------------
static bool k_from_limits(const Limits &l, size_t &k) {  // prefilter.rs:388-405
    if (l.edits >= 0) { k = (l.swp == 0) ? (size_t)l.edits : 2 * (size_t)l.edits; return true; }
    if (l.ins < 0 || l.del < 0 || l.sub < 0 || l.swp < 0) return false;
    k = (size_t)l.ins + (size_t)l.del + (size_t)l.sub + 2 * (size_t)l.swp;
    return true;
}

static void build_prefilter(Engine &E) {  // BitapFilter::build, prefilter.rs:161-245
    E.pf_built = true; E.pf_active = false;
    if (!E.mappings.empty()) return;
    if (E.patterns.empty()) return;
    float max_sim = E.similarity.max_off_diagonal();
    float p_sub_min = E.pen_sub * (1.0f - max_sim);
    float mults[4] = {1.0f / E.pen_ins, 1.0f / E.pen_del, 1.0f / p_sub_min, 2.0f / E.pen_swap};
    for (float m : mults) if (!std::isfinite(m) || m <= 0.0f) return;
    float mult = 0.f;
    for (float m : mults) mult = std::max(mult, m);
    E.pf_edit_cost_mult = mult;
    std::vector<std::vector<uint32_t>> all_ids;
    for (auto &pat : E.patterns) {
        auto gs = fold_graphemes(pat.text, E.ci);
        size_t m = gs.size();
        if (m == 0 || m > 63) return;
        std::vector<uint32_t> ids;
        for (auto &g : gs) {
            uint32_t next_id = (uint32_t)E.pf_symbol_ids.size() + 1;
            auto it = E.pf_symbol_ids.find(g);
            uint32_t id;
            if (it == E.pf_symbol_ids.end()) { E.pf_symbol_ids[g] = next_id; id = next_id; } else id = it->second;
            if (id > 255) { E.pf_symbol_ids.clear(); E.pf_patterns.clear(); return; }
            ids.push_back(id);
        }
        Engine::BitapPattern bp; bp.m = m; bp.weight = pat.weight;
        const Limits *L = pat.has_limits ? &pat.limits : (E.has_limits ? &E.limits : nullptr);
        bp.has_k_limit = false; bp.k_limit = 0;
        if (L) bp.has_k_limit = k_from_limits(*L, bp.k_limit);
        E.pf_patterns.push_back(bp);
        all_ids.push_back(ids);
    }
    memset(E.pf_ascii_id, 0, sizeof(E.pf_ascii_id));
    for (int b = 0; b < 128; b++) {
        std::string f; lower_cp((uint32_t)b, f);
        if (!E.ci) { f.assign(1, (char)b); }
        auto it = E.pf_symbol_ids.find(f);
        if (it != E.pf_symbol_ids.end()) E.pf_ascii_id[b] = (uint8_t)it->second;
    }
    size_t alphabet = E.pf_symbol_ids.size();
    for (size_t i = 0; i < E.pf_patterns.size(); i++) {
        E.pf_patterns[i].mask.assign(alphabet + 1, 0);
        for (size_t k = 0; k < all_ids[i].size(); k++) E.pf_patterns[i].mask[all_ids[i][k]] |= 1ULL << k;
    }
    E.pf_active = true;
}

static bool k_for(const Engine &E, const Engine::BitapPattern &pat, float thr, size_t &k) {  // prefilter.rs:285-302
    float n = (float)pat.m;
    float p_max = n * (1.0f - thr / pat.weight);
    size_t k_pen;
    if (p_max <= 0.0f) k_pen = 0;
    else {
        float v = std::floor(p_max * E.pf_edit_cost_mult);
        // Rust `as usize` saturates; NaN -> 0
        if (std::isnan(v)) k_pen = 0; else if (v >= 1.8e19f) k_pen = SIZE_MAX; else k_pen = (size_t)v;
    }
    k = pat.has_k_limit ? std::min(k_pen, pat.k_limit) : k_pen;
    return k <= 24;
}

static void bitap_windows(const std::vector<uint64_t> &mask, size_t m, size_t k, const std::vector<uint8_t> &ids,
                          std::vector<std::pair<size_t, size_t>> &out) {  // prefilter.rs:410-435
    uint64_t match_bit = 1ULL << (m - 1);
    std::vector<uint64_t> r(k + 1), nr(k + 1);
    for (size_t d = 0; d <= k; d++) r[d] = (1ULL << d) - 1;
    size_t span = m + k;
    for (size_t i = 0; i < ids.size(); i++) {
        uint64_t bc = mask[ids[i]];
        nr[0] = ((r[0] << 1) | 1) & bc;
        for (size_t d = 1; d <= k; d++) nr[d] = ((r[d] << 1) & bc) | ((r[d - 1] | nr[d - 1]) << 1) | r[d - 1] | 1;
        if (nr[k] & match_bit) { size_t end = i + 1; out.push_back({end >= span ? end - span : 0, end}); }
        std::swap(r, nr);
    }
}

// BitapFilter::search_unsorted, prefilter.rs:304-374
static void prefiltered_raw(Engine &E, const uint8_t *s, size_t len, float thr, std::vector<Match> &out, uint64_t *states) {
    if (!E.pf_built) build_prefilter(E);
    if (!E.pf_active) { search_raw(E, s, len, thr, out, states); return; }
    std::vector<size_t> ks;
    for (auto &p : E.pf_patterns) { size_t k; if (!k_for(E, p, thr, k)) { search_raw(E, s, len, thr, out, states); return; } ks.push_back(k); }
    // transcode, prefilter.rs:251-281
    std::vector<uint8_t> ids; std::vector<size_t> offsets; bool identity = true;
    for (size_t i = 0; i < len; i++) if (s[i] >= 0x80) { identity = false; break; }
    if (identity) { ids.resize(len); for (size_t i = 0; i < len; i++) ids[i] = E.pf_ascii_id[s[i]]; }
    else {
        std::vector<size_t> st; grapheme_starts(s, len, st);
        for (size_t k = 0; k < st.size(); k++) {
            size_t b = st[k], e = (k + 1 < st.size()) ? st[k + 1] : len;
            offsets.push_back(b);
            std::string g = E.ci ? to_lowercase(s + b, e - b) : std::string((const char *)s + b, e - b);
            auto it = E.pf_symbol_ids.find(g);
            ids.push_back(it == E.pf_symbol_ids.end() ? 0 : (uint8_t)it->second);
        }
        offsets.push_back(len);
    }
    size_t n = ids.size();
    std::vector<std::pair<size_t, size_t>> windows;
    for (size_t i = 0; i < E.pf_patterns.size(); i++) bitap_windows(E.pf_patterns[i].mask, E.pf_patterns[i].m, ks[i], ids, windows);
    if (states) *states = 0;
    if (windows.empty()) return;
    std::sort(windows.begin(), windows.end());
    std::vector<std::pair<size_t, size_t>> merged;
    for (auto &w : windows) {
        if (!merged.empty() && w.first <= merged.back().second) merged.back().second = std::max(merged.back().second, w.second);
        else merged.push_back(w);
    }
    std::map<BKey, Match> best;
    uint64_t tot = 0;
    for (auto &w : merged) {
        size_t bstart = identity ? w.first : offsets[w.first];
        size_t ge = std::min(w.second, n);
        size_t bend = identity ? ge : offsets[ge];
        std::vector<Match> sub; uint64_t st = 0;
        search_raw(E, s + bstart, bend - bstart, thr, sub, &st);
        tot += st;
        for (auto &m : sub) {
            Match mm = m; mm.start += bstart; mm.end += bstart;
            BKey bk{mm.start, mm.end, mm.pattern_index};
            auto it = best.find(bk);
            if (it == best.end()) best.emplace(bk, mm);
            else if (mm.similarity > it->second.similarity) it->second = mm;
        }
    }
    for (auto &kv : best) out.push_back(kv.second);
    if (states) *states = tot;
}

// ------------------------------------------------------------------------------------
// Streaming (src/stream.rs)
// ------------------------------------------------------------------------------------
struct Window { uint64_t base; std::string text; size_t commit; };

struct WindowReader {  // stream.rs:77-159
    fac_read_fn read; void *user;
    std::vector<uint8_t> buf, chunk;
    uint64_t base = 0, total = 0;
    size_t window, overlap;
    bool done = false;
    WindowReader(fac_read_fn r, void *u, size_t w, size_t ov) : read(r), user(u), chunk(64 * 1024), window(w), overlap(ov) {}
    // returns 1 window, 0 end, -1 io error
    int next(Window &out) {
        if (done) return 0;
        for (;;) {
            while (buf.size() < window) {
                int64_t n = read(user, chunk.data(), chunk.size());
                if (n < 0) return -1;
                if (n == 0) break;
                buf.insert(buf.end(), chunk.begin(), chunk.begin() + n);
                total += (uint64_t)n;
            }
            bool eof = buf.size() < window;
            size_t valid = utf8_valid_up_to(buf.data(), buf.size());
            if (eof) {
                done = true;
                out.base = base; out.commit = valid; out.text.assign((const char *)buf.data(), valid);
                return 1;
            }
            // commit = byte offset of the overlap-th grapheme from the end, stream.rs:133-147
            std::vector<size_t> st; grapheme_starts(buf.data(), valid, st);
            size_t commit = 0; bool ok = false;
            if (st.size() >= overlap) { commit = st[st.size() - overlap]; ok = commit > 0; }
            if (!ok) { window += std::max<size_t>(window, 64 * 1024); continue; }
            out.base = base; out.text.assign((const char *)buf.data(), valid); out.commit = commit;
            buf.erase(buf.begin(), buf.begin() + commit);
            base += commit;
            return 1;
        }
    }
};

// window_matches, stream.rs:262-297
static void window_matches(const Engine &E, const Window &w, float thr, std::vector<Match> &out, std::vector<uint64_t> &abs_start,
                           std::vector<uint64_t> &abs_end) {
    std::vector<Match> v; uint64_t st;
    search_raw(E, (const uint8_t *)w.text.data(), w.text.size(), thr, v, &st);
    apply(E, v, FAC_ORDER_DEFAULT, FAC_OVERLAP_NON_OVERLAPPING);
    for (auto &m : v)
        if (m.start < w.commit) { out.push_back(m); abs_start.push_back(w.base + m.start); abs_end.push_back(w.base + m.end); }
}

}  // namespace orc

// =====================================================================================
// C interface for ctypes (tests / bench only).  Mirrors include/fac.h with an orc_ prefix.
// =====================================================================================
using namespace orc;

struct orc_matches { std::vector<fac_match> v; uint64_t states = 0; };

static void to_c(const std::vector<Match> &v, orc_matches *o, uint64_t base = 0) {
    for (auto &m : v) {
        fac_match c; memset(&c, 0, sizeof(c));
        c.start = m.start + base; c.end = m.end + base; c.pattern_index = (uint32_t)m.pattern_index; c.similarity = m.similarity;
        c.insertions = m.ins; c.deletions = m.del; c.substitutions = m.sub; c.swaps = m.swp; c.edits = m.edits;
        o->v.push_back(c);
    }
}

extern "C" {

void *orc_engine_create(const fac_config *cfg, const fac_pattern *pats, size_t n) { return build(cfg, pats, n); }
void orc_engine_free(void *e) { delete (Engine *)e; }
size_t orc_engine_max_match_graphemes(void *e) { return max_match_graphemes(*(Engine *)e); }
size_t orc_engine_num_nodes(void *e) { return ((Engine *)e)->nodes.size(); }
int orc_engine_max_edits_fast(void *e) { return ((Engine *)e)->max_edits_fast; }
int orc_engine_prefilter_active(void *e) { Engine *E = (Engine *)e; if (!E->pf_built) build_prefilter(*E); return E->pf_active; }

// returns 0 ok, 1 too large, 2 invalid utf8
int orc_search(void *e, const uint8_t *hay, size_t len, float thr, int order, int overlap, int use_prefilter, orc_matches **out) {
    Engine *E = (Engine *)e;
    if (utf8_valid_up_to(hay, len) != len) return FAC_INVALID_UTF8;
    orc_matches *o = new orc_matches();
    std::vector<Match> v;
    if (use_prefilter) prefiltered_raw(*E, hay, len, thr, v, &o->states);
    else search_raw(*E, hay, len, thr, v, &o->states);
    apply(*E, v, order, overlap);
    to_c(v, o);
    *out = o;
    return 0;
}

// per-window pushed-state counts (queue.len()), for kernel accounting tests
int orc_window_states(void *e, const uint8_t *hay, size_t len, float thr, uint32_t *counts, size_t cap) {
    Engine *E = (Engine *)e;
    std::vector<Match> v; uint64_t st; std::vector<uint32_t> pw;
    search_raw(*E, hay, len, thr, v, &st, &pw);
    for (size_t i = 0; i < pw.size() && i < cap; i++) counts[i] = pw[i];
    return (int)pw.size();
}

// FuzzyMatches::apply on a caller-provided list
int orc_apply(void *e, const fac_match *in, size_t n, int order, int overlap, orc_matches **out) {
    Engine *E = (Engine *)e;
    std::vector<Match> v;
    for (size_t i = 0; i < n; i++) v.push_back(Match{in[i].insertions, in[i].deletions, in[i].substitutions, in[i].swaps, in[i].edits, in[i].pattern_index, (size_t)in[i].start, (size_t)in[i].end, in[i].similarity});
    apply(*E, v, order, overlap);
    orc_matches *o = new orc_matches(); to_c(v, o); *out = o;
    return 0;
}

// The reference's own parallel decomposition (search_stream_parallel, stream.rs:378-429) on an
// in-memory input: 256 KiB windows + overlap, `threads` workers, matches with absolute offsets
// returned in window order.  Used as the multi-core CPU baseline.
struct MemReader { const uint8_t *p; size_t len, pos, block; };
static int64_t mem_read(void *u, uint8_t *buf, size_t cap) {
    MemReader *r = (MemReader *)u;
    size_t n = std::min(cap, r->len - r->pos);
    if (r->block) n = std::min(n, r->block - (r->pos % r->block));
    memcpy(buf, r->p + r->pos, n); r->pos += n;
    return (int64_t)n;
}

int orc_search_stream(void *e, const uint8_t *data, size_t len, size_t read_block, float thr, int threads, orc_matches **out) {
    Engine *E = (Engine *)e;
    MemReader mr{data, len, 0, read_block};
    WindowReader wr(mem_read, &mr, 256 * 1024, max_match_graphemes(*E) + 1);
    std::vector<Window> wins; Window w;
    while (wr.next(w) == 1) wins.push_back(w);
    std::vector<std::vector<fac_match>> res(wins.size());
    auto work = [&](size_t tid, size_t nt) {
        for (size_t i = tid; i < wins.size(); i += nt) {
            std::vector<Match> v; std::vector<uint64_t> as, ae;
            window_matches(*E, wins[i], thr, v, as, ae);
            for (size_t k = 0; k < v.size(); k++) {
                fac_match c; memset(&c, 0, sizeof(c));
                c.start = as[k]; c.end = ae[k]; c.pattern_index = (uint32_t)v[k].pattern_index; c.similarity = v[k].similarity;
                c.insertions = v[k].ins; c.deletions = v[k].del; c.substitutions = v[k].sub; c.swaps = v[k].swp; c.edits = v[k].edits;
                res[i].push_back(c);
            }
        }
    };
    size_t nt = (size_t)std::max(1, threads);
    if (nt == 1) work(0, 1);
    else { std::vector<std::thread> th; for (size_t t = 0; t < nt; t++) th.emplace_back(work, t, nt); for (auto &t : th) t.join(); }
    orc_matches *o = new orc_matches();
    for (auto &r : res) o->v.insert(o->v.end(), r.begin(), r.end());
    *out = o;
    return 0;
}

// Window cutting only (WindowReader): returns (base, len, commit) triples.
int orc_cut_windows(void *e, const uint8_t *data, size_t len, size_t read_block, uint64_t *triples, size_t cap) {
    Engine *E = (Engine *)e;
    MemReader mr{data, len, 0, read_block};
    WindowReader wr(mem_read, &mr, 256 * 1024, max_match_graphemes(*E) + 1);
    Window w; size_t k = 0;
    while (wr.next(w) == 1) { if (k < cap) { triples[3 * k] = w.base; triples[3 * k + 1] = w.text.size(); triples[3 * k + 2] = w.commit; } k++; }
    return (int)k;
}

// replace_stream (stream.rs:465-531, 654-704) with the same callback shape as fac_replace_fn.
// Output is appended to a growable buffer the caller frees with orc_free_buf.
int orc_replace_stream(void *e, const uint8_t *data, size_t len, size_t read_block, float thr, fac_replace_fn cb, void *cb_user,
                       uint8_t **out_buf, size_t *out_len) {
    Engine *E = (Engine *)e;
    MemReader mr{data, len, 0, read_block};
    WindowReader wr(mem_read, &mr, 256 * 1024, max_match_graphemes(*E) + 1);
    std::string out; uint64_t emitted = 0;
    Window w;
    while (wr.next(w) == 1) {
        std::vector<Match> v; uint64_t st;
        search_raw(*E, (const uint8_t *)w.text.data(), w.text.size(), thr, v, &st);
        apply(*E, v, FAC_ORDER_DEFAULT, FAC_OVERLAP_NON_OVERLAPPING);
        std::vector<Match> keep;
        for (auto &m : v) if (m.start < w.commit) keep.push_back(m);
        std::stable_sort(keep.begin(), keep.end(), [](const Match &a, const Match &b) { return a.start != b.start ? a.start < b.start : a.end < b.end; });
        for (auto &m : keep) {  // ReplaceCursor::emit_window, stream.rs:654-704
            uint64_t ms = w.base + m.start;
            if (ms < emitted) continue;
            if (emitted < ms) { size_t lo = (size_t)(emitted - w.base); out.append(w.text, lo, m.start - lo); }
            fac_match c; memset(&c, 0, sizeof(c));
            c.start = w.base + m.start; c.end = w.base + m.end; c.pattern_index = (uint32_t)m.pattern_index; c.similarity = m.similarity;
            c.insertions = m.ins; c.deletions = m.del; c.substitutions = m.sub; c.swaps = m.swp; c.edits = m.edits;
            const uint8_t *rp = nullptr; size_t rl = 0;
            if (cb && cb(cb_user, &c, w.base, (const uint8_t *)w.text.data() + m.start, m.end - m.start, &rp, &rl)) out.append((const char *)rp, rl);
            else out.append(w.text, m.start, m.end - m.start);
            emitted = w.base + m.end;
        }
        uint64_t commit_abs = w.base + w.commit;
        if (emitted < commit_abs) { size_t lo = (size_t)(emitted - w.base); out.append(w.text, lo, w.commit - lo); emitted = commit_abs; }
    }
    *out_buf = (uint8_t *)malloc(out.size() ? out.size() : 1);
    memcpy(*out_buf, out.data(), out.size());
    *out_len = out.size();
    return 0;
}
void orc_free_buf(uint8_t *p) { free(p); }

const fac_match *orc_matches_data(orc_matches *m) { return m->v.data(); }
size_t orc_matches_len(orc_matches *m) { return m->v.size(); }
uint64_t orc_matches_states_pushed(orc_matches *m) { return m->states; }
void orc_matches_free(orc_matches *m) { delete m; }

// Unicode helpers exposed for table-pinning tests.
int orc_grapheme_starts(const uint8_t *s, size_t n, uint64_t *out, size_t cap) {
    std::vector<size_t> st; grapheme_starts(s, n, st);
    for (size_t i = 0; i < st.size() && i < cap; i++) out[i] = st[i];
    return (int)st.size();
}
size_t orc_to_lowercase(const uint8_t *s, size_t n, uint8_t *out, size_t cap) {
    std::string r = to_lowercase(s, n);
    memcpy(out, r.data(), std::min(cap, r.size()));
    return r.size();
}
size_t orc_utf8_valid_up_to(const uint8_t *s, size_t n) { return utf8_valid_up_to(s, n); }

}  // extern "C"

// ---- multi-core CPU baseline helper (bench.py only) ------------------------------------------------
// Whole-input raw search split over `threads` host threads: contiguous shards of start positions,
// each searched with a right halo of max_match_graphemes()+1 graphemes and ownership by start
// offset -- the same decomposition rule the reference's streaming API uses (stream.rs:262-297),
// applied to an in-memory ASCII input so that small samples still use every core.
extern "C" int orc_search_parallel(void *e, const uint8_t *hay, size_t len, float thr, int threads, orc_matches **out) {
    Engine *E = (Engine *)e;
    for (size_t i = 0; i < len; i++) if (hay[i] >= 0x80) return -1;  // ASCII only (byte == grapheme)
    const size_t nt = (size_t)std::max(1, threads);
    const size_t halo = max_match_graphemes(*E) + 1;
    std::vector<std::vector<Match>> res(nt);
    std::vector<uint64_t> st(nt, 0);
    auto work = [&](size_t t) {
        const size_t a = len * t / nt, b = len * (t + 1) / nt;
        if (a >= b) return;
        const size_t end = std::min(len, b + halo);
        std::vector<Match> v;
        search_raw(*E, hay + a, end - a, thr, v, &st[t]);
        for (auto &m : v) if (m.start < b - a) { m.start += a; m.end += a; res[t].push_back(m); }
    };
    std::vector<std::thread> th;
    for (size_t t = 0; t < nt; t++) th.emplace_back(work, t);
    for (auto &t : th) t.join();
    orc_matches *o = new orc_matches();
    for (size_t t = 0; t < nt; t++) { to_c(res[t], o); o->states += st[t]; }
    *out = o;
    return 0;
}
