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
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "../../fuzzy-aho-corasick-rs_b200/csrc/fac_builder.h"
#include "../../fuzzy-aho-corasick-rs_b200/csrc/fac_core.h"
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

extern "C" {

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
                fac_make_ctx(A, T, maxpen, start, text_end, S, C);
                for (uint32_t s = 0; s < C.nslots; s++) {
                    FacState child;
                    if (fac_eval_slot(A, T, maxpen, start, text_end, C, s, child)) {
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

void emu_free(void *p) { free(p); }

}  // extern "C"
