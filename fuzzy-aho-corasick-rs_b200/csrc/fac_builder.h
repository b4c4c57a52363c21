// fac_builder.h -- host-side automaton construction and flattening.
//
// Restates FuzzyAhoCorasickBuilder::build (src/builder.rs:181-484) for the B200 layout: the trie,
// the output merge along fail links, derived limits, prune coefficients and mapping transitions
// are computed on the host (build is not on the timed path) and emitted directly as the CSR /
// dense arrays of fac_types.h, ready for one bulk upload into HBM.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/fac.h"
#include "fac_types.h"
#include "fac_unicode.h"

namespace fac {

const UnicodeTables &host_unicode_tables();

// `graphemes(true)` + optional per-grapheme `to_lowercase` of a UTF-8 string.
std::vector<std::string> fold_graphemes(const std::string &s, bool case_insensitive);
size_t count_graphemes(const std::string &s);
// Strict UTF-8 validation (what `str::from_utf8` accepts): returns valid_up_to.
size_t utf8_valid_up_to(const uint8_t *s, size_t n);

struct HostPattern {
    std::string text;
    uint32_t glen = 0;     // graphemes of the original text (src/structs.rs:664)
    uint32_t byte_len = 0; // Pattern::len() (src/structs.rs:628-630)
    float weight = 1.f;
    bool has_limits = false;
    FacLimits limits{};    // finalized
    int64_t unique_id = -1;
};

// One symbol of the grapheme-id table (engines with mappings): folded grapheme text -> id >= 1.
struct HostSymbol {  // layout must equal FacSymbol (fac_unicode.h)
    uint64_t hash;
    uint32_t gid;
    uint32_t pool_off;
    uint32_t len;
    uint32_t pad;
};

// Bitap pre-filter model (BitapFilter, src/prefilter.rs:66-93 / build :161-245).
struct HostBitap {
    bool active = false;
    float edit_cost_mult = 0.f;
    uint32_t alphabet = 0;                 // symbol ids are 1..=alphabet, 0 = other
    uint8_t ascii_id[128] = {0};
    std::vector<HostSymbol> symbols;       // folded grapheme -> id (hash table content, ids <= 255)
    std::vector<uint8_t> symbol_pool;
    std::vector<uint32_t> m;               // per pattern: length in graphemes
    std::vector<float> weight;             // per pattern
    std::vector<int32_t> k_limit;          // per pattern: -1 = none
    std::vector<uint64_t> masks;           // [P * (alphabet+1)]
};

// Succinct BFS-ordered trie for the fast expansion kernel (fac_succinct.h).  `ok` is false when the
// engine is outside that kernel's domain (mappings, multi-byte edges, > 63 symbols, edit bounds above 6).
struct HostSuccinct {
    bool ok = false;
    // engines without any FuzzyLimits: the reference still expands one substitution per start window
    // (src/search.rs:143-145) but rejects every output whose edit counts are not all zero (:166-168), so the
    // results are exactly the outputs along the exact chain from the root
    bool exact_only = false;
    // non-ASCII haystacks (grapheme stream of first chars from K1) are inside the kernel's domain when no similarity
    // entry involves a non-ASCII char
    bool unicode_text_ok = false;
    // per-pattern / per-type limits (the reference's generic MAX_EDITS_FAST = 255 path): permissions are evaluated
    // per state from node_lim; edit_bound = an upper bound on the edits any state can accumulate
    bool limits_mode = false;
    uint32_t edit_bound = 0;
    std::vector<uint32_t> node_lim;      // [N] BFS order: index into HostAutomaton::lim or FAC_NONE
    uint32_t n_syms = 0;
    bool wide = false;                   // 32..63 symbols: 64-bit child bitmaps (SuccW<true>), else 32-bit
    uint8_t sym_of[256];                 // folded text byte -> dense symbol, NOSYM (31 / 63) = not in the alphabet
    std::vector<uint64_t> bm;            // [N] child bitmap over symbols
    std::vector<uint32_t> fc;            // [N] first child (children are contiguous in BFS order, sorted by symbol)
    std::vector<uint8_t> insym;          // [N] symbol of the edge leading into the node
    std::vector<float> prune_len, prune_low;  // [N] in BFS numbering
    std::vector<uint32_t> out_idx;       // [N] first entry in out2 or FAC_NONE
    std::vector<uint32_t> out2;          // 4 words per entry: pat | last << 31, glen f32 bits, weight f32 bits, limits index
    std::vector<float> sub_pen;          // [ROW][SUCC_SP_STRIDE] pen_sub * (1 - sim(edge char, text first char)), +inf below min_symbol_similarity
    std::vector<uint32_t> old_of;        // [N] reference node index of BFS node i
    uint64_t first_mask = 0, second_mask = 0;  // 2-gram window skip (search.rs:504-521) in symbol space
    // grandchild masks of the first gm_nodes BFS nodes (fac_succinct.h): [gm_nodes * ROW]
    std::vector<uint64_t> gmask;
    uint32_t gm_nodes = 0;
    // two-deep masks of the first gm2_nodes BFS nodes: [gm2_nodes * ROW * ROW]
    std::vector<uint64_t> gmask2;
    uint32_t gm2_nodes = 0;
    // narrow layout only (fac_succinct.h, "deep tables"): three-deep survivor masks and productivity masks over the
    // compact symbol row r3 = n_syms + 1 (index n_syms = "no symbol")
    uint32_t r3 = 0;
    uint32_t n3 = 0;                     // first n3 BFS nodes: gmask3 / pmask3 [n3][r3][r3][r3]
    uint32_t np2 = 0;                    // first np2 BFS nodes (>= n3): pmask2 [np2][r3][r3] (rows of the first n3 nodes unused)
    uint32_t n4 = 0;                     // first n4 BFS nodes (<= n3): pmask4 [n4][r3][r3][r3][r3]
    std::vector<uint32_t> gmask3, pmask3, pmask2, pmask4;
};

struct HostAutomaton {
    // flattened arrays, exactly the members of AutomatonView
    std::vector<uint32_t> node_edge_off, node_out_off, node_bitmap, node_lim, node_map_off;
    std::vector<float> node_prune_len, node_prune_low;
    std::vector<uint32_t> edge_char, edge_next, edge_sym;
    std::vector<FacTrans> trans;
    std::vector<uint32_t> out_pat;
    std::vector<float> pat_glen, pat_weight;
    std::vector<uint32_t> pat_lim, pat_bytes;
    std::vector<int64_t> pat_uid;  // UniqueId ordering key: (custom? 1:0)<<62 | id
    std::vector<FacLimits> lim;
    std::vector<float> sim_ascii;
    std::vector<uint64_t> sim_keys;
    std::vector<float> sim_vals;
    std::vector<uint32_t> map_hay_off, map_hay_gid, map_next;
    std::vector<float> map_pen;
    std::vector<uint32_t> ascii_gid;
    std::vector<HostSymbol> symbols;  // open-addressing table (size power of two), gid 0 = empty
    std::vector<uint8_t> symbol_pool;
    // scalars
    float pen_sub, pen_ins, pen_del, pen_swap, min_sym;
    int32_t mef = 255;              // after dispatch (1..6 or 255)
    int32_t max_edits_fast_raw = 0; // the reference's field (0, 1.., 255)
    bool has_mappings = false, has_pattern_limits = false, has_global_limits = false, ci = false;
    bool wskip = false;
    uint32_t ws_first[4] = {0, 0, 0, 0}, ws_second[4] = {0, 0, 0, 0};
    uint64_t beam_width = 0;  // 0 = None
    bool has_auto_beam = false;
    uint64_t ab_budget = 0, ab_width = 0;
    size_t max_match_graphemes = 0;  // src/stream.rs:213-253
    uint32_t max_map_hay = 1;
    std::vector<HostPattern> patterns;
    HostBitap bitap;
    HostSuccinct succ;
    // merged 16-byte records of the general stack-machine kernel (fac_flat.h); flat_ok = the packed fields fit
    bool flat_ok = false;
    std::vector<uint32_t> flat_nrec, flat_erec;   // 4 words per node / edge (ceilings are filled per call)
    std::vector<uint32_t> flat_ooff, flat_olist;  // per node: node-relative edges whose child has outputs (build order)
    std::vector<uint32_t> flat_gm_row;            // [N] survivor-mask row of a branching node (3..64 edges) or FAC_NONE
    std::vector<uint64_t> flat_pm;                // root productivity masks [(a * flat_pm_g + b) * flat_pm_words + w] (fac_flat.h); flat_pm_g == 0: none
    uint32_t flat_pm_g = 0, flat_pm_words = 0, flat_pm_k = 0;   // k = symbols per row (2 or 3)
    std::vector<uint32_t> flat_px_row;            // [N] row of the exact-child table or FAC_NONE (empty: no table)
    std::vector<uint64_t> flat_px;                // [rows][ceil(g * g / 64)] bit (b * g + c3)
    std::vector<uint64_t> flat_gm;                // [rows * 128] per ASCII look-ahead char: edges whose child has an output or that byte edge

    uint32_t n_nodes() const { return (uint32_t)node_prune_len.size(); }
    // View over the host vectors (used by the CPU-side emulator in tests).
    AutomatonView host_view() const;
};

// Returns FAC_OK or an error status with `err` filled.
fac_status build_automaton(const fac_config *cfg, const fac_pattern *pats, size_t n, HostAutomaton &out, std::string &err);

}  // namespace fac
