// fac_types.h -- flattened automaton layout shared by the host builder and the CUDA kernels.
//
// The reference keeps the automaton as Vec<Node> with per-node Vec<Edge>, Vec<u32> outputs and a
// FxHashMap<String,u32> (src/structs.rs:248-281).  Here the builder flattens it into CSR / dense
// arrays that live in HBM for the lifetime of the engine; `AutomatonView` is the POD bundle of raw
// pointers that kernels (and the host-side emulator used by the CPU tests) read.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FAC_HD __host__ __device__ __forceinline__
#else
#define FAC_HD inline
#endif

#define FAC_NONE 0xFFFFFFFFu

// FuzzyLimits with -1 == None (src/structs.rs:292-299), padded to 16 B.
struct FacLimits {
    int16_t ins, del, sub, swp, edits;
    int16_t pad[3];
};

// One slot of the (node, symbol) -> next transition table (open addressing, linear probing).
struct FacTrans {
    uint32_t node;  // FAC_NONE == empty
    uint32_t sym;   // first char (no-mappings engines) or grapheme id (engines with mappings)
    uint32_t next;
    uint32_t pad;
};

struct AutomatonView {
    uint32_t n_nodes, n_edges, n_patterns, n_outputs;
    // ---- nodes (index = reference node index) ----
    const uint32_t *node_edge_off;  // [N+1] CSR into edge_* (build order, src/builder.rs:336-342)
    const float *node_prune_len;    // [N]   Node::prune_len
    const float *node_prune_low;    // [N]   Node::prune_len_over_weight
    const uint32_t *node_out_off;   // [N+1] CSR into out_pat
    const uint32_t *node_bitmap;    // [N*4] bit c set iff an edge is the single ASCII byte c
    const uint32_t *node_lim;       // [N]   limits index of patterns[node.pattern_index] or FAC_NONE
    const uint32_t *node_map_off;   // [N+1] CSR into map_* (only when has_mappings)
    // ---- edges ----
    const uint32_t *edge_char;  // [E] Edge::first_char
    const uint32_t *edge_next;  // [E] bit31 = child has a non-empty output; low 31 bits = Edge::next()
    const uint32_t *edge_sym;   // [E] exact-match symbol of the edge: first char (no mappings) or grapheme id (mappings)
    // ---- exact transition lookup ----
    const FacTrans *trans;  // [trans_mask+1]
    uint32_t trans_mask;
    // ---- outputs / patterns ----
    const uint32_t *out_pat;   // [O]
    const float *pat_glen;     // [P] grapheme_len as f32
    const float *pat_weight;   // [P]
    const uint32_t *pat_lim;   // [P] limits index of the pattern's own limits or FAC_NONE
    const FacLimits *lim;      // [L] index 0 = global limits (valid iff has_global_limits)
    // ---- similarity ----
    const float *sim_ascii;     // [128*128] Similarity::ascii_table
    const uint64_t *sim_keys;   // [n_sim] sorted (a<<32|b) for pairs with a non-ASCII member
    const float *sim_vals;      // [n_sim]
    uint32_t n_sim;
    // ---- multi-character mappings (MappingTransition, src/structs.rs:234-242) ----
    const uint32_t *map_hay_off;  // [M+1] CSR into map_hay_gid
    const uint32_t *map_hay_gid;  // grapheme ids of the haystack side
    const uint32_t *map_next;     // [M]
    const float *map_pen;         // [M]
    const uint32_t *ascii_gid;    // [128] grapheme id of each (already folded) ASCII byte, 0 = unknown
    // ---- scalars ----
    float pen_sub, pen_ins, pen_del, pen_swap, min_sym;
    int32_t mef;                 // MAX_EDITS_FAST after the dispatch of src/search.rs:205-247: 1..6 or 255
    int32_t has_mappings;        // MAPPINGS const generic (per-node map non-empty, src/search.rs:204)
    int32_t has_pattern_limits;  // src/structs.rs:544
    int32_t has_global_limits;   // self.limits.is_some()
    int32_t ci;                  // case_insensitive
    // window skip (src/search.rs:504-521): valid iff wskip != 0
    int32_t wskip;
    uint32_t ws_first[4], ws_second[4];
};

// Haystack as the kernels see it.  ASCII haystacks are searched straight from the bytes
// (AsciiGraphemes, src/grapheme.rs:76-125); non-ASCII ones from the grapheme streams the
// segmentation kernel produced (build_unicode_graphemes, src/search.rs:398-416).
struct TextView {
    const uint8_t *bytes;   // ascii: the haystack bytes
    const uint32_t *first;  // unicode: first char of every folded grapheme (text_chars)
    const uint32_t *gid;    // unicode + mappings: grapheme id of every folded grapheme (0 = unknown)
    const uint32_t *off32;  // unicode: byte offset of every grapheme (+ sentinel) when len < 4 GiB
    const uint64_t *off64;  // unicode: same, when len >= 4 GiB
    uint64_t n_bytes;       // haystack length in bytes
    uint32_t n;             // grapheme count
    int32_t ascii;
};

// A search state (State, src/structs.rs:166-179) packed to 16 bytes.
//   pos = window-in-tile << 20 | (j - start) << 10 | (matched_end - start)
// matched_start always equals the window start (SURVEY invariant I1), `edits` is the byte sum of cnt.
struct FacState {
    uint32_t node;
    float pen;
    uint32_t cnt;  // packed_counts: ins | del<<8 | sub<<16 | swap<<24
    uint32_t pos;
};
#define FAC_POS_W_SHIFT 20
#define FAC_POS_J_SHIFT 10
#define FAC_POS_MASK 0x3FFu
#define FAC_MAX_SPAN 1000u   // max_match_graphemes()+2 must stay below this (10-bit fields)
#define FAC_MAX_TILE 4096u   // windows per tile (12-bit field)

// A raw output candidate (one `best.entry(key)` visit, src/search.rs:705-735).
struct FacCand {
    uint32_t sg;     // start grapheme (absolute in the call's grapheme stream)
    uint32_t eg;     // end grapheme (may equal text_end)
    uint32_t pat;
    float sim;
    uint32_t cnt;
    uint32_t seq;    // FIFO queue position of the emitting state within its tile
    uint32_t tile;   // tile index within the launch (failed tiles are superseded by the retry pass)
    uint32_t tag;    // window id | pass << 31
};

// One haystack window of a call: a whole-haystack search has exactly one; the streaming API
// (StreamWindow, src/stream.rs:67-73) and the pre-filter slices (src/prefilter.rs:346-350) have many.
struct FacWindow {
    uint64_t byte_begin, byte_end;  // in the device text buffer
    uint64_t base;                  // absolute offset of byte_begin in the caller's stream
    uint64_t commit;                // window-relative: the window owns matches with start < commit
    uint32_t g_begin, g_end;        // grapheme range in the call's grapheme stream
    uint32_t pad[2];
};

// Internal match record (window-relative byte offsets); converted to fac_match on the way out.
struct WMatch {
    uint64_t start, end;
    uint32_t pat;
    float sim;
    uint32_t cnt;  // ins | del<<8 | sub<<16 | swap<<24
    uint32_t win;
};
