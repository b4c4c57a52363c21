/*
 * fac.h -- C ABI of the B200-native fuzzy Aho-Corasick search engine (libfacgpu.so).
 *
 * The reference (kakserpom/fuzzy-aho-corasick-rs v0.5.0) is a pure-Rust rlib with no FFI of
 * its own; its drop-in boundary is the public Rust API.  Each entry point below is what a
 * `fuzzy-aho-corasick-gpu-sys` crate (`extern "C"` + build.rs nvcc step) binds underneath that
 * unchanged API; the citation on each declaration is the reference interface it replaces
 * (paths under /root/reference).  See INTEGRATION.md for the Rust-side binding.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ / torch types.
 *   - every function returns a fac_status (0 = FAC_OK); on failure
 *     fac_last_error_string() returns a thread-local, human-readable description.
 *   - there is NO CPU fallback: if no CUDA device is usable, creation fails with
 *     FAC_CUDA_ERROR.
 *   - an engine handle is immutable after creation and may be searched from many host
 *     threads concurrently (src/structs.rs:522-529, src/stream.rs:395-402).
 */
#ifndef FAC_H_
#define FAC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FAC_ABI_VERSION 2

typedef enum fac_status {
    FAC_OK = 0,
    FAC_HAYSTACK_TOO_LARGE = 1, /* SearchError::HaystackTooLarge, src/error.rs:9-17 */
    FAC_INVALID_UTF8 = 2,       /* Rust `&str` guarantees validity; C callers do not */
    FAC_CUDA_ERROR = 3,
    FAC_OOM = 4,
    FAC_INVALID_ARGUMENT = 5,
    FAC_UNSUPPORTED = 6,        /* a size outside the device layout's fixed-width fields */
    FAC_IO_ERROR = 7            /* reader / writer callback failed (stream API) */
} fac_status;

/* Order / Overlap: src/options.rs:10-36 */
typedef enum fac_order {
    FAC_ORDER_UNSORTED = 0,
    FAC_ORDER_DEFAULT = 1,
    FAC_ORDER_GREEDY = 2,
    FAC_ORDER_COVERAGE_WEIGHTED = 3
} fac_order;

typedef enum fac_overlap {
    FAC_OVERLAP_KEEP = 0,
    FAC_OVERLAP_NON_OVERLAPPING = 1,
    FAC_OVERLAP_NON_OVERLAPPING_UNIQUE = 2
} fac_overlap;

/* FuzzyLimits (src/structs.rs:292-363): each field is Option<u8>; -1 encodes None.
 * The library applies `finalize()` (src/structs.rs:319-335) itself, exactly where the
 * reference does (builder.fuzzy(), Pattern::fuzzy()). */
typedef struct fac_limits {
    int16_t insertions;
    int16_t deletions;
    int16_t substitutions;
    int16_t swaps;
    int16_t edits;
} fac_limits;

/* Pattern (src/structs.rs:597-610).  `text` need not be NUL-terminated. */
typedef struct fac_pattern {
    const char *text;
    size_t len;
    float weight;        /* default 1.0 */
    int32_t has_limits;  /* Pattern::fuzzy() was called */
    fac_limits limits;
    int64_t unique_id;   /* custom_unique_id; -1 = automatic (pattern index) */
} fac_pattern;

/* One `(char, char) -> f32` entry of Similarity::from_map (src/structs.rs:30-54). */
typedef struct fac_sim_pair {
    uint32_t a; /* pattern-side scalar value */
    uint32_t b; /* haystack-side scalar value */
    float similarity;
} fac_sim_pair;

/* builder.mapping_scored(a, b, score) (src/builder.rs:116-132). */
typedef struct fac_mapping {
    const char *a;
    size_t a_len;
    const char *b;
    size_t b_len;
    float score;
} fac_mapping;

/* FuzzyAhoCorasickBuilder (src/builder.rs:22-143). Zero-initialise, then set fields. */
typedef struct fac_config {
    int32_t case_insensitive;
    int32_t has_limits; /* builder.fuzzy() was called */
    fac_limits limits;
    int32_t has_penalties; /* 0 = FuzzyPenalties::default() (src/structs.rs:381-393) */
    float penalty_insertion;
    float penalty_deletion;
    float penalty_substitution;
    float penalty_swap;
    uint64_t beam_width; /* 0 = None */
    int32_t has_auto_beam;
    uint64_t auto_beam_budget;
    uint64_t auto_beam_width;
    float min_symbol_similarity;
    int32_t has_similarity; /* 0 = DEFAULT_SIMILARITY (src/builder.rs:492-526) */
    const fac_sim_pair *similarity;
    size_t n_similarity;
    const fac_mapping *mappings;
    size_t n_mappings;
} fac_config;

/* FuzzyMatch (src/structs.rs:757-781) without the borrows: `text` is
 * haystack[start..end] and `pattern` is patterns[pattern_index]; the Rust shim rebuilds
 * both from these offsets.  32 bytes. */
typedef struct fac_match {
    uint64_t start; /* inclusive byte offset */
    uint64_t end;   /* exclusive byte offset */
    uint32_t pattern_index;
    float similarity;
    uint8_t insertions;
    uint8_t deletions;
    uint8_t substitutions;
    uint8_t swaps;
    uint8_t edits;
    uint8_t pad_[3];
} fac_match;

typedef struct fac_engine fac_engine;
typedef struct fac_matches fac_matches;

/* Thread-local description of the last failure on this thread ("" if none). */
const char *fac_last_error_string(void);
int fac_abi_version(void);
/* SHA-256 (hex) of the library's source files at build time; the loaders compare it with the sources in the tree
 * and refuse a stale binary. */
const char *fac_build_source_hash(void);

/* FuzzyAhoCorasickBuilder::build (src/builder.rs:181-484).  Builds the trie on the host,
 * flattens it and uploads the automaton to CUDA device `device` (fac_engine_create uses
 * the current device). */
fac_status fac_engine_create(const fac_config *cfg, const fac_pattern *patterns, size_t n_patterns,
                             fac_engine **out);
fac_status fac_engine_create_on(int device, const fac_config *cfg, const fac_pattern *patterns,
                                size_t n_patterns, fac_engine **out);
/* One engine replicated on several CUDA devices (SURVEY 8b `fac_engine_create_on(devices[], n)`, 8e "windows
 * round-robin over GPUs"): the automaton is built once on the host and uploaded to every device.  The stream entry
 * points deal their window batches round-robin over the devices (the GPU analogue of search_stream_parallel /
 * replace_stream_parallel, src/stream.rs:378-429, 533-638) and reassemble the results in stream order; the
 * whole-haystack entry points use devices[0].  fac_engine_free releases every replica. */
fac_status fac_engine_create_multi(const int *devices, size_t n_devices, const fac_config *cfg,
                                   const fac_pattern *patterns, size_t n_patterns, fac_engine **out);
void fac_engine_free(fac_engine *engine);

/* FuzzyAhoCorasick::max_match_graphemes (src/stream.rs:213-253). */
size_t fac_engine_max_match_graphemes(const fac_engine *engine);
/* Prefiltered::is_active (src/prefilter.rs:125-127). */
int fac_engine_prefilter_active(const fac_engine *engine);
/* Introspection used by tests / the bench. */
size_t fac_engine_num_nodes(const fac_engine *engine);
size_t fac_engine_num_patterns(const fac_engine *engine);
int fac_engine_device(const fac_engine *engine);
size_t fac_engine_num_devices(const fac_engine *engine);

/* FuzzyAhoCorasick::search (src/query.rs:30-38) and Prefiltered::search
 * (src/prefilter.rs:135-143) when use_prefilter != 0.  `haystack` is host memory holding
 * UTF-8.  On FAC_HAYSTACK_TOO_LARGE, *graphemes_out (if non-NULL) receives the count. */
fac_status fac_search(const fac_engine *engine, const uint8_t *haystack, size_t len, float threshold,
                      fac_order order, fac_overlap overlap, int use_prefilter, fac_matches **out);
/* Same call with the haystack already resident in device memory on the engine's device
 * (used to time the kernels without the host->device copy). */
fac_status fac_search_device(const fac_engine *engine, const uint8_t *d_haystack, size_t len,
                             float threshold, fac_order order, fac_overlap overlap, int use_prefilter,
                             fac_matches **out);
/* The grapheme count reported by the last FAC_HAYSTACK_TOO_LARGE on this thread. */
uint64_t fac_last_haystack_graphemes(void);

/* Shard-with-halo search: report only matches whose start byte lies in
 * [own_begin, own_end) of `haystack`; bytes after own_end are the halo.  Offsets in the
 * result are relative to `haystack` + base.  Order/overlap are not applied (the caller
 * gathers shards and applies them globally with fac_matches_apply). */
fac_status fac_search_shard(const fac_engine *engine, const uint8_t *haystack, size_t len,
                            size_t own_begin, size_t own_end, uint64_t base, float threshold,
                            int on_device, fac_matches **out);

/* FuzzyMatches::apply (src/matches.rs:7-22) on a caller-assembled list (e.g. the gathered
 * shards): sorts + overlap resolution run on the engine's device. */
fac_status fac_matches_apply(const fac_engine *engine, const fac_match *in, size_t n, fac_order order,
                             fac_overlap overlap, fac_matches **out);

/* ---- multi-GPU sharding of ONE haystack (ownership rule of src/stream.rs:262-297) ---- */

/* Flags of fac_search_args.flags / fac_matches_apply_device. */
#define FAC_HAYSTACK_ON_DEVICE 1u /* `haystack` is device memory on the engine's device */
#define FAC_RESULT_ON_DEVICE 2u   /* keep the match list in device memory (fac_matches_device_data); no D2H copy */
#define FAC_TEXT_IS_UNICODE 4u    /* the slice belongs to a non-ASCII haystack: use the Unicode grapheme storage even
                                     when the slice itself is ASCII (`\r\n` is one grapheme there, src/search.rs:196,
                                     src/grapheme.rs:92-98) */
#define FAC_APPLY_PRESORTED 8u    /* fac_matches_apply_device: the input already is in `order` (e.g. Order::Unsorted
                                     shard lists concatenated in shard order) */

/* One shard of a haystack: the rank searches haystack[own_begin, read_end) and owns the matches that start in
 * [own_begin, own_end).  Cuts lie on extended-grapheme-cluster boundaries; read_end - own_end covers
 * max_match_graphemes() + 3 clusters (the reference's streaming overlap, src/stream.rs:213-258, plus the
 * look-ahead of the dead-end filter). */
typedef struct fac_shard {
    size_t own_begin;
    size_t own_end;
    size_t read_end;
} fac_shard;

/* Cut `haystack` (host memory, valid UTF-8; NULL = the caller guarantees an ASCII haystack of `len` bytes)
 * into n_shards contiguous shards of about equal byte length.  Pure host function (no device needed):
 * `max_match_graphemes` is fac_engine_max_match_graphemes() of the engine that will search the shards. */
fac_status fac_plan_shards(size_t max_match_graphemes, const uint8_t *haystack, size_t len, size_t n_shards,
                           fac_shard *out);

/* The general search entry point: engine.search / Prefiltered::search restricted to the start positions in
 * [own_begin, own_end) of `haystack` (own_end > len means len), offsets reported as `base` + slice offset.
 * `order` is applied to the shard's own list (with overlap == FAC_OVERLAP_KEEP a global ranking is then a merge of
 * the shard lists; for Order::Unsorted -- ascending (start, end, pattern) -- it is their concatenation); overlap
 * modes other than KEEP are only meaningful when the owned range is the whole haystack. */
typedef struct fac_search_args {
    const uint8_t *haystack;
    size_t len;
    size_t own_begin;
    size_t own_end;
    uint64_t base;
    float threshold;
    fac_order order;
    fac_overlap overlap;
    int32_t use_prefilter;
    uint32_t flags;
} fac_search_args;
fac_status fac_search_ex(const fac_engine *engine, const fac_search_args *args, fac_matches **out);

/* FuzzyMatches::apply on a device-resident list of fac_match records (the NCCL-gathered shard lists on rank 0). */
fac_status fac_matches_apply_device(const fac_engine *engine, const fac_match *d_in, size_t n, fac_order order,
                                    fac_overlap overlap, uint32_t flags, fac_matches **out);
/* Device pointer of a list produced with FAC_RESULT_ON_DEVICE (NULL otherwise); valid until fac_matches_free. */
const fac_match *fac_matches_device_data(const fac_matches *m);

/* One reader-cut streaming window (StreamWindow, src/stream.rs:67-73). */
typedef struct fac_window {
    const uint8_t *text; /* valid UTF-8 */
    size_t len;
    uint64_t base;   /* absolute offset of text[0] */
    size_t commit;   /* the window owns matches with start < commit */
} fac_window;

/* window_matches (src/stream.rs:262-297) over a batch of windows: each window is searched
 * with threshold/sorted()/non_overlapping(), filtered to start < commit and rebased by
 * `base`; results are concatenated in window order. */
fac_status fac_search_windows(const fac_engine *engine, const fac_window *windows, size_t n_windows,
                              float threshold, fac_matches **out);

const fac_match *fac_matches_data(const fac_matches *m);
size_t fac_matches_len(const fac_matches *m);
/* Number of search states of the call.  On the order-faithful kernels (engines with a beam,
 * mappings, ... or FAC_FAITHFUL=1) this is exactly the sum over start windows of the
 * reference's `queue.len()` (src/search.rs:1099); on the fast kernel it is the number of states
 * that kernel visited, which is smaller (it skips children that cannot emit).  A statistic for
 * the states/s figure, never part of the result. */
uint64_t fac_matches_states_pushed(const fac_matches *m);
/* Device time (ms, CUDA events on the library's stream) of the whole call and of the
 * frontier-expansion kernel launches inside it, and the number of kernel launches. */
double fac_matches_device_ms(const fac_matches *m);
double fac_matches_expand_ms(const fac_matches *m);
uint32_t fac_matches_kernel_launches(const fac_matches *m);
void fac_matches_free(fac_matches *m);

/* ---- streaming (src/stream.rs:319-638; FuzzyReplacer::replace_stream src/replacer.rs:35) ---- */

/* Read up to `cap` bytes into `buf`; return the count, 0 at EOF, <0 on error (io::Read). */
typedef int64_t (*fac_read_fn)(void *user, uint8_t *buf, size_t cap);
/* Write exactly `len` bytes; return 0 on success (io::Write::write_all). */
typedef int (*fac_write_fn)(void *user, const uint8_t *buf, size_t len);
/* Called once per match in stream order with absolute offsets (StreamMatch). */
typedef void (*fac_match_fn)(void *user, const fac_match *m);

/* search_stream / search_stream_parallel (src/stream.rs:319-429): windows are cut exactly
 * as WindowReader does (src/stream.rs:102-158) and searched in batches on the device.
 * *bytes_read receives the total read from the reader. */
fac_status fac_search_stream(const fac_engine *engine, fac_read_fn read, void *read_user, float threshold,
                             fac_match_fn on_match, void *match_user, uint64_t *bytes_read);

/* Replacement callback of replace_stream (src/stream.rs:465-480): `m` carries ABSOLUTE offsets
 * (the reference's FuzzyMatch is window-relative: subtract `window_base`), `text` is the matched
 * slice.  Return 1 and set *repl / *repl_len to substitute, 0 to keep the original text. */
typedef int (*fac_replace_fn)(void *user, const fac_match *m, uint64_t window_base, const uint8_t *text,
                              size_t text_len, const uint8_t **repl, size_t *repl_len);

/* replace_stream / replace_stream_parallel (src/stream.rs:465-638) and
 * FuzzyReplacer::replace_stream (src/replacer.rs:35-46).  Windows are searched in batches on the
 * device; output is assembled in stream order by ReplaceCursor::emit_window semantics
 * (src/stream.rs:654-704).  *bytes_written receives the total written. */
fac_status fac_replace_stream(const fac_engine *engine, fac_read_fn read, void *read_user,
                              fac_write_fn write, void *write_user, float threshold,
                              fac_replace_fn replace, void *replace_user, uint64_t *bytes_written);

/* Counters of one stream call (all optional outputs). */
typedef struct fac_stream_stats {
    uint64_t bytes_read;
    uint64_t bytes_written;
    uint64_t windows;       /* reader-cut windows searched */
    uint64_t matches;       /* matches owned by their windows */
    uint64_t states;        /* search states (see fac_matches_states_pushed) */
    double device_ms;       /* sum over window batches, CUDA events on the library's streams */
    double expand_ms;
    uint32_t kernel_launches;
    uint32_t devices;       /* devices the windows were spread over */
} fac_stream_stats;

/* FuzzyReplacer::replace_stream (src/replacer.rs:35-46): replace_stream with the replacement taken from a table indexed
 * by pattern index (entries beyond n_replacements, or NULL entries, keep the matched text).  `stats` may be NULL. */
fac_status fac_replace_stream_table(const fac_engine *engine, fac_read_fn read, void *read_user, fac_write_fn write,
                                    void *write_user, float threshold, const uint8_t *const *replacements,
                                    const size_t *replacement_lens, size_t n_replacements, fac_stream_stats *stats);
/* search_stream with the counters of the call (on_match may be NULL: count only). */
fac_status fac_search_stream_stats(const fac_engine *engine, fac_read_fn read, void *read_user, float threshold,
                                   fac_match_fn on_match, void *match_user, fac_stream_stats *stats);

#ifdef __cplusplus
}
#endif
#endif /* FAC_H_ */
