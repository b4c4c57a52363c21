/* conformance.c -- plain-C conformance test of include/fac.h (no Python, no C++): what a maintainer binding the library
 * from Rust / C would call, in the order they would call it.  Built by tests/test_abi_c.py with
 *     gcc conformance.c facio.c -I include -L csrc -lfacgpu -lcudart
 * Usage: conformance [--no-device] [--big]
 *   --no-device : expect engine creation to fail loudly with FAC_CUDA_ERROR (CPU box; there is no CPU fallback)
 *   --big       : also run the > 4 GiB cases (FAC_HAYSTACK_TOO_LARGE on a 4 GiB + 1 byte device buffer; a 4.5 GiB stream
 *                 whose matches must carry absolute u64 offsets past 2^32, src/stream.rs:262-297)
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/fac.h"

/* from facio.c */
typedef struct facio_block_reader { const uint8_t *block; size_t block_len; uint64_t total, pos; size_t max_read; int64_t fail_at; } facio_block_reader;
typedef struct facio_sink { uint64_t bytes, fnv; uint8_t *keep; size_t keep_cap, keep_len; int64_t fail_at; } facio_sink;
typedef struct facio_match_stats { uint64_t count, max_start, hash, beyond_4g, last_start, order_violations; } facio_match_stats;
int64_t facio_block_read(void *user, uint8_t *buf, size_t cap);
void facio_sink_init(facio_sink *s, uint8_t *keep, size_t keep_cap);
int facio_sink_write(void *user, const uint8_t *buf, size_t len);
void facio_on_match(void *user, const fac_match *m);

/* the two CUDA runtime calls the --big case needs (declared here to stay plain C without cuda headers) */
extern int cudaMalloc(void **p, size_t n);
extern int cudaMemset(void *p, int v, size_t n);
extern int cudaFree(void *p);

static int failures = 0;
#define CHECK(cond, name)                                                                     \
    do {                                                                                      \
        if (cond) printf("OK   %s\n", name);                                                  \
        else { printf("FAIL %s (line %d): %s\n", name, __LINE__, fac_last_error_string()); failures++; } \
    } while (0)

static fac_limits no_limits(void) { fac_limits l = {-1, -1, -1, -1, -1}; return l; }

static fac_engine *make_engine(const char **pats, size_t n, int edits, int ci, uint64_t ab_budget, uint64_t ab_width) {
    fac_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.case_insensitive = ci;
    cfg.limits = no_limits();
    if (edits >= 0) { cfg.has_limits = 1; cfg.limits.edits = (int16_t)edits; }
    if (ab_budget) { cfg.has_auto_beam = 1; cfg.auto_beam_budget = ab_budget; cfg.auto_beam_width = ab_width; }
    fac_pattern *p = (fac_pattern *)calloc(n, sizeof(fac_pattern));
    for (size_t i = 0; i < n; i++) { p[i].text = pats[i]; p[i].len = strlen(pats[i]); p[i].weight = 1.0f; p[i].limits = no_limits(); p[i].unique_id = -1; }
    fac_engine *e = NULL;
    fac_status st = fac_engine_create(&cfg, p, n, &e);
    free(p);
    return st == FAC_OK ? e : NULL;
}

static int64_t failing_read(void *u, uint8_t *b, size_t c) { (void)u; (void)b; (void)c; return -1; }

int main(int argc, char **argv) {
    int no_device = 0, big = 0;
    for (int i = 1; i < argc; i++) { if (!strcmp(argv[i], "--no-device")) no_device = 1; if (!strcmp(argv[i], "--big")) big = 1; }
    CHECK(fac_abi_version() == FAC_ABI_VERSION, "abi version");
    CHECK(strlen(fac_build_source_hash()) >= 9, "source stamp");
    CHECK(sizeof(fac_match) == 32, "fac_match is 32 bytes");

    const char *pats[] = {"saddam", "hussein", "needle"};
    if (no_device) {
        fac_config cfg; memset(&cfg, 0, sizeof(cfg)); cfg.limits = no_limits();
        fac_pattern p; memset(&p, 0, sizeof(p)); p.text = "abc"; p.len = 3; p.weight = 1.0f; p.limits = no_limits(); p.unique_id = -1;
        fac_engine *e = NULL;
        fac_status st = fac_engine_create(&cfg, &p, 1, &e);
        CHECK(st == FAC_CUDA_ERROR && e == NULL && strstr(fac_last_error_string(), "no CPU fallback") != NULL, "no device: FAC_CUDA_ERROR, no fallback");
        /* the shard planner is a pure host function */
        fac_shard sh[3];
        CHECK(fac_plan_shards(4, NULL, 10, 3, sh) == FAC_OK && sh[0].own_end == 3 && sh[1].own_end == 6 && sh[2].own_end == 10 && sh[0].read_end == 10, "fac_plan_shards without a device");
        printf("%s\n", failures ? "FAILED" : "ALL OK");
        return failures ? 1 : 0;
    }

    fac_engine *e = make_engine(pats, 3, 2, 1, 0, 0);
    CHECK(e != NULL, "engine create");
    if (!e) return 1;
    CHECK(fac_engine_num_patterns(e) == 3 && fac_engine_max_match_graphemes(e) == 7 + 2, "engine introspection");

    /* search: tests.rs:187-207 style */
    const char *hay = "this is a saddamhu example with saddam and husein and a neeedle";
    fac_matches *m = NULL;
    fac_status st = fac_search(e, (const uint8_t *)hay, strlen(hay), 0.7f, FAC_ORDER_DEFAULT, FAC_OVERLAP_NON_OVERLAPPING, 0, &m);
    CHECK(st == FAC_OK && m && fac_matches_len(m) >= 3, "fac_search sorted non_overlapping");
    if (m) {
        const fac_match *d = fac_matches_data(m);
        int seen[3] = {0, 0, 0}, ordered = 1;
        for (size_t i = 0; i < fac_matches_len(m); i++) {
            if (d[i].pattern_index < 3) seen[d[i].pattern_index] = 1;
            if (i && d[i].start < d[i - 1].end) ordered = 0;
            if (!(d[i].similarity >= 0.7f && d[i].similarity <= 1.0f)) ordered = 0;
        }
        CHECK(seen[0] && seen[1] && seen[2] && ordered, "all three patterns found, disjoint, ascending");
        CHECK(fac_matches_kernel_launches(m) > 0 && fac_matches_device_ms(m) > 0.0, "device statistics");
        fac_matches_free(m);
    }
    /* errors */
    const uint8_t bad[] = {'a', 0xFF, 'b'};
    m = NULL;
    CHECK(fac_search(e, bad, 3, 0.8f, FAC_ORDER_UNSORTED, FAC_OVERLAP_KEEP, 0, &m) == FAC_INVALID_UTF8 && m == NULL, "FAC_INVALID_UTF8");
    CHECK(fac_search(e, NULL, 5, 0.8f, FAC_ORDER_UNSORTED, FAC_OVERLAP_KEEP, 0, &m) == FAC_INVALID_ARGUMENT, "FAC_INVALID_ARGUMENT (null haystack)");
    CHECK(fac_search(e, (const uint8_t *)"", 0, 0.8f, FAC_ORDER_UNSORTED, FAC_OVERLAP_KEEP, 0, &m) == FAC_OK && fac_matches_len(m) == 0, "empty haystack");
    fac_matches_free(m);

    /* shard + global apply == whole */
    {
        size_t n = strlen(hay);
        fac_shard sh[3];
        CHECK(fac_plan_shards(fac_engine_max_match_graphemes(e), (const uint8_t *)hay, n, 3, sh) == FAC_OK, "fac_plan_shards");
        fac_match all[256];
        size_t na = 0;
        for (int r = 0; r < 3; r++) {
            fac_search_args a;
            memset(&a, 0, sizeof(a));
            a.haystack = (const uint8_t *)hay + sh[r].own_begin; a.len = sh[r].read_end - sh[r].own_begin; a.own_begin = 0;
            a.own_end = sh[r].own_end - sh[r].own_begin; a.base = sh[r].own_begin; a.threshold = 0.7f; a.order = FAC_ORDER_UNSORTED; a.overlap = FAC_OVERLAP_KEEP;
            fac_matches *pm = NULL;
            if (fac_search_ex(e, &a, &pm) != FAC_OK) { failures++; printf("FAIL shard %d: %s\n", r, fac_last_error_string()); continue; }
            for (size_t i = 0; i < fac_matches_len(pm) && na < 256; i++) all[na++] = fac_matches_data(pm)[i];
            fac_matches_free(pm);
        }
        fac_matches *fin = NULL, *whole = NULL;
        st = fac_matches_apply(e, all, na, FAC_ORDER_DEFAULT, FAC_OVERLAP_NON_OVERLAPPING, &fin);
        fac_status st2 = fac_search(e, (const uint8_t *)hay, n, 0.7f, FAC_ORDER_DEFAULT, FAC_OVERLAP_NON_OVERLAPPING, 0, &whole);
        int same = st == FAC_OK && st2 == FAC_OK && fac_matches_len(fin) == fac_matches_len(whole) &&
                   memcmp(fac_matches_data(fin), fac_matches_data(whole), fac_matches_len(fin) * sizeof(fac_match)) == 0;
        CHECK(same, "3 shards + fac_matches_apply == whole search");
        fac_matches_free(fin); fac_matches_free(whole);
    }

    /* streams: search_stream, replace_stream_table (FuzzyReplacer), io errors */
    {
        static uint8_t block[1 << 16];
        memset(block, 'x', sizeof(block));
        for (size_t i = 0; i < sizeof(block); i += 64) block[i] = ' ';
        memcpy(block + 1000, " needle ", 8);
        memcpy(block + 40000, " neeedle ", 9);
        facio_block_reader rd = {block, sizeof(block), 20ull * sizeof(block), 0, 60000, -1};   /* 1.25 MiB: several 256 KiB windows */
        facio_match_stats ms;
        memset(&ms, 0, sizeof(ms));
        fac_stream_stats ss;
        st = fac_search_stream_stats(e, facio_block_read, &rd, 0.8f, facio_on_match, &ms, &ss);
        CHECK(st == FAC_OK && ss.bytes_read == rd.total && ms.count == 40 && ss.matches == 40 && ss.windows >= 4 && ms.order_violations == 0, "fac_search_stream_stats over 5 windows");
        const uint8_t *repl[3] = {(const uint8_t *)"S", (const uint8_t *)"H", (const uint8_t *)"<N>"};
        const size_t repl_len[3] = {1, 1, 3};
        rd.pos = 0;
        static uint8_t keep[1 << 12];
        facio_sink sink;
        facio_sink_init(&sink, keep, sizeof(keep));
        st = fac_replace_stream_table(e, facio_block_read, &rd, facio_sink_write, &sink, 0.8f, repl, repl_len, 3, &ss);
        /* "needle" (6) -> "<N>" (3) and "neeedle" (7) -> "<N>" per block */
        CHECK(st == FAC_OK && ss.bytes_written == sink.bytes && sink.bytes == rd.total - 20ull * (3 + 4) && memcmp(keep + 1000, " <N> ", 5) == 0, "fac_replace_stream_table (FuzzyReplacer::replace_stream)");
        uint64_t nread = 0;
        CHECK(fac_search_stream(e, failing_read, NULL, 0.8f, NULL, NULL, &nread) == FAC_IO_ERROR, "FAC_IO_ERROR (reader)");
        rd.pos = 0;
        facio_sink_init(&sink, NULL, 0);
        sink.fail_at = 100000;
        CHECK(fac_replace_stream_table(e, facio_block_read, &rd, facio_sink_write, &sink, 0.8f, repl, repl_len, 3, NULL) == FAC_IO_ERROR, "FAC_IO_ERROR (writer)");
    }

    /* auto_beam engines cannot be sharded (the budget is cumulative, src/search.rs:1096-1103) */
    {
        fac_engine *ab = make_engine(pats, 3, 2, 1, 100, 10);
        fac_search_args a;
        memset(&a, 0, sizeof(a));
        a.haystack = (const uint8_t *)hay; a.len = strlen(hay); a.own_begin = 0; a.own_end = 10; a.threshold = 0.8f;
        m = NULL;
        CHECK(ab && fac_search_ex(ab, &a, &m) == FAC_UNSUPPORTED, "FAC_UNSUPPORTED (shard of an auto_beam engine)");
        a.own_end = a.len;
        CHECK(ab && fac_search_ex(ab, &a, &m) == FAC_OK, "auto_beam engine on the whole haystack");
        fac_matches_free(m);
        fac_engine_free(ab);
    }

    if (big) {
        /* SearchError::HaystackTooLarge (src/search.rs:198-202): 4 GiB + 1 zero bytes, device resident */
        void *d = NULL;
        const size_t n = (1ull << 32) + 1;
        if (cudaMalloc(&d, n) == 0 && cudaMemset(d, 'a', n) == 0) {
            m = NULL;
            st = fac_search_device(e, (const uint8_t *)d, n, 0.8f, FAC_ORDER_UNSORTED, FAC_OVERLAP_KEEP, 0, &m);
            CHECK(st == FAC_HAYSTACK_TOO_LARGE && fac_last_haystack_graphemes() == n, "FAC_HAYSTACK_TOO_LARGE at 2^32 + 1 graphemes");
            cudaFree(d);
        } else CHECK(0, "cudaMalloc 4 GiB");
        /* absolute u64 offsets past 4 GiB: 4.5 GiB stream of a repeating 1 MiB block with one planted hit per block */
        static uint8_t blk[1 << 20];
        for (size_t i = 0; i < sizeof(blk); i++) blk[i] = (uint8_t)("lorem ipsum dolor sit amet "[i % 27]);
        memcpy(blk + 333333, " neadle ", 8);
        fac_engine *e1 = make_engine(pats + 2, 1, 1, 1, 0, 0);
        facio_block_reader rd = {blk, sizeof(blk), 4608ull << 20, 0, 65536, -1};
        facio_match_stats ms;
        memset(&ms, 0, sizeof(ms));
        fac_stream_stats ss;
        st = fac_search_stream_stats(e1, facio_block_read, &rd, 0.8f, facio_on_match, &ms, &ss);
        CHECK(st == FAC_OK && ss.bytes_read == (4608ull << 20) && ms.count == 4608 && ms.beyond_4g == 512 && ms.max_start == (4607ull << 20) + 333334 && ms.order_violations == 0,
              "4.5 GiB stream: 4608 matches, 512 of them past 2^32 with exact absolute offsets");
        printf("     stream: %.1f MB/s device side (%.0f ms device, %u launches)\n", ss.bytes_read / ss.device_ms / 1e3, ss.device_ms, ss.kernel_launches);
        fac_engine_free(e1);
    }
    fac_engine_free(e);
    printf("%s\n", failures ? "FAILED" : "ALL OK");
    return failures ? 1 : 0;
}
