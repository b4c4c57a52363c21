/* facio.c -- TEST / BENCH INFRASTRUCTURE.  Native io::Read / io::Write / match-callback implementations for the
 * streaming entry points of include/fac.h, so that stream benchmarks and the >4 GiB offset tests do not measure Python
 * callbacks.  Built as libfacio.so (gcc -shared) and handed to libfacgpu.so through ctypes function pointers, or linked
 * into tests/abi_c/conformance.c.
 *
 *   facio_block_reader : a Read source that repeats one block and hands out at most max_read bytes per call, with a
 *                        short read at every block end (the contract of examples/streaming.rs:43-82)
 *   facio_sink         : a Write sink that counts, hashes (FNV-1a 64) and keeps the first keep_cap bytes
 *   facio_match_stats  : a match callback that counts, hashes and tracks the largest start offset
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "../../include/fac.h"

typedef struct facio_block_reader {
    const uint8_t *block;
    size_t block_len;
    uint64_t total, pos;
    size_t max_read;
    int64_t fail_at; /* >= 0: return -1 once pos reaches this offset (io::Error) */
} facio_block_reader;

int64_t facio_block_read(void *user, uint8_t *buf, size_t cap) {
    facio_block_reader *r = (facio_block_reader *)user;
    if (r->fail_at >= 0 && r->pos >= (uint64_t)r->fail_at) return -1;
    if (r->pos >= r->total) return 0;
    const size_t off = (size_t)(r->pos % r->block_len);
    size_t n = r->block_len - off;
    if ((uint64_t)n > r->total - r->pos) n = (size_t)(r->total - r->pos);
    if (n > r->max_read) n = r->max_read;
    if (n > cap) n = cap;
    memcpy(buf, r->block + off, n);
    r->pos += n;
    return (int64_t)n;
}

typedef struct facio_sink {
    uint64_t bytes, fnv;
    uint8_t *keep;
    size_t keep_cap, keep_len;
    int64_t fail_at; /* >= 0: fail (return 1) once `bytes` reaches this count */
} facio_sink;

void facio_sink_init(facio_sink *s, uint8_t *keep, size_t keep_cap) {
    s->bytes = 0; s->fnv = 0xCBF29CE484222325ull; s->keep = keep; s->keep_cap = keep_cap; s->keep_len = 0; s->fail_at = -1;
}

int facio_sink_write(void *user, const uint8_t *buf, size_t len) {
    facio_sink *s = (facio_sink *)user;
    if (s->fail_at >= 0 && s->bytes + len > (uint64_t)s->fail_at) return 1;
    uint64_t h = s->fnv;
    for (size_t i = 0; i < len; i++) { h ^= buf[i]; h *= 0x100000001B3ull; }
    s->fnv = h;
    if (s->keep && s->keep_len < s->keep_cap) {
        size_t n = s->keep_cap - s->keep_len;
        if (n > len) n = len;
        memcpy(s->keep + s->keep_len, buf, n);
        s->keep_len += n;
    }
    s->bytes += len;
    return 0;
}

typedef struct facio_match_stats {
    uint64_t count, max_start, hash, beyond_4g;
    uint64_t last_start; /* stream order: starts never decrease by more than a window */
    uint64_t order_violations;
} facio_match_stats;

void facio_on_match(void *user, const fac_match *m) {
    facio_match_stats *s = (facio_match_stats *)user;
    s->count++;
    if (m->start > s->max_start) s->max_start = m->start;
    if (m->start >= (1ull << 32)) s->beyond_4g++;
    if (m->start + (1u << 20) < s->last_start) s->order_violations++;
    s->last_start = m->start;
    uint64_t h = s->hash ^ (m->start * 0x9E3779B97F4A7C15ull) ^ (m->end << 17) ^ ((uint64_t)m->pattern_index << 40);
    s->hash = h * 0x100000001B3ull + m->edits;
}
