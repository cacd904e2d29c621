/* ta_synth.c -- deterministic synthetic batches for bench.py and the full-size tests, in C because
 * a million independently edited pairs are too slow to make in Python.  Test / bench infrastructure, not product code.
 *
 * Modelled on the reference's bench generators (benches/rand_benchmarks.rs:126-260: a random string, then a bounded
 * number of random edits; a needle planted in random haystacks), seeded like them with 1234 by the callers.  Every
 * unit (pair / haystack) draws from its own splitmix64 stream keyed by (seed, global unit index), so a unit's bytes do
 * not depend on how many units are generated, by how many threads, or by which rank: rank r of N can generate units
 * [lo, hi) of ONE batch on its own (strong scaling in bench.py) and 1 M pairs are 1 M different edit scripts.
 *
 *   gcc -O2 -shared -fPIC -pthread -o tools/libta_synth.so tools/ta_synth.c
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    uint64_t s;
} rng_t;

static inline uint64_t rng_next(rng_t *r) {
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline rng_t rng_for(uint64_t seed, uint64_t unit, uint64_t stream) {
    rng_t r = {seed * 0xD1342543DE82EF95ull + unit * 0x9E3779B97F4A7C15ull + stream * 0xC2B2AE3D27D4EB4Full};
    (void)rng_next(&r);
    return r;
}
static inline uint32_t rng_below(rng_t *r, uint32_t n) { /* uniform in [0, n), n > 0 */
    return (uint32_t)(((rng_next(r) >> 32) * (uint64_t)n) >> 32);
}

typedef struct {
    uint64_t seed, first, count;
    uint32_t len_lo, len_hi, max_edits, alphabet;
    int exact, allow_swap;
    uint32_t *la, *lb;                /* pass 1 */
    uint8_t *a, *b;                   /* pass 2 */
    const uint64_t *a_off, *b_off;
} pairs_job;

/* The edit script of one pair, replayed identically by the length pass and the fill pass.  b == NULL: lengths only. */
static uint32_t edit_script(const pairs_job *j, uint64_t unit, const uint8_t *a, uint32_t la, uint8_t *b) {
    rng_t r = rng_for(j->seed, unit, 1);
    uint32_t len = la;
    if (b) memcpy(b, a, la);
    const uint32_t e = j->exact ? j->max_edits : rng_below(&r, j->max_edits + 1);
    const uint32_t kinds = j->allow_swap ? 4 : 3;
    for (uint32_t i = 0; i < e; i++) {
        const uint32_t kind = rng_below(&r, kinds);
        if (kind == 0) { /* substitute with a different symbol */
            const uint32_t p = len ? rng_below(&r, len) : 0;
            const uint32_t d = 1 + rng_below(&r, j->alphabet - 1);
            if (len && b) b[p] = (uint8_t)((b[p] + d) % j->alphabet);
        } else if (kind == 1) { /* insert */
            const uint32_t p = rng_below(&r, len + 1);
            const uint8_t c = (uint8_t)rng_below(&r, j->alphabet);
            if (b) {
                memmove(b + p + 1, b + p, len - p);
                b[p] = c;
            }
            len++;
        } else if (kind == 2) { /* delete */
            const uint32_t p = len ? rng_below(&r, len) : 0;
            if (len) {
                if (b) memmove(b + p, b + p + 1, len - p - 1);
                len--;
            }
        } else { /* swap two adjacent symbols */
            const uint32_t p = len > 1 ? rng_below(&r, len - 1) : 0;
            if (len > 1 && b) {
                const uint8_t t = b[p];
                b[p] = b[p + 1];
                b[p + 1] = t;
            }
        }
    }
    return len;
}

static uint32_t pair_len_a(const pairs_job *j, uint64_t unit) {
    rng_t r = rng_for(j->seed, unit, 2);
    return j->len_lo + (j->len_hi > j->len_lo ? rng_below(&r, j->len_hi - j->len_lo + 1) : 0);
}

typedef struct {
    void (*fn)(void *, uint64_t, uint64_t);
    void *arg;
    uint64_t lo, hi;
} slice_t;
static void *slice_main(void *p) {
    slice_t *s = (slice_t *)p;
    s->fn(s->arg, s->lo, s->hi);
    return NULL;
}
static void parallel_for(void (*fn)(void *, uint64_t, uint64_t), void *arg, uint64_t n, int threads) {
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > n) threads = n ? (int)n : 1;
    pthread_t th[256];
    slice_t sl[256];
    if (threads > 256) threads = 256;
    for (int t = 0; t < threads; t++) {
        sl[t].fn = fn, sl[t].arg = arg, sl[t].lo = n * (uint64_t)t / threads, sl[t].hi = n * (uint64_t)(t + 1) / threads;
        if (t + 1 < threads) pthread_create(&th[t], NULL, slice_main, &sl[t]);
    }
    slice_main(&sl[threads - 1]);
    for (int t = 0; t + 1 < threads; t++) pthread_join(th[t], NULL);
}

static void lengths_range(void *arg, uint64_t lo, uint64_t hi) {
    const pairs_job *j = (const pairs_job *)arg;
    for (uint64_t i = lo; i < hi; i++) {
        const uint32_t la = pair_len_a(j, j->first + i);
        j->la[i] = la;
        j->lb[i] = edit_script(j, j->first + i, NULL, la, NULL);
    }
}
static void fill_range(void *arg, uint64_t lo, uint64_t hi) {
    const pairs_job *j = (const pairs_job *)arg;
    for (uint64_t i = lo; i < hi; i++) {
        const uint64_t unit = j->first + i;
        const uint32_t la = (uint32_t)(j->a_off[i + 1] - j->a_off[i]);
        uint8_t *a = j->a + j->a_off[i];
        rng_t r = rng_for(j->seed, unit, 0);
        uint32_t p = 0;
        if (j->alphabet == 256) {
            for (; p + 8 <= la; p += 8) {
                const uint64_t x = rng_next(&r);
                memcpy(a + p, &x, 8);
            }
        }
        for (; p < la; p++) a[p] = (uint8_t)rng_below(&r, j->alphabet);
        /* b is edited in place inside its final slot: the slot is exactly the final length, intermediate lengths may
         * exceed it by up to max_edits bytes, so edit in a scratch buffer and copy */
        uint8_t stack[8192 + 512];
        const uint32_t need = la + j->max_edits + 8;
        uint8_t *tmp = need <= sizeof stack ? stack : (uint8_t *)malloc(need);
        const uint32_t lb = edit_script(j, unit, a, la, tmp);
        memcpy(j->b + j->b_off[i], tmp, lb);
        if (tmp != stack) free(tmp);
    }
}

/* pass 1: |a_i| and |b_i| of units first .. first + count */
void synth_pair_lengths(uint64_t seed, uint64_t first, uint64_t count, uint32_t len_lo, uint32_t len_hi,
                        uint32_t max_edits, int exact, int allow_swap, uint32_t alphabet, uint32_t *la, uint32_t *lb,
                        int threads) {
    pairs_job j = {seed, first, count, len_lo, len_hi, max_edits, alphabet ? alphabet : 256, exact, allow_swap,
                   la, lb, NULL, NULL, NULL, NULL};
    parallel_for(lengths_range, &j, count, threads);
}
/* pass 2: the bytes, into CSR buffers whose offsets are the prefix sums of pass 1 */
void synth_pair_fill(uint64_t seed, uint64_t first, uint64_t count, uint32_t len_lo, uint32_t len_hi, uint32_t max_edits,
                     int exact, int allow_swap, uint32_t alphabet, uint8_t *a, const uint64_t *a_off, uint8_t *b,
                     const uint64_t *b_off, int threads) {
    pairs_job j = {seed, first, count, len_lo, len_hi, max_edits, alphabet ? alphabet : 256, exact, allow_swap,
                   NULL, NULL, a, b, a_off, b_off};
    parallel_for(fill_range, &j, count, threads);
}

/* ---- haystacks with a planted, mutated needle (model: benches/rand_benchmarks.rs:175-198) ---------------------- */
typedef struct {
    uint64_t seed, first;
    uint32_t hay_len, needle_len, plant_per_million, max_edits;
    const uint8_t *needle;
    uint8_t *hay;
} hay_job;
static void hay_range(void *arg, uint64_t lo, uint64_t hi) {
    const hay_job *j = (const hay_job *)arg;
    for (uint64_t i = lo; i < hi; i++) {
        const uint64_t unit = j->first + i;
        uint8_t *h = j->hay + i * (uint64_t)j->hay_len;
        rng_t r = rng_for(j->seed, unit, 3);
        for (uint32_t p = 0; p < j->hay_len; p++) h[p] = (uint8_t)(1 + rng_below(&r, 255)); /* bytes 1..255 */
        rng_t q = rng_for(j->seed, unit, 4);
        if (rng_below(&q, 1000000) < j->plant_per_million && j->needle_len && j->hay_len >= j->needle_len + j->max_edits) {
            uint8_t mut[1024];
            pairs_job pj = {j->seed ^ 0x5bd1e995u, 0, 0, 0, 0, j->max_edits, 256, 0, 0, NULL, NULL, NULL, NULL, NULL, NULL};
            const uint32_t nl = j->needle_len < 512 ? j->needle_len : 512;
            const uint32_t ml = edit_script(&pj, unit, j->needle, nl, mut);
            for (uint32_t p = 0; p < ml; p++)
                if (mut[p] == 0) mut[p] = 1;
            const uint32_t at = rng_below(&q, j->hay_len - ml + 1);
            memcpy(h + at, mut, ml);
        }
    }
}
void synth_haystacks(uint64_t seed, uint64_t first, uint64_t count, uint32_t hay_len, const uint8_t *needle,
                     uint32_t needle_len, uint32_t plant_per_million, uint32_t max_edits, uint8_t *hay, int threads) {
    hay_job j = {seed, first, hay_len, needle_len, plant_per_million, max_edits, needle, hay};
    parallel_for(hay_range, &j, count, threads);
}
