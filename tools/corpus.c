/*
 * corpus.c -- deterministic synthetic corpora for tests and bench.py
 * (SURVEY.md section 8d).  Bench/test infrastructure, not part of the codec.
 *
 * PRNG: splitmix64, seed = 0x5EED000000000000 + (config_id << 32) + block.
 * Classes:
 *   0 text-like   4096-word vocabulary (length 1+geometric(0.25) capped 12,
 *                 letters by English unigram frequency), Zipf(s=1.1) word
 *                 choice, separators " " 85% / ", " 5% / ". "+capital 6% /
 *                 "\n" 4%
 *   1 binary      32-byte records: u32 counter, u32 counter*stride, 8 bytes
 *                 of a 16-symbol alphabet, u64 random walk, 8 zero bytes with
 *                 10% random mutations
 *   2 random      raw PRNG bytes
 *   3 repetitive  the data model of the reference's benchmark generator
 *                 (LzmaBench.java:63-128: literals vs. copies from a
 *                 log-distributed offset with short lengths and rep0 reuse),
 *                 its two MWC seeds offset by the block seed
 *   4 mixed       class = block_index mod 4
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint64_t s; } sm64;
static inline uint64_t sm_next(sm64 *r) {
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double sm_unit(sm64 *r) { return (double)(sm_next(r) >> 11) * (1.0 / 9007199254740992.0); }

/* ---- text ---------------------------------------------------------------- */
#define VOCAB 4096
static char g_words[VOCAB][13];
static uint8_t g_wlen[VOCAB];
static double g_zipf_cdf[VOCAB];
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

static void init_vocab(void) {
    /* English letter frequencies (per mille), a..z */
    static const int freq[26] = {82, 15, 28, 43, 127, 22, 20, 61, 70, 2, 8, 40, 24,
                                 67, 75, 19, 1, 60, 63, 91, 28, 10, 24, 2, 20, 1};
    int cdf[26], tot = 0;
    for (int i = 0; i < 26; i++) { tot += freq[i]; cdf[i] = tot; }
    sm64 r = {0x5EED0000C0FFEE00ull};
    for (int w = 0; w < VOCAB; w++) {
        int len = 1;
        while (len < 12 && sm_unit(&r) >= 0.25) len++;
        for (int k = 0; k < len; k++) {
            int x = (int)(sm_next(&r) % (uint64_t)tot), c = 0;
            while (cdf[c] <= x) c++;
            g_words[w][k] = (char)('a' + c);
        }
        g_words[w][len] = 0;
        g_wlen[w] = (uint8_t)len;
    }
    double sum = 0;
    for (int i = 0; i < VOCAB; i++) sum += 1.0 / pow((double)(i + 1), 1.1);
    double acc = 0;
    for (int i = 0; i < VOCAB; i++) {
        acc += 1.0 / pow((double)(i + 1), 1.1) / sum;
        g_zipf_cdf[i] = acc;
    }
    g_zipf_cdf[VOCAB - 1] = 1.0;
}

static void gen_text(uint8_t *out, size_t n, uint64_t seed) {
    pthread_once(&g_once, init_vocab);
    sm64 r = {seed};
    size_t pos = 0;
    int capital = 1;
    while (pos < n) {
        double u = sm_unit(&r);
        int lo = 0, hi = VOCAB - 1;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (g_zipf_cdf[mid] < u) lo = mid + 1; else hi = mid;
        }
        const char *w = g_words[lo];
        int len = g_wlen[lo];
        for (int k = 0; k < len && pos < n; k++) {
            char c = w[k];
            if (k == 0 && capital) c = (char)(c - 32);
            out[pos++] = (uint8_t)c;
        }
        capital = 0;
        double s = sm_unit(&r);
        if (s < 0.85) {
            if (pos < n) out[pos++] = ' ';
        } else if (s < 0.90) {
            if (pos < n) out[pos++] = ',';
            if (pos < n) out[pos++] = ' ';
        } else if (s < 0.96) {
            if (pos < n) out[pos++] = '.';
            if (pos < n) out[pos++] = ' ';
            capital = 1;
        } else {
            if (pos < n) out[pos++] = '\n';
        }
    }
}

/* ---- binary records ------------------------------------------------------ */
static void gen_binary(uint8_t *out, size_t n, uint64_t seed) {
    sm64 r = {seed};
    static const uint8_t alphabet[16] = {0x00, 0x01, 0x02, 0x04, 0x08, 0x10, 0x20, 0x40,
                                         0x80, 0xFF, 0x7F, 0x55, 0xAA, 0x0F, 0xF0, 0x3C};
    uint32_t counter = (uint32_t)sm_next(&r) & 0xFFFF;
    uint32_t stride = 1 + (uint32_t)(sm_next(&r) % 997);
    uint64_t walk = sm_next(&r);
    size_t pos = 0;
    while (pos < n) {
        uint8_t rec[32];
        uint32_t c2 = counter * stride;
        memcpy(rec, &counter, 4);
        memcpy(rec + 4, &c2, 4);
        uint64_t a = sm_next(&r);
        for (int k = 0; k < 8; k++) rec[8 + k] = alphabet[(a >> (4 * k)) & 15];
        walk += (uint64_t)((int64_t)(sm_next(&r) % 33) - 16);
        memcpy(rec + 16, &walk, 8);
        uint64_t m = sm_next(&r), v = sm_next(&r);
        for (int k = 0; k < 8; k++) {
            /* 10% of the padding bytes mutate */
            rec[24 + k] = ((m >> (8 * k)) & 0xFF) < 26 ? (uint8_t)(v >> (8 * k)) : 0;
        }
        counter++;
        size_t take = n - pos < 32 ? n - pos : 32;
        memcpy(out + pos, rec, take);
        pos += take;
    }
}

/* ---- random --------------------------------------------------------------- */
static void gen_random(uint8_t *out, size_t n, uint64_t seed) {
    sm64 r = {seed};
    size_t pos = 0;
    while (pos + 8 <= n) {
        uint64_t v = sm_next(&r);
        memcpy(out + pos, &v, 8);
        pos += 8;
    }
    if (pos < n) {
        uint64_t v = sm_next(&r);
        memcpy(out + pos, &v, n - pos);
    }
}

/* ---- repetitive: data model of LzmaBench.java:63-128 ---------------------- */
typedef struct { uint32_t a1, a2, value; int num_bits; } bitrng;
static uint32_t mwc(bitrng *g) { /* two 16-bit multiply-with-carry lanes, LzmaBench.java:27-31 */
    g->a1 = 36969u * (g->a1 & 0xffff) + (g->a1 >> 16);
    g->a2 = 18000u * (g->a2 & 0xffff) + (g->a2 >> 16);
    return (g->a1 << 16) ^ g->a2;
}
static uint32_t bits(bitrng *g, int nb) { /* LzmaBench.java:44-60 */
    uint32_t result;
    if (g->num_bits > nb) {
        result = g->value & ((1u << nb) - 1);
        g->value >>= nb;
        g->num_bits -= nb;
        return result;
    }
    nb -= g->num_bits;
    result = g->value << nb;
    g->value = mwc(g);
    result |= g->value & ((1u << nb) - 1);
    g->value = nb < 32 ? g->value >> nb : 0;
    g->num_bits = 32 - nb;
    return result;
}
static uint32_t log_bits(bitrng *g, int nb) { uint32_t len = bits(g, nb); return bits(g, (int)len); }

static void gen_repetitive(uint8_t *out, size_t n, uint64_t seed) {
    bitrng g = {362436069u + (uint32_t)seed, 521288629u + (uint32_t)(seed >> 32), 0, 0};
    if ((g.a1 & 0xffff) == 0) g.a1 += 1;
    if ((g.a2 & 0xffff) == 0) g.a2 += 1;
    size_t pos = 0;
    uint32_t rep0 = 1;
    while (pos < n) {
        if (bits(&g, 1) == 0 || pos < 1) {
            out[pos++] = (uint8_t)bits(&g, 8);
        } else {
            uint32_t len;
            if (bits(&g, 3) == 0) {
                len = 1 + bits(&g, 1 + (int)bits(&g, 2));
            } else {
                do {
                    if (bits(&g, 1) == 0) rep0 = log_bits(&g, 4);
                    else rep0 = (log_bits(&g, 4) << 10) | bits(&g, 10);
                } while (rep0 >= pos);
                rep0++;
                len = 2 + bits(&g, 2 + (int)bits(&g, 2));
            }
            for (uint32_t i = 0; i < len && pos < n; i++, pos++) out[pos] = out[pos - rep0];
        }
    }
}

/* ---- driver --------------------------------------------------------------- */
static void gen_block(uint8_t *out, size_t n, int cls, uint64_t seed, uint64_t block) {
    if (cls == 4) cls = (int)(block & 3);
    switch (cls) {
        case 0: gen_text(out, n, seed); break;
        case 1: gen_binary(out, n, seed); break;
        case 2: gen_random(out, n, seed); break;
        default: gen_repetitive(out, n, seed); break;
    }
}

typedef struct {
    uint8_t *out;
    size_t block_size;
    uint64_t n_blocks, first_block;
    int cls;
    uint64_t config_id;
    volatile uint64_t next;
} job;

static void *worker(void *arg) {
    job *j = (job *)arg;
    for (;;) {
        uint64_t b = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (b >= j->n_blocks) break;
        uint64_t gb = j->first_block + b;
        uint64_t seed = 0x5EED000000000000ull + (j->config_id << 32) + gb;
        gen_block(j->out + b * j->block_size, j->block_size, j->cls, seed, gb);
    }
    return NULL;
}

/* Fill out[0 .. n_blocks*block_size) with blocks first_block .. of the corpus
 * (cls, config_id). */
void corpus_generate(uint8_t *out, size_t block_size, uint64_t n_blocks, uint64_t first_block, int cls,
                     uint64_t config_id, int threads) {
    job j = {out, block_size, n_blocks, first_block, cls, config_id, 0};
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t tids[256];
    int started = 0;
    for (int t = 0; t < threads; t++)
        if (pthread_create(&tids[t], NULL, worker, &j) == 0) started++; else break;
    if (!started) worker(&j);
    for (int t = 0; t < started; t++) pthread_join(tids[t], NULL);
}
