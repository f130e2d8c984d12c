/*
 * bm_synth.c -- TEST INFRASTRUCTURE ONLY (part of oracle/liboracle.so).
 *
 * Host restatement of the benchmark's synthetic input generator (the product's synth_*_kernel family,
 * bitmagic_b200/csrc/aux_kernels.cuh): bit p of vector v is set iff u16(splitmix64(seed_v, p)) < round(density_v * 65536),
 * one 64-bit hash yields four bits, and with optimize != 0 every block is stored the way bvector::optimize(opt_compress)
 * stores it (optimize_bit_block, reference src/bmblocks.h:1414-1437: all-zero -> NULL, all-one -> FULL,
 * runs < 1276 -> GAP via bit_block_to_gap src/bmfunc.h:5540, else bit-block).
 *
 * Why it exists: bench.py's reference arm and its all-column parity check need the SAME inputs as the GPU arm without
 * touching the product library or a GPU, and an independent second implementation of the generator is itself a check
 * (tests compare the two bit for bit, and push the blocks through the real bvector::optimize()).
 * The output is a host bmb200_packed_set (include/bmb200.h) in the flat-streamable GAP form, identical to what
 * bmb200_synth_set + bmb200_set_download deliver.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "bm_oracle.h"

typedef struct {
    uint16_t* gap; size_t gap_n, gap_cap;      /* GAP units of this thread's column range, in layout order */
    uint32_t* bit; size_t bit_n, bit_cap;      /* bit-blocks of this thread's column range, in layout order (words) */
} synth_local;

struct orc_synth {
    uint32_t n_vec, n_blocks;
    uint32_t* desc; uint64_t* bit_base; uint64_t* gap_base; uint32_t* bit_pool; uint16_t* gap_pool;
};

typedef struct {
    struct orc_synth* s; const uint32_t* thr; const uint64_t* seed; int optimize;
    uint32_t lo, hi; synth_local loc; int rc;
    uint64_t bit_off, gap_off;                 /* pass 2: where this thread's data goes (blocks / units) */
} synth_job;

static inline uint64_t mix64(uint64_t x)
{   /* splitmix64 finalizer */
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27; x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
/* the eight hashes of one word are one zmm register, and the sixteen-bit compares of all 32 lanes ARE the word (bit 4h+q =
 * lane q of hash h): ~12x the scalar loop.  Same arithmetic, checked against the scalar form by the tests. */
__attribute__((target("avx512f,avx512bw,avx512dq")))
static void synth_block_avx512(uint64_t seed, uint32_t nb, uint32_t thr, uint32_t* w)
{
    const uint64_t G = 0x9e3779b97f4a7c15ull;
    const __m512i lane = _mm512_set_epi64((long long)(7 * G), (long long)(6 * G), (long long)(5 * G), (long long)(4 * G),
                                          (long long)(3 * G), (long long)(2 * G), (long long)G, 0);
    const __m512i c1 = _mm512_set1_epi64((long long)0xbf58476d1ce4e5b9ull), c2 = _mm512_set1_epi64((long long)0x94d049bb133111ebull);
    const __m512i step = _mm512_set1_epi64((long long)(8 * G)), vthr = _mm512_set1_epi16((short)thr);
    __m512i x0 = _mm512_add_epi64(_mm512_set1_epi64((long long)(seed + G * ((((uint64_t)nb << 11) << 3) + 1))), lane);
    for (uint32_t wi = 0; wi < BMB200_BLOCK_WORDS; ++wi) {
        __m512i x = x0;
        x = _mm512_xor_si512(x, _mm512_srli_epi64(x, 30)); x = _mm512_mullo_epi64(x, c1);
        x = _mm512_xor_si512(x, _mm512_srli_epi64(x, 27)); x = _mm512_mullo_epi64(x, c2);
        x = _mm512_xor_si512(x, _mm512_srli_epi64(x, 31));
        w[wi] = (uint32_t)_mm512_cmplt_epu16_mask(x, vthr);
        x0 = _mm512_add_epi64(x0, step);
    }
}
static int have_avx512(void)
{
    static int v = -1;
    if (v < 0) v = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512dq");
    return v;
}
#else
static int have_avx512(void) { return 0; }
static void synth_block_avx512(uint64_t seed, uint32_t nb, uint32_t thr, uint32_t* w) { (void)seed; (void)nb; (void)thr; (void)w; }
#endif

static int g_force_scalar = 0;
void orc_synth_force_scalar(int on) { g_force_scalar = on; }      /* tests: compare the two forms */

static void synth_block(uint64_t seed, uint32_t nb, uint32_t thr, uint32_t* w)
{
    if (thr == 0) { memset(w, 0, BMB200_BLOCK_BYTES); return; }
    if (thr >= 65536u) { memset(w, 0xff, BMB200_BLOCK_BYTES); return; }
    if (!g_force_scalar && have_avx512()) { synth_block_avx512(seed, nb, thr, w); return; }
    for (uint32_t wi = 0; wi < BMB200_BLOCK_WORDS; ++wi) {
        const uint64_t ctr = ((((uint64_t)nb << 11) | wi) << 3);      /* 8 hashes per word */
        uint32_t x = 0;
        for (int h = 0; h < 8; ++h) {
            const uint64_t r = mix64(seed + 0x9e3779b97f4a7c15ull * (ctr + (uint64_t)h + 1));
            x |= (uint32_t)((uint32_t)(r & 0xffffu) < thr) << (4 * h);
            x |= (uint32_t)((uint32_t)((r >> 16) & 0xffffu) < thr) << (4 * h + 1);
            x |= (uint32_t)((uint32_t)((r >> 32) & 0xffffu) < thr) << (4 * h + 2);
            x |= (uint32_t)((uint32_t)(r >> 48) < thr) << (4 * h + 3);
        }
        w[wi] = x;
    }
}

/* thread-local output grows in few, large steps (x2, at least 32 MB worth of elements): every realloc of a big array is an
 * mremap that takes the process-wide mmap lock, and 128 threads doing that often stall each other's page faults */
static int grow(void** p, size_t* cap, size_t need, size_t elem)
{
    if (need <= *cap) return 0;
    size_t nc = *cap ? *cap * 2 : (size_t)(32u << 20) / elem;
    while (nc < need) nc *= 2;
    void* q = realloc(*p, nc * elem);
    if (!q) return 1;
    *p = q; *cap = nc;
    return 0;
}

static void* synth_pass1(void* arg)
{
    synth_job* j = (synth_job*)arg;
    struct orc_synth* s = j->s;
    uint32_t w[BMB200_BLOCK_WORDS];
    uint16_t g[BMB200_GAP_MAX_WORDS + 8];
    for (uint32_t nb = j->lo; nb < j->hi; ++nb) {
        uint64_t nbit = 0, ngap = 0;
        for (uint32_t v = 0; v < s->n_vec; ++v) {
            synth_block(j->seed[v], nb, j->thr[v], w);
            uint32_t pc = 0, tr = 0;
            for (uint32_t i = 0; i < BMB200_BLOCK_WORDS; ++i) {
                const uint32_t x = w[i], nxt = (i + 1 < BMB200_BLOCK_WORDS) ? (w[i + 1] & 1u) : (x >> 31);
                pc += (uint32_t)__builtin_popcount(x);
                tr += (uint32_t)__builtin_popcount(x ^ ((x >> 1) | (nxt << 31)));
            }
            const uint32_t runs = tr + 1u;
            uint32_t kd;
            if (pc == 0) kd = BMB200_BLK_NULL;
            else if (!j->optimize) kd = BMB200_BLK_BIT;
            else if (pc == 65536u) kd = BMB200_BLK_FULL;
            else if (runs < BMB200_GAP_THRESHOLD) kd = BMB200_BLK_GAP;
            else kd = BMB200_BLK_BIT;
            uint32_t d = kd;
            if (kd == BMB200_BLK_BIT) {
                if (grow((void**)&j->loc.bit, &j->loc.bit_cap, j->loc.bit_n + BMB200_BLOCK_WORDS, 4)) { j->rc = BMB200_ERR_BADALLOC; return 0; }
                memcpy(j->loc.bit + j->loc.bit_n, w, BMB200_BLOCK_BYTES);
                j->loc.bit_n += BMB200_BLOCK_WORDS;
                d |= (uint32_t)nbit++ << 2;
            } else if (kd == BMB200_BLK_GAP) {
                /* bit_block_to_gap (src/bmfunc.h:5540): run ends = positions whose successor differs, then 65535 */
                uint32_t len = 0;
                for (uint32_t i = 0; i < BMB200_BLOCK_WORDS; ++i) {
                    const uint32_t x = w[i], nxt = (i + 1 < BMB200_BLOCK_WORDS) ? (w[i + 1] & 1u) : (x >> 31);
                    uint32_t m = x ^ ((x >> 1) | (nxt << 31));
                    while (m) { g[++len] = (uint16_t)(32u * i + (uint32_t)__builtin_ctz(m)); m &= m - 1u; }
                }
                g[++len] = 65535u;                                         /* len = runs */
                g[0] = (uint16_t)((w[0] & 1u) | ((len <= 124u ? 0u : len <= 252u ? 1u : len <= 508u ? 2u : 3u) << 1) | (len << 3));   /* gap_calc_level src/bmfunc.h:5418 */
                const uint32_t pad = (g[0] & 1u) ? 0u : 1u;               /* flat form: lead pad iff the first run is 0 */
                const uint32_t words = len + 1u + pad, units = (words + BMB200_GAP_UNIT_WORDS - 1) / BMB200_GAP_UNIT_WORDS;
                if (grow((void**)&j->loc.gap, &j->loc.gap_cap, j->loc.gap_n + (size_t)units * BMB200_GAP_UNIT_WORDS, 2)) { j->rc = BMB200_ERR_BADALLOC; return 0; }
                uint16_t* dst = j->loc.gap + j->loc.gap_n;
                memset(dst, 0, (size_t)units * 16u);
                if (pad) dst[0] = 0xffffu;
                memcpy(dst + pad, g, (size_t)(len + 1u) * 2u);
                j->loc.gap_n += (size_t)units * BMB200_GAP_UNIT_WORDS;
                d |= ((uint32_t)ngap << 2) | (pad ? BMB200_DESC_GAP_PAD : 0u) | BMB200_DESC_GAP_FLAT;
                ngap += units;
            }
            s->desc[(size_t)nb * s->n_vec + v] = d;
        }
        s->bit_base[nb + 1] = nbit; s->gap_base[nb + 1] = ngap;       /* per-column sizes, scanned by the caller */
    }
    return 0;
}

static void* synth_pass2(void* arg)
{
    synth_job* j = (synth_job*)arg;
    if (j->loc.bit_n) memcpy(j->s->bit_pool + j->bit_off * BMB200_BLOCK_WORDS, j->loc.bit, j->loc.bit_n * 4);
    if (j->loc.gap_n) memcpy(j->s->gap_pool + j->gap_off * BMB200_GAP_UNIT_WORDS, j->loc.gap, j->loc.gap_n * 2);
    free(j->loc.bit); free(j->loc.gap);
    j->loc.bit = 0; j->loc.gap = 0;
    return 0;
}

void orc_synth_free(struct orc_synth* s)
{
    if (!s) return;
    free(s->desc); free(s->bit_base); free(s->gap_base); free(s->bit_pool); free(s->gap_pool);
    free(s);
}

int orc_synth_create(uint32_t n_vec, uint32_t n_blocks, const double* density, const uint64_t* seed, int optimize, int threads,
                     struct orc_synth** out)
{
    if (!n_vec || !n_blocks || !density || !seed || !out) return BMB200_ERR_BADARG;
    if (threads < 1) threads = 1;
    if ((uint32_t)threads > n_blocks) threads = (int)n_blocks;
    struct orc_synth* s = (struct orc_synth*)calloc(1, sizeof *s);
    uint32_t* thr = (uint32_t*)malloc((size_t)n_vec * 4);
    synth_job* jobs = (synth_job*)calloc((size_t)threads, sizeof *jobs);
    pthread_t* th = (pthread_t*)malloc((size_t)threads * sizeof *th);
    if (!s || !thr || !jobs || !th) { free(s); free(thr); free(jobs); free(th); return BMB200_ERR_BADALLOC; }
    s->n_vec = n_vec; s->n_blocks = n_blocks;
    s->desc = (uint32_t*)malloc((size_t)n_vec * n_blocks * 4);
    s->bit_base = (uint64_t*)calloc((size_t)n_blocks + 1, 8);
    s->gap_base = (uint64_t*)calloc((size_t)n_blocks + 1, 8);
    int rc = (s->desc && s->bit_base && s->gap_base) ? BMB200_OK : BMB200_ERR_BADALLOC;
    for (uint32_t v = 0; v < n_vec; ++v) {
        const double t = density[v] * 65536.0 + 0.5;                       /* same rounding as bmb200_synth_set */
        thr[v] = t <= 0 ? 0u : t >= 65536.0 ? 65536u : (uint32_t)t;
    }
    if (!rc) {
        for (int t = 0; t < threads; ++t) {
            jobs[t].s = s; jobs[t].thr = thr; jobs[t].seed = seed; jobs[t].optimize = optimize;
            jobs[t].lo = (uint32_t)((uint64_t)n_blocks * (uint64_t)t / (uint64_t)threads);
            jobs[t].hi = (uint32_t)((uint64_t)n_blocks * ((uint64_t)t + 1) / (uint64_t)threads);
            pthread_create(&th[t], 0, synth_pass1, &jobs[t]);
        }
        for (int t = 0; t < threads; ++t) { pthread_join(th[t], 0); if (jobs[t].rc) rc = jobs[t].rc; }
    }
    if (!rc) {
        for (uint32_t nb = 0; nb < n_blocks; ++nb) { s->bit_base[nb + 1] += s->bit_base[nb]; s->gap_base[nb + 1] += s->gap_base[nb]; }
        const uint64_t n_bit = s->bit_base[n_blocks], n_gap = s->gap_base[n_blocks];
        s->bit_pool = (uint32_t*)malloc(n_bit ? n_bit * BMB200_BLOCK_BYTES : 16);
        s->gap_pool = (uint16_t*)malloc(n_gap ? n_gap * 16u + 64 : 64);
        if (!s->bit_pool || !s->gap_pool) rc = BMB200_ERR_BADALLOC;
    }
    if (!rc) {
        for (int t = 0; t < threads; ++t) {
            jobs[t].bit_off = s->bit_base[jobs[t].lo]; jobs[t].gap_off = s->gap_base[jobs[t].lo];
            pthread_create(&th[t], 0, synth_pass2, &jobs[t]);
        }
        for (int t = 0; t < threads; ++t) pthread_join(th[t], 0);
    } else {
        for (int t = 0; t < threads; ++t) { free(jobs[t].loc.bit); free(jobs[t].loc.gap); }
    }
    free(thr); free(jobs); free(th);
    if (rc) { orc_synth_free(s); return rc; }
    *out = s;
    return BMB200_OK;
}

void orc_synth_packed(const struct orc_synth* s, bmb200_packed_set* out)
{
    out->n_vec = s->n_vec; out->n_blocks = s->n_blocks;
    out->desc = s->desc; out->bit_base = s->bit_base; out->gap_base = s->gap_base;
    out->bit_pool = s->bit_pool; out->gap_pool = s->gap_pool;
}
