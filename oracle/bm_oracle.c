/*
 * bm_oracle.c -- TEST INFRASTRUCTURE ONLY (see bm_oracle.h).
 *
 * CPU restatement, in plain C, of the reference algorithms on the hot path.  Every
 * function cites the reference lines (relative to the reference tree) it follows.
 * Arithmetic is u16/u32/u64 integer only; results must be bit-exact.
 */
#include "bm_oracle_int.h"
#include <string.h>
#include <stdlib.h>

#define BW   BMB200_BLOCK_WORDS
#define GMAX BMB200_GAP_MAX_WORDS

static inline uint32_t popc32(uint32_t x) { return (uint32_t)__builtin_popcount(x); }

/* ------------------------------------------------------------------ */
/* bit_block_count: src/bmfunc.h:5808 (-> avx2_bit_count src/bmavx2.h:156) */
uint32_t orc_bit_block_count(const uint32_t* blk)
{
    uint32_t c = 0;
    for (unsigned i = 0; i < BW; ++i) c += popc32(blk[i]);
    return c;
}

/* calc_block_digest0: src/bmfunc.h:1239 -- bit w set iff wave w (32 words) has a set bit */
uint64_t orc_block_digest(const uint32_t* blk)
{
    uint64_t d = 0;
    for (unsigned w = 0; w < 64; ++w) {
        uint32_t acc = 0;
        for (unsigned k = 0; k < 32; ++k) acc |= blk[w * 32 + k];
        if (acc) d |= (1ull << w);
    }
    return d;
}

/* bit_block_calc_change: src/bmfunc.h:6040 (bit_block_change32 :5876) -- number of runs */
uint32_t orc_bit_block_calc_change(const uint32_t* blk)
{
    uint32_t runs = 1;
    uint32_t prev = blk[0] & 1u;
    for (unsigned i = 0; i < BW; ++i) {
        uint32_t w = blk[i];
        /* transitions inside the word: bit k vs bit k-1 (k>=1), plus bit0 vs previous word's bit31 */
        uint32_t x = w ^ ((w << 1) | prev);
        runs += popc32(x);
        prev = w >> 31;
    }
    return runs;
}

/* gap_calc_level: src/bmfunc.h:5418 with the default table {128,256,512,1280} (src/bmconst.h:396-403) */
static int gap_calc_level(uint32_t len)
{
    if (len <= 128 - 4) return 0;
    if (len <= 256 - 4) return 1;
    if (len <= 512 - 4) return 2;
    if (len <= 1280 - 4) return 3;
    return -1;
}

/* bit_block_to_gap: src/bmfunc.h:5540-5617.  Emits header | run-ends; returns len (number of runs).
 * The capacity level bits are set like blocks_manager::allocate_gap_block + set_gap_level
 * (src/bmblocks.h:1394-1403). dest must hold >= 65537 u16 in the worst case; callers only use it
 * when calc_change < 1276. */
uint32_t orc_bit_to_gap(uint16_t* dest, const uint32_t* blk)
{
    uint32_t bitval = blk[0] & 1u;
    uint32_t first = bitval;
    uint32_t len = 0;              /* number of run-ends written so far */
    for (uint32_t pos = 0; pos < 65536; ) {
        uint32_t w = blk[pos >> 5];
        if ((pos & 31) == 0 && (w == 0 || w == ~0u)) {
            uint32_t v = w & 1u;
            if (v != bitval) { dest[++len] = (uint16_t)(pos - 1); bitval = v; }
            pos += 32;
            continue;
        }
        uint32_t v = (w >> (pos & 31)) & 1u;
        if (v != bitval) { dest[++len] = (uint16_t)(pos - 1); bitval = v; }
        ++pos;
    }
    dest[++len] = 65535;
    int level = gap_calc_level(len);
    if (level < 0) level = 3;
    dest[0] = (uint16_t)(first | ((uint32_t)level << 1) | (len << 3));
    return len;
}

/* ---- bit-range primitives: or_bit_block / sub_bit_block / xor_bit_block src/bmfunc.h:4526,4568,4611 ---- */
static void range_or(uint32_t* blk, uint32_t from, uint32_t to)   /* inclusive */
{
    for (uint32_t w = from >> 5; w <= (to >> 5); ++w) {
        uint32_t lo = (w == (from >> 5)) ? (from & 31) : 0;
        uint32_t hi = (w == (to >> 5)) ? (to & 31) : 31;
        uint32_t m = (~0u << lo) & (~0u >> (31 - hi));
        blk[w] |= m;
    }
}
static void range_sub(uint32_t* blk, uint32_t from, uint32_t to)
{
    for (uint32_t w = from >> 5; w <= (to >> 5); ++w) {
        uint32_t lo = (w == (from >> 5)) ? (from & 31) : 0;
        uint32_t hi = (w == (to >> 5)) ? (to & 31) : 31;
        uint32_t m = (~0u << lo) & (~0u >> (31 - hi));
        blk[w] &= ~m;
    }
}
static void range_xor(uint32_t* blk, uint32_t from, uint32_t to)
{
    for (uint32_t w = from >> 5; w <= (to >> 5); ++w) {
        uint32_t lo = (w == (from >> 5)) ? (from & 31) : 0;
        uint32_t hi = (w == (to >> 5)) ? (to & 31) : 31;
        uint32_t m = (~0u << lo) & (~0u >> (31 - hi));
        blk[w] ^= m;
    }
}

/* iterate runs of a GAP block (format: src/bmfunc.h:1696-1725, 3079-3098):
 * run k (1-based) covers (buf[k-1], buf[k]] with buf[0] := -1; value = first ^ ((k-1)&1) */
#define GAP_FOR_RUNS(gap, VAL, FROM, TO, BODY)                              \
    do {                                                                    \
        uint32_t _len = (uint32_t)((gap)[0] >> 3);                          \
        uint32_t _first = (gap)[0] & 1u;                                    \
        uint32_t _prev_end = 0xFFFFFFFFu;                                   \
        for (uint32_t _k = 1; _k <= _len; ++_k) {                           \
            uint32_t FROM = _prev_end + 1u;                                 \
            uint32_t TO = (gap)[_k];                                        \
            uint32_t VAL = _first ^ ((_k - 1u) & 1u);                       \
            BODY;                                                           \
            _prev_end = TO;                                                 \
        }                                                                   \
    } while (0)

/* gap_convert_to_bitset: src/bmfunc.h:5232 */
void orc_gap_convert_to_bitset(uint32_t* blk, const uint16_t* gap)
{
    memset(blk, 0, BMB200_BLOCK_BYTES);
    orc_gap_add_to_bitset(blk, gap);
}
/* gap_add_to_bitset: src/bmfunc.h:4795-4820 -- OR every 1-run */
void orc_gap_add_to_bitset(uint32_t* blk, const uint16_t* gap)
{
    GAP_FOR_RUNS(gap, v, a, b, { if (v) range_or(blk, a, b); });
}
/* gap_and_to_bitset: src/bmfunc.h:4847,4884-4943 -- clear every 0-run (digest only skips work) */
void orc_gap_and_to_bitset(uint32_t* blk, const uint16_t* gap)
{
    GAP_FOR_RUNS(gap, v, a, b, { if (!v) range_sub(blk, a, b); });
}
/* gap_sub_to_bitset: src/bmfunc.h:4669,4700-4756 -- clear every 1-run */
void orc_gap_sub_to_bitset(uint32_t* blk, const uint16_t* gap)
{
    GAP_FOR_RUNS(gap, v, a, b, { if (v) range_sub(blk, a, b); });
}
/* gap_xor_to_bitset: src/bmfunc.h:4768 -- flip every 1-run */
void orc_gap_xor_to_bitset(uint32_t* blk, const uint16_t* gap)
{
    GAP_FOR_RUNS(gap, v, a, b, { if (v) range_xor(blk, a, b); });
}
/* gap_bit_count: src/bmfunc.h:3079-3098 */
uint32_t orc_gap_bit_count(const uint16_t* gap)
{
    uint32_t c = 0;
    GAP_FOR_RUNS(gap, v, a, b, { if (v) c += b - a + 1u; });
    return c;
}
/* gap_bfind: src/bmfunc.h:1844-1893 -- smallest k >= 1 with buf[k] >= pos; is_set = value of run k */
uint32_t orc_gap_bfind(const uint16_t* gap, uint32_t pos, uint32_t* is_set)
{
    uint32_t lo = 1, hi = (uint32_t)(gap[0] >> 3);
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (gap[mid] < pos) lo = mid + 1; else hi = mid;
    }
    if (is_set) *is_set = (gap[0] & 1u) ^ ((lo - 1u) & 1u);
    return lo;
}
/* gap_bit_count_range: src/bmfunc.h:3173-3209 (closed range) */
uint32_t orc_gap_bit_count_range(const uint16_t* gap, uint32_t left, uint32_t right)
{
    uint32_t c = 0;
    GAP_FOR_RUNS(gap, v, a, b, {
        if (v && b >= left && a <= right) {
            uint32_t lo = a < left ? left : a;
            uint32_t hi = b > right ? right : b;
            c += hi - lo + 1u;
        }
    });
    return c;
}
/* bit_block_calc_count_range / _count_to: src/bmfunc.h:6146-6281 (closed range) */
uint32_t orc_bit_block_count_range(const uint32_t* blk, uint32_t left, uint32_t right)
{
    uint32_t c = 0;
    for (uint32_t w = left >> 5; w <= (right >> 5); ++w) {
        uint32_t lo = (w == (left >> 5)) ? (left & 31) : 0;
        uint32_t hi = (w == (right >> 5)) ? (right & 31) : 31;
        uint32_t m = (~0u << lo) & (~0u >> (31 - hi));
        c += popc32(blk[w] & m);
    }
    return c;
}

/* ------------------------------------------------------------------ */
/* packed-set accessors */
static inline uint32_t set_desc(const bmb200_packed_set* s, uint32_t v, uint32_t nb)
{
    return s->desc[(size_t)nb * s->n_vec + v];
}
static inline const uint32_t* set_bit_ptr(const bmb200_packed_set* s, uint32_t nb, uint32_t rel)
{
    return s->bit_pool + (s->bit_base[nb] + rel) * (size_t)BW;
}
static inline const uint16_t* set_gap_ptr(const bmb200_packed_set* s, uint32_t nb, uint32_t rel)
{   /* rel = desc >> 2: low 29 bits = 16-byte unit, bit 29 (= desc bit 31) = one u16 of lead padding */
    return s->gap_pool + (s->gap_base[nb] + (rel & BMB200_DESC_REL_MASK)) * (size_t)BMB200_GAP_UNIT_WORDS + (rel >> 29);
}

void orc_expand_block(const bmb200_packed_set* s, uint32_t vec, uint32_t nb, uint32_t* out)
{
    uint32_t d = set_desc(s, vec, nb);
    uint32_t kind = d & 3u, rel = d >> 2;
    switch (kind) {
    case BMB200_BLK_NULL: memset(out, 0, BMB200_BLOCK_BYTES); break;
    case BMB200_BLK_FULL: memset(out, 0xFF, BMB200_BLOCK_BYTES); break;
    case BMB200_BLK_BIT:  memcpy(out, set_bit_ptr(s, nb, rel), BMB200_BLOCK_BYTES); break;
    default:              orc_gap_convert_to_bitset(out, set_gap_ptr(s, nb, rel)); break;
    }
}

static int is_all_one(const uint32_t* blk)
{
    for (unsigned i = 0; i < BW; ++i) if (blk[i] != ~0u) return 0;
    return 1;
}

/*
 * One block column of aggregator::combine_or (src/bmaggregator.h:1626-1663):
 *   sort_input_blocks_or (:2278-2310): GAP sources / bit sources, NULL skipped, any FULL => FULL
 *   process_bit_blocks_or (:1924-1988): copy first, OR the rest; the all-ones test only runs inside
 *       the OR calls, i.e. when there are >= 2 bit sources
 *   process_gap_blocks_or (:1808-1814): gap_add_to_bitset per GAP source
 * returns: 0 = no block stored (NULL), 1 = FULL, 2 = block in tb
 */
static int column_or(const bmb200_packed_set* s, uint32_t nb, const uint32_t* g, uint32_t n, uint32_t* tb)
{
    uint32_t nbit = 0, ngap = 0;
    for (uint32_t k = 0; k < n; ++k) {
        uint32_t kind = set_desc(s, g[k], nb) & 3u;
        if (kind == BMB200_BLK_FULL) return 1;
        if (kind == BMB200_BLK_BIT) ++nbit;
        if (kind == BMB200_BLK_GAP) ++ngap;
    }
    if (!nbit && !ngap) return 0;
    memset(tb, 0, BMB200_BLOCK_BYTES);
    for (uint32_t k = 0; k < n; ++k) {
        uint32_t d = set_desc(s, g[k], nb);
        if ((d & 3u) != BMB200_BLK_BIT) continue;
        const uint32_t* b = set_bit_ptr(s, nb, d >> 2);
        for (unsigned i = 0; i < BW; ++i) tb[i] |= b[i];
    }
    if (nbit >= 2 && is_all_one(tb)) return 1;
    for (uint32_t k = 0; k < n; ++k) {
        uint32_t d = set_desc(s, g[k], nb);
        if ((d & 3u) != BMB200_BLK_GAP) continue;
        orc_gap_add_to_bitset(tb, set_gap_ptr(s, nb, d >> 2));
    }
    return 2;
}

/*
 * One block column of combine_and / combine_and_sub (src/bmaggregator.h:1668-1716, 1720-1803):
 *   sort_input_blocks_and (:2315-2366): any NULL AND source => empty; FULL sources dropped;
 *       only-FULL sources => one "real" all-ones block
 *   SUB group through sort_input_blocks_or: any FULL => empty; NULL skipped
 *   only-FULL AND sources and no SUB group given => FULL result (:1751-1758, :1683-1695)
 *   process_bit_blocks_and/_sub (:1994-2205), process_gap_blocks_and/_sub (:1820-1890)
 * returns 0 = empty, 1 = FULL (is_result_full), 2 = block in tb (may still be all-zero: digest decides)
 */
static int column_and_sub(const bmb200_packed_set* s, uint32_t nb,
                          const uint32_t* ga, uint32_t na, const uint32_t* gs, uint32_t ns, uint32_t* tb)
{
    uint32_t nbit = 0, ngap = 0;
    for (uint32_t k = 0; k < na; ++k) {
        uint32_t kind = set_desc(s, ga[k], nb) & 3u;
        if (kind == BMB200_BLK_NULL) return 0;
        if (kind == BMB200_BLK_BIT) ++nbit;
        if (kind == BMB200_BLK_GAP) ++ngap;
    }
    if (na == 0) return 0;
    for (uint32_t k = 0; k < ns; ++k)
        if ((set_desc(s, gs[k], nb) & 3u) == BMB200_BLK_FULL) return 0;
    if (!nbit && !ngap && ns == 0) return 1;
    memset(tb, 0xFF, BMB200_BLOCK_BYTES);
    for (uint32_t k = 0; k < na; ++k) {
        uint32_t d = set_desc(s, ga[k], nb);
        if ((d & 3u) == BMB200_BLK_BIT) {
            const uint32_t* b = set_bit_ptr(s, nb, d >> 2);
            for (unsigned i = 0; i < BW; ++i) tb[i] &= b[i];
        }
    }
    for (uint32_t k = 0; k < ns; ++k) {
        uint32_t d = set_desc(s, gs[k], nb);
        if ((d & 3u) == BMB200_BLK_BIT) {
            const uint32_t* b = set_bit_ptr(s, nb, d >> 2);
            for (unsigned i = 0; i < BW; ++i) tb[i] &= ~b[i];
        }
    }
    for (uint32_t k = 0; k < na; ++k) {
        uint32_t d = set_desc(s, ga[k], nb);
        if ((d & 3u) == BMB200_BLK_GAP) orc_gap_and_to_bitset(tb, set_gap_ptr(s, nb, d >> 2));
    }
    for (uint32_t k = 0; k < ns; ++k) {
        uint32_t d = set_desc(s, gs[k], nb);
        if ((d & 3u) == BMB200_BLK_GAP) orc_gap_sub_to_bitset(tb, set_gap_ptr(s, nb, d >> 2));
    }
    return 2;
}

/* N-way XOR column: bit_block_xor src/bmfunc.h:9191, gap_xor_to_bitset :4768; FULL = invert, NULL skipped
 * (combine_operation_block_xor src/bm.h:7359). returns 0 when no non-NULL source, else 2 */
static int column_xor(const bmb200_packed_set* s, uint32_t nb, const uint32_t* g, uint32_t n, uint32_t* tb)
{
    int any = 0;
    memset(tb, 0, BMB200_BLOCK_BYTES);
    for (uint32_t k = 0; k < n; ++k) {
        uint32_t d = set_desc(s, g[k], nb);
        switch (d & 3u) {
        case BMB200_BLK_NULL: break;
        case BMB200_BLK_FULL: for (unsigned i = 0; i < BW; ++i) tb[i] = ~tb[i]; any = 1; break;
        case BMB200_BLK_BIT: {
            const uint32_t* b = set_bit_ptr(s, nb, d >> 2);
            for (unsigned i = 0; i < BW; ++i) tb[i] ^= b[i];
            any = 1; break; }
        default: orc_gap_xor_to_bitset(tb, set_gap_ptr(s, nb, d >> 2)); any = 1; break;
        }
    }
    return any ? 2 : 0;
}

/* aggregator::combine_shift_right_and, src/bmaggregator.h:2494-2669, restated with its own data flow: blocks in order,
 * one carry bit per source handed from block to block (carry_overs[], :2485-2489); per block the first source is
 * copied, every further source k does  blk = shift_r1(blk, carry_overs[k]) & arg  (process_shift_right_and :2634-2693,
 * bit_block_shift_r1 src/bmfunc.h:6391-6410: w = (w << 1) | carry).  The digest shortcuts of the reference only skip
 * work on empty blocks and are not restated.  A non-empty block is stored through opt_copy_bit_block(opt_mode). */
static int orc_shift_right_and(const bmb200_packed_set* s, const bmb200_agg_args* a,
                               uint8_t* kind, uint32_t* popcnt, uint64_t* digest, uint32_t* nruns, uint32_t* blocks, uint16_t* gaps)
{
    uint32_t nb_to = a->nb_to ? a->nb_to : s->n_blocks;
    int compress = (a->flags & BMB200_F_OPT_COMPRESS) != 0;
    uint32_t n = a->n0;
    uint32_t* tb = (uint32_t*)malloc(BMB200_BLOCK_BYTES);
    uint32_t* arg = (uint32_t*)malloc(BMB200_BLOCK_BYTES);
    uint16_t* tg = (uint16_t*)malloc(sizeof(uint16_t) * 65540);
    uint8_t* carry = (uint8_t*)calloc(n ? n : 1, 1);
    if (!tb || !arg || !tg || !carry) { free(tb); free(arg); free(tg); free(carry); return BMB200_ERR_BADALLOC; }
    /* blocks before nb_from still feed carries into the range: walk from block 0 */
    for (uint32_t nb = 0; nb < nb_to; ++nb) {
        if (n) orc_expand_block(s, a->group0[0], nb, tb); else memset(tb, 0, BMB200_BLOCK_BYTES);
        carry[0] = 0;
        for (uint32_t k = 1; k < n; ++k) {
            uint32_t co = carry[k];
            for (uint32_t i = 0; i < BW; ++i) { uint32_t w = tb[i]; uint32_t co1 = w >> 31; tb[i] = (w << 1) | co; co = co1; }
            carry[k] = (uint8_t)co;
            orc_expand_block(s, a->group0[k], nb, arg);
            for (uint32_t i = 0; i < BW; ++i) tb[i] &= arg[i];
        }
        if (nb < a->nb_from) continue;
        uint32_t c = nb - a->nb_from;
        uint32_t pc = orc_bit_block_count(tb), nr = orc_bit_block_calc_change(tb);
        uint64_t dg = orc_block_digest(tb);
        uint8_t kd;
        if (dg == 0) kd = BMB200_BLK_NULL;                     /* if (digest) ... opt_copy_bit_block, :2617-2632 */
        else if (!compress) kd = BMB200_BLK_BIT;
        else if (nr == 1) kd = BMB200_BLK_FULL;
        else if (nr < BMB200_GAP_THRESHOLD) kd = BMB200_BLK_GAP;
        else kd = BMB200_BLK_BIT;
        if (kind) kind[c] = kd;
        if (popcnt) popcnt[c] = pc;
        if (digest) digest[c] = dg;
        if (nruns) nruns[c] = nr;
        if (blocks) memcpy(blocks + (size_t)c * BW, tb, BMB200_BLOCK_BYTES);
        if (gaps) {
            uint16_t* gout = gaps + (size_t)c * GMAX;
            memset(gout, 0, sizeof(uint16_t) * GMAX);
            if (kd == BMB200_BLK_GAP) { uint32_t len = orc_bit_to_gap(tg, tb); memcpy(gout, tg, sizeof(uint16_t) * (len + 1)); }
        }
    }
    free(tb); free(arg); free(tg); free(carry);
    return BMB200_OK;
}

int orc_aggregate(const bmb200_packed_set* s, const bmb200_agg_args* a,
                  uint8_t* kind, uint32_t* popcnt, uint64_t* digest, uint32_t* nruns,
                  uint32_t* blocks, uint16_t* gaps)
{
    if (!s || !a) return BMB200_ERR_BADARG;
    if (a->op == BMB200_OP_SHIFT_R_AND) {
        uint32_t nbt = a->nb_to ? a->nb_to : s->n_blocks;
        if (a->nb_from > nbt || nbt > s->n_blocks) return BMB200_ERR_RANGE;
        for (uint32_t k = 0; k < a->n0; ++k) if (a->group0[k] >= s->n_vec) return BMB200_ERR_RANGE;
        return orc_shift_right_and(s, a, kind, popcnt, digest, nruns, blocks, gaps);
    }
    uint32_t nb_to = a->nb_to ? a->nb_to : s->n_blocks;
    if (a->nb_from > nb_to || nb_to > s->n_blocks) return BMB200_ERR_RANGE;
    for (uint32_t k = 0; k < a->n0; ++k) if (a->group0[k] >= s->n_vec) return BMB200_ERR_RANGE;
    if (a->op == BMB200_OP_AND_SUB)
        for (uint32_t k = 0; k < a->n1; ++k) if (a->group1[k] >= s->n_vec) return BMB200_ERR_RANGE;
    int compress = (a->flags & BMB200_F_OPT_COMPRESS) != 0;
    uint32_t* tb = (uint32_t*)malloc(BMB200_BLOCK_BYTES);
    uint16_t* tg = (uint16_t*)malloc(sizeof(uint16_t) * 65540);
    if (!tb || !tg) { free(tb); free(tg); return BMB200_ERR_BADALLOC; }

    for (uint32_t nb = a->nb_from; nb < nb_to; ++nb) {
        uint32_t c = nb - a->nb_from;
        int r;
        switch (a->op) {
        case BMB200_OP_OR:      r = column_or(s, nb, a->group0, a->n0, tb); break;
        case BMB200_OP_AND:     r = column_and_sub(s, nb, a->group0, a->n0, 0, 0, tb); break;
        case BMB200_OP_AND_SUB: r = column_and_sub(s, nb, a->group0, a->n0, a->group1, a->n1, tb); break;
        case BMB200_OP_XOR:     r = column_xor(s, nb, a->group0, a->n0, tb); break;
        default: free(tb); free(tg); return BMB200_ERR_BADARG;
        }
        if (r == 0) memset(tb, 0, BMB200_BLOCK_BYTES);
        if (r == 1) memset(tb, 0xFF, BMB200_BLOCK_BYTES);
        uint32_t pc = orc_bit_block_count(tb);
        uint64_t dg = orc_block_digest(tb);
        uint32_t nr = orc_bit_block_calc_change(tb);
        uint8_t kd;
        if (r == 0) kd = BMB200_BLK_NULL;
        else if (r == 1) kd = BMB200_BLK_FULL;
        else if (a->op != BMB200_OP_OR && dg == 0) kd = BMB200_BLK_NULL; /* AND/SUB/XOR: digest==0 => nothing stored */
        else if (!compress) kd = BMB200_BLK_BIT;                          /* copy_bit_block src/bmblocks.h:1340 */
        else {                                                            /* opt_copy_bit_block src/bmblocks.h:1355-1409 */
            if (nr == 1) kd = tb[0] ? BMB200_BLK_FULL : BMB200_BLK_NULL;
            else if (nr < BMB200_GAP_THRESHOLD) kd = BMB200_BLK_GAP;
            else kd = BMB200_BLK_BIT;
        }
        if (kind) kind[c] = kd;
        if (popcnt) popcnt[c] = pc;
        if (digest) digest[c] = dg;
        if (nruns) nruns[c] = nr;
        if (blocks) memcpy(blocks + (size_t)c * BW, tb, BMB200_BLOCK_BYTES);
        if (gaps) {
            uint16_t* gout = gaps + (size_t)c * GMAX;
            memset(gout, 0, sizeof(uint16_t) * GMAX);
            if (kd == BMB200_BLK_GAP) {
                uint32_t len = orc_bit_to_gap(tg, tb);
                memcpy(gout, tg, sizeof(uint16_t) * (len + 1));
            }
        }
    }
    free(tb); free(tg);
    return BMB200_OK;
}

/* ------------------------------------------------------------------ */
/* bm::deserialize (src/bmserial.h:4152, deserializer::deserialize :5578-6090) into an EMPTY vector, restated for the
 * tokens whose length is explicit in the stream (what serializer levels 0..2 emit, plus the array tokens of level 3):
 *   zero / one runs (set_block_1zero .. set_block_aone, the 0x80 | n short zero run :5716-5722), set_block_bit :5493,
 *   set_block_bit_1bit :5825, set_block_bit_0runs :4738, set_block_bit_interval :5511, set_block_bit_digest0 (read_digest0_block),
 *   set_block_arrbit :5539, set_block_arrbit_inv :5424, set_block_gap / _gapbit :5243-5300, set_block_arrgap(_inv) :4833-4845,
 *   set_block_gap_egamma_v3 with plain 16-bit values :5050-5082 (bit stream: 32-bit words, LSB first, src/encoding.h:1313,2506).
 * Block kinds follow the reference: bit tokens -> bit-blocks, GAP / array-of-GAP tokens and single bits -> GAP blocks
 * (new_blocks_strat BM_GAP, :5687), capacity level from gap_calc_level(gap_length).  Returns BMB200_ERR_UNSUPPORTED for
 * every other token (gamma / interpolative / XOR / super-block / bookmark encodings). */
static void gap_from_sorted(uint16_t* g, const uint16_t* a, uint32_t n, int invert)
{   /* gap_set_array (src/bmfunc.h) + optional gap_invert: positions a[0..n) ascending */
    uint32_t len = 0, first = (n && a[0] == 0) ? 1u : 0u;
    for (uint32_t k = 0; k < n; ) {
        uint32_t s = a[k], e = s;
        while (k + 1 < n && a[k + 1] == e + 1) { ++k; ++e; }
        ++k;
        if (s > 0) g[++len] = (uint16_t)(s - 1);
        if (e < 65535) g[++len] = (uint16_t)e;
    }
    g[++len] = 65535;
    int level = gap_calc_level(len + 1); if (level < 0) level = 3;
    g[0] = (uint16_t)((first ^ (invert ? 1u : 0u)) | ((uint32_t)level << 1) | (len << 3));
}

static uint32_t* g_token_hist = 0;      /* optional token-type histogram (256 counters) filled by orc_deserialize: test coverage evidence */
void orc_set_token_hist(uint32_t* hist) { g_token_hist = hist; }

int orc_deserialize(const uint8_t* blob, uint64_t size, uint32_t n_cols, uint8_t* kind, uint32_t* blocks, uint16_t* gaps)
{
    rd_t r = { blob, blob + size };
    memset(kind, 0, n_cols);
    if (blocks) memset(blocks, 0, (size_t)n_cols * BMB200_BLOCK_BYTES);
    if (gaps) memset(gaps, 0, (size_t)n_cols * GMAX * 2);
    uint32_t hf = rd8(&r);
    if (!(hf & (1u << 3))) rd8(&r);                       /* byte order (BM_HM_NO_BO) */
    if (hf & (1u << 2)) return BMB200_ERR_UNSUPPORTED;    /* BM_HM_ID_LIST */
    if (hf & (1u << 6)) return BMB200_ERR_UNSUPPORTED;    /* BM_HM_HXOR */
    if (!(hf & (1u << 4))) for (int k = 0; k < 4; ++k) rd16(&r);   /* GAP levels */
    if (hf & (1u << 1)) { if (hf & (1u << 5)) rd64(&r); else rd32(&r); }   /* BM_HM_RESIZE: size (64-bit in a BM_HM_64_BIT stream) */
    uint32_t* tb = (uint32_t*)malloc(BMB200_BLOCK_BYTES);
    uint16_t* tg = (uint16_t*)malloc(sizeof(uint16_t) * 65540);
    uint16_t* arr = (uint16_t*)malloc(sizeof(uint16_t) * 65536);
    if (!tb || !tg || !arr) { free(tb); free(tg); free(arr); return BMB200_ERR_BADALLOC; }
    int rc = BMB200_OK;
    uint64_t nb = 0;
    while (rc == BMB200_OK && r.p <= r.end) {
        uint32_t bt = rd8(&r);
        if (r.p > r.end) { rc = BMB200_ERR_BADARG; break; }
        if (g_token_hist) g_token_hist[bt & 0x80u ? 0x80u : bt]++;
        if (bt & 0x80u) { nb += bt & 0x7fu; continue; }
        uint64_t ones = 0; int is_bit = 0, is_gap = 0;
        switch (bt) {
        case 0: case 9: nb = 1ull << 40; break;                           /* set_block_end / set_block_azero */
        case 1: break;                                                    /* set_block_1zero */
        case 3: nb += rd8(&r); continue;
        case 5: nb += rd16(&r); continue;
        case 7: nb += rd32(&r); continue;
        case 25: nb += rd64(&r); continue;                                /* set_block_64zero (BM64ADDR streams) */
        case 10: ones = (nb < n_cols) ? n_cols - nb : 0; for (uint64_t c = nb; c < n_cols; ++c) kind[c] = BMB200_BLK_FULL; nb = 1ull << 40; break;
        case 2: ones = 1; break;
        case 4: ones = rd8(&r); break;
        case 6: ones = rd16(&r); break;
        case 8: ones = rd32(&r); break;
        case 26: ones = rd64(&r); break;                                  /* set_block_64one */
        case 11: for (uint32_t i = 0; i < BW; ++i) tb[i] = rd32(&r); is_bit = 1; break;
        case 19: { arr[0] = (uint16_t)rd16(&r); gap_from_sorted(tg, arr, 1, 0); is_gap = 1; break; }
        case 22: {
            memset(tb, 0, BMB200_BLOCK_BYTES);
            uint32_t run_type = rd8(&r);
            for (uint32_t j = 0; j < BW && r.p <= r.end; run_type = !run_type) {
                uint32_t run_len = rd16(&r);
                if (run_type) { uint32_t e = j + run_len; if (e > BW) { rc = BMB200_ERR_BADARG; break; } for (; j < e; ++j) tb[j] = rd32(&r); }
                else j += run_len;
            }
            is_bit = 1; break; }
        case 17: {
            uint32_t head = rd16(&r), tail = rd16(&r);
            memset(tb, 0, BMB200_BLOCK_BYTES);
            if (tail >= BW || head > tail) { rc = BMB200_ERR_BADARG; break; }
            for (uint32_t i = head; i <= tail; ++i) tb[i] = rd32(&r);
            is_bit = 1; break; }
        case 34: {
            uint64_t d0 = rd64(&r);
            memset(tb, 0, BMB200_BLOCK_BYTES);
            for (uint32_t w = 0; w < 64; ++w) if ((d0 >> w) & 1u) for (uint32_t i = 0; i < 32; ++i) tb[w * 32 + i] = rd32(&r);
            is_bit = 1; break; }
        case 16: case 30: {
            uint32_t n = rd16(&r);
            memset(tb, bt == 30 ? 0xFF : 0, BMB200_BLOCK_BYTES);
            for (uint32_t k = 0; k < n; ++k) { uint32_t b = rd16(&r); if (bt == 30) tb[b >> 5] &= ~(1u << (b & 31)); else tb[b >> 5] |= 1u << (b & 31); }
            is_bit = 1; break; }
        case 14: case 15: {
            uint32_t hdr = rd16(&r), len = hdr >> 3;
            if (len < 1 || len + 1 > GMAX) { rc = BMB200_ERR_UNSUPPORTED; break; }
            for (uint32_t k = 1; k < len; ++k) tg[k] = (uint16_t)rd16(&r);
            tg[len] = 65535;
            int level = gap_calc_level(len + 1); if (level < 0) { rc = BMB200_ERR_UNSUPPORTED; break; }
            tg[0] = (uint16_t)((hdr & 1u) | ((uint32_t)level << 1) | (len << 3));
            is_gap = 1; break; }
        case 18: case 24: {
            uint32_t n = rd16(&r);
            for (uint32_t k = 0; k < n; ++k) arr[k] = (uint16_t)rd16(&r);
            gap_from_sorted(tg, arr, n, bt == 24);
            if ((tg[0] >> 3) + 1u > GMAX - 4) { rc = BMB200_ERR_UNSUPPORTED; break; }
            is_gap = 1; break; }
        case 47: rd16(&r); continue;                                      /* set_nb_bookmark16/24/32: skip offsets (:5897-5920) */
        case 48: rd16(&r); rd8(&r); continue;
        case 49: rd32(&r); continue;
        case 50: rd8(&r); continue;                                       /* set_nb_sync_mark8..64 */
        case 51: rd16(&r); continue;
        case 52: rd16(&r); rd8(&r); continue;
        case 53: rd32(&r); continue;
        case 54: rd32(&r); rd16(&r); continue;
        case 55: rd64(&r); continue;
        case 56: case 68: {                                               /* super-block position lists (decode_arr_sblock :5458) */
            uint32_t* sarr = (uint32_t*)malloc(sizeof(uint32_t) * 65536u);
            uint32_t slen = 0, sb = 0;
            if (!sarr) { rc = BMB200_ERR_BADALLOC; break; }
            rc = orc_sblock_token(&r, bt, sarr, &slen, &sb);
            if (rc == BMB200_OK && (uint64_t)sb * 256u != (nb & ~255ull)) rc = BMB200_ERR_BADARG;
            if (rc == BMB200_OK) {
                /* bv.set_bit_no_check into empty blocks under BM_GAP strategy: every touched block becomes a GAP block */
                for (uint32_t k = 0; k < slen; ) {
                    uint32_t c = sarr[k] >> 16, k2 = k;
                    if (c >= 256u) { rc = BMB200_ERR_BADARG; break; }
                    memset(tb, 0, BMB200_BLOCK_BYTES);
                    while (k2 < slen && (sarr[k2] >> 16) == c) { tb[(sarr[k2] & 65535u) >> 5] |= 1u << (sarr[k2] & 31u); ++k2; }
                    uint64_t col = (uint64_t)sb * 256u + c;
                    if (col < n_cols) {
                        uint32_t len = orc_bit_to_gap(tg, tb);
                        /* gap_block_set_no_ret (src/bm.h:4800): the block is extended when the run count exceeds gap_limit = glen[level] - 4,
                         * so it stays GAP while len <= 1276 and sits on the smallest level with len <= glen - 4 */
                        int level = gap_calc_level(len);
                        if (level < 0) { kind[col] = BMB200_BLK_BIT; }
                        else { kind[col] = BMB200_BLK_GAP; tg[0] = (uint16_t)((tg[0] & 1u) | ((uint32_t)level << 1) | (len << 3));
                               if (gaps) memcpy(gaps + col * GMAX, tg, ((size_t)len + 1) * 2); }
                        if (blocks) memcpy(blocks + col * BW, tb, BMB200_BLOCK_BYTES);
                    }
                    k = k2;
                }
            }
            free(sarr);
            if (rc != BMB200_OK) break;
            nb = (nb & ~255ull) + 256u;
            continue; }
        case 20: case 21: case 23: case 27: case 28: case 29: case 31: case 32: case 33: case 43: case 44: case 45: case 57:
        case 61: case 62: case 63: case 64: case 65: case 66: case 67: { /* entropy-coded blocks: bm_oracle_entropy.c */
            int gapf = 0;
            memset(tb, 0, BMB200_BLOCK_BYTES);
            rc = orc_entropy_token(&r, bt, tb, &gapf);
            if (rc != BMB200_OK) break;
            if (gapf) {                                                   /* clone_gap_block: GAP unless it does not fit (src/bmblocks.h:838) */
                uint32_t len = orc_bit_to_gap(tg, tb);
                int level = gap_calc_level(len + 1);
                if (level < 0) is_bit = 1;
                else { tg[0] = (uint16_t)((tg[0] & 1u) | ((uint32_t)level << 1) | (len << 3)); is_gap = 1; }
            } else is_bit = 1;
            break; }
        default: rc = BMB200_ERR_UNSUPPORTED; break;
        }
        if (rc != BMB200_OK) break;
        if (r.p > r.end) { rc = BMB200_ERR_BADARG; break; }
        if (ones && bt != 10) { for (uint64_t c = nb; c < nb + ones && c < n_cols; ++c) kind[c] = BMB200_BLK_FULL; nb += ones; continue; }
        if (nb < n_cols) {
            if (is_bit) { kind[nb] = BMB200_BLK_BIT; if (blocks) memcpy(blocks + nb * BW, tb, BMB200_BLOCK_BYTES); }
            if (is_gap) {
                kind[nb] = BMB200_BLK_GAP;
                if (gaps) memcpy(gaps + nb * GMAX, tg, ((size_t)(tg[0] >> 3) + 1) * 2);
                if (blocks) orc_gap_convert_to_bitset(blocks + nb * BW, tg);
            }
        }
        if (nb >= (1ull << 40)) break;
        ++nb;
    }
    if (blocks) for (uint32_t c = 0; c < n_cols; ++c) if (kind[c] == BMB200_BLK_FULL) memset(blocks + (size_t)c * BW, 0xFF, BMB200_BLOCK_BYTES);
    free(tb); free(tg); free(arr);
    return rc;
}

/* ------------------------------------------------------------------ */
/* sparse_vector_scanner searches, restated from their contract (src/bmsparsevec_algo.h:1083-1176: "search result is a
 * vector of 1s when sv[i] == / > / >= / < / <= value", find_range: closed interval [from, to]; NULL elements and
 * indexes >= size() never match, :2426 finalize_search_result, :1686 invert_internal).  Element i of the sparse
 * vector is rebuilt from its bit-planes (sparse_vector::get_slice(j) holds bit j, src/bmsparsevec.h) and compared as
 * an unsigned integer.  Pinned against the real scanner through oracle/_ref (tests/test_oracle_vs_reference.py). */
int orc_scan(const bmb200_packed_set* s, const bmb200_scan_args* a,
             uint8_t* kind, uint32_t* popcnt, uint64_t* digest, uint32_t* nruns, uint32_t* blocks, uint16_t* gaps)
{
    if (!s || !a || !a->values || !a->n_values || !a->n_planes || a->n_planes > 64) return BMB200_ERR_BADARG;
    if ((uint64_t)a->plane0 + a->n_planes > s->n_vec) return BMB200_ERR_RANGE;
    if (a->universe != 0xffffffffu && a->universe >= s->n_vec) return BMB200_ERR_RANGE;
    uint32_t nb_to = a->nb_to ? a->nb_to : s->n_blocks;
    if (a->nb_from >= nb_to || nb_to > s->n_blocks) return BMB200_ERR_RANGE;
    int compress = (a->flags & BMB200_F_OPT_COMPRESS) != 0;
    uint32_t n_cols = nb_to - a->nb_from;
    uint32_t* planes = (uint32_t*)malloc((size_t)BMB200_BLOCK_BYTES * (a->n_planes + 1));
    uint64_t* elem = (uint64_t*)malloc(sizeof(uint64_t) * BMB200_BLOCK_BITS);
    uint32_t* tb = (uint32_t*)malloc(BMB200_BLOCK_BYTES);
    uint16_t* tg = (uint16_t*)malloc(sizeof(uint16_t) * 65540);
    if (!planes || !elem || !tb || !tg) { free(planes); free(elem); free(tb); free(tg); return BMB200_ERR_BADALLOC; }
    for (uint32_t nb = a->nb_from; nb < nb_to; ++nb) {
        uint32_t* uni = planes + (size_t)a->n_planes * BW;
        for (uint32_t j = 0; j < a->n_planes; ++j) orc_expand_block(s, a->plane0 + j, nb, planes + (size_t)j * BW);
        if (a->universe == 0xffffffffu) memset(uni, 0xFF, BMB200_BLOCK_BYTES); else orc_expand_block(s, a->universe, nb, uni);
        for (uint32_t i = 0; i < BMB200_BLOCK_BITS; ++i) {
            uint64_t e = 0;
            for (uint32_t j = 0; j < a->n_planes; ++j) e |= (uint64_t)((planes[(size_t)j * BW + (i >> 5)] >> (i & 31)) & 1u) << j;
            elem[i] = e;
        }
        for (uint32_t k = 0; k < a->n_values; ++k) {
            uint64_t va = a->values[a->pred == BMB200_SCAN_RANGE ? 2 * k : k];
            uint64_t vb = a->pred == BMB200_SCAN_RANGE ? a->values[2 * k + 1] : 0;
            if (a->pred == BMB200_SCAN_RANGE && vb < va) { uint64_t t = va; va = vb; vb = t; }   /* find_range swaps, :2871-2872 */
            memset(tb, 0, BMB200_BLOCK_BYTES);
            for (uint32_t i = 0; i < BMB200_BLOCK_BITS; ++i) {
                if (!((uni[i >> 5] >> (i & 31)) & 1u)) continue;
                uint64_t e = elem[i]; int hit;
                switch (a->pred) {
                case BMB200_SCAN_EQ: hit = e == va; break;
                case BMB200_SCAN_GT: hit = e >  va; break;
                case BMB200_SCAN_GE: hit = e >= va; break;
                case BMB200_SCAN_LT: hit = e <  va; break;
                case BMB200_SCAN_LE: hit = e <= va; break;
                case BMB200_SCAN_RANGE: hit = e >= va && e <= vb; break;
                default: free(planes); free(elem); free(tb); free(tg); return BMB200_ERR_BADARG;
                }
                if (hit) tb[i >> 5] |= 1u << (i & 31);
            }
            size_t c = (size_t)k * n_cols + (nb - a->nb_from);
            uint32_t pc = orc_bit_block_count(tb), nr = orc_bit_block_calc_change(tb);
            uint64_t dg = orc_block_digest(tb);
            uint8_t kd;                                   /* stored like an AND-SUB result: empty digest => nothing */
            if (dg == 0) kd = BMB200_BLK_NULL;
            else if (!compress) kd = BMB200_BLK_BIT;
            else if (nr == 1) kd = BMB200_BLK_FULL;
            else if (nr < BMB200_GAP_THRESHOLD) kd = BMB200_BLK_GAP;
            else kd = BMB200_BLK_BIT;
            if (kind) kind[c] = kd;
            if (popcnt) popcnt[c] = pc;
            if (digest) digest[c] = dg;
            if (nruns) nruns[c] = nr;
            if (blocks) memcpy(blocks + c * BW, tb, BMB200_BLOCK_BYTES);
            if (gaps) {
                uint16_t* gout = gaps + c * GMAX;
                memset(gout, 0, sizeof(uint16_t) * GMAX);
                if (kd == BMB200_BLK_GAP) { uint32_t len = orc_bit_to_gap(tg, tb); memcpy(gout, tg, sizeof(uint16_t) * (len + 1)); }
            }
        }
    }
    free(planes); free(elem); free(tb); free(tg);
    return BMB200_OK;
}

/* ------------------------------------------------------------------ */
/* build_rs_index: src/bm.h:2531-2660; borders src/bmconst.h:120-124 */
#define RS3_B0   21824u
#define RS3_B1   43648u
#define RS3_B0_1 32736u
#define RS3_B1_1 54560u

int orc_rs_build(const bmb200_packed_set* s, uint32_t vec,
                 uint32_t* bcount, uint64_t* sub_count, uint64_t* sb_count)
{
    if (!s || vec >= s->n_vec) return BMB200_ERR_BADARG;
    uint32_t* tb = (uint32_t*)malloc(BMB200_BLOCK_BYTES);
    if (!tb) return BMB200_ERR_BADALLOC;
    uint64_t run = 0;
    uint32_t nsb = (s->n_blocks + 255u) / 256u;
    if (sb_count) sb_count[0] = 0;
    for (uint32_t nb = 0; nb < s->n_blocks; ++nb) {
        uint32_t d = set_desc(s, vec, nb);
        uint32_t kind = d & 3u;
        uint32_t first = 0, second = 0, third = 0; uint64_t aux0 = 0, aux1 = 0;
        if (kind == BMB200_BLK_GAP) {
            const uint16_t* g = set_gap_ptr(s, nb, d >> 2);
            uint32_t is_set;
            first  = orc_gap_bit_count_range(g, 0, RS3_B0);
            second = orc_gap_bit_count_range(g, RS3_B0 + 1, RS3_B1);
            third  = orc_gap_bit_count_range(g, RS3_B1 + 1, 65535);
            aux0 = orc_gap_bfind(g, RS3_B0 + 1, &is_set); aux0 = (aux0 << 1) | (is_set ? 1 : 0);
            aux1 = orc_gap_bfind(g, RS3_B1 + 1, &is_set); aux1 = (aux1 << 1) | (is_set ? 1 : 0);
        } else if (kind != BMB200_BLK_NULL) {
            orc_expand_block(s, vec, nb, tb);
            first  = orc_bit_block_count_range(tb, 0, RS3_B0);
            second = orc_bit_block_count_range(tb, RS3_B0 + 1, RS3_B1);
            third  = orc_bit_block_count_range(tb, RS3_B1 + 1, 65535);
            aux0 = orc_bit_block_count_range(tb, 0, RS3_B0_1);
            aux1 = orc_bit_block_count_range(tb, 0, RS3_B1_1);
        }
        uint32_t cnt = first + second + third;
        if (bcount) bcount[nb] = cnt;
        if (sub_count) sub_count[nb] = (uint64_t)(first | (second << 16)) | (aux0 << 32) | (aux1 << 48);
        run += cnt;
        if (sb_count && ((nb & 255u) == 255u || nb + 1 == s->n_blocks)) sb_count[(nb >> 8) + 1] = run;
    }
    (void)nsb;
    free(tb);
    return BMB200_OK;
}

/* count_to: src/bm.h:3120-3167 -- bits set in [0, pos] (the anchor scheme :2686-2869 only shortens the scan) */
int orc_rank_batch(const bmb200_packed_set* s, uint32_t vec,
                   const uint64_t* pos, uint64_t n, uint64_t* out)
{
    if (!s || vec >= s->n_vec) return BMB200_ERR_BADARG;
    uint64_t* pre = (uint64_t*)malloc(sizeof(uint64_t) * ((size_t)s->n_blocks + 1));
    uint32_t* tb = (uint32_t*)malloc(BMB200_BLOCK_BYTES);
    if (!pre || !tb) { free(pre); free(tb); return BMB200_ERR_BADALLOC; }
    pre[0] = 0;
    for (uint32_t nb = 0; nb < s->n_blocks; ++nb) {
        uint32_t d = set_desc(s, vec, nb), c = 0;
        switch (d & 3u) {
        case BMB200_BLK_FULL: c = 65536; break;
        case BMB200_BLK_BIT:  c = orc_bit_block_count(set_bit_ptr(s, nb, d >> 2)); break;
        case BMB200_BLK_GAP:  c = orc_gap_bit_count(set_gap_ptr(s, nb, d >> 2)); break;
        default: break;
        }
        pre[nb + 1] = pre[nb] + c;
    }
    for (uint64_t q = 0; q < n; ++q) {
        uint64_t nb = pos[q] >> 16;
        if (nb >= s->n_blocks) { out[q] = pre[s->n_blocks]; continue; }
        uint32_t d = set_desc(s, vec, (uint32_t)nb), in = (uint32_t)(pos[q] & 65535u), c = 0;
        switch (d & 3u) {
        case BMB200_BLK_FULL: c = in + 1; break;
        case BMB200_BLK_BIT:  c = orc_bit_block_count_range(set_bit_ptr(s, (uint32_t)nb, d >> 2), 0, in); break;
        case BMB200_BLK_GAP:  c = orc_gap_bit_count_range(set_gap_ptr(s, (uint32_t)nb, d >> 2), 0, in); break;
        default: break;
        }
        out[q] = pre[nb] + c;
    }
    free(pre); free(tb);
    return BMB200_OK;
}

/* select: src/bm.h:5350-5385 -- position of the rank-th (1-based) set bit; bit_find_rank src/bmfunc.h:9673 */
int orc_select_batch(const bmb200_packed_set* s, uint32_t vec,
                     const uint64_t* rank, uint64_t n, uint64_t* pos, uint8_t* found)
{
    if (!s || vec >= s->n_vec) return BMB200_ERR_BADARG;
    uint64_t* pre = (uint64_t*)malloc(sizeof(uint64_t) * ((size_t)s->n_blocks + 1));
    uint32_t* tb = (uint32_t*)malloc(BMB200_BLOCK_BYTES);
    if (!pre || !tb) { free(pre); free(tb); return BMB200_ERR_BADALLOC; }
    pre[0] = 0;
    for (uint32_t nb = 0; nb < s->n_blocks; ++nb) {
        uint32_t d = set_desc(s, vec, nb), c = 0;
        switch (d & 3u) {
        case BMB200_BLK_FULL: c = 65536; break;
        case BMB200_BLK_BIT:  c = orc_bit_block_count(set_bit_ptr(s, nb, d >> 2)); break;
        case BMB200_BLK_GAP:  c = orc_gap_bit_count(set_gap_ptr(s, nb, d >> 2)); break;
        default: break;
        }
        pre[nb + 1] = pre[nb] + c;
    }
    for (uint64_t q = 0; q < n; ++q) {
        uint64_t r = rank[q];
        if (r == 0 || r > pre[s->n_blocks]) { found[q] = 0; pos[q] = 0; continue; }
        /* first block with pre[nb+1] >= r */
        uint32_t lo = 0, hi = s->n_blocks - 1;
        while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (pre[mid + 1] < r) lo = mid + 1; else hi = mid; }
        uint32_t rr = (uint32_t)(r - pre[lo]);
        orc_expand_block(s, vec, lo, tb);
        uint32_t bit = 0;
        for (uint32_t w = 0; w < BW; ++w) {
            uint32_t c = popc32(tb[w]);
            if (rr > c) { rr -= c; continue; }
            uint32_t x = tb[w];
            for (uint32_t k = 1; k < rr; ++k) x &= x - 1;
            bit = w * 32 + (uint32_t)__builtin_ctz(x);
            break;
        }
        found[q] = 1; pos[q] = ((uint64_t)lo << 16) | bit;
    }
    free(pre); free(tb);
    return BMB200_OK;
}
