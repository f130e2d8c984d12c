/*
 * bm_oracle_entropy.c -- TEST INFRASTRUCTURE ONLY (part of liboracle.so, see bm_oracle.h).
 *
 * Plain-C restatement of the entropy-coded block encodings of the reference serialization format: the bit stream
 * (bm::bit_in, src/encoding.h:299-404), Elias gamma (:2506), gamma8 (:2441), delta16 / delta16s (:2463, :2488),
 * centered-minimal binary interpolative coding (bic_decode_u16_cm :2213, bic_decode_u32_cm :2168), the selective array
 * codec (decode_array :2698, arr_restore_min / arr_restore_min_w src/bmfunc.h:2657, :2526), and the block tokens built
 * from them (deseriaizer_base::read_gap_block src/bmserial.h:4812-5090, read_id_list :4307, read_bic_arr :4363-4477,
 * read_bic_gap :4675, read_bic_sb_arr :4480-4650).  Written recursively, the way the reference states it; the device
 * decoder (bitmagic_b200/csrc/blob_entropy.cuh) is an independent iterative formulation.
 * Pinned against bm::deserialize through oracle/_ref (tests/test_oracle_vs_reference.py) and tests/golden/blobs_entropy.npz.
 */
#include <stdlib.h>
#include <string.h>
#include "bm_oracle_int.h"

/* ---- bm::bit_in: 32-bit little-endian words fetched lazily, bits consumed LSB first ---- */
typedef struct { rd_t* r; uint32_t acc; uint32_t used; } bin_t;
static void bin_init(bin_t* b, rd_t* r) { b->r = r; b->acc = 0; b->used = 32; }

static uint32_t bin_bit(bin_t* b)
{
    if (b->used == 32) { b->acc = rd32(b->r); b->used = 0; }
    uint32_t v = b->acc & 1u; b->acc >>= 1; ++b->used;
    return v;
}
static uint32_t bin_bits(bin_t* b, uint32_t count)          /* get_bits(count), count 1..32 */
{
    uint32_t v = 0;
    for (uint32_t i = 0; i < count; ++i) v |= bin_bit(b) << i;
    return v;
}
static uint32_t bin_16(bin_t* b) { uint32_t lo = bin_bits(b, 8); return lo | (bin_bits(b, 8) << 8); }           /* get_16_no */
static uint32_t bin_24(bin_t* b) { uint32_t v = bin_16(b); return v | (bin_bits(b, 8) << 16); }
static uint32_t bin_32(bin_t* b) { uint32_t v = bin_16(b); return v | (bin_16(b) << 16); }

static uint32_t bin_gamma(bin_t* b)                          /* zeros, a 1, then `zeros` value bits; value | 1 << zeros */
{
    uint32_t zeros = 0;
    while (!bin_bit(b)) { if (++zeros > 32 || b->r->p > b->r->end) return 0; }
    uint32_t v = zeros ? bin_bits(b, zeros) : 0u;
    return zeros >= 32 ? v : (v | (1u << zeros));
}
static uint32_t bin_delta16(bin_t* b)
{
    switch (bin_gamma(b)) {
    case 1: return 511u - bin_bits(b, 8);
    case 2: return 512u + 255u - bin_bits(b, 8);
    case 3: return 512u + 256u + 255u - bin_bits(b, 8);
    default: return bin_16(b);
    }
}
static uint32_t bin_delta16s(bin_t* b) { return bin_bit(b) ? bin_delta16(b) : bin_bits(b, 8); }
static uint32_t bin_gamma8(bin_t* b)
{
    switch (bin_gamma(b)) {
    case 1: return bin_gamma(b);
    case 2: return bin_bits(b, 8);
    case 3: return bin_delta16(b);
    default: return 0;                                       /* 4: zero */
    }
}

/* one centered-minimal code word for the range size r = hi - lo - sz + 1 (src/encoding.h:2224-2237) */
static uint32_t bic_read(bin_t* b, uint32_t r)
{
    if (!r) return 0;
    uint32_t logv = 31u - (uint32_t)__builtin_clz(r + 1u);
    uint32_t c = (uint32_t)((1ull << (logv + 1)) - r - 1);
    int64_t half_c = c >> 1, half_r = r >> 1;
    int64_t lo1 = half_r - half_c - ((r + 1) & 1), hi1 = half_r + half_c + 1;
    uint32_t val = logv ? bin_bits(b, logv) : 0u;
    if ((int64_t)val <= lo1 || (int64_t)val >= hi1) val += bin_bit(b) << logv;
    return val;
}
/* bic_decode_u16_cm: arr[0..sz) strictly increasing values in [lo, hi]; 16-bit wrap-around kept */
static void bic_u16(bin_t* b, uint16_t* arr, uint32_t sz, uint16_t lo, uint16_t hi)
{
    while (sz) {
        if (b->r->p > b->r->end) return;
        uint32_t val = bic_read(b, (uint32_t)hi - lo - sz + 1u);
        uint32_t mid = sz >> 1;
        val += (uint32_t)lo + mid;
        arr[mid] = (uint16_t)val;
        if (sz <= 1) return;
        bic_u16(b, arr, mid, lo, (uint16_t)(val - 1u));
        arr += mid + 1; sz -= mid + 1; lo = (uint16_t)(val + 1u);
    }
}
static void bic_u32(bin_t* b, uint32_t* arr, uint32_t sz, uint32_t lo, uint32_t hi)
{
    while (sz) {
        if (b->r->p > b->r->end) return;
        uint32_t val = bic_read(b, hi - lo - sz + 1u);
        uint32_t mid = sz >> 1;
        val += lo + mid;
        arr[mid] = val;
        if (sz <= 1) return;
        bic_u32(b, arr, mid, lo, val - 1u);
        arr += mid + 1; sz -= mid + 1; lo = val + 1u;
    }
}

/* arr_restore_min_w (src/bmfunc.h:2526-2581), T = u16 */
static void restore_min_w(uint16_t* arr, uint32_t n, uint32_t wlen, uint16_t min0, const uint32_t* wflags)
{
    uint16_t dacc = 0; uint32_t min_w_prev = ~0u;
    for (uint32_t i = 1; i < wlen && i < n; ++i) {
        arr[i] = (uint16_t)(arr[i] + min0 + dacc); dacc = (uint16_t)(dacc + min0);
        uint16_t d = (uint16_t)(arr[i] - arr[i - 1]); if (d < min_w_prev) min_w_prev = d;
    }
    min_w_prev -= (min_w_prev != 0);
    uint32_t wave = 1;
    for (uint32_t i = wlen; i < n; ++wave, i += wlen) {
        if (i + wlen > n) wlen = n % wlen;
        if (!wlen) break;
        int recalc = (wflags[(wave >> 5) & 2047u] >> (wave & 31)) & 1u;
        uint32_t min_w = ~0u;
        for (uint32_t j = 0; j < wlen; ++j) {
            if (recalc) { arr[i + j] = (uint16_t)(arr[i + j] + (uint16_t)(min_w_prev + dacc)); dacc = (uint16_t)(dacc + (uint16_t)min_w_prev); }
            else        { arr[i + j] = (uint16_t)(arr[i + j] + min0 + dacc); dacc = (uint16_t)(dacc + min0); }
            uint16_t d = (uint16_t)(arr[i + j] - arr[i + j - 1]); if (d < min_w) min_w = d;
        }
        min_w_prev = (min_w > min0) ? min_w - 1 : min0;
    }
}

/* bit_in::decode_array (src/encoding.h:2698-2798); arr has room for 65536 values; returns the flag byte, -1 on a bad size */
static int decode_array(bin_t* b, uint16_t* arr, uint32_t* wflags, uint32_t* sz, uint32_t default_sz)
{
    uint32_t h = bin_bits(b, 8);
    if ((h & 3u) == 3u && (h & 0x80u)) { *sz = 0; return (int)h; }
    if ((h & 3u) == 3u) { *sz = 1; arr[0] = (h & 0x40u) ? 0 : (uint16_t)((h & 8u) ? bin_gamma(b) : bin_16(b)); return (int)h; }
    uint32_t n = default_sz ? default_sz : ((h & 8u) ? bin_gamma8(b) + 1u : bin_delta16(b));
    if (n > 65536u) return -1;
    *sz = n;
    uint16_t min0 = (h & 0x40u) ? 0 : (uint16_t)bin_gamma(b);
    if ((h & 3u) == 0) {                                      /* delta-gamma */
        arr[0] = (h & 0x80u) ? 0 : (uint16_t)bin_gamma(b);
        for (uint32_t i = 1; i < n; ++i) arr[i] = (uint16_t)(arr[i - 1] + bin_gamma(b) + min0);
    } else if (h & 2u) {                                      /* gamma */
        uint32_t zc = (h & 0x80u) ? 1u : 0u;
        for (uint32_t i = 0; i < n; ++i) arr[i] = (uint16_t)(bin_gamma(b) - zc + min0);
    } else {                                                  /* BIC with delta-range reduction */
        uint16_t min_v = 0, max_v = 65535; uint32_t s = n; uint16_t* p = arr;
        if (h & 0x80u) {
            min_v = (uint16_t)bin_16(b); max_v = (uint16_t)bin_16(b);
            if (n < 2) return -1;
            arr[0] = min_v; arr[n - 1] = max_v;
            if (n == 2) return (int)h;
            ++min_v; --max_v; s -= 2; ++p;
        }
        if (s) bic_u16(b, p, s, min_v, max_v);
        if (bin_bit(b)) {                                     /* windowed restore */
            memset(wflags, 0, BMB200_BLOCK_BYTES);
            uint32_t win = bin_gamma(b), wcnt = bin_gamma(b);
            wcnt += 15u - 1u; win = (win + 9u) * 2u;
            uint32_t max_wd = n / win + 1u;
            if (wcnt > 65536u) return -1;
            uint16_t* tmp = (uint16_t*)malloc(sizeof(uint16_t) * 65536u);
            if (!tmp) return -1;
            bic_u16(b, tmp, wcnt, 1, (uint16_t)max_wd);
            for (uint32_t k = 0; k < wcnt; ++k) wflags[tmp[k] >> 5] |= 1u << (tmp[k] & 31);
            free(tmp);
            restore_min_w(arr, n, win, min0, wflags);
        } else if (min0) {
            uint16_t dacc = 0;
            for (uint32_t i = 1; i < n; ++i) { arr[i] = (uint16_t)(arr[i] + min0 + dacc); dacc = (uint16_t)(dacc + min0); }
        }
    }
    return (int)h;
}

static void set_pos(uint32_t* tb, uint32_t p, int v) { if (v) tb[(p & 65535u) >> 5] |= 1u << (p & 31); else tb[(p & 65535u) >> 5] &= ~(1u << (p & 31)); }
static void set_run(uint32_t* tb, uint32_t from, uint32_t cnt) { for (uint32_t p = from; p < from + cnt && p < 65536u; ++p) tb[p >> 5] |= 1u << (p & 31); }

/* gap_restore_mins (src/bmfunc.h:3009-3041) */
static void gap_restore_mins(uint16_t* buf, uint16_t min0, uint16_t min1)
{
    uint32_t dsize = buf[0] >> 3, p = 1;
    buf[p] = (uint16_t)(buf[p] + min0);
    uint16_t dacc = min0;
    for (++p; p <= dsize; ) {
        if (p == dsize) break;
        buf[p] = (uint16_t)(buf[p] + min1 + dacc); dacc = (uint16_t)(dacc + min1);
        if (++p < dsize) { buf[p] = (uint16_t)(buf[p] + min0 + dacc); dacc = (uint16_t)(dacc + min0); ++p; }
        else break;
    }
}

/* tg[0] = header (first-run bit in bit 0, len in bits 3..), tg[1..len] run ends -> bits */
static int gap_to_bits(uint32_t* tb, uint16_t* tg, uint32_t len)
{
    if (len < 1 || len > 65536u) return BMB200_ERR_BADARG;
    tg[len] = 65535;
    uint32_t on = tg[0] & 1u, prev = 0;                      /* run k covers [prev, tg[k]] */
    for (uint32_t k = 1; k <= len; ++k) {
        uint32_t e = tg[k];
        if (k > 1 && e < prev) return BMB200_ERR_BADARG;
        if (on) set_run(tb, prev, e - prev + 1u);
        prev = e + 1u; on ^= 1u;
    }
    return BMB200_OK;
}

/* GAP-family and bit-family entropy tokens: fills tb (zeroed by the caller) with the bits of the block.
 * *is_gap = 1 for tokens the reference materialises through a GAP block (deserialize_gap, src/bmserial.h:5222-5395). */
int orc_entropy_token(rd_t* r, uint32_t bt, uint32_t* tb, int* is_gap)
{
    uint16_t* tg = (uint16_t*)malloc(sizeof(uint16_t) * (65536u + 8u));
    uint16_t* a1 = (uint16_t*)malloc(sizeof(uint16_t) * 65536u);
    uint16_t* a2 = (uint16_t*)malloc(sizeof(uint16_t) * 65536u);
    uint32_t* wf = (uint32_t*)malloc(BMB200_BLOCK_BYTES);
    int rc = BMB200_OK;
    if (!tg || !a1 || !a2 || !wf) { rc = BMB200_ERR_BADALLOC; goto done; }
    bin_t b; bin_init(&b, r);
    *is_gap = 0;
    switch (bt) {
    case 20: {                                                /* set_block_gap_egamma (:4856) */
        uint32_t head = rd16(r), len = head >> 3;
        if (len < 1) { rc = BMB200_ERR_BADARG; break; }
        tg[0] = (uint16_t)head;
        uint32_t sum = 0;
        for (uint32_t i = 1; i < len; ++i) { uint32_t v = bin_gamma(&b); sum = (i == 1) ? v - 1u : sum + v; tg[i] = (uint16_t)sum; }
        rc = gap_to_bits(tb, tg, len); *is_gap = 1; break; }
    case 21: case 23: {                                       /* set_block_arrgap_egamma(_inv) (read_id_list :4331) */
        uint32_t n = bin_gamma(&b) & 0xffffu, prev = 0;
        for (uint32_t k = 0; k < n; ++k) { uint32_t v = bin_gamma(&b); if (!k) --v; prev = (prev + v) & 0xffffu; set_pos(tb, prev, 1); }
        if (bt == 23) for (uint32_t i = 0; i < BMB200_BLOCK_WORDS; ++i) tb[i] = ~tb[i];
        *is_gap = 1; break; }
    case 28: case 29: case 44: case 45: {                     /* set_block_arrgap_bienc(_inv)(_v2) (:4345-4375) */
        uint32_t n, min_v, max_v;
        if (bt <= 29) { min_v = rd16(r); max_v = rd16(r); n = (bin_gamma(&b) + 4u) & 0xffffu; }
        else { n = rd16(r); min_v = (n & 1u) ? rd8(r) : rd16(r); max_v = (n & 2u) ? rd8(r) : rd16(r); max_v = (min_v + max_v) & 0xffffu; n >>= 2; }
        if (n < 2) { rc = BMB200_ERR_BADARG; break; }
        a1[0] = (uint16_t)min_v; a1[n - 1] = (uint16_t)max_v;
        if (n > 2) bic_u16(&b, a1 + 1, n - 2, (uint16_t)min_v, (uint16_t)max_v);
        for (uint32_t k = 0; k < n; ++k) set_pos(tb, a1[k], 1);
        if (bt == 29 || bt == 45) for (uint32_t i = 0; i < BMB200_BLOCK_WORDS; ++i) tb[i] = ~tb[i];
        *is_gap = 1; break; }
    case 27: {                                                /* set_block_gap_bienc (:4874) */
        uint32_t head = rd16(r), len = head >> 3, min_v = rd16(r);
        if (len < 2) { rc = BMB200_ERR_BADARG; break; }
        tg[0] = (uint16_t)head; tg[1] = (uint16_t)min_v;
        if (len > 2) bic_u16(&b, tg + 2, len - 2, (uint16_t)min_v, 65535);
        rc = gap_to_bits(tb, tg, len); *is_gap = 1; break; }
    case 43: case 62: {                                       /* set_block_gap_bienc_v2 (:4885), _v3s (:4910) */
        uint32_t head = (bt == 43) ? rd16(r) : bin_delta16s(&b), len = head >> 3, min_v, max_v;
        if (bt == 43) { min_v = (head & 2u) ? rd8(r) : rd16(r); max_v = (head & 4u) ? rd8(r) : rd16(r); }
        else { min_v = (head & 2u) ? bin_gamma8(&b) : bin_16(&b); max_v = (head & 4u) ? bin_gamma8(&b) : bin_16(&b); }
        max_v = (65535u - max_v) & 0xffffu;
        if (len < 3) { rc = BMB200_ERR_BADARG; break; }
        tg[0] = (uint16_t)(head & ~6u); tg[1] = (uint16_t)min_v;
        if (len > 3) bic_u16(&b, tg + 2, len - 3, (uint16_t)min_v, (uint16_t)max_v);
        tg[len - 1] = (uint16_t)max_v;
        rc = gap_to_bits(tb, tg, len); *is_gap = 1; break; }
    case 61: {                                                /* set_block_gap_bienc_v3 (:4934-5020) */
        uint32_t h3 = bin_bits(&b, 8), head = bin_delta16s(&b), len = head >> 3;
        if (len < 1) { rc = BMB200_ERR_BADARG; break; }
        tg[0] = (uint16_t)(head & ~6u);
        if ((h3 & 0x80u) && len < 4) {
            if (len > 1) { tg[1] = (uint16_t)bin_delta16s(&b); for (uint32_t k = 2; k < len; ++k) tg[k] = (uint16_t)(tg[k - 1] + bin_delta16s(&b)); }
        } else {
            if (len < 3) { rc = BMB200_ERR_BADARG; break; }
            uint32_t min_v = (head & 2u) ? bin_bits(&b, 8) : bin_16(&b), max_v;          /* decode_min_max (:4780) */
            if (head & 4u) { max_v = bin_bits(&b, 8); max_v = ((max_v << 3) | (h3 & 7u)) & 0xffffu; } else max_v = bin_16(&b);
            max_v = (65535u - max_v) & 0xffffu;
            tg[1] = (uint16_t)min_v;
            uint32_t min0 = 0, min1 = 0;                                                   /* decode_mins (:4760) */
            if (!(h3 & 8u))    min0 = (h3 & 0x10u) ? bin_gamma8(&b) : bin_delta16(&b);
            if (!(h3 & 0x40u)) min1 = (h3 & 0x20u) ? bin_gamma8(&b) : bin_delta16(&b);
            if (len > 3) bic_u16(&b, tg + 2, len - 3, (uint16_t)(min_v + 1u), (uint16_t)max_v);
            tg[len - 1] = (uint16_t)(max_v + 1u); tg[len] = 65535;
            if ((h3 & 0x80u) || min0 || min1) gap_restore_mins(tg, (uint16_t)min0, (uint16_t)min1);
        }
        rc = gap_to_bits(tb, tg, len);
        if (rc == BMB200_OK && (h3 & 0x80u)) {                /* exception lists: single bits restored after the runs */
            for (int pass = 0; pass < 2; ++pass) {
                uint32_t cnt = 0; int h = decode_array(&b, a1, wf, &cnt, 0);
                if (h < 0) { rc = BMB200_ERR_BADARG; break; }
                for (uint32_t k = 0; k < cnt; ++k) set_pos(tb, a1[k], (h & 0x10) != 0);
                if (h & 0x20) break;
            }
        }
        *is_gap = 1; break; }
    case 67: {                                                /* set_block_gap_egamma_v3 (:5042) */
        uint32_t len = bin_gamma(&b) + 1u, start = bin_bit(&b), use_gamma = bin_bit(&b);
        if (len < 1 || len > 65536u) { rc = BMB200_ERR_BADARG; break; }
        tg[0] = (uint16_t)((len << 3) | start);
        if (use_gamma) { if (len > 1) { tg[1] = (uint16_t)bin_gamma8(&b); for (uint32_t i = 2; i < len; ++i) tg[i] = (uint16_t)(tg[i - 1] + bin_gamma8(&b)); } }
        else for (uint32_t i = 1; i < len; ++i) tg[i] = (uint16_t)bin_16(&b);
        rc = gap_to_bits(tb, tg, len); *is_gap = 1; break; }
    case 31: case 32: case 57: {                              /* set_block_arr_bienc(_inv), _8bh (read_bic_arr :4373-4385, :4466) */
        uint32_t min_v, max_v;
        if (bt == 57) { min_v = rd8(r); max_v = (65536u - rd8(r)) & 0xffffu; } else { min_v = rd16(r); max_v = rd16(r); }
        uint32_t n = rd16(r);
        if (n < 2) { rc = BMB200_ERR_BADARG; break; }
        set_pos(tb, min_v, 1); set_pos(tb, max_v, 1);
        if (n > 2) { bic_u16(&b, a1, n - 2, (uint16_t)min_v, (uint16_t)max_v); for (uint32_t k = 0; k < n - 2; ++k) set_pos(tb, a1[k], 1); }
        if (bt == 32) for (uint32_t i = 0; i < BMB200_BLOCK_WORDS; ++i) tb[i] = ~tb[i];
        break; }
    case 65: case 66: {                                       /* set_block_arr_bienc(_inv)_v3s (:4432-4462) */
        uint32_t n = bin_delta16s(&b) & 0xffffu, min_v = 0, max_v = 65535;
        if (bin_bits(&b, 1)) {
            min_v = bin_delta16s(&b) & 0xffffu; n = (n - 2u) & 0xffffu; max_v = (65536u - bin_delta16s(&b)) & 0xffffu;
            set_pos(tb, min_v, 1); set_pos(tb, max_v, 1);
            min_v = (min_v + 1u) & 0xffffu; max_v = (max_v - 1u) & 0xffffu;
        }
        if (n) { bic_u16(&b, a1, n, (uint16_t)min_v, (uint16_t)max_v); for (uint32_t k = 0; k < n; ++k) set_pos(tb, a1[k], 1); }
        if (bt == 66) for (uint32_t i = 0; i < BMB200_BLOCK_WORDS; ++i) tb[i] = ~tb[i];
        break; }
    case 63: case 64: {                                       /* set_block_arr_bienc(_inv)_v3 (:4386-4430): singles + runs */
        uint32_t s_cnt = 0, r_cnt = 0, l_cnt = 0;
        int h = decode_array(&b, a1, wf, &s_cnt, 0);
        if (h < 0) { rc = BMB200_ERR_BADARG; break; }
        for (uint32_t k = 0; k < s_cnt; ++k) set_pos(tb, a1[k], 1);
        if (!(h & 0x20)) {
            h = decode_array(&b, a1, wf, &r_cnt, 0);
            if (h < 0 || !r_cnt) { rc = BMB200_ERR_BADARG; break; }
            h = decode_array(&b, a2, wf, &l_cnt, r_cnt);
            if (h < 0) { rc = BMB200_ERR_BADARG; break; }
            if (l_cnt > r_cnt) l_cnt = r_cnt;
            if ((h & 3) == 1) for (uint32_t i = 0; i < l_cnt; ++i) a2[i] = (uint16_t)(a2[i] - a1[i]);
            for (uint32_t i = 0; i < l_cnt; ++i) set_run(tb, a1[i], (uint32_t)a2[i] + 1u);
        }
        if (bt == 64) for (uint32_t i = 0; i < BMB200_BLOCK_WORDS; ++i) tb[i] = ~tb[i];
        break; }
    case 33: {                                                /* set_block_bitgap_bienc (read_bic_gap :4675) */
        uint32_t head = rd8(r), len = rd16(r), min_v = rd16(r);
        if (len < 2) { rc = BMB200_ERR_BADARG; break; }
        tg[0] = (uint16_t)head; tg[1] = (uint16_t)min_v;
        if (len > 2) bic_u16(&b, tg + 2, len - 2, (uint16_t)min_v, 65535);
        rc = gap_to_bits(tb, tg, len); break; }
    default: rc = BMB200_ERR_UNSUPPORTED; break;
    }
    if (rc == BMB200_OK && r->p > r->end) rc = BMB200_ERR_BADARG;
done:
    free(tg); free(a1); free(a2); free(wf);
    return rc;
}

/* super-block tokens (read_bic_sb_arr :4480-4650): positions relative to super-block *sb, ascending, arr has room for 65536 */
int orc_sblock_token(rd_t* r, uint32_t bt, uint32_t* arr, uint32_t* len_out, uint32_t* sb)
{
    bin_t b; bin_init(&b, r);
    uint32_t len, min_v, max_v, min0 = 0, flag;
    if (bt == 56) {                                           /* set_sblock_bienc: byte header */
        flag = rd8(r);
        *sb = (flag & 2u) ? rd32(r) : (flag & 1u) ? rd16(r) : rd8(r);
        len = (flag & 0x10u) ? rd16(r) : rd8(r);
        if (flag & 8u) min_v = (flag & 4u) ? rd32(r) : (rd16(r) | (rd8(r) << 16)); else min_v = (flag & 4u) ? rd16(r) : rd8(r);
        if (flag & 0x40u) max_v = (flag & 0x20u) ? rd32(r) : (rd16(r) | (rd8(r) << 16)); else max_v = (flag & 0x20u) ? rd16(r) : rd8(r);
        max_v = 256u * 65536u - max_v;
        if (flag & 0x80u) min0 = bin_bit(&b) ? bin_gamma(&b) : bin_16(&b);
        if (len < 2 || len > 65536u) return BMB200_ERR_BADARG;
        arr[0] = min_v; arr[len - 1] = max_v;
        if (len > 2) bic_u32(&b, arr + 1, len - 2, min_v, max_v);
    } else if (bt == 68) {                                    /* set_sblock_bienc_v3: everything in the bit stream */
        flag = bin_bits(&b, 8);
        len = (flag & 0x10u) ? bin_delta16(&b) : bin_bits(&b, 8);
        if (flag & 8u) { uint32_t j = bin_gamma(&b), nbit = bin_16(&b); min_v = j * 65536u + nbit; }
        else min_v = (flag & 4u) ? bin_16(&b) : bin_bits(&b, 8);
        if (flag & 0x40u) max_v = bin_24(&b); else max_v = (flag & 0x20u) ? bin_16(&b) : bin_bits(&b, 8);
        max_v = 256u * 65536u - max_v;
        if (flag & 0x80u) { switch (bin_gamma(&b)) { case 1: min0 = bin_gamma(&b); break; case 2: min0 = bin_bits(&b, 8); break; default: min0 = bin_16(&b); break; } }
        if ((flag & 3u) == 3u) *sb = bin_gamma(&b) - 1u; else *sb = (flag & 2u) ? bin_32(&b) : (flag & 1u) ? bin_16(&b) : bin_bits(&b, 8);
        if (len < 2 || len > 65536u) return BMB200_ERR_BADARG;
        arr[0] = min_v; arr[len - 1] = max_v;
        if (len > 2) bic_u32(&b, arr + 1, len - 2, min_v + 1u, max_v - 1u);
    } else return BMB200_ERR_UNSUPPORTED;
    if (min0) { uint32_t dacc = 0; for (uint32_t i = 1; i < len; ++i) { arr[i] += min0 + dacc; dacc += min0; } }   /* arr_restore_min, T = u32 */
    *len_out = len;
    return r->p > r->end ? BMB200_ERR_BADARG : BMB200_OK;
}
