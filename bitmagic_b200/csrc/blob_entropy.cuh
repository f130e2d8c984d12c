// blob_entropy.cuh -- deserialize-to-device, part 2: the entropy-coded block encodings of bm::serializer<> (levels 3..6).
//
// These tokens carry no length: the byte where the next token starts is only known once the current one is decoded
// (Elias gamma, binary interpolative coding).  So the work is split in two launches:
//   blob_walk_kernel     one warp per serialized vector: lane 0 walks the token stream of its BLOB (src/bmserial.h:5578-6090) and
//                        decodes the entropy-coded tokens on the way, the warp turns each decoded block into a bitmap in shared
//                        memory and MEASURES it (kind, GAP length, first-run value, payload offset) -> one BlobTok per block.
//                        The host builds the column-major arena layout from the BlobToks (prefix sums, as for explicit tokens).
//   blob_entropy_kernel  one warp per entropy-coded TOKEN (all tokens of all vectors in parallel now that their offsets are
//                        known): decode again, build the bitmap, write the bit-block / the GAP block (flat-streamable form)
//                        into its arena slot.
// A token is decoded by ONE lane (the codes are sequential by construction); everything after that -- scattering runs and
// positions into the bitmap, counting runs, bit -> GAP -- is done by the 32 lanes together.
//
// Reference routines restated here (file:line under the reference tree):
//   bit_in::gamma src/encoding.h:2506, gamma8 :2441, delta16 :2463, delta16s :2488, get_bits :2638, get_16_no :2591
//   bit_in::bic_decode_u16_cm :2213 / bic_decode_u32_cm :2168 (centered-minimal binary interpolative coding), recursion
//       replaced by an explicit stack (depth <= 17)
//   bit_in::decode_array :2698 (delta-gamma | gamma | BIC with delta-range reduction | single value), arr_restore_min
//       src/bmfunc.h:2657, arr_restore_min_w :2526
//   read_gap_block src/bmserial.h:4812 (set_block_gap_bienc_v3 :4934 incl. decode_min_max :4780, decode_mins :4760,
//       gap_restore_mins src/bmfunc.h:3009 and the exception lists; set_block_gap_bienc_v3s :4910; set_block_gap_egamma_v3 :5042)
//   read_id_list :4307 (set_block_arrgap_egamma(_inv)), read_bic_arr :4363 (set_block_arr_bienc(_inv)_v3 / _v3s),
//   read_bic_gap :4675 (set_block_bitgap_bienc), read_bic_sb_arr :4568 + decode_arr_sblock :5458 (set_sblock_bienc_v3)
// Covered = every entropy-coded token the serializer of this reference version emits (measured over levels 3..6, see
// tests/test_oracle_vs_reference.py); the legacy encodings it can still READ but no longer writes (gap_egamma 20, gap_bienc
// 27/43, arrgap_bienc 28/29/44/45, arr_bienc 31/32/57, sblock_bienc 56) and the XOR-reference tokens are refused.
//
// The decode logic is plain C++ behind BME_HD so that a host build of this header (oracle/blob_host_check.cpp, test
// infrastructure) can run the SAME walker + decoder against bm::deserialize on the CPU box; the product only ever runs
// it inside the two kernels at the bottom.
#pragma once
#include <stdint.h>
#include <string.h>
#include <vector>

#if defined(__CUDACC__)
#define BME_HD __host__ __device__ __forceinline__
#define BME_HDN __host__ __device__
#else
#define BME_HD inline
#define BME_HDN
#endif

namespace bmb200 {

constexpr uint32_t kEntListCap   = 65536u + 16u;          // u16 entries per scratch list
constexpr uint32_t kEntLists     = 3u;
constexpr uint32_t kEntWords     = 2048u;                 // words of one block bitmap
constexpr size_t   kEntScratchBytes = (size_t)kEntLists * kEntListCap * 2u + kEntWords * 4u;   // per warp: 3 lists + window flags
constexpr uint32_t kTokEntropy   = 0x100u;                // BlobTok / BlobRec type = kTokEntropy | serializer token code
constexpr uint32_t kTokSbMember  = 0x200u;                // BlobTok only: one block of a super-block token (layout, no rec)
constexpr uint32_t kGapFitWords  = 1276u;                 // gap_calc_level(len) >= 0  <=>  len <= 1280 - 4 (src/bmfunc.h:5418)

// one block found by the walk (host walker for explicit-length streams, blob_walk_kernel otherwise)
struct BlobTok {
    uint32_t nb;         // block column
    uint32_t type;       // DB_* (explicit tokens), kTokEntropy | code, kTokSbMember
    uint64_t off;        // byte offset of the token payload inside its BLOB
    uint32_t aux;        // DB_* specific (see blob_kernel.cuh); super-block: first column of the super-block
    uint32_t first;      // GAP kinds: value of the first run
    uint32_t gap_words;  // GAP kinds: header + run ends (u16 words) the block needs in the arena
    uint32_t kind;       // BMB200_BLK_BIT / BMB200_BLK_GAP
};

// set_block_arrgap(_inv): n ascending u16 positions at p (byte aligned).  A well-formed token has strictly ascending positions and the
// GAP block gap_set_array builds from them fits the largest GAP capacity; anything else (crafted / corrupt streams: up to 2n+1 runs
// for n <= 2048 isolated bits) is refused instead of being written past its arena slot.  Returns the u16 words the block needs
// (header + run ends), or 0 for a malformed / oversized list.
BME_HDN inline uint32_t arrgap_measure(const uint8_t* p, uint32_t n)
{
    uint32_t ends = 0, prev = 0;
    for (uint32_t k = 0; k < n; ++k) {
        const uint32_t s = (uint32_t)p[2u * k] | ((uint32_t)p[2u * k + 1u] << 8);
        if (k && s <= prev) return 0u;
        const bool st = (k == 0u) || prev + 1u != s;
        if (st) { if (k && prev < 65535u) ++ends; if (s > 0u) ++ends; }      // close the previous 1-run, open this one
        prev = s;
    }
    if (n && prev < 65535u) ++ends;
    const uint32_t words = ends + 2u;                                          // header + ends + the final 65535
    return words - 1u <= kGapFitWords ? words : 0u;      // longer lists are bit-blocks in the reference (gap_calc_level < 0); no serializer writes them
}

// ------------------------------------------------------------------------------------------------------------------
// lane helpers: on the device a "team" is one warp; on the host the same code runs with one lane
// ------------------------------------------------------------------------------------------------------------------
struct EntTeam { uint32_t lane, nl; };

BME_HD void bme_sync()
{
#ifdef __CUDA_ARCH__
    __syncwarp();
#endif
}
BME_HD uint32_t bme_bcast(uint32_t v)
{
#ifdef __CUDA_ARCH__
    return __shfl_sync(0xffffffffu, v, 0);
#else
    return v;
#endif
}
BME_HD uint64_t bme_bcast64(uint64_t v) { const uint32_t lo = bme_bcast((uint32_t)v), hi = bme_bcast((uint32_t)(v >> 32)); return lo | ((uint64_t)hi << 32); }
BME_HD uint32_t bme_sum(uint32_t v)
{
#ifdef __CUDA_ARCH__
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
#endif
    return v;
}
// exclusive prefix of v over the team, *total = sum
BME_HD uint32_t bme_excl_scan(uint32_t v, uint32_t lane, uint32_t* total)
{
#ifdef __CUDA_ARCH__
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (uint32_t)o) inc += y; }
    *total = __shfl_sync(0xffffffffu, inc, 31);
    return inc - v;
#else
    (void)lane; *total = v; return 0u;
#endif
}
BME_HD void bme_or(uint32_t* p, uint32_t v)
{
#ifdef __CUDA_ARCH__
    atomicOr(p, v);
#else
    *p |= v;
#endif
}
BME_HD void bme_and(uint32_t* p, uint32_t v)
{
#ifdef __CUDA_ARCH__
    atomicAnd(p, v);
#else
    *p &= v;
#endif
}
BME_HD uint32_t bme_popc(uint32_t x)
{
#ifdef __CUDA_ARCH__
    return (uint32_t)__popc(x);
#else
    return (uint32_t)__builtin_popcount(x);
#endif
}
BME_HD uint32_t bme_clz(uint32_t x)          // x != 0
{
#ifdef __CUDA_ARCH__
    return (uint32_t)__clz((int)x);
#else
    return (uint32_t)__builtin_clz(x);
#endif
}
BME_HD uint32_t bme_ctz64(uint64_t x)        // x != 0
{
#ifdef __CUDA_ARCH__
    return (uint32_t)(__ffsll((long long)x) - 1);
#else
    return (uint32_t)__builtin_ctzll(x);
#endif
}

// ------------------------------------------------------------------------------------------------------------------
// byte reader (bm::decoder, little endian, any alignment) and bit reader (bm::bit_in) over the staged BLOB bytes
// ------------------------------------------------------------------------------------------------------------------
struct EntRd {
    const uint8_t* s;      // staging base
    uint64_t p, end;       // current / one-past-last byte offset of this BLOB inside the staging buffer
    uint32_t bad;          // sticky: a read ran past the end or a code was malformed
    BME_HD uint32_t u8()  { if (p + 1 > end) { bad = 1; return 0; } return s[p++]; }
    BME_HD uint32_t u16() { if (p + 2 > end) { bad = 1; p = end; return 0; } const uint32_t v = (uint32_t)s[p] | ((uint32_t)s[p + 1] << 8); p += 2; return v; }
    BME_HD uint32_t u32()
    {
        if (p + 4 > end) { bad = 1; p = end; return 0; }
#ifdef __CUDA_ARCH__
        // any alignment: two aligned loads + funnel shift (the staging buffer is 256-byte aligned and carries 64 bytes of slack)
        const uint32_t* w = reinterpret_cast<const uint32_t*>(s + (p & ~3ull));
        const uint32_t sh = (uint32_t)(p & 3ull) * 8u, lo = w[0];
        const uint32_t v = sh ? __funnelshift_r(lo, w[1], sh) : lo;
#else
        const uint32_t v = (uint32_t)s[p] | ((uint32_t)s[p + 1] << 8) | ((uint32_t)s[p + 2] << 16) | ((uint32_t)s[p + 3] << 24);
#endif
        p += 4; return v;
    }
    BME_HD uint64_t u64() { const uint64_t lo = u32(); return lo | ((uint64_t)u32() << 32); }
    // 32 bits at byte offset `at` (at + 4 <= end), no state change: the fast readers keep their own position
    BME_HD uint32_t peek32(uint64_t at) const
    {
#ifdef __CUDA_ARCH__
        const uint32_t* w = reinterpret_cast<const uint32_t*>(s + (at & ~3ull));
        const uint32_t sh = (uint32_t)(at & 3ull) * 8u, lo = w[0];
        return sh ? __funnelshift_r(lo, w[1], sh) : lo;
#else
        return (uint32_t)s[at] | ((uint32_t)s[at + 1] << 8) | ((uint32_t)s[at + 2] << 16) | ((uint32_t)s[at + 3] << 24);
#endif
    }
    BME_HD void skip(uint64_t k) { if (p + k > end) { bad = 1; p = end; } else p += k; }
};

// bits are consumed LSB first from 32-bit little-endian words; a word is fetched only when a bit of it is needed, so the
// byte position after a token is start + 4 * ceil(bits / 32) exactly like bm::bit_in (src/encoding.h:2638-2696)
struct EntBits {
    EntRd* r; uint64_t acc; uint32_t have;
    BME_HD void init(EntRd* rd) { r = rd; acc = 0; have = 0; }
    BME_HD uint32_t bits(uint32_t n)               // n = 0..32
    {
        if (!n) return 0u;
        if (have < n) { acc |= (uint64_t)r->u32() << have; have += 32u; }
        const uint32_t v = (uint32_t)(acc & (n >= 32u ? 0xffffffffull : ((1ull << n) - 1ull)));
        acc >>= n; have -= n;
        return v;
    }
    BME_HD uint32_t bit() { return bits(1u); }
    BME_HD uint32_t g16() { const uint32_t lo = bits(8u); return lo | (bits(8u) << 8); }
    BME_HD uint32_t g24() { const uint32_t v = g16(); return v | (bits(8u) << 16); }
    BME_HD uint32_t g32() { const uint32_t v = g16(); return v | (g16() << 16); }
    BME_HD uint32_t gamma()                         // Elias gamma: z zeros, a one, z value bits -> value | 1 << z
    {
        uint32_t zeros = 0;
        for (;;) {
            if (!have) { acc = r->u32(); have = 32u; if (r->bad) return 0u; }
            if (!acc) { zeros += have; have = 0; if (zeros > 32u) { r->bad = 1; return 0u; } continue; }
            const uint32_t tz = bme_ctz64(acc);
            zeros += tz; acc >>= (tz + 1u); have -= (tz + 1u);
            break;
        }
        if (zeros > 31u) { r->bad = 1; return 0u; }
        return bits(zeros) | (1u << zeros);
    }
    BME_HD uint32_t delta16()
    {
        switch (gamma()) {
        case 1: return 511u - bits(8u);
        case 2: return 512u + 255u - bits(8u);
        case 3: return 512u + 256u + 255u - bits(8u);
        default: return g16();
        }
    }
    BME_HD uint32_t delta16s() { return bit() ? delta16() : bits(8u); }
    BME_HD uint32_t gamma8()
    {
        switch (gamma()) {
        case 1: return gamma();
        case 2: return bits(8u);
        case 3: return delta16();
        default: return 0u;
        }
    }
    // one centered-minimal code word for a range of size r + 1 (src/encoding.h:2224-2237)
    BME_HD uint32_t bic(uint32_t r)
    {
        if (!r) return 0u;
        if (r > 0x7ffffff0u) { this->r->bad = 1; return 0u; }
        const uint32_t logv = 31u - bme_clz(r + 1u);                   // <= 30: everything below fits 32-bit signed arithmetic
        const uint32_t c = (1u << (logv + 1u)) - r - 1u;
        const int32_t half_c = (int32_t)(c >> 1), half_r = (int32_t)(r >> 1);
        const int32_t lo1 = half_r - half_c - (int32_t)((r + 1u) & 1u), hi1 = half_r + half_c + 1;
        uint32_t val = bits(logv);
        if ((int32_t)val <= lo1 || (int32_t)val >= hi1) val += bit() << logv;
        return val;
    }
};

// bic_decode_u16_cm / _u32_cm: out[0..sz) ascending in [lo, hi]; pre-order (node, left half, right half) with an explicit stack
template <typename T>
BME_HD void ent_bic_decode(EntBits& b, T* out, uint32_t sz, uint32_t lo, uint32_t hi)
{
    constexpr bool k16 = (sizeof(T) == 2);
    constexpr uint32_t kMask = k16 ? 0xffffu : 0xffffffffu;                    // the reference narrows lo / hi to T at every call
    // pending right halves: (first index | size << 16) -- a right half of at most 65536 values holds < 32768 -- and their value range
    // (16-bit: lo | hi << 16 in one word)
    uint32_t st_seg[20], st_lo[20], st_hi[k16 ? 1 : 20]; int sp = 0;
    EntRd* const r = b.r;
    if (sz > 65536u) { r->bad = 1; return; }
    if (r->bad) return;
    // This loop is the hot spot of both passes (one lane, one dependent chain), so it runs on a private copy of the bit reader that
    // keeps >= 32 bits in the window: one refill test per value, no branch inside a code word.  The window may run ONE word ahead
    // of what bm::bit_in would have fetched; that word is handed back on exit, so the stream position after the array is the lazy
    // one (start + 4 * ceil(bits / 32)) and the bits left in the accumulator are the ones a lazy reader would hold.
    uint64_t acc = b.acc; uint32_t have = b.have, malformed = 0;
    uint64_t p = r->p; const uint64_t end = r->end;
    uint32_t off = 0;
    for (;;) {
        while (sz) {
            if (have < 32u) { const uint32_t w = (p + 4u <= end) ? r->peek32(p) : 0u; p += 4u; acc |= (uint64_t)w << have; have += 32u; }
            const uint32_t rr = hi - lo - sz + 1u;
            uint32_t val = 0;
            if (rr) {                                   // one centered-minimal code word (src/encoding.h:2224-2237)
                if (rr > 0x7ffffff0u) { malformed = 1; sp = 0; break; }
                const uint32_t logv = 31u - bme_clz(rr + 1u);                  // <= 30
                const uint32_t c = (1u << (logv + 1u)) - rr - 1u;
                const int32_t half_c = (int32_t)(c >> 1), half_r = (int32_t)(rr >> 1);
                const int32_t lo1 = half_r - half_c - (int32_t)((rr + 1u) & 1u), hi1 = half_r + half_c + 1;
                val = (uint32_t)acc & ((1u << logv) - 1u);
                acc >>= logv; have -= logv;
                if ((int32_t)val <= lo1 || (int32_t)val >= hi1) { val += ((uint32_t)acc & 1u) << logv; acc >>= 1; have -= 1u; }
            }
            const uint32_t mid = sz >> 1;
            val += lo + mid;
            out[off + mid] = (T)val;
            if (sz <= 1u) break;
            if (sz - mid - 1u) {                         // a non-empty right half waits on the stack (its first index is <= 65535 then)
                if (sp >= 20) { malformed = 1; sp = 0; break; }
                st_seg[sp] = (off + mid + 1u) | ((sz - mid - 1u) << 16);
                if (k16) st_lo[sp] = ((val + 1u) & kMask) | (hi << 16); else { st_lo[sp] = val + 1u; st_hi[k16 ? 0 : sp] = hi; }
                ++sp;
            }
            sz = mid; hi = (val - 1u) & kMask;          // left half next; off and lo stay
        }
        if (!sp) break;
        --sp; off = st_seg[sp] & 0xffffu; sz = st_seg[sp] >> 16;
        if (k16) { lo = st_lo[sp] & 0xffffu; hi = st_lo[sp] >> 16; } else { lo = st_lo[sp]; hi = st_hi[k16 ? 0 : sp]; }
    }
    if (have >= 32u) { p -= 4u; have -= 32u; acc &= have ? ((1ull << have) - 1ull) : 0ull; }     // hand the look-ahead word back
    if (p > end || malformed) { r->bad = 1; p = p > end ? end : p; }
    r->p = p; b.acc = acc; b.have = have;
}

// arr_restore_min_w (src/bmfunc.h:2526-2581), T = u16: per-window minimal delta put back
BME_HDN void ent_restore_min_w(uint16_t* arr, uint32_t n, uint32_t wlen, uint32_t min0, const uint32_t* wflags)
{
    uint32_t dacc = 0, min_w_prev = ~0u;
    for (uint32_t i = 1; i < wlen && i < n; ++i) {
        arr[i] = (uint16_t)(arr[i] + min0 + dacc); dacc = (dacc + min0) & 0xffffu;
        const uint32_t d = (uint16_t)(arr[i] - arr[i - 1]); if (d < min_w_prev) min_w_prev = d;
    }
    min_w_prev -= (min_w_prev != 0u);
    uint32_t wave = 1;
    for (uint32_t i = wlen; i < n; ++wave, i += wlen) {
        if (i + wlen > n) wlen = n % wlen;
        if (!wlen) break;
        const uint32_t recalc = (wflags[(wave >> 5) & (kEntWords - 1u)] >> (wave & 31u)) & 1u;
        uint32_t min_w = ~0u;
        for (uint32_t j = 0; j < wlen; ++j) {
            if (recalc) { arr[i + j] = (uint16_t)(arr[i + j] + ((min_w_prev + dacc) & 0xffffu)); dacc = (dacc + (min_w_prev & 0xffffu)) & 0xffffu; }
            else        { arr[i + j] = (uint16_t)(arr[i + j] + min0 + dacc); dacc = (dacc + min0) & 0xffffu; }
            const uint32_t d = (uint16_t)(arr[i + j] - arr[i + j - 1]); if (d < min_w) min_w = d;
        }
        min_w_prev = (min_w > min0) ? min_w - 1u : min0;
    }
}

// bit_in::decode_array (src/encoding.h:2698-2798).  out / tmp: lists of kEntListCap u16, wf: 2048 words.  Returns the flag byte
// (>= 0) with *sz set, or -1.
BME_HD int ent_decode_array(EntBits& b, uint16_t* out, uint16_t* tmp, uint32_t* wf, uint32_t* sz, uint32_t default_sz)
{
    const uint32_t h = b.bits(8u);
    if ((h & 3u) == 3u && (h & 0x80u)) { *sz = 0; return (int)h; }                                     // no-op
    if ((h & 3u) == 3u) { *sz = 1; out[0] = (h & 0x40u) ? (uint16_t)0 : (uint16_t)((h & 8u) ? b.gamma() : b.g16()); return b.r->bad ? -1 : (int)h; }
    const uint32_t n = default_sz ? default_sz : ((h & 8u) ? b.gamma8() + 1u : b.delta16());
    if (n > 65536u || b.r->bad) return -1;
    *sz = n;
    const uint32_t min0 = (h & 0x40u) ? 0u : (b.gamma() & 0xffffu);
    if ((h & 3u) == 0u) {                                       // delta-gamma
        uint32_t prev = (h & 0x80u) ? 0u : b.gamma();
        if (n) out[0] = (uint16_t)prev;
        for (uint32_t i = 1; i < n; ++i) { prev = (prev + b.gamma() + min0) & 0xffffu; out[i] = (uint16_t)prev; if (b.r->bad) return -1; }
    } else if (h & 2u) {                                        // gamma
        const uint32_t zc = (h & 0x80u) ? 1u : 0u;
        for (uint32_t i = 0; i < n; ++i) { out[i] = (uint16_t)(b.gamma() - zc + min0); if (b.r->bad) return -1; }
    } else {                                                    // interpolative, with delta-range reduction
        uint32_t min_v = 0, max_v = 65535u, s = n; uint16_t* p = out;
        if (h & 0x80u) {
            min_v = b.g16(); max_v = b.g16();
            if (n < 2u) return -1;
            out[0] = (uint16_t)min_v; out[n - 1u] = (uint16_t)max_v;
            if (n == 2u) return b.r->bad ? -1 : (int)h;
            min_v = (min_v + 1u) & 0xffffu; max_v = (max_v - 1u) & 0xffffu; s -= 2u; ++p;
        }
        if (s) ent_bic_decode<uint16_t>(b, p, s, min_v, max_v);
        if (b.bit()) {                                          // windowed restore of the minimal deltas
            for (uint32_t i = 0; i < kEntWords; ++i) wf[i] = 0u;
            uint32_t win = b.gamma(), wcnt = b.gamma();
            wcnt += 15u - 1u; win = (win + 9u) * 2u;
            const uint32_t max_wd = n / win + 1u;
            if (wcnt > 65536u || b.r->bad) return -1;
            ent_bic_decode<uint16_t>(b, tmp, wcnt, 1u, max_wd & 0xffffu);
            for (uint32_t k = 0; k < wcnt; ++k) wf[tmp[k] >> 5] |= 1u << (tmp[k] & 31u);
            ent_restore_min_w(out, n, win, min0, wf);
        } else if (min0) {
            uint32_t dacc = 0;
            for (uint32_t i = 1; i < n; ++i) { out[i] = (uint16_t)(out[i] + min0 + dacc); dacc = (dacc + min0) & 0xffffu; }
        }
    }
    return b.r->bad ? -1 : (int)h;
}

// gap_restore_mins (src/bmfunc.h:3009-3041) on g[1..len]
BME_HDN void ent_gap_restore_mins(uint16_t* g, uint32_t len, uint32_t min0, uint32_t min1)
{
    uint32_t p = 1;
    g[p] = (uint16_t)(g[p] + min0);
    uint32_t dacc = min0 & 0xffffu;
    for (++p; p <= len; ) {
        if (p == len) break;
        g[p] = (uint16_t)(g[p] + min1 + dacc); dacc = (dacc + min1) & 0xffffu;
        if (++p < len) { g[p] = (uint16_t)(g[p] + min0 + dacc); dacc = (dacc + min0) & 0xffffu; ++p; }
        else break;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// team context: bitmap (shared memory on the device) + scratch lists (global memory) + the byte / bit readers of lane 0
// ------------------------------------------------------------------------------------------------------------------
struct EntCtx {
    EntTeam t;
    uint32_t* bm;                    // 2048-word bitmap of the block being built
    uint16_t *la, *lb, *lc;          // scratch lists
    uint32_t* wf;                    // window flags of decode_array
};

BME_HD void ent_set_range(uint32_t* bm, uint32_t s, uint32_t e)      // bits [s, e], s <= e <= 65535
{
    const uint32_t ws = s >> 5, we = e >> 5;
    if (ws == we) { bme_or(&bm[ws], (0xffffffffu << (s & 31u)) & (0xffffffffu >> (31u - (e & 31u)))); return; }
    bme_or(&bm[ws], 0xffffffffu << (s & 31u));
    for (uint32_t w = ws + 1u; w < we; ++w) bm[w] = 0xffffffffu;
    bme_or(&bm[we], 0xffffffffu >> (31u - (e & 31u)));
}
BME_HD void ent_clear(const EntCtx& c) { for (uint32_t w = c.t.lane; w < kEntWords; w += c.t.nl) c.bm[w] = 0u; bme_sync(); }
BME_HD void ent_invert(const EntCtx& c) { bme_sync(); for (uint32_t w = c.t.lane; w < kEntWords; w += c.t.nl) c.bm[w] = ~c.bm[w]; bme_sync(); }
// g[0] = value of the first run, g[1..len] inclusive run ends (g[len] = 65535): OR the 1-runs into the bitmap
BME_HD void ent_apply_gap(const EntCtx& c, const uint16_t* g, uint32_t len)
{
    bme_sync();
    const uint32_t first = g[0] & 1u;
    for (uint32_t k = 1u + c.t.lane; k <= len; k += c.t.nl) {
        if (!((first ^ (k - 1u)) & 1u)) continue;
        const uint32_t s = (k == 1u) ? 0u : (uint32_t)g[k - 1u] + 1u, e = g[k];
        if (s <= e && s < 65536u) ent_set_range(c.bm, s, e);
    }
    bme_sync();
}
BME_HD void ent_apply_pos(const EntCtx& c, const uint16_t* a, uint32_t n, uint32_t val)
{
    bme_sync();
    for (uint32_t k = c.t.lane; k < n; k += c.t.nl) {
        const uint32_t p = a[k];
        if (val) bme_or(&c.bm[p >> 5], 1u << (p & 31u)); else bme_and(&c.bm[p >> 5], ~(1u << (p & 31u)));
    }
    bme_sync();
}
BME_HD void ent_apply_runs(const EntCtx& c, const uint16_t* r, const uint16_t* rl, uint32_t n)      // or_bit_block(blk, r[i], rl[i] + 1)
{
    bme_sync();
    for (uint32_t k = c.t.lane; k < n; k += c.t.nl) {
        const uint32_t s = r[k]; uint32_t e = s + (uint32_t)rl[k]; if (e > 65535u) e = 65535u;
        ent_set_range(c.bm, s, e);
    }
    bme_sync();
}
// number of runs of the bitmap (bit_block_calc_change, src/bmfunc.h:6040): team-wide result
BME_HD uint32_t ent_count_runs(const EntCtx& c)
{
    bme_sync();
    uint32_t cnt = 0;
    for (uint32_t w = c.t.lane; w < kEntWords; w += c.t.nl) {
        const uint32_t x = c.bm[w], nxt = (w + 1u < kEntWords) ? (c.bm[w + 1u] & 1u) : (x >> 31);
        cnt += bme_popc(x ^ ((x >> 1) | (nxt << 31)));
    }
    const uint32_t runs = bme_sum(cnt) + 1u;
    bme_sync();                                  // the bitmap may be rewritten right after (next token)
    return runs;
}
// bit_block_to_gap (src/bmfunc.h:5540): out[0] = header, out[1..len] run ends; len = runs (computed by ent_count_runs)
// sb_member: the block was built bit by bit under BM_GAP (bvector::gap_block_set_no_ret, src/bm.h:4800): its capacity level only
// rises when the run count exceeds glen[level] - 4, so level = the smallest one with len <= glen - 4; every other GAP token goes
// through deserialize_gap, which takes gap_calc_level(gap_length = len + 1)
BME_HD void ent_write_gap(const EntCtx& c, uint16_t* out, uint32_t len, bool sb_member = false)
{
    bme_sync();
    uint32_t base = 0;
    for (uint32_t w0 = 0; w0 < kEntWords; w0 += c.t.nl) {
        const uint32_t w = w0 + c.t.lane;
        const uint32_t x = c.bm[w], nxt = (w + 1u < kEntWords) ? (c.bm[w + 1u] & 1u) : (x >> 31);
        uint32_t m = x ^ ((x >> 1) | (nxt << 31));
        uint32_t total; uint32_t off = base + bme_excl_scan(bme_popc(m), c.t.lane, &total);
        while (m) {
#ifdef __CUDA_ARCH__
            const uint32_t bpos = (uint32_t)__ffs((int)m) - 1u;
#else
            const uint32_t bpos = (uint32_t)__builtin_ctz(m);
#endif
            m &= m - 1u;
            if (off + 1u < len) out[1u + off] = (uint16_t)(w * 32u + bpos);
            ++off;
        }
        base += total;
    }
    if (c.t.lane == 0) {
        out[len] = 65535u;
        const uint32_t ll = sb_member ? len : len + 1u;
        const uint32_t lvl = ll <= 124u ? 0u : ll <= 252u ? 1u : ll <= 508u ? 2u : 3u;   // gap_calc_level
        out[0] = (uint16_t)((c.bm[0] & 1u) | (lvl << 1) | (len << 3));
    }
    bme_sync();
}
BME_HD void ent_write_bits(const EntCtx& c, uint32_t* dst)
{
    bme_sync();
    for (uint32_t w = c.t.lane; w < kEntWords; w += c.t.nl) dst[w] = c.bm[w];
    bme_sync();
}

// ------------------------------------------------------------------------------------------------------------------
// one entropy-coded block token -> bitmap.  `code` = serializer token (team-uniform), rd positioned at its payload (lane 0's
// copy is the one that counts).  Returns 0 or a BMB200_ERR_* value (team-uniform); *gap_family = the reference materialises
// the token through a GAP block (deserialize_gap) and keeps it a GAP block when it fits.
// ------------------------------------------------------------------------------------------------------------------
BME_HD int ent_decode_block_impl(const EntCtx& c, uint32_t code, EntRd& rd, uint32_t* gap_family)
{
    const bool lead = (c.t.lane == 0u);
    EntBits b; b.init(&rd);
    uint32_t err = 0, n = 0, flags = 0;
    ent_clear(c);
    *gap_family = 0u;
    switch (code) {
    case 61u: case 62u: case 67u: case 33u: {            // run lists: GAP v3 / v3s / gamma v3, bit-block stored as interpolated runs
        *gap_family = (code != 33u);
        if (lead) {
            uint16_t* g = c.la;
            uint32_t len = 0, h3 = 0;
            if (code == 67u) {
                len = b.gamma() + 1u; const uint32_t start = b.bit(), use_gamma = b.bit();
                if (len > 65536u || rd.bad) err = 1;
                else {
                    g[0] = (uint16_t)start;
                    if (use_gamma) { uint32_t prev = 0; for (uint32_t i = 1; i < len && !rd.bad; ++i) { prev = (i == 1u) ? b.gamma8() : prev + b.gamma8(); g[i] = (uint16_t)prev; } }
                    else for (uint32_t i = 1; i < len && !rd.bad; ++i) g[i] = (uint16_t)b.g16();
                }
            } else if (code == 33u) {
                const uint32_t head = rd.u8(); len = rd.u16(); const uint32_t min_v = rd.u16();
                if (len < 2u || rd.bad) err = 1;
                else { g[0] = (uint16_t)(head & 1u); g[1] = (uint16_t)min_v; if (len > 2u) ent_bic_decode<uint16_t>(b, g + 2, len - 2u, min_v, 65535u); }
            } else if (code == 62u) {
                const uint32_t head = b.delta16s(); len = head >> 3;
                uint32_t min_v = (head & 2u) ? b.gamma8() : b.g16(), max_v = (head & 4u) ? b.gamma8() : b.g16();
                max_v = (65535u - max_v) & 0xffffu; min_v &= 0xffffu;
                if (len < 3u || rd.bad) err = 1;
                else { g[0] = (uint16_t)(head & 1u); g[1] = (uint16_t)min_v; if (len > 3u) ent_bic_decode<uint16_t>(b, g + 2, len - 3u, min_v, max_v); g[len - 1u] = (uint16_t)max_v; }
            } else {                                     // 61: header byte, head word, [min / max / min0 / min1, interpolated ends] or plain deltas
                h3 = b.bits(8u); const uint32_t head = b.delta16s(); len = head >> 3;
                g[0] = (uint16_t)(head & 1u);
                if (len < 1u || rd.bad) err = 1;
                else if ((h3 & 0x80u) && len < 4u) {
                    uint32_t prev = 0;
                    for (uint32_t k = 1; k < len; ++k) { prev = (k == 1u) ? b.delta16s() : prev + b.delta16s(); g[k] = (uint16_t)prev; }
                } else if (len < 3u) err = 1;
                else {
                    uint32_t min_v = (head & 2u) ? b.bits(8u) : b.g16(), max_v;
                    if (head & 4u) { max_v = b.bits(8u); max_v = ((max_v << 3) | (h3 & 7u)) & 0xffffu; } else max_v = b.g16();
                    max_v = (65535u - max_v) & 0xffffu;
                    g[1] = (uint16_t)min_v;
                    uint32_t min0 = 0, min1 = 0;
                    if (!(h3 & 8u))    min0 = ((h3 & 0x10u) ? b.gamma8() : b.delta16()) & 0xffffu;
                    if (!(h3 & 0x40u)) min1 = ((h3 & 0x20u) ? b.gamma8() : b.delta16()) & 0xffffu;
                    if (len > 3u) ent_bic_decode<uint16_t>(b, g + 2, len - 3u, (min_v + 1u) & 0xffffu, max_v);
                    g[len - 1u] = (uint16_t)(max_v + 1u); g[len] = 65535u;
                    if ((h3 & 0x80u) || min0 || min1) ent_gap_restore_mins(g, len, min0, min1);
                }
            }
            if (!err) g[len] = 65535u;
            if (rd.bad) err = 1;
            n = len; flags = h3;
        }
        err = bme_bcast(err); n = bme_bcast(n); flags = bme_bcast(flags);
        if (err) return BMB200_ERR_BADARG;
        ent_apply_gap(c, c.la, n);
        if (code == 61u && (flags & 0x80u)) {            // exception lists: single bits put back after the runs (gap_set_value)
            for (int pass = 0; pass < 2; ++pass) {
                uint32_t cnt = 0; int h = 0;
                if (lead) h = ent_decode_array(b, c.lb, c.lc, c.wf, &cnt, 0u);
                h = (int)bme_bcast((uint32_t)h); cnt = bme_bcast(cnt);
                if (h < 0) return BMB200_ERR_BADARG;
                ent_apply_pos(c, c.lb, cnt, (h & 0x10) ? 1u : 0u);
                if (h & 0x20) break;
            }
        }
        break; }
    case 21u: case 23u: {                                // gamma-coded position deltas (GAP block as an array), optionally inverted
        *gap_family = 1u;
        if (lead) {
            n = b.gamma() & 0xffffu; uint32_t prev = 0;
            for (uint32_t k = 0; k < n && !rd.bad; ++k) { uint32_t v = b.gamma(); if (!k) --v; prev = (prev + v) & 0xffffu; c.la[k] = (uint16_t)prev; }
            err = rd.bad;
        }
        err = bme_bcast(err); n = bme_bcast(n);
        if (err) return BMB200_ERR_BADARG;
        ent_apply_pos(c, c.la, n, 1u);
        if (code == 23u) ent_invert(c);
        break; }
    case 65u: case 66u: {                                // interpolated positions of a bit-block (v3s), optionally inverted
        if (lead) {
            uint32_t cnt = b.delta16s() & 0xffffu, min_v = 0, max_v = 65535u, k = 0;
            if (b.bits(1u)) {
                min_v = b.delta16s() & 0xffffu; cnt = (cnt - 2u) & 0xffffu; max_v = (65536u - b.delta16s()) & 0xffffu;
                c.la[k++] = (uint16_t)min_v; c.la[k++] = (uint16_t)max_v;
                min_v = (min_v + 1u) & 0xffffu; max_v = (max_v - 1u) & 0xffffu;
            }
            if (rd.bad) err = 1;
            else if (cnt) ent_bic_decode<uint16_t>(b, c.la + k, cnt, min_v, max_v);
            n = k + cnt; if (rd.bad) err = 1;
        }
        err = bme_bcast(err); n = bme_bcast(n);
        if (err) return BMB200_ERR_BADARG;
        ent_apply_pos(c, c.la, n, 1u);
        if (code == 66u) ent_invert(c);
        break; }
    case 63u: case 64u: {                                // v3 bit-block: single bits, then runs (start, length - 1), optionally inverted
        int h = 0;
        if (lead) h = ent_decode_array(b, c.la, c.lc, c.wf, &n, 0u);
        h = (int)bme_bcast((uint32_t)h); n = bme_bcast(n);
        if (h < 0) return BMB200_ERR_BADARG;
        ent_apply_pos(c, c.la, n, 1u);
        if (!(h & 0x20)) {
            uint32_t r_cnt = 0, l_cnt = 0; int h2 = 0;
            if (lead) {
                h2 = ent_decode_array(b, c.la, c.lc, c.wf, &r_cnt, 0u);
                if (h2 >= 0 && !r_cnt) h2 = -1;
                if (h2 >= 0) h2 = ent_decode_array(b, c.lb, c.lc, c.wf, &l_cnt, r_cnt);
                if (h2 >= 0) {
                    if (l_cnt > r_cnt) l_cnt = r_cnt;
                    if ((h2 & 3) == 1) for (uint32_t i = 0; i < l_cnt; ++i) c.lb[i] = (uint16_t)(c.lb[i] - c.la[i]);
                }
            }
            h2 = (int)bme_bcast((uint32_t)h2); l_cnt = bme_bcast(l_cnt);
            if (h2 < 0) return BMB200_ERR_BADARG;
            ent_apply_runs(c, c.la, c.lb, l_cnt);
        }
        if (code == 64u) ent_invert(c);
        break; }
    default:
        return BMB200_ERR_UNSUPPORTED;
    }
    err = bme_bcast(lead ? rd.bad : 0u);
    return err ? BMB200_ERR_BADARG : BMB200_OK;
}

// the reader state is copied into locals so that, with everything above inlined, the bit accumulator and the stream position live in
// registers during the (sequential, latency-bound) decode loops instead of in the caller's stack frame
BME_HDN int ent_decode_block(const EntCtx& c, uint32_t code, EntRd& rd_io, uint32_t* gap_family)
{
    EntRd rd = rd_io;
    const int rc = ent_decode_block_impl(c, code, rd, gap_family);
    rd_io = rd;
    return rc;
}

// super-block token 68 (set_sblock_bienc_v3): lane 0 decodes the ascending 24-bit positions into `arr` (65536 u32 = lists a + b)
BME_HD int ent_decode_sblock_impl(EntRd& rd, uint32_t* arr, uint32_t* len_out, uint32_t* sb_out, uint32_t* min0_out)
{
    EntBits b; b.init(&rd);
    const uint32_t flag = b.bits(8u);
    const uint32_t len = (flag & 0x10u) ? b.delta16() : b.bits(8u);
    uint32_t min_v, max_v, min0 = 0, sb;
    if (flag & 8u) { const uint32_t j = b.gamma(), nbit = b.g16(); min_v = j * 65536u + nbit; }
    else min_v = (flag & 4u) ? b.g16() : b.bits(8u);
    if (flag & 0x40u) max_v = b.g24(); else max_v = (flag & 0x20u) ? b.g16() : b.bits(8u);
    max_v = 256u * 65536u - max_v;
    if (flag & 0x80u) { switch (b.gamma()) { case 1: min0 = b.gamma(); break; case 2: min0 = b.bits(8u); break; default: min0 = b.g16(); break; } }
    if ((flag & 3u) == 3u) sb = b.gamma() - 1u; else sb = (flag & 2u) ? b.g32() : (flag & 1u) ? b.g16() : b.bits(8u);
    if (len < 2u || len > 65536u || rd.bad) return BMB200_ERR_BADARG;
    arr[0] = min_v; arr[len - 1u] = max_v;
    if (len > 2u) ent_bic_decode<uint32_t>(b, arr + 1, len - 2u, min_v + 1u, max_v - 1u);
    if (rd.bad) return BMB200_ERR_BADARG;
    *len_out = len; *sb_out = sb; *min0_out = min0;
    return BMB200_OK;
}
// Team-wide: lane 0 decodes the interpolative list (the only sequential part); putting the minimal delta back (arr_restore_min,
// src/bmfunc.h:2657: arr[i] += i * min0 in closed form) and the validity check (strictly ascending, inside the super-block) are
// done by all lanes -- as lane-0 loops over the list in global memory they cost more than the decode itself.
BME_HDN int ent_decode_sblock(const EntCtx& c, EntRd& rd_io, uint32_t* arr, uint32_t* len_out, uint32_t* sb_out)
{
    uint32_t rc = 0, len = 0, sb = 0, min0 = 0;
    if (c.t.lane == 0u) {
        EntRd rd = rd_io;
        rc = (uint32_t)ent_decode_sblock_impl(rd, arr, &len, &sb, &min0);
        rd_io = rd;
    }
    rc = bme_bcast(rc); len = bme_bcast(len); sb = bme_bcast(sb); min0 = bme_bcast(min0);
    if (rc) return (int)rc;
    bme_sync();
    if (min0) { for (uint32_t i = 1u + c.t.lane; i < len; i += c.t.nl) arr[i] += i * min0; bme_sync(); }
    uint32_t bad = 0;
    for (uint32_t i = 1u + c.t.lane; i < len; i += c.t.nl) bad |= (arr[i] <= arr[i - 1u]) ? 1u : 0u;
    if (c.t.lane == 0u && arr[len - 1u] >= 256u * 65536u) bad = 1u;
    if (bme_sum(bad)) return BMB200_ERR_BADARG;
    bme_sync();
    *len_out = len; *sb_out = sb;
    return BMB200_OK;
}
// positions [k0, k1) of a decoded super-block list that fall into block `blk` -> bitmap (cleared first)
BME_HD void ent_sblock_fill(const EntCtx& c, const uint32_t* arr, uint32_t k0, uint32_t k1)
{
    ent_clear(c);
    for (uint32_t k = k0 + c.t.lane; k < k1; k += c.t.nl) { const uint32_t p = arr[k] & 65535u; bme_or(&c.bm[p >> 5], 1u << (p & 31u)); }
    bme_sync();
}

// ------------------------------------------------------------------------------------------------------------------
// the walk of one serialized vector (deserializer<BV>::deserialize token loop, src/bmserial.h:5704-6072)
// ------------------------------------------------------------------------------------------------------------------
struct EntWalkOut {
    BlobTok* toks; uint32_t cap; uint32_t n;      // block records of this vector, in block order
    uint8_t* full; uint32_t full_stride;          // full[nb * full_stride] = 1 for all-ones blocks (pre-zeroed)
};

BME_HD void ent_push(EntWalkOut& o, const BlobTok& t, uint32_t* err) { if (o.n < o.cap) o.toks[o.n++] = t; else *err = BMB200_ERR_RANGE; }

// explicit-length tokens are only measured here (type + extent, exactly what the host walker of capi.cu records); their payload
// is decoded by blob_decode_kernel.  Entropy-coded tokens are decoded to find their end and their block shape.
//
// One call walks one SEGMENT of a vector's stream: the bytes [seg.start, seg.end) of the BLOB, beginning at block seg.nb0.  A BLOB
// without bookmarks is a single segment that starts at its header (seg.start == 0, header parsed here) and runs to its end token;
// a BLOB written with serializer::set_bookmarks (src/bmserial.h:1487, process_bookmark :3567) is cut at its sync marks by
// ent_find_segments, and the segments of one vector are walked by different warps at the same time.
struct EntSeg {
    uint64_t blob_off, blob_size;   // the BLOB inside the staging buffer
    uint64_t start, end;            // segment bytes, relative to the BLOB; start == 0: begin at the header
    uint32_t nb0;                   // block index at `start`
    uint32_t bounded;               // 1: the segment ends exactly at `end` (a sync mark follows), 0: it ends with an end-of-stream token
    uint32_t vec;                   // vector the segment belongs to
    uint32_t tok_base;              // first BlobTok slot of this segment in the token table
};

BME_HDN int ent_walk_segment(const EntCtx& c, const uint8_t* stg, const EntSeg& sg, uint32_t n_blocks, EntWalkOut& o)
{
    const bool lead = (c.t.lane == 0u);
    const uint64_t blob_off = sg.blob_off;
    EntRd rd{stg, blob_off + sg.start, blob_off + (sg.bounded ? sg.end : sg.blob_size), 0u};
    uint32_t err = 0;
    if (lead && sg.start == 0u) {
        const uint32_t hf = rd.u8();
        if (!(hf & (1u << 3))) rd.u8();                                              // byte order
        if (hf & ((1u << 2) | (1u << 6))) err = BMB200_ERR_UNSUPPORTED;              // id list / XOR compression
        if (!(hf & (1u << 4))) rd.skip(8);                                           // GAP levels
        if (hf & (1u << 1)) { rd.u32(); if (hf & (1u << 5)) rd.u32(); }              // size (64-bit in a BM64ADDR stream)
        if (rd.bad && !err) err = BMB200_ERR_BADARG;
    }
    err = bme_bcast(err);
    if (err) return (int)err;
    uint64_t nb = sg.nb0;
    for (;;) {
        // ---- lane 0 reads the token byte and settles everything that needs no team work ----
        uint32_t act = 0, code = 0, cnt = 0;      // act: 0 next token, 1 end, 2 all-ones run of cnt blocks, 3 entropy block, 4 super-block, 5 error (code)
        if (lead && sg.bounded && rd.p == rd.end) act = 1;                           // reached the sync mark that closes this segment
        else if (lead) {
            const uint32_t bt = rd.u8();
            BlobTok t; t.nb = (uint32_t)nb; t.type = 0; t.off = rd.p - blob_off; t.aux = 0; t.first = 0; t.gap_words = 0; t.kind = BMB200_BLK_BIT;
            bool blk = false;
            if (rd.bad) { act = 5; code = BMB200_ERR_BADARG; }
            else if (bt & 0x80u) nb += bt & 0x7fu;
            else switch (bt) {
            case 0: case 9: act = 1; break;
            case 1: ++nb; break;
            case 3: nb += rd.u8(); break;
            case 5: nb += rd.u16(); break;
            case 7: nb += rd.u32(); break;
            case 25: nb += rd.u64(); break;                                          // set_block_64zero (BM64ADDR streams)
            case 10: act = 2; cnt = 0xffffffffu; break;
            case 2: act = 2; cnt = 1; break;
            case 4: act = 2; cnt = rd.u8(); break;
            case 6: act = 2; cnt = rd.u16(); break;
            case 8: act = 2; cnt = rd.u32(); break;
            case 26: { const uint64_t c64 = rd.u64(); act = 2; cnt = c64 > 0xfffffffeull ? 0xfffffffeu : (uint32_t)c64; break; }   // set_block_64one
            case 47: rd.skip(2); break; case 48: rd.skip(3); break; case 49: rd.skip(4); break;      // bookmarks: skip offsets
            case 50: rd.skip(1); break; case 51: rd.skip(2); break; case 52: rd.skip(3); break;      // sync marks
            case 53: rd.skip(4); break; case 54: rd.skip(6); break; case 55: rd.skip(8); break;
            case 11: t.type = 0 /*DB_BIT*/; rd.skip(BMB200_BLOCK_BYTES); blk = true; break;
            case 17: { const uint32_t head = rd.u16(), tail = rd.u16();
                       if (tail >= BMB200_BLOCK_WORDS || head > tail) { act = 5; code = BMB200_ERR_BADARG; break; }
                       t.type = 1 /*DB_BIT_INTERVAL*/; rd.skip(4ull * (tail - head + 1u)); blk = true; break; }
            case 22: { uint32_t rt = rd.u8(), j = 0;
                       while (j < BMB200_BLOCK_WORDS && !rd.bad) { const uint32_t len = rd.u16(); if (rt) rd.skip(4ull * len); j += len; rt ^= 1u; }
                       if (j != BMB200_BLOCK_WORDS) { act = 5; code = BMB200_ERR_BADARG; break; }
                       t.type = 2 /*DB_BIT_0RUNS*/; blk = true; break; }
            case 34: { const uint64_t d0 = rd.u64(); uint32_t pc = bme_popc((uint32_t)d0) + bme_popc((uint32_t)(d0 >> 32));
                       t.type = 3 /*DB_BIT_DIGEST0*/; rd.skip(128ull * pc); blk = true; break; }
            case 16: case 30: { const uint32_t n = rd.u16(); t.type = bt == 16 ? 4u : 5u /*DB_ARRBIT(_INV)*/; rd.skip(2ull * n); blk = true; break; }
            case 14: case 15: { const uint32_t hdr = rd.u16(), len = hdr >> 3;
                       if (len < 1u || len > BMB200_GAP_MAX_WORDS - 5u) { act = 5; code = BMB200_ERR_UNSUPPORTED; break; }
                       t.type = 6 /*DB_GAP16*/; t.kind = BMB200_BLK_GAP; t.first = hdr & 1u; t.gap_words = len + 1u; rd.skip(2ull * (len - 1u)); blk = true; break; }
            case 19: { const uint32_t pos = rd.u16(); t.type = 8 /*DB_ARRGAP*/; t.kind = BMB200_BLK_GAP; t.aux = 1; t.off = rd.p - 2u - blob_off;
                       t.first = pos == 0u; t.gap_words = 4; blk = true; break; }
            case 18: case 24: { const uint32_t n = rd.u16();
                       if (!n || n > 2048u) { act = 5; code = BMB200_ERR_UNSUPPORTED; break; }
                       const uint64_t a0 = rd.p; rd.skip(2ull * n);
                       if (rd.bad) break;
                       const uint32_t first_pos = (uint32_t)stg[a0] | ((uint32_t)stg[a0 + 1] << 8);
                       t.type = bt == 18 ? 8u : 9u /*DB_ARRGAP(_INV)*/; t.kind = BMB200_BLK_GAP; t.aux = n; t.off = a0 - blob_off;
                       t.first = (uint32_t)(first_pos == 0u) ^ (uint32_t)(bt == 24);
                       t.gap_words = arrgap_measure(stg + a0, n);
                       if (!t.gap_words) { act = 5; code = BMB200_ERR_BADARG; break; }
                       blk = true; break; }
            case 67: {                                   // plain 16-bit run ends -> explicit token, gamma-coded -> entropy token
                       const uint64_t w0 = rd.p; EntBits pb; pb.init(&rd);
                       const uint32_t len = pb.gamma() + 1u, start = pb.bit(), use_gamma = pb.bit();
                       if (rd.bad) break;
                       if (use_gamma) { rd.p = w0; act = 3; code = 67u; break; }
                       if (len > BMB200_GAP_MAX_WORDS - 5u) { act = 5; code = BMB200_ERR_UNSUPPORTED; break; }
                       const uint32_t used = (uint32_t)((rd.p - w0) * 8u) - pb.have;
                       rd.p = w0; rd.skip(4ull * (((uint64_t)used + 16ull * (len - 1u) + 31u) / 32u));
                       t.type = 7 /*DB_GAP_V3*/; t.kind = BMB200_BLK_GAP; t.off = w0 - blob_off; t.aux = used | (len << 8); t.first = start; t.gap_words = len + 1u;
                       blk = true; break; }
            case 21: case 23: case 33: case 61: case 62: case 63: case 64: case 65: case 66: act = 3; code = bt; break;
            case 68: act = 4; break;
            default: act = 5; code = BMB200_ERR_UNSUPPORTED; break;
            }
            if (rd.bad && act != 5u) { act = 5; code = BMB200_ERR_BADARG; }
            if (blk && act == 0u) { if (nb < n_blocks) ent_push(o, t, &err); ++nb; if (err) { act = 5; code = err; } }
        }
        act = bme_bcast(act);
        if (act == 0u) continue;
        code = bme_bcast(code); cnt = bme_bcast(cnt); nb = bme_bcast64(nb);
        if (act == 1u) return BMB200_OK;
        if (act == 5u) return (int)code;
        if (act == 2u) {                                 // all-ones blocks
            uint64_t e = (cnt == 0xffffffffu) ? (uint64_t)n_blocks : nb + cnt; const uint64_t e_clip = e < n_blocks ? e : n_blocks;
            for (uint64_t q = nb + c.t.lane; q < e_clip; q += c.t.nl) o.full[q * o.full_stride] = 1;
            if (cnt == 0xffffffffu) return BMB200_OK;
            nb = e;
            continue;
        }
        if (act == 3u) {                                 // entropy-coded block: decode, measure
            const uint64_t off = rd.p - blob_off;
            uint32_t gap_family = 0;
            const int rc = ent_decode_block(c, code, rd, &gap_family);
            if (rc) return rc;
            if (nb < n_blocks) {
                const uint32_t runs = ent_count_runs(c);
                if (lead) {
                    BlobTok t; t.nb = (uint32_t)nb; t.type = kTokEntropy | code; t.off = off; t.aux = 0; t.first = c.bm[0] & 1u;
                    const bool as_gap = gap_family && runs + 1u <= kGapFitWords;
                    t.kind = as_gap ? BMB200_BLK_GAP : BMB200_BLK_BIT; t.gap_words = as_gap ? runs + 1u : 0u;
                    ent_push(o, t, &err);
                }
                err = bme_bcast(err);
                if (err) return (int)err;
            }
            ++nb;
            continue;
        }
        // act == 4: super-block position list; every touched block becomes a GAP block (set_bit_no_check under BM_GAP)
        {
            const uint64_t off = rd.p - blob_off;
            uint32_t* arr = reinterpret_cast<uint32_t*>(c.la);            // lists a + b are contiguous: 65536 u32
            uint32_t len = 0, sb = 0;
            const uint32_t rc = (uint32_t)ent_decode_sblock(c, rd, arr, &len, &sb);
            if (rc) return (int)rc;
            const uint64_t nb0 = nb & ~255ull;
            if ((uint64_t)sb * 256u != nb0) return BMB200_ERR_BADARG;
            bme_sync();
            if (lead && nb0 < n_blocks) {
                BlobTok t; t.nb = (uint32_t)nb0; t.type = kTokEntropy | 68u; t.off = off; t.aux = (uint32_t)nb0; t.first = 0; t.gap_words = 0; t.kind = 0;
                ent_push(o, t, &err);
            }
            for (uint32_t k = 0; k < len; ) {
                const uint32_t blk = arr[k] >> 16; uint32_t k2 = k;
                if (blk >= 256u) return BMB200_ERR_BADARG;
                while (k2 < len && (arr[k2] >> 16) == blk) ++k2;
                if (nb0 + blk < n_blocks) {
                    ent_sblock_fill(c, arr, k, k2);
                    const uint32_t runs = ent_count_runs(c);
                    if (lead) {
                        BlobTok t; t.nb = (uint32_t)(nb0 + blk); t.type = kTokSbMember; t.off = off; t.aux = (uint32_t)nb0; t.first = c.bm[0] & 1u;
                        const bool as_gap = runs <= kGapFitWords;              // stays GAP while runs <= glen(3) - 4 (gap_block_set_no_ret, src/bm.h:4800)
                        t.kind = as_gap ? BMB200_BLK_GAP : BMB200_BLK_BIT; t.gap_words = as_gap ? runs + 1u : 0u;
                        ent_push(o, t, &err);
                    }
                }
                k = k2;
            }
            err = bme_bcast(err);
            if (err) return (int)err;
            nb = nb0 + 256u;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// pass 2, one entropy-coded token: decode again and write the block(s) into the arena.  code 68 (super-block) looks its
// member blocks up in the descriptor table (dst = first column of the super-block), every other token has its slot in dst
// (bit kinds: block index in bit_pool; GAP kinds: absolute 16-byte unit in gap_pool, aux2 bit 0 = lead pad).
// ------------------------------------------------------------------------------------------------------------------
struct EntSetView { uint32_t n_vec, n_blocks; const uint32_t* desc; const uint64_t* bit_base; const uint64_t* gap_base; };

BME_HD void ent_store_gap(const EntCtx& c, uint16_t* unit, uint32_t pad, bool sb_member = false)
{
    const uint32_t runs = ent_count_runs(c);
    if (c.t.lane == 0u && pad) unit[0] = 0xffffu;
    ent_write_gap(c, unit + pad, runs, sb_member);
}

BME_HDN int ent_emit(const EntCtx& c, const uint8_t* stg, uint64_t src, uint64_t end, uint32_t code, uint32_t v, uint64_t dst, uint32_t kind,
                     uint32_t aux2, const EntSetView& set, uint32_t* bit_pool, uint16_t* gap_pool)
{
    EntRd rd{stg, src, end, 0u};
    bme_sync();
    if (code == 68u) {
        uint32_t* arr = reinterpret_cast<uint32_t*>(c.la);
        uint32_t len = 0, sb = 0;
        const uint32_t rc = (uint32_t)ent_decode_sblock(c, rd, arr, &len, &sb);
        if (rc) return (int)rc;
        bme_sync();
        for (uint32_t k = 0; k < len; ) {
            const uint32_t blk = arr[k] >> 16; uint32_t k2 = k;
            while (k2 < len && (arr[k2] >> 16) == blk) ++k2;
            const uint64_t col = dst + blk;
            if (blk < 256u && col < set.n_blocks) {
                const uint32_t d = set.desc[col * set.n_vec + v], kd = d & 3u, rel = (d >> 2) & BMB200_DESC_REL_MASK;
                if (kd == BMB200_BLK_BIT || kd == BMB200_BLK_GAP) {
                    ent_sblock_fill(c, arr, k, k2);
                    if (kd == BMB200_BLK_BIT) ent_write_bits(c, bit_pool + (set.bit_base[col] + rel) * (size_t)kEntWords);
                    else ent_store_gap(c, gap_pool + (set.gap_base[col] + rel) * (size_t)BMB200_GAP_UNIT_WORDS, (d & BMB200_DESC_GAP_PAD) ? 1u : 0u, true);
                }
            }
            k = k2;
        }
        return BMB200_OK;
    }
    uint32_t gap_family = 0;
    const int rc = ent_decode_block(c, code, rd, &gap_family);
    if (rc) return rc;
    if (kind == BMB200_BLK_BIT) ent_write_bits(c, bit_pool + dst * (size_t)kEntWords);
    else ent_store_gap(c, gap_pool + dst * (size_t)BMB200_GAP_UNIT_WORDS, aux2 & 1u);
    return BMB200_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// host side of pass 1: cut a BLOB at its bookmarks.  Pure byte parsing (no decoding): header, then the chain
//   [set_nb_bookmark16/24/32, offset] ... tokens ... [set_nb_sync_mark8/16/24/32, blocks since the bookmark] [bookmark] ...
// (serializer::process_bookmark, src/bmserial.h:3567-3665; reader :5897-5950).  offset = bytes from the end of the offset field
// to the sync mark; 0 = the bookmark was never closed (the last one).  Returns segments relative to the BLOB; a BLOB without a
// leading bookmark is one unbounded segment starting at its header.
// ------------------------------------------------------------------------------------------------------------------
inline int ent_find_segments(const uint8_t* blob, uint64_t size, uint32_t vec, uint64_t blob_off, std::vector<EntSeg>& out)
{
    auto whole = [&]() { EntSeg s{}; s.blob_off = blob_off; s.blob_size = size; s.start = 0; s.end = size; s.nb0 = 0; s.bounded = 0; s.vec = vec; out.push_back(s); return BMB200_OK; };
    if (size < 2) return BMB200_ERR_BADARG;
    uint64_t p = 0;
    const uint32_t hf = blob[p++];
    if (!(hf & (1u << 3))) ++p;
    if (!(hf & (1u << 4))) p += 8;
    if (hf & (1u << 1)) p += (hf & (1u << 5)) ? 8 : 4;
    if (p >= size) return BMB200_ERR_BADARG;
    if (blob[p] < 47u || blob[p] > 49u) return whole();              // no bookmark right after the header
    const size_t first = out.size();
    uint64_t nb = 0; bool head = true;
    for (;;) {
        if (p >= size) return BMB200_ERR_BADARG;
        const uint32_t bt = blob[p];
        if (bt < 47u || bt > 49u) {                                   // the chain ends: the rest is one unbounded segment
            EntSeg s{}; s.blob_off = blob_off; s.blob_size = size; s.start = p; s.end = size; s.nb0 = (uint32_t)nb; s.bounded = 0; s.vec = vec;
            out.push_back(s); break;
        }
        const uint32_t fsz = bt == 47u ? 2u : bt == 48u ? 3u : 4u;
        if (p + 1 + fsz > size) return BMB200_ERR_BADARG;
        uint64_t ofs = 0; for (uint32_t i = 0; i < fsz; ++i) ofs |= (uint64_t)blob[p + 1 + i] << (8 * i);
        const uint64_t body = p + 1 + fsz;
        EntSeg s{}; s.blob_off = blob_off; s.blob_size = size; s.vec = vec; s.nb0 = (uint32_t)nb;
        s.start = head ? 0 : body;                                    // the first segment starts at the header (and meets its own bookmark token)
        head = false;
        if (!ofs) { s.end = size; s.bounded = 0; out.push_back(s); break; }
        const uint64_t sync = body + ofs;
        if (sync + 2 > size) return BMB200_ERR_BADARG;
        const uint32_t st = blob[sync];
        if (st < 50u || st > 53u) { out.resize(first); return whole(); }   // not a mark this walker follows: fall back to one segment
        const uint32_t dsz = st - 49u;
        if (sync + 1 + dsz > size) return BMB200_ERR_BADARG;
        uint64_t delta = 0; for (uint32_t i = 0; i < dsz; ++i) delta |= (uint64_t)blob[sync + 1 + i] << (8 * i);
        s.end = sync; s.bounded = 1; out.push_back(s);
        nb += delta; if (nb > 0xffffffffull) return BMB200_ERR_BADARG;
        p = sync + 1 + dsz;
    }
    return BMB200_OK;
}

}  // namespace bmb200

// ======================================================================================================================
// kernels (device build only)
// ======================================================================================================================
#if defined(__CUDACC__)
#include "common.cuh"
#include "blob_kernel.cuh"
namespace bmb200 {

constexpr int kEntThreads = 32;        // one warp per CTA: the bitmap (8 KB of shared memory) is the only per-CTA resource

__device__ __forceinline__ EntCtx ent_make_ctx(uint32_t* s_bm, uint8_t* scratch_base, uint32_t slot)
{
    uint8_t* sc = scratch_base + (size_t)slot * kEntScratchBytes;
    EntCtx c;
    c.t.lane = threadIdx.x & 31u; c.t.nl = 32u;
    c.bm = s_bm;
    c.la = reinterpret_cast<uint16_t*>(sc);
    c.lb = c.la + kEntListCap;
    c.lc = c.lb + kEntListCap;
    c.wf = reinterpret_cast<uint32_t*>(c.lc + kEntListCap);
    return c;
}

// pass 1: one warp per segment (a whole vector, or one bookmark interval of it).  toks: token table, segment i writes from
// segs[i].tok_base (at most tok_cap[i] records); n_toks, status: [n_segs]; full: [n_blocks][n_vec] (zeroed by the caller)
// Work items are pulled from a counter in the order the host sorted them (longest stream first): the items are sequential decodes
// of very different lengths, and the launch ends with its slowest warp.
__device__ __forceinline__ uint32_t ent_next_item(uint32_t* counter)
{
    uint32_t i = 0;
    if ((threadIdx.x & 31u) == 0u) i = atomicAdd(counter, 1u);
    return __shfl_sync(0xffffffffu, i, 0);
}

__global__ void __launch_bounds__(kEntThreads) blob_walk_kernel(const uint8_t* __restrict__ stg, const EntSeg* __restrict__ segs, const uint32_t* __restrict__ tok_cap,
                                                                const uint32_t* __restrict__ order, uint32_t* __restrict__ counter,
                                                                uint32_t n_segs, uint32_t n_vec, uint32_t n_blocks, BlobTok* __restrict__ toks,
                                                                uint32_t* __restrict__ n_toks, int* __restrict__ status, uint8_t* __restrict__ full,
                                                                uint8_t* __restrict__ scratch)
{
    __shared__ __align__(16) uint32_t s_bm[kEntWords];
    const EntCtx c = ent_make_ctx(s_bm, scratch, blockIdx.x);
    for (uint32_t q = ent_next_item(counter); q < n_segs; q = ent_next_item(counter)) {
        const uint32_t i = order[q];
        const EntSeg sg = segs[i];
        EntWalkOut o; o.toks = toks + sg.tok_base; o.cap = tok_cap[i]; o.n = 0; o.full = full + sg.vec; o.full_stride = n_vec;
        const int rc = ent_walk_segment(c, stg, sg, n_blocks, o);
        if (c.t.lane == 0) { n_toks[i] = o.n; status[i] = rc; }
        __syncwarp();
    }
}

// pass 2: one warp per entropy-coded token (grid-stride).  A BlobRec of type kTokEntropy | code carries: src = payload offset in the
// staging buffer, aux = vector index, kind = BMB200_BLK_*, dst / aux2 as for the explicit GAP kinds (blob_kernel.cuh); a
// super-block rec (code 68) carries dst = first column of the super-block and finds the slot of every member block in the
// descriptor table.
__global__ void __launch_bounds__(kEntThreads) blob_entropy_kernel(const uint8_t* __restrict__ stg, const uint64_t* __restrict__ blob_off,
                                                                   const uint64_t* __restrict__ blob_size, const BlobRec* __restrict__ recs,
                                                                   uint32_t n_recs, uint32_t* __restrict__ counter, const SetView set, uint32_t* __restrict__ bit_pool,
                                                                   uint16_t* __restrict__ gap_pool, int* __restrict__ status, uint8_t* __restrict__ scratch,
                                                                   unsigned long long* __restrict__ dur /* BMB200_TRACE: clocks per rec, or null */)
{
    __shared__ __align__(16) uint32_t s_bm[kEntWords];
    const EntCtx c = ent_make_ctx(s_bm, scratch, blockIdx.x);
    const EntSetView sv{set.n_vec, set.n_blocks, set.desc, set.bit_base, set.gap_base};
    for (uint32_t ri = ent_next_item(counter); ri < n_recs; ri = ent_next_item(counter)) {      // recs arrive sorted, longest payload first
        const BlobRec r = recs[ri];
        const uint32_t v = r.aux;
        const long long t0 = dur ? clock64() : 0;
        const int rc = ent_emit(c, stg, r.src, blob_off[v] + blob_size[v], r.type & 0xffu, v, r.dst, r.kind, r.aux2, sv, bit_pool, gap_pool);
        if (rc && c.t.lane == 0u) atomicCAS(status, 0, rc);
        if (dur && c.t.lane == 0u) dur[ri] = (unsigned long long)(clock64() - t0);
        __syncwarp();
    }
}

}  // namespace bmb200
#endif
