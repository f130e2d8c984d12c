// blob_kernel.cuh -- deserialize-to-device: serialized bvector BLOBs decoded straight into the HBM arena (sm_100a).
//
// Replaces, for operands that arrive as BitMagic serialization BLOBs, bm::deserialize + the upload of the materialised
// blocks (src/bmserial.h:4152, deserializer::deserialize :5578-6090): the host only WALKS the token stream (block type +
// explicit payload length, capi.cu), the compressed bytes cross PCIe once, and one CTA per block token writes the block
// in its arena slot (bit-block, or GAP block in the flat-streamable form).  Token decoders (file:line = the reference
// routine each one restates):
//   DB_BIT          set_block_bit            decode_block_bit            :5493   2048 raw words
//   DB_BIT_INTERVAL set_block_bit_interval   decode_block_bit_interval   :5511   head, tail, words[head..tail]
//   DB_BIT_0RUNS    set_block_bit_0runs      read_0runs_block            :4738   alternating zero / data word runs
//   DB_BIT_DIGEST0  set_block_bit_digest0    read_digest0_block                  64-bit wave digest + the non-zero waves
//   DB_ARRBIT(_INV) set_block_arrbit(_inv)   decode_arrbit :5539, :5424          list of bits ON (OFF)
//   DB_GAP16        set_block_gap / _gapbit  deserialize_gap             :5243   header + (len-1) u16 run ends
//   DB_GAP_V3       set_block_gap_egamma_v3  read_gap_block              :5050   bit stream: gamma(len-1), start, 0, 16-bit run ends
//   DB_ARRGAP(_INV) set_block_arrgap(_inv), set_block_bit_1bit  :4833-4845, gap_set_array  sorted positions -> GAP runs
// Encodings that need a sequential entropy decoder (gamma values, binary interpolative coding, super-block lists) make the host
// walker give up (BMB200_ERR_UNSUPPORTED internally): bmb200_set_upload_blobs then walks and decodes the streams on the device,
// blob_entropy.cuh.  There is no CPU fallback.
#pragma once
#include "common.cuh"

namespace bmb200 {

enum : uint32_t { DB_BIT = 0, DB_BIT_INTERVAL, DB_BIT_0RUNS, DB_BIT_DIGEST0, DB_ARRBIT, DB_ARRBIT_INV, DB_GAP16, DB_GAP_V3, DB_ARRGAP, DB_ARRGAP_INV };

struct BlobRec {
    uint64_t src;        // byte offset of the token payload in the staging buffer
    uint64_t dst;        // bit kinds: block index in bit_pool; GAP kinds: absolute 16-byte unit in gap_pool
    uint32_t type;       // DB_*
    uint32_t aux;        // DB_GAP_V3: bit offset of the first run end | len << 8 ... see capi.cu; DB_ARRGAP*: element count (1bit: 1)
    uint32_t aux2;       // GAP kinds: bit 0 = lead pad, bit 1 = first-run value; explicit-length GAP tokens: u16 words of the arena slot << 8
    uint32_t kind;       // entropy-coded tokens (blob_entropy.cuh): BMB200_BLK_BIT / BMB200_BLK_GAP of the decoded block
};

constexpr int kBlobThreads = 256;

__device__ __forceinline__ uint32_t b_rd8(const uint8_t* __restrict__ s, uint64_t o)  { return s[o]; }
__device__ __forceinline__ uint32_t b_rd16(const uint8_t* __restrict__ s, uint64_t o) { return (uint32_t)s[o] | ((uint32_t)s[o + 1] << 8); }
// 32-bit little-endian word at an arbitrary byte offset: two aligned loads + funnel shift
__device__ __forceinline__ uint32_t b_rd32(const uint8_t* __restrict__ s, uint64_t o)
{
    const uint64_t a = o & ~3ull; const uint32_t sh = (uint32_t)(o & 3ull) * 8u;
    const uint32_t lo = *reinterpret_cast<const uint32_t*>(s + a);
    if (!sh) return lo;
    const uint32_t hi = *reinterpret_cast<const uint32_t*>(s + a + 4);
    return __funnelshift_r(lo, hi, sh);
}
// 16 bits at an arbitrary BIT offset of a stream of little-endian 32-bit words that starts at byte offset o (bit_in order)
__device__ __forceinline__ uint32_t b_bits16(const uint8_t* __restrict__ s, uint64_t o, uint32_t bit)
{
    const uint32_t w = bit >> 5, sh = bit & 31u;
    const uint32_t lo = b_rd32(s, o + 4ull * w);
    const uint32_t hi = (sh > 16u) ? b_rd32(s, o + 4ull * w + 4ull) : 0u;
    return __funnelshift_r(lo, hi, sh) & 0xffffu;
}
__device__ __forceinline__ uint32_t gap_level_of(uint32_t gap_length)      // gap_calc_level(gap_length), src/bmfunc.h:5418
{
    return gap_length <= 124u ? 0u : gap_length <= 252u ? 1u : gap_length <= 508u ? 2u : 3u;
}

// staging must be readable for 8 bytes past the last payload byte (the allocation carries the slack)
__global__ void __launch_bounds__(kBlobThreads) blob_decode_kernel(const uint8_t* __restrict__ stg, const BlobRec* __restrict__ recs, uint32_t n_recs,
                                                                    uint32_t* __restrict__ bit_pool, uint16_t* __restrict__ gap_pool)
{
    __shared__ uint32_t s_blk[kBlockWords];                 // bit assembly / position list (u16 pairs)
    __shared__ uint32_t s_run[3 * 1024 + 8];                // 0-runs table: (first word, words, payload byte offset)
    __shared__ uint32_t s_scan[kBlobThreads / 32 + 1];
    __shared__ uint32_t s_n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t ri = blockIdx.x; ri < n_recs; ri += gridDim.x) {
        const BlobRec r = recs[ri];
        __syncthreads();
        if (r.type <= DB_ARRBIT_INV) {
            uint32_t* dst = bit_pool + r.dst * (size_t)kBlockWords;
            switch (r.type) {
            case DB_BIT:
                for (uint32_t i = tid; i < kBlockWords; i += kBlobThreads) dst[i] = b_rd32(stg, r.src + 4ull * i);
                break;
            case DB_BIT_INTERVAL: {
                const uint32_t head = b_rd16(stg, r.src), tail = b_rd16(stg, r.src + 2);
                for (uint32_t i = tid; i < kBlockWords; i += kBlobThreads)
                    dst[i] = (i >= head && i <= tail) ? b_rd32(stg, r.src + 4ull + 4ull * (i - head)) : 0u;
                break; }
            case DB_BIT_DIGEST0: {
                const uint64_t d0 = (uint64_t)b_rd32(stg, r.src) | ((uint64_t)b_rd32(stg, r.src + 4) << 32);
                for (uint32_t i = tid; i < kBlockWords; i += kBlobThreads) {
                    const uint32_t wave = i >> 5;
                    uint32_t v = 0;
                    if ((d0 >> wave) & 1ull) {
                        const uint32_t rank = __popcll(d0 & ((1ull << wave) - 1ull));
                        v = b_rd32(stg, r.src + 8ull + 4ull * (rank * 32u + (i & 31u)));
                    }
                    dst[i] = v;
                }
                break; }
            case DB_BIT_0RUNS: {
                if (tid == 0) {                              // run headers are sequential: (type, len16 [, len words]) ...
                    uint64_t o = r.src; uint32_t run_type = b_rd8(stg, o++), j = 0, n = 0;
                    while (j < kBlockWords && n < 1024u) {
                        const uint32_t len = b_rd16(stg, o); o += 2;
                        if (run_type) { s_run[3 * n] = j; s_run[3 * n + 1] = len; s_run[3 * n + 2] = (uint32_t)(o - r.src); ++n; o += 4ull * len; }
                        j += len; run_type ^= 1u;
                    }
                    s_n = n;
                }
                for (uint32_t i = tid; i < kBlockWords; i += kBlobThreads) s_blk[i] = 0u;
                __syncthreads();
                const uint32_t n = s_n;
                for (uint32_t q = warp; q < n; q += kBlobThreads / 32) {
                    const uint32_t j0 = s_run[3 * q], len = s_run[3 * q + 1]; const uint64_t o = r.src + s_run[3 * q + 2];
                    for (uint32_t w = lane; w < len && j0 + w < kBlockWords; w += 32) s_blk[j0 + w] = b_rd32(stg, o + 4ull * w);
                }
                __syncthreads();
                for (uint32_t i = tid; i < kBlockWords; i += kBlobThreads) dst[i] = s_blk[i];
                break; }
            default: {                                       // DB_ARRBIT / DB_ARRBIT_INV
                const bool inv = (r.type == DB_ARRBIT_INV);
                for (uint32_t i = tid; i < kBlockWords; i += kBlobThreads) s_blk[i] = inv ? 0xffffffffu : 0u;
                __syncthreads();
                const uint32_t n = b_rd16(stg, r.src);
                for (uint32_t k = tid; k < n; k += kBlobThreads) {
                    const uint32_t b = b_rd16(stg, r.src + 2ull + 2ull * k);
                    if (inv) atomicAnd(&s_blk[b >> 5], ~(1u << (b & 31u))); else atomicOr(&s_blk[b >> 5], 1u << (b & 31u));
                }
                __syncthreads();
                for (uint32_t i = tid; i < kBlockWords; i += kBlobThreads) dst[i] = s_blk[i];
                break; }
            }
            continue;
        }
        // ---- GAP kinds: out = header + run ends in the arena slot (lead pad 0xFFFF iff the first run is 0; the pool is pre-zeroed)
        const uint32_t pad = r.aux2 & 1u;
        uint16_t* unit = gap_pool + r.dst * (size_t)kGapUnit;
        uint16_t* out = unit + pad;
        if (tid == 0 && pad) unit[0] = 0xffffu;
        if (r.type == DB_GAP16 || r.type == DB_GAP_V3) {
            uint32_t len, first;
            if (r.type == DB_GAP16) { const uint32_t hdr = b_rd16(stg, r.src); len = hdr >> 3; first = hdr & 1u; }
            else { len = r.aux >> 8; first = (r.aux2 >> 1) & 1u; }
            const uint32_t bit0 = r.aux & 0xffu;             // DB_GAP_V3: bit offset of run end #1 inside the word stream
            for (uint32_t k = 1 + tid; k < len; k += kBlobThreads)
                out[k] = (uint16_t)(r.type == DB_GAP16 ? b_rd16(stg, r.src + 2ull * k) : b_bits16(stg, r.src, bit0 + 16u * (k - 1u)));
            if (tid == 0) { out[len] = 65535u; out[0] = (uint16_t)(first | (gap_level_of(len + 1u) << 1) | (len << 3)); }
            continue;
        }
        // DB_ARRGAP / DB_ARRGAP_INV: ascending positions a[0..n) -> run ends (gap_set_array); inverted = same ends, first flipped
        {
            const uint32_t n = r.aux;
            const uint64_t a0 = r.src;                       // u16 positions, byte aligned
            uint16_t* a = reinterpret_cast<uint16_t*>(s_blk);
            for (uint32_t k = tid; k < n; k += kBlobThreads) a[k] = (uint16_t)b_rd16(stg, a0 + 2ull * k);
            __syncthreads();
            // element k opens a 1-run if a[k-1]+1 != a[k], closes one if a[k]+1 != a[k+1]; emitted ends: (s-1 if s > 0), (e if e < 65535)
            constexpr uint32_t kPer = 8;                     // consecutive elements per thread (n <= 2048)
            uint32_t cnt = 0;
            for (uint32_t k = tid * kPer; k < min(n, (tid + 1) * kPer); ++k) {
                const uint32_t s = a[k];
                const bool st = (k == 0) || (uint32_t)a[k - 1] + 1u != s, en = (k + 1 == n) || s + 1u != (uint32_t)a[k + 1];
                cnt += (st && s > 0u) + (en && s < 65535u);
            }
            uint32_t inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
            if (lane == 31) s_scan[warp] = inc;
            __syncthreads();
            uint32_t woff = 0, total = 0;
            for (int w = 0; w < kBlobThreads / 32; ++w) { const uint32_t t = s_scan[w]; if (w < warp) woff += t; total += t; }
            uint32_t pos = 1u + woff + inc - cnt;
            if (total + 2u > (r.aux2 >> 8)) continue;        // never write past the slot the walker measured (uniform; the walkers reject such lists)
            for (uint32_t k = tid * kPer; k < min(n, (tid + 1) * kPer); ++k) {
                const uint32_t s = a[k];
                const bool st = (k == 0) || (uint32_t)a[k - 1] + 1u != s, en = (k + 1 == n) || s + 1u != (uint32_t)a[k + 1];
                if (st && s > 0u) out[pos++] = (uint16_t)(s - 1u);
                if (en && s < 65535u) out[pos++] = (uint16_t)s;
            }
            if (tid == 0) {
                const uint32_t len = total + 1u;
                const uint32_t first = ((n && a[0] == 0) ? 1u : 0u) ^ (r.type == DB_ARRGAP_INV ? 1u : 0u);
                out[len] = 65535u;
                out[0] = (uint16_t)(first | (gap_level_of(len + 1u) << 1) | (len << 3));
            }
        }
    }
}

}  // namespace bmb200
