// aux_kernels.cuh -- result post-step (bit -> GAP), rs_index build + block-scan, batched rank/select,
// and the synthetic-set generator.  sm_100a only.
#pragma once
#include "common.cuh"

namespace bmb200 {

// ---------------------------------------------------------------------------------------------
// bit-block (in shared memory) -> GAP block.  Restates bit_block_to_gap (src/bmfunc.h:5540-5617) as
// "find every position p with bit[p] != bit[p+1], prefix-scan the per-thread counts, write the
// run ends in order".  256 threads, thread t owns words [8t, 8t+8).  Returns len (number of runs);
// out[0] = header with the capacity level set like allocate_gap_block (src/bmblocks.h:1902-1921).
// ---------------------------------------------------------------------------------------------
constexpr int kPostThreads = 256;

__device__ __forceinline__ uint32_t gap_level_for(uint32_t len)
{   // gap_calc_level src/bmfunc.h:5418 with the default table {128,256,512,1280}
    return len <= 124u ? 0u : len <= 252u ? 1u : len <= 508u ? 2u : 3u;
}

// s_blk: 2048 words in smem (+1 readable pad word not required), s_scan: 8 words, out: >= len+1 u16
__device__ uint32_t block_to_gap_256(const uint32_t* s_blk, uint32_t* s_scan, uint16_t* out, uint32_t max_len)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t x[8]; uint32_t cnt = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t wi = 8u * tid + i;
        const uint32_t w = s_blk[wi];
        const uint32_t nxt = (wi + 1 < kBlockWords) ? (s_blk[wi + 1] & 1u) : (w >> 31);  // last bit: no successor
        x[i] = w ^ ((w >> 1) | (nxt << 31));
        cnt += __popc(x[i]);
    }
    // exclusive scan of cnt over 256 threads
    uint32_t inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
    if (lane == 31) s_scan[warp] = inc;
    __syncthreads();
    uint32_t woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kPostThreads / 32; ++w) { const uint32_t v = s_scan[w]; if (w < warp) woff += v; total += v; }
    uint32_t off = woff + inc - cnt;          // run ends before this thread
    const uint32_t len = total + 1u;          // + the final run end 65535
    if (len <= max_len) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint32_t m = x[i];
            const uint32_t base = (8u * tid + i) * 32u;
            while (m) { const uint32_t b = __ffs(m) - 1u; m &= m - 1u; out[1u + off++] = (uint16_t)(base + b); }
        }
        if (tid == 0) {
            out[len] = 65535u;
            out[0] = (uint16_t)((s_blk[0] & 1u) | (gap_level_for(len) << 1) | (len << 3));
        }
    }
    __syncthreads();
    return len;
}

// result post-step: columns whose kind is GAP get their stored bit-block converted (opt_copy_bit_block's
// bit_to_gap branch, src/bmblocks.h:1389-1403).  gaps: [n_cols][1280] u16.
__global__ void __launch_bounds__(kPostThreads) result_to_gap_kernel(const uint32_t* __restrict__ blocks,
                                                                     const uint8_t* __restrict__ kind,
                                                                     uint16_t* __restrict__ gaps, uint32_t n_cols)
{
    __shared__ __align__(16) uint32_t s_blk[kBlockWords];
    __shared__ uint32_t s_scan[8];
    for (uint32_t col = blockIdx.x; col < n_cols; col += gridDim.x) {
        if (kind[col] != BMB200_BLK_GAP) continue;
        const uint4* src = reinterpret_cast<const uint4*>(blocks + (size_t)col * kBlockWords);
        uint4* dst = reinterpret_cast<uint4*>(s_blk);
        dst[threadIdx.x] = src[threadIdx.x];
        dst[threadIdx.x + kPostThreads] = src[threadIdx.x + kPostThreads];
        __syncthreads();
        block_to_gap_256(s_blk, s_scan, gaps + (size_t)col * kGapMax, kGapMax - 1u);
    }
}

// compaction of an optimized result into per-vector flat form: off[c] computed on the host from kinds
// and gap lengths; this kernel only moves data.  bits_out: compacted BIT blocks, gaps_out: compacted GAPs.
__global__ void __launch_bounds__(256) result_compact_kernel(const uint32_t* __restrict__ blocks,
                                                             const uint16_t* __restrict__ gaps,
                                                             const uint8_t* __restrict__ kind,
                                                             const uint64_t* __restrict__ off,
                                                             uint32_t* __restrict__ bits_out,
                                                             uint16_t* __restrict__ gaps_out, uint32_t n_cols)
{
    for (uint32_t col = blockIdx.x; col < n_cols; col += gridDim.x) {
        const uint32_t kd = kind[col];
        if (kd == BMB200_BLK_BIT) {
            const uint4* src = reinterpret_cast<const uint4*>(blocks + (size_t)col * kBlockWords);
            uint4* dst = reinterpret_cast<uint4*>(bits_out + off[col] * (size_t)kBlockWords);
            dst[threadIdx.x] = src[threadIdx.x];
            dst[threadIdx.x + 256] = src[threadIdx.x + 256];
        } else if (kd == BMB200_BLK_GAP) {
            const uint16_t* src = gaps + (size_t)col * kGapMax;
            uint16_t* dst = gaps_out + off[col];
            const uint32_t n = (uint32_t)(src[0] >> 3) + 1u;
            const uint32_t npad = (n + kGapUnit - 1u) / kGapUnit * kGapUnit;
            for (uint32_t i = threadIdx.x; i < npad; i += 256) dst[i] = (i < n) ? src[i] : (uint16_t)0;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// rs_index build (bvector::build_rs_index src/bm.h:2531-2660): one warp per block computes
//   bcount, first = bits in [0,21824], second = bits in (21824,43648], aux0, aux1
// bit-block : aux0/aux1 = bits in [0,32736] / [0,54560]              (src/bm.h:2626-2641)
// GAP block : aux = (gap_bfind(border+1) << 1) | is_set               (src/bm.h:2601-2625)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cnt_le(uint32_t w, uint32_t wi, uint32_t B)
{   // bits of word wi (bit positions 32*wi .. 32*wi+31) at positions <= B
    const uint32_t bw = B >> 5;
    if (wi < bw) return __popc(w);
    if (wi > bw) return 0u;
    return __popc(w & (0xffffffffu >> (31u - (B & 31u))));
}

__global__ void __launch_bounds__(256) rs_block_kernel(const SetView set, uint32_t vec,
                                                       uint32_t* __restrict__ bcount,
                                                       uint64_t* __restrict__ sub_count)
{
    const int lane = threadIdx.x & 31;
    const uint32_t warps_total = gridDim.x * (blockDim.x >> 5);
    for (uint32_t nb = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); nb < set.n_blocks; nb += warps_total) {
        const uint32_t d = set.desc[(size_t)nb * set.n_vec + vec];
        const uint32_t kd = d & 3u, rel = d >> 2;
        uint32_t tot = 0, le0 = 0, le1 = 0, a0 = 0, a1 = 0;
        if (kd == BMB200_BLK_GAP) {
            const uint16_t* g = set.gap_pool + (set.gap_base[nb] + (rel & BMB200_DESC_REL_MASK)) * (size_t)kGapUnit + (rel >> 29);
            const uint32_t hdr = g[0], len = hdr >> 3, first = hdr & 1u;
            uint32_t lt0 = 0, lt1 = 0;   // run ends < border+1  (for gap_bfind)
            for (uint32_t k = 1 + lane; k <= len; k += 32) {
                const uint32_t e = g[k];
                const uint32_t s = (k == 1) ? 0u : (uint32_t)g[k - 1] + 1u;
                lt0 += (e < kRs3B0 + 1u); lt1 += (e < kRs3B1 + 1u);
                if (first ^ ((k - 1u) & 1u)) {
                    tot += e - s + 1u;
                    if (s <= kRs3B0) le0 += min(e, kRs3B0) - s + 1u;
                    if (s <= kRs3B1) le1 += min(e, kRs3B1) - s + 1u;
                }
            }
            tot = warp_sum(tot); le0 = warp_sum(le0); le1 = warp_sum(le1);
            lt0 = warp_sum(lt0); lt1 = warp_sum(lt1);
            const uint32_t i0 = lt0 + 1u, i1 = lt1 + 1u;          // gap_bfind src/bmfunc.h:1844
            a0 = (i0 << 1) | (first ^ ((i0 - 1u) & 1u));
            a1 = (i1 << 1) | (first ^ ((i1 - 1u) & 1u));
        } else if (kd != BMB200_BLK_NULL) {
            if (kd == BMB200_BLK_BIT) {
                // 16 coalesced 512-byte warp loads per block, all issued before the popcounts (8 KB in flight per warp)
                const uint4* b4 = reinterpret_cast<const uint4*>(set.bit_pool + (set.bit_base[nb] + rel) * (size_t)kBlockWords);
                uint4 v[16];
#pragma unroll
                for (int it = 0; it < 16; ++it) v[it] = ld_stream_v4(b4 + it * 32 + lane);
#pragma unroll
                for (int it = 0; it < 16; ++it) {
                    const uint32_t wi = (uint32_t)(it * 32 + lane) * 4u;
                    const uint32_t w4[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t w = w4[q];
                        tot += __popc(w);
                        le0 += cnt_le(w, wi + q, kRs3B0);   le1 += cnt_le(w, wi + q, kRs3B1);
                        a0  += cnt_le(w, wi + q, kRs3B0_1); a1  += cnt_le(w, wi + q, kRs3B1_1);
                    }
                }
            } else {   // FULL block: an all-ones block, like BLOCK_ADDR_SAN in src/bm.h:2628
                for (uint32_t wi = lane; wi < kBlockWords; wi += 32) {
                    tot += 32u;
                    le0 += cnt_le(0xffffffffu, wi, kRs3B0);   le1 += cnt_le(0xffffffffu, wi, kRs3B1);
                    a0  += cnt_le(0xffffffffu, wi, kRs3B0_1); a1  += cnt_le(0xffffffffu, wi, kRs3B1_1);
                }
            }
            tot = warp_sum(tot); le0 = warp_sum(le0); le1 = warp_sum(le1); a0 = warp_sum(a0); a1 = warp_sum(a1);
        }
        if (lane == 0) {
            const uint32_t firstc = le0, secondc = le1 - le0;
            bcount[nb] = tot;
            sub_count[nb] = (uint64_t)(firstc | (secondc << 16)) | ((uint64_t)a0 << 32) | ((uint64_t)a1 << 48);
        }
    }
}

// block-scan, level 1: one CTA per 256-block superblock -> running counts inside the superblock
// (rs_index::register_super_block src/bmrs.h:688-715) and the superblock total
__global__ void __launch_bounds__(256) rs_scan_rows_kernel(const uint32_t* __restrict__ bcount, uint32_t n_blocks,
                                                           uint32_t* __restrict__ row_cum, uint64_t* __restrict__ sb_tot)
{
    __shared__ uint32_t s_w[8];
    const uint32_t sb = blockIdx.x, nb = sb * 256u + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t c = (nb < n_blocks) ? bcount[nb] : 0u;
    uint32_t inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    uint32_t woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { const uint32_t v = s_w[w]; if (w < warp) woff += v; total += v; }
    if (nb < n_blocks) row_cum[nb] = woff + inc;
    if (threadIdx.x == 0) sb_tot[sb] = total;
}

// block-scan, level 2: running totals over superblocks (sblock_count_, src/bmrs.h:586-620); one CTA
__global__ void __launch_bounds__(1024) rs_scan_sb_kernel(const uint64_t* __restrict__ sb_tot, uint32_t nsb,
                                                          uint64_t* __restrict__ sb_cum)
{
    __shared__ uint64_t s_w[32];
    __shared__ uint64_t s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { s_carry = 0; sb_cum[0] = 0; }
    __syncthreads();
    for (uint32_t base = 0; base < nsb; base += 1024u) {
        const uint32_t i = base + threadIdx.x;
        const uint64_t c = (i < nsb) ? sb_tot[i] : 0ull;
        uint64_t inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint64_t y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        uint64_t woff = 0, total = 0;
        for (int w = 0; w < 32; ++w) { const uint64_t v = s_w[w]; if (w < warp) woff += v; total += v; }
        const uint64_t carry = s_carry;
        if (i < nsb) sb_cum[i + 1] = carry + woff + inc;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + total;
        __syncthreads();
    }
}

struct RsView {
    SetView set;
    uint32_t vec;
    uint32_t nsb;
    const uint32_t* bcount;     // [n_blocks]
    const uint64_t* sub_count;  // [n_blocks]
    const uint32_t* row_cum;    // [n_blocks] inclusive running count inside the superblock
    const uint64_t* sb_cum;     // [nsb+1]
};

// bits set in words of a bit-block at positions [from, to] (inclusive, from <= to)
__device__ __forceinline__ uint32_t bit_count_range(const uint32_t* __restrict__ b, uint32_t from, uint32_t to)
{
    const uint32_t wf = from >> 5, wt = to >> 5;
    if (wf == wt) return __popc(b[wf] & bit_range_mask(from & 31u, to & 31u));
    uint32_t c = __popc(b[wf] & (0xffffffffu << (from & 31u)));
    for (uint32_t w = wf + 1; w < wt; ++w) c += __popc(b[w]);
    return c + __popc(b[wt] & (0xffffffffu >> (31u - (to & 31u))));
}
// 1-bits of a GAP block in [from, to], scanning from run index k (run k must contain `from`)
__device__ __forceinline__ uint32_t gap_count_from(const uint16_t* __restrict__ g, uint32_t k, uint32_t from, uint32_t to)
{
    const uint32_t first = g[0] & 1u;
    uint32_t c = 0, s = from;
    for (;; ++k) {
        const uint32_t e = g[k];
        const uint32_t v = first ^ ((k - 1u) & 1u);
        if (e >= to) { if (v) c += to - s + 1u; break; }
        if (v) c += e - s + 1u;
        s = e + 1u;
    }
    return c;
}

// inclusive rank, bvector::count_to src/bm.h:3120-3167; block part follows block_count_to's
// nearest-anchor idea (src/bm.h:2686-2869) with the five anchors stored in sub_count.
__global__ void __launch_bounds__(256) rs_rank_kernel(const RsView rs, const uint64_t* __restrict__ pos, uint64_t n,
                                                      uint64_t* __restrict__ out)
{
    for (uint64_t q = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; q < n; q += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t p = pos[q];
        const uint64_t nb64 = p >> 16;
        if (nb64 >= rs.set.n_blocks) { out[q] = rs.sb_cum[rs.nsb]; continue; }
        const uint32_t nb = (uint32_t)nb64, in = (uint32_t)(p & 0xffffu);
        uint64_t r = rs.sb_cum[nb >> 8] + ((nb & 255u) ? rs.row_cum[nb - 1] : 0u);
        const uint32_t d = rs.set.desc[(size_t)nb * rs.set.n_vec + rs.vec];
        const uint32_t kd = d & 3u, rel = d >> 2;
        if (kd == BMB200_BLK_FULL) r += in + 1u;
        else if (kd != BMB200_BLK_NULL) {
            const uint64_t sub = rs.sub_count[nb];
            const uint32_t first = (uint32_t)(sub & 0xffffu), second = (uint32_t)((sub >> 16) & 0xffffu);
            const uint32_t a0 = (uint32_t)((sub >> 32) & 0xffffu), a1 = (uint32_t)(sub >> 48);
            if (kd == BMB200_BLK_BIT) {
                const uint32_t* b = rs.set.bit_pool + (rs.set.bit_base[nb] + rel) * (size_t)kBlockWords;
                const uint32_t bc = rs.bcount[nb];
                // anchors: (position, bits in [0,position])
                const int32_t  ap[6] = { -1, (int32_t)kRs3B0, (int32_t)kRs3B0_1, (int32_t)kRs3B1, (int32_t)kRs3B1_1, 65535 };
                const uint32_t ac[6] = { 0u, first, a0, first + second, a1, bc };
                int best = 0; uint32_t bd = in + 1u;
#pragma unroll
                for (int a = 1; a < 6; ++a) {
                    const uint32_t dist = (uint32_t)abs((int32_t)in - ap[a]);
                    if (dist < bd) { bd = dist; best = a; }
                }
                uint32_t c = ac[best];
                if ((int32_t)in > ap[best])      c += bit_count_range(b, (uint32_t)(ap[best] + 1), in);
                else if ((int32_t)in < ap[best]) c -= bit_count_range(b, in + 1u, (uint32_t)ap[best]);
                r += c;
            } else {
                const uint16_t* g = rs.set.gap_pool + (rs.set.gap_base[nb] + (rel & BMB200_DESC_REL_MASK)) * (size_t)kGapUnit + (rel >> 29);
                uint32_t c;
                if (in <= kRs3B0)      c = gap_count_from(g, 1u, 0u, in);
                else if (in <= kRs3B1) c = first + gap_count_from(g, a0 >> 1, kRs3B0 + 1u, in);
                else                   c = first + second + gap_count_from(g, a1 >> 1, kRs3B1 + 1u, in);
                r += c;
            }
        }
        out[q] = r;
    }
}

// 1-based select, bvector::select src/bm.h:5350-5385; rs_index::find src/bmrs.h:398-460;
// bit_find_rank src/bmfunc.h:9673-9768 (word select via __fns instead of PDEP, src/bmbmi2.h:56-71)
__global__ void __launch_bounds__(256) rs_select_kernel(const RsView rs, const uint64_t* __restrict__ rank, uint64_t n,
                                                        uint64_t* __restrict__ pos, uint8_t* __restrict__ found)
{
    const uint64_t total = rs.sb_cum[rs.nsb];
    for (uint64_t q = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; q < n; q += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t r = rank[q];
        if (r == 0 || r > total) { found[q] = 0; pos[q] = 0; continue; }
        // smallest superblock i with sb_cum[i+1] >= r
        uint32_t lo = 0, hi = rs.nsb - 1u;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (rs.sb_cum[mid + 1] < r) lo = mid + 1; else hi = mid; }
        const uint32_t sb = lo;
        uint32_t rr = (uint32_t)(r - rs.sb_cum[sb]);
        // smallest block j in the superblock with row_cum >= rr
        const uint32_t b0 = sb * 256u;
        lo = 0; hi = min(255u, rs.set.n_blocks - 1u - b0);
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (rs.row_cum[b0 + mid] < rr) lo = mid + 1; else hi = mid; }
        const uint32_t nb = b0 + lo;
        if (lo) rr -= rs.row_cum[nb - 1];
        const uint32_t d = rs.set.desc[(size_t)nb * rs.set.n_vec + rs.vec];
        const uint32_t kd = d & 3u, rel = d >> 2;
        uint32_t bit = 0;
        if (kd == BMB200_BLK_FULL) bit = rr - 1u;
        else {
            const uint64_t sub = rs.sub_count[nb];
            const uint32_t first = (uint32_t)(sub & 0xffffu), second = (uint32_t)((sub >> 16) & 0xffffu);
            const uint32_t a0 = (uint32_t)((sub >> 32) & 0xffffu), a1 = (uint32_t)(sub >> 48);
            if (kd == BMB200_BLK_BIT) {
                const uint32_t* b = rs.set.bit_pool + (rs.set.bit_base[nb] + rel) * (size_t)kBlockWords;
                // last anchor with count < rr
                uint32_t start = 0, c = 0;
                if (first < rr)          { start = kRs3B0 + 1u;   c = first; }
                if (a0 < rr)             { start = kRs3B0_1 + 1u; c = a0; }
                if (first + second < rr) { start = kRs3B1 + 1u;   c = first + second; }
                if (a1 < rr)             { start = kRs3B1_1 + 1u; c = a1; }
                uint32_t need = rr - c;
                uint32_t wi = start >> 5;
                uint32_t w = b[wi] & (0xffffffffu << (start & 31u));
                for (;;) {
                    const uint32_t pc = __popc(w);
                    if (need <= pc) break;
                    need -= pc; w = b[++wi];
                }
                bit = wi * 32u + __fns(w, 0, (int)need);
            } else {
                const uint16_t* g = rs.set.gap_pool + (rs.set.gap_base[nb] + (rel & BMB200_DESC_REL_MASK)) * (size_t)kGapUnit + (rel >> 29);
                const uint32_t firstv = g[0] & 1u;
                uint32_t k = 1u, s = 0u, need = rr;
                if (first + second < rr) { k = a1 >> 1; s = kRs3B1 + 1u; need = rr - first - second; }
                else if (first < rr)     { k = a0 >> 1; s = kRs3B0 + 1u; need = rr - first; }
                for (;; ++k) {
                    const uint32_t e = g[k];
                    if (firstv ^ ((k - 1u) & 1u)) {
                        const uint32_t rl = e - s + 1u;
                        if (need <= rl) { bit = s + need - 1u; break; }
                        need -= rl;
                    }
                    s = e + 1u;
                }
            }
        }
        found[q] = 1; pos[q] = ((uint64_t)nb << 16) | bit;
    }
}

// ---------------------------------------------------------------------------------------------
// synthetic set generator.  Bit (v, p) of vector v is set iff u16(hash(seed_v, p)) < thr_v with
// thr_v = round(density_v * 65536): a counter-based generator, so the classify pass and the write pass
// regenerate identical blocks.  One 64-bit hash yields four bits.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t mix64(uint64_t x)
{   // splitmix64 finalizer
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27; x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}
__device__ __forceinline__ uint32_t synth_word(uint64_t seed, uint32_t nb, uint32_t wi, uint32_t thr)
{
    if (thr == 0) return 0u;
    if (thr >= 65536u) return 0xffffffffu;
    uint32_t w = 0;
    const uint64_t ctr = (((uint64_t)nb << 11) | wi) << 3;      // 8 hashes per word
#pragma unroll
    for (int h = 0; h < 8; ++h) {
        const uint64_t r = mix64(seed + 0x9e3779b97f4a7c15ull * (ctr + h + 1));
        w |= (uint32_t)((uint32_t)(r & 0xffffu) < thr) << (4 * h);
        w |= (uint32_t)((uint32_t)((r >> 16) & 0xffffu) < thr) << (4 * h + 1);
        w |= (uint32_t)((uint32_t)((r >> 32) & 0xffffu) < thr) << (4 * h + 2);
        w |= (uint32_t)((uint32_t)(r >> 48) < thr) << (4 * h + 3);
    }
    return w;
}

// pass 1: CTA per (nb, v): popcount + run count -> kind and stored size
__global__ void __launch_bounds__(kPostThreads) synth_classify_kernel(uint32_t n_vec, uint32_t n_blocks,
                                                                      const uint64_t* __restrict__ seed,
                                                                      const uint32_t* __restrict__ thr, int optimize,
                                                                      uint8_t* __restrict__ kind8,
                                                                      uint16_t* __restrict__ glen)
{
    __shared__ uint32_t s_pc[8], s_tr[8], s_last[kPostThreads];
    const uint64_t item = blockIdx.x;
    const uint32_t nb = (uint32_t)(item / n_vec), v = (uint32_t)(item % n_vec);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t sd = seed[v]; const uint32_t th = thr[v];
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = synth_word(sd, nb, 8u * tid + i, th);
    s_last[tid] = w[7] >> 31;
    __syncthreads();
    uint32_t prev = tid ? s_last[tid - 1] : (w[0] & 1u);
    uint32_t pc = 0, tr = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { pc += __popc(w[i]); tr += __popc(w[i] ^ ((w[i] << 1) | prev)); prev = w[i] >> 31; }
    pc = warp_sum(pc); tr = warp_sum(tr);
    if (lane == 0) { s_pc[warp] = pc; s_tr[warp] = tr; }
    __syncthreads();
    if (tid == 0) {
        uint32_t tpc = 0, ttr = 0;
        for (int i = 0; i < 8; ++i) { tpc += s_pc[i]; ttr += s_tr[i]; }
        const uint32_t runs = ttr + 1u;
        uint32_t kd;
        if (tpc == 0) kd = BMB200_BLK_NULL;
        else if (!optimize) kd = BMB200_BLK_BIT;
        else if (tpc == 65536u) kd = BMB200_BLK_FULL;
        else if (runs < BMB200_GAP_THRESHOLD) kd = BMB200_BLK_GAP;   // optimize_bit_block src/bmblocks.h:1414-1437
        else kd = BMB200_BLK_BIT;
        kind8[item] = (uint8_t)kd;
        glen[item] = (uint16_t)((runs & 0x7fffu) | ((w[0] & 1u) << 15));   // bit 15 = first bit (tid 0 owns word 0)
    }
}

// per column: exclusive scans over the vectors -> descriptors + column totals
__global__ void __launch_bounds__(256) synth_layout_kernel(uint32_t n_vec, uint32_t flat_form, const uint8_t* __restrict__ kind8,
                                                           const uint16_t* __restrict__ glen,
                                                           uint32_t* __restrict__ desc,
                                                           uint64_t* __restrict__ col_bits, uint64_t* __restrict__ col_gaps)
{
    __shared__ uint32_t s_wb[8], s_wg[8];
    __shared__ uint32_t s_cb, s_cg;
    const uint32_t nb = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_cb = 0; s_cg = 0; }
    __syncthreads();
    for (uint32_t base = 0; base < n_vec; base += 256u) {
        const uint32_t v = base + tid;
        uint32_t kd = 0, nbit = 0, ngap = 0, pad = 0;
        if (v < n_vec) {
            kd = kind8[(size_t)nb * n_vec + v];
            nbit = (kd == BMB200_BLK_BIT);
            if (kd == BMB200_BLK_GAP) {
                const uint32_t gl = glen[(size_t)nb * n_vec + v];
                pad = ((gl >> 15) || !flat_form) ? 0u : 1u;     // BMB200_DESC_GAP_FLAT: lead pad iff the first run is 0
                ngap = ((gl & 0x7fffu) + 1u + pad + kGapUnit - 1u) / kGapUnit;
            }
        }
        uint32_t ib = nbit, ig = ngap;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t yb = __shfl_up_sync(0xffffffffu, ib, o), yg = __shfl_up_sync(0xffffffffu, ig, o);
            if (lane >= o) { ib += yb; ig += yg; }
        }
        if (lane == 31) { s_wb[warp] = ib; s_wg[warp] = ig; }
        __syncthreads();
        uint32_t ob = s_cb, og = s_cg, tb = 0, tg = 0;
        for (int w = 0; w < 8; ++w) { if (w < warp) { ob += s_wb[w]; og += s_wg[w]; } tb += s_wb[w]; tg += s_wg[w]; }
        if (v < n_vec) {
            const uint32_t rel = (kd == BMB200_BLK_BIT) ? ob + ib - nbit : (kd == BMB200_BLK_GAP) ? og + ig - ngap : 0u;
            desc[(size_t)nb * n_vec + v] = kd | (rel << 2) | (pad << 31) | ((kd == BMB200_BLK_GAP && flat_form) ? BMB200_DESC_GAP_FLAT : 0u);
        }
        __syncthreads();
        if (tid == 0) { s_cb += tb; s_cg += tg; }
        __syncthreads();
    }
    if (tid == 0) { col_bits[nb] = s_cb; col_gaps[nb] = s_cg; }
}

// exclusive scan of per-column totals into bit_base / gap_base ([n+1]); one CTA
__global__ void __launch_bounds__(1024) scan_u64_kernel(const uint64_t* __restrict__ in, uint32_t n, uint64_t* __restrict__ out)
{
    __shared__ uint64_t s_w[32];
    __shared__ uint64_t s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { s_carry = 0; out[0] = 0; }
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024u) {
        const uint32_t i = base + threadIdx.x;
        const uint64_t c = (i < n) ? in[i] : 0ull;
        uint64_t inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint64_t y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        uint64_t woff = 0, total = 0;
        for (int w = 0; w < 32; ++w) { const uint64_t v = s_w[w]; if (w < warp) woff += v; total += v; }
        const uint64_t carry = s_carry;
        if (i < n) out[i + 1] = carry + woff + inc;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + total;
        __syncthreads();
    }
}

// pass 2: regenerate and store every BIT / GAP block in its arena slot
__global__ void __launch_bounds__(kPostThreads) synth_write_kernel(uint32_t n_vec, uint32_t n_blocks,
                                                                   const uint64_t* __restrict__ seed,
                                                                   const uint32_t* __restrict__ thr,
                                                                   const uint32_t* __restrict__ desc,
                                                                   const uint64_t* __restrict__ bit_base,
                                                                   const uint64_t* __restrict__ gap_base,
                                                                   uint32_t* __restrict__ bit_pool,
                                                                   uint16_t* __restrict__ gap_pool)
{
    __shared__ __align__(16) uint32_t s_blk[kBlockWords];
    __shared__ uint32_t s_scan[8];
    const uint64_t item = blockIdx.x;
    const uint32_t nb = (uint32_t)(item / n_vec), v = (uint32_t)(item % n_vec);
    const uint32_t d = desc[item], kd = d & 3u, rel = (d >> 2) & BMB200_DESC_REL_MASK, pad = d >> 31;
    if (kd != BMB200_BLK_BIT && kd != BMB200_BLK_GAP) return;
    const int tid = threadIdx.x;
    const uint64_t sd = seed[v]; const uint32_t th = thr[v];
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = synth_word(sd, nb, 8u * tid + i, th);
    if (kd == BMB200_BLK_BIT) {
        uint4* dst = reinterpret_cast<uint4*>(bit_pool + (bit_base[nb] + rel) * (size_t)kBlockWords) + 2 * tid;
        dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
        dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) s_blk[8 * tid + i] = w[i];
        __syncthreads();
        uint16_t* unit = gap_pool + (gap_base[nb] + rel) * (size_t)kGapUnit;
        uint16_t* out = unit + pad;
        const uint32_t len = block_to_gap_256(s_blk, s_scan, out, kGapMax - 1u);
        // lead pad = 0xFFFF, zeros from buf[len] to the end of the 16-byte unit (BMB200_DESC_GAP_FLAT contract)
        const uint32_t n = len + 1u + pad, npad = (n + kGapUnit - 1u) / kGapUnit * kGapUnit;
        if (tid < (int)(npad - n)) unit[n + tid] = 0;
        if (tid == 0 && pad) unit[0] = 0xffffu;
    }
}

}  // namespace bmb200
