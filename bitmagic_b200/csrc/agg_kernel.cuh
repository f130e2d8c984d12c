// agg_kernel.cuh -- N-way block-column aggregation kernel (OR / AND / AND-SUB / XOR) for sm_100a.
//
// Replaces, for one block column (i,j) per CTA iteration, the reference's
//   sort_input_blocks_or/_and      src/bmaggregator.h:2278-2366   (classification pass, warp ballots)
//   process_bit_blocks_or/_and/_sub src/bmaggregator.h:1924-2205  (bit phase, register accumulator)
//   process_gap_blocks_or/_and/_sub src/bmaggregator.h:1808-1890  (GAP phase, run scatter into smem)
//   bit_block_count / calc_block_digest0 / bit_block_calc_change  src/bmfunc.h:5808,1239,6040 (epilogue)
//   the classification half of opt_copy_bit_block                 src/bmblocks.h:1355-1409
//
// Layout / mapping:
//   * persistent CTAs (grid = SMs x occupancy) pull block columns from an atomic counter, so skewed
//     columns (NULL / GAP / bit mixes) balance themselves;
//   * 512 threads own the 8 KB accumulator in registers: thread t holds words [4t, 4t+4) as one uint4,
//     every source bit-block is consumed with one coalesced 128-bit streaming load per thread,
//     8 blocks in flight per thread (64 KB per CTA);
//   * GAP sources never get expanded on their own: each selected run is applied to an 8 KB kill/union
//     mask K in shared memory with red.shared (one warp per GAP block, one run per lane), and K meets the
//     register accumulator only once, in the epilogue:
//         OR      : R = U | K            (K = union of 1-runs)
//         AND-SUB : R = P & ~U & ~K      (K = 0-runs of AND-group GAPs  U  1-runs of SUB-group GAPs)
//         XOR     : R = X ^ K
//   * epilogue fuses popcount, 64-wave digest, run count and the result-kind decision.
#pragma once
#include "common.cuh"

namespace bmb200 {

constexpr int kAggThreads = 512;
constexpr int kAggWarps   = kAggThreads / 32;
constexpr int kAggChunk   = 1024;   // group members classified per pass

struct AggParams {
    SetView   set;
    const uint32_t* group;     // device: n0 + n1 member vector ids (group0 then group1)
    uint32_t  n0, n1;
    uint32_t  nb_from, n_cols;
    uint32_t  compress;        // classify like opt_copy_bit_block(opt_compress)
    uint32_t  store_blocks;    // 0 = counts only
    uint32_t* blocks;          // [n_cols][2048]
    uint32_t* popcnt;          // [n_cols]
    uint64_t* digest;          // [n_cols]
    uint32_t* nruns;           // [n_cols]
    uint8_t*  kind;            // [n_cols]
    unsigned long long* total; // sum of popcounts
    uint32_t* work_counter;    // zeroed before launch
};

// flag bits collected while classifying
constexpr uint32_t kFlNull0 = 1u;   // a NULL block in group0
constexpr uint32_t kFlFull0 = 2u;   // a FULL block in group0
constexpr uint32_t kFlFull1 = 4u;   // a FULL block in group1 (SUB)

template <int OP>
__device__ __forceinline__ void acc_apply0(uint4& a, const uint4& v)
{
    if (OP == BMB200_OP_OR)       { a.x |= v.x; a.y |= v.y; a.z |= v.z; a.w |= v.w; }
    else if (OP == BMB200_OP_XOR) { a.x ^= v.x; a.y ^= v.y; a.z ^= v.z; a.w ^= v.w; }
    else                          { a.x &= v.x; a.y &= v.y; a.z &= v.z; a.w &= v.w; }
}
__device__ __forceinline__ void acc_or(uint4& a, const uint4& v)
{
    a.x |= v.x; a.y |= v.y; a.z |= v.z; a.w |= v.w;
}

// stream `n` bit-blocks of this column (indices in lst[], relative to the column's bit segment)
template <int OP, bool ROLE1>
__device__ __forceinline__ void bit_phase(const uint4* __restrict__ seg, const uint32_t* lst, uint32_t n, uint4& acc)
{
    const uint4 ident = (!ROLE1 && (OP == BMB200_OP_AND || OP == BMB200_OP_AND_SUB))
                            ? make_uint4(~0u, ~0u, ~0u, ~0u) : make_uint4(0u, 0u, 0u, 0u);
    for (uint32_t i = 0; i < n; i += 8) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (i + u < n) v[u] = ld_stream_v4(seg + (size_t)lst[i + u] * (kBlockWords / 4));
            else           v[u] = ident;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (ROLE1) acc_or(acc, v[u]); else acc_apply0<OP>(acc, v[u]);
        }
    }
}

template <bool XOR>
__device__ __forceinline__ void apply_run(uint32_t* K, uint32_t s, uint32_t e)
{
    const uint32_t ws = s >> 5, we = e >> 5;
    if (ws == we) {
        const uint32_t m = bit_range_mask(s & 31u, e & 31u);
        if (XOR) red_xor_shared(K + ws, m); else red_or_shared(K + ws, m);
    } else {
        const uint32_t m0 = 0xffffffffu << (s & 31u);
        const uint32_t m1 = 0xffffffffu >> (31u - (e & 31u));
        if (XOR) { red_xor_shared(K + ws, m0); red_xor_shared(K + we, m1); }
        else     { red_or_shared(K + ws, m0);  red_or_shared(K + we, m1); }
        for (uint32_t w = ws + 1; w < we; ++w) {
            if (XOR) red_xor_shared(K + w, 0xffffffffu); else red_or_shared(K + w, 0xffffffffu);
        }
    }
}

// One warp scatters the runs of value `want` of one GAP block into K.
// GAP format (src/bmfunc.h:1696-1725): buf[0] = header (bit0 first-run value, len = hdr>>3),
// buf[k] k=1..len inclusive run ends; run k = (buf[k-1], buf[k]], value = first ^ ((k-1)&1).
// Lane j reads the aligned pair word W[j] = (buf[2j], buf[2j+1]); w_first is W[lane] of the first
// 32 words (already loaded by the caller so the next block's load overlaps this block's scatter).
template <bool XOR>
__device__ __forceinline__ void gap_scatter(uint32_t* K, const uint32_t* __restrict__ g32, uint32_t w_first,
                                            uint32_t want, int lane)
{
    const uint32_t hdr = __shfl_sync(0xffffffffu, w_first, 0) & 0xffffu;
    const uint32_t len = hdr >> 3;
    const bool odd = ((hdr & 1u) == want);       // selected runs are k = 1,3,5.. else k = 2,4,6..
    const uint32_t k0 = odd ? 1u : 2u;
    const uint32_t stride = odd ? 32u : 31u;
    uint32_t w = w_first;
    for (uint32_t base = 0; 2u * base + k0 <= len; base += stride) {
        const uint32_t j = base + lane;
        // prefetch the next iteration's pair word before scattering this one
        const uint32_t jn = j + stride;
        uint32_t wn = 0;
        if (2u * (base + stride) + k0 <= len && 2u * jn <= len) wn = ld_nc_u32(g32 + jn);
        const uint32_t lo = w & 0xffffu, hi = w >> 16;
        const uint32_t nlo = __shfl_down_sync(0xffffffffu, lo, 1);
        uint32_t s, e; bool valid;
        if (odd) { s = j ? lo + 1u : 0u; e = hi; valid = (2u * j + 1u <= len); }
        else     { s = hi + 1u; e = nlo; valid = (2u * j + 2u <= len) && (lane < 31); }
        if (valid) apply_run<XOR>(K, s, e);
        w = wn;
    }
}

template <int OP>
__global__ void __launch_bounds__(kAggThreads, 2) agg_kernel(const AggParams p)
{
    __shared__ __align__(16) uint32_t K[kBlockWords];
    __shared__ uint32_t lst_bit0[kAggChunk];
    __shared__ uint32_t lst_bit1[kAggChunk];
    __shared__ uint32_t lst_gap[kAggChunk];      // group0 GAPs from the front, group1 GAPs from the back
    __shared__ uint32_t s_cnt[4];                // nbit0, nbit1, ngap0, ngap1 of the current chunk
    __shared__ uint32_t s_stat[4];               // flags, total nbit0, total ngap0, nfull0
    __shared__ uint32_t s_col, s_gap_next;
    __shared__ uint32_t s_pc[kAggWarps], s_tr[kAggWarps], s_dg[kAggWarps];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t M = p.set.n_vec;
    const uint32_t ntot = p.n0 + ((OP == BMB200_OP_AND_SUB) ? p.n1 : 0u);
    uint4* K4 = reinterpret_cast<uint4*>(K);

    for (;;) {
        if (tid == 0) s_col = atomicAdd(p.work_counter, 1u);
        __syncthreads();
        const uint32_t col = s_col;
        if (col >= p.n_cols) break;
        const uint32_t nb = p.nb_from + col;

        K4[tid] = make_uint4(0u, 0u, 0u, 0u);
        if (tid < 4) s_stat[tid] = 0u;
        uint4 acc0 = (OP == BMB200_OP_AND || OP == BMB200_OP_AND_SUB) ? make_uint4(~0u, ~0u, ~0u, ~0u)
                                                                      : make_uint4(0u, 0u, 0u, 0u);
        uint4 acc1 = make_uint4(0u, 0u, 0u, 0u);   // union of SUB-group bit-blocks

        const uint32_t* drow = p.set.desc + (size_t)nb * M;
        const uint4* bseg = reinterpret_cast<const uint4*>(p.set.bit_pool)
                            + p.set.bit_base[nb] * (size_t)(kBlockWords / 4) + tid;
        const uint16_t* gseg = p.set.gap_pool + p.set.gap_base[nb] * (size_t)kGapUnit;

        for (uint32_t cs = 0; cs < ntot; cs += kAggChunk) {
            if (tid < 4) s_cnt[tid] = 0u;
            if (tid == 0) s_gap_next = 0u;
            __syncthreads();
            // ---- classification (sort_input_blocks_*) ----
            const uint32_t ce = min(cs + (uint32_t)kAggChunk, ntot);
            uint32_t fl = 0, nfull0 = 0;
            for (uint32_t kb = cs; kb < ce; kb += kAggThreads) {   // warp-uniform trip count
                const uint32_t k = kb + tid;
                uint32_t kind = 0xffu, rel = 0; bool g1 = false;
                if (k < ce) {
                    const uint32_t d = drow[p.group[k]];
                    kind = d & 3u; rel = d >> 2; g1 = (k >= p.n0);
                }
                const bool b0 = (kind == BMB200_BLK_BIT) && !g1, b1 = (kind == BMB200_BLK_BIT) && g1;
                const bool q0 = (kind == BMB200_BLK_GAP) && !g1, q1 = (kind == BMB200_BLK_GAP) && g1;
                if (kind == BMB200_BLK_NULL && !g1) fl |= kFlNull0;
                if (kind == BMB200_BLK_FULL) { if (g1) fl |= kFlFull1; else { fl |= kFlFull0; ++nfull0; } }
                const uint32_t lt = (1u << lane) - 1u;
                uint32_t m, basei;
                m = __ballot_sync(0xffffffffu, b0);
                if (m) { if (lane == 0) basei = atomicAdd(&s_cnt[0], __popc(m)); basei = __shfl_sync(0xffffffffu, basei, 0);
                         if (b0) lst_bit0[basei + __popc(m & lt)] = rel; }
                m = __ballot_sync(0xffffffffu, b1);
                if (m) { if (lane == 0) basei = atomicAdd(&s_cnt[1], __popc(m)); basei = __shfl_sync(0xffffffffu, basei, 0);
                         if (b1) lst_bit1[basei + __popc(m & lt)] = rel; }
                m = __ballot_sync(0xffffffffu, q0);
                if (m) { if (lane == 0) basei = atomicAdd(&s_cnt[2], __popc(m)); basei = __shfl_sync(0xffffffffu, basei, 0);
                         if (q0) lst_gap[basei + __popc(m & lt)] = rel; }
                m = __ballot_sync(0xffffffffu, q1);
                if (m) { if (lane == 0) basei = atomicAdd(&s_cnt[3], __popc(m)); basei = __shfl_sync(0xffffffffu, basei, 0);
                         if (q1) lst_gap[kAggChunk - 1 - (basei + __popc(m & lt))] = rel; }
            }
            fl = __reduce_or_sync(0xffffffffu, fl);
            nfull0 = warp_sum(nfull0);
            if (lane == 0) { if (fl) atomicOr(&s_stat[0], fl); if (nfull0) atomicAdd(&s_stat[3], nfull0); }
            __syncthreads();
            const uint32_t nbit0 = s_cnt[0], nbit1 = s_cnt[1], ngap0 = s_cnt[2], ngap1 = s_cnt[3];
            if (tid == 0) { s_stat[1] += nbit0; s_stat[2] += ngap0; }

            // ---- bit phase: registers <- streamed bit-blocks ----
            bit_phase<OP, false>(bseg, lst_bit0, nbit0, acc0);
            if (OP == BMB200_OP_AND_SUB) bit_phase<OP, true>(bseg, lst_bit1, nbit1, acc1);

            // ---- GAP phase: one warp per GAP block, runs scattered into K ----
            const uint32_t ngap = ngap0 + ngap1;
            const uint32_t want0 = (OP == BMB200_OP_AND || OP == BMB200_OP_AND_SUB) ? 0u : 1u;
            uint32_t g = 0;
            if (lane == 0) g = atomicAdd(&s_gap_next, 1u);
            g = __shfl_sync(0xffffffffu, g, 0);
            const uint32_t* g32 = nullptr; uint32_t wf = 0;
            if (g < ngap) {
                const uint32_t rel = (g < ngap0) ? lst_gap[g] : lst_gap[kAggChunk - 1 - (g - ngap0)];
                g32 = reinterpret_cast<const uint32_t*>(gseg + (size_t)rel * kGapUnit);
                wf = ld_nc_u32(g32 + lane);
            }
            while (g < ngap) {
                uint32_t gn = 0;
                if (lane == 0) gn = atomicAdd(&s_gap_next, 1u);
                gn = __shfl_sync(0xffffffffu, gn, 0);
                const uint32_t* g32n = nullptr; uint32_t wfn = 0;
                if (gn < ngap) {
                    const uint32_t rel = (gn < ngap0) ? lst_gap[gn] : lst_gap[kAggChunk - 1 - (gn - ngap0)];
                    g32n = reinterpret_cast<const uint32_t*>(gseg + (size_t)rel * kGapUnit);
                    wfn = ld_nc_u32(g32n + lane);
                }
                const uint32_t want = (g < ngap0) ? want0 : 1u;
                gap_scatter<OP == BMB200_OP_XOR>(K, g32, wf, want, lane);
                g = gn; g32 = g32n; wf = wfn;
            }
            __syncthreads();
        }

        // ---- epilogue ----
        const uint32_t flags = s_stat[0], tot_bit0 = s_stat[1], tot_gap0 = s_stat[2], tot_full0 = s_stat[3];
        const uint4 k4 = K4[tid];
        uint4 R;
        int state;   // 0 = NULL (nothing stored), 1 = FULL, 2 = computed block R
        if (OP == BMB200_OP_OR) {
            if (flags & kFlFull0) state = 1;
            else if (tot_bit0 + tot_gap0 == 0) state = 0;
            else {
                state = 2;
                // all-ones is only detected inside the bit-block OR calls (>= 2 bit sources), :1948-1957
                const int ones = __syncthreads_and((acc0.x & acc0.y & acc0.z & acc0.w) == 0xffffffffu);
                if (ones && tot_bit0 >= 2) state = 1;
            }
            R = make_uint4(acc0.x | k4.x, acc0.y | k4.y, acc0.z | k4.z, acc0.w | k4.w);
        } else if (OP == BMB200_OP_XOR) {
            state = (tot_bit0 + tot_gap0 + tot_full0) ? 2 : 0;
            const uint32_t inv = (tot_full0 & 1u) ? 0xffffffffu : 0u;
            R = make_uint4(acc0.x ^ k4.x ^ inv, acc0.y ^ k4.y ^ inv, acc0.z ^ k4.z ^ inv, acc0.w ^ k4.w ^ inv);
        } else {
            if ((flags & kFlNull0) || p.n0 == 0) state = 0;
            else if (flags & kFlFull1) state = 0;
            else if (tot_bit0 + tot_gap0 == 0 && (OP == BMB200_OP_AND || p.n1 == 0)) state = 1;
            else state = 2;
            R = make_uint4(acc0.x & ~(acc1.x | k4.x), acc0.y & ~(acc1.y | k4.y),
                           acc0.z & ~(acc1.z | k4.z), acc0.w & ~(acc1.w | k4.w));
        }
        if (state == 0) R = make_uint4(0u, 0u, 0u, 0u);
        if (state == 1) R = make_uint4(~0u, ~0u, ~0u, ~0u);

        // popcount, digest (4 waves per warp: 8 threads x 4 words = one 32-word wave), transitions
        K4[tid] = R;                 // reuse K so each thread can see its left neighbour's last word
        const uint32_t nz = (R.x | R.y | R.z | R.w) != 0u;
        const uint32_t bal = __ballot_sync(0xffffffffu, nz);
        uint32_t dg4 = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) if ((bal >> (8 * q)) & 0xffu) dg4 |= (1u << q);
        uint32_t pc = warp_sum(popc4(R));
        __syncthreads();
        uint32_t prev = tid ? (K[4 * tid - 1] >> 31) : (R.x & 1u);
        uint32_t tr = __popc(R.x ^ ((R.x << 1) | prev));
        tr += __popc(R.y ^ ((R.y << 1) | (R.x >> 31)));
        tr += __popc(R.z ^ ((R.z << 1) | (R.y >> 31)));
        tr += __popc(R.w ^ ((R.w << 1) | (R.z >> 31)));
        tr = warp_sum(tr);
        if (lane == 0) { s_pc[warp] = pc; s_tr[warp] = tr; s_dg[warp] = dg4; }
        __syncthreads();
        uint32_t tpc = 0, ttr = 0; uint64_t dg = 0;
#pragma unroll
        for (int w = 0; w < kAggWarps; ++w) { tpc += s_pc[w]; ttr += s_tr[w]; dg |= (uint64_t)s_dg[w] << (4 * w); }
        const uint32_t runs = ttr + 1u;

        // result kind: aggregator stores nothing when the AND/SUB/XOR digest is empty; otherwise
        // copy_bit_block (opt_none) or the opt_copy_bit_block classification (src/bmblocks.h:1355-1409)
        uint32_t kd;
        if (state == 0) kd = BMB200_BLK_NULL;
        else if (state == 1) kd = BMB200_BLK_FULL;
        else if (OP != BMB200_OP_OR && dg == 0) kd = BMB200_BLK_NULL;
        else if (!p.compress) kd = BMB200_BLK_BIT;
        else if (runs == 1u) kd = tpc ? BMB200_BLK_FULL : BMB200_BLK_NULL;
        else if (runs < BMB200_GAP_THRESHOLD) kd = BMB200_BLK_GAP;
        else kd = BMB200_BLK_BIT;

        if (p.store_blocks && (kd == BMB200_BLK_BIT || kd == BMB200_BLK_GAP))
            st_stream_v4(reinterpret_cast<uint4*>(p.blocks) + (size_t)col * (kBlockWords / 4) + tid, R);
        if (tid == 0) {
            p.popcnt[col] = tpc;
            p.digest[col] = dg;
            p.nruns[col]  = runs;
            p.kind[col]   = (uint8_t)kd;
            if (tpc) atomicAdd(p.total, (unsigned long long)tpc);
        }
    }
}

}  // namespace bmb200
