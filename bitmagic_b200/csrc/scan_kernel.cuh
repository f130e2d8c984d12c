// scan_kernel.cuh -- bit-sliced comparison scan over the planes of a sparse vector (sm_100a).
//
// Replaces the reference's sparse_vector_scanner<SV> searches for unsigned sparse vectors
//   find_eq   src/bmsparsevec_algo.h:4360-4395 (prepare_and_sub_aggregator :2593-2632 + aggregator::combine_and_sub)
//   find_gt / find_ge / find_lt / find_le / find_range   src/bmsparsevec_algo.h:1135-1182 (find_gt_horizontal & co.)
// which the reference builds out of many aggregator / bvector passes over the planes.  Here ONE pass over the
// planes of a block column answers any of them: walking the planes from the most significant bit down, every thread
// keeps, for its 128 bits of the column,
//     EQ = elements equal to the search value on all planes seen so far      (starts as the index universe U)
//     GT = elements already known to be greater
//   plane bit of the value == 1 :  EQ &= P                      (elements with a 0 here are smaller: they leave EQ)
//   plane bit of the value == 0 :  GT |= EQ & P ;  EQ &= ~P
// and the predicates fall out at the end:  eq = EQ, gt = GT, ge = GT | EQ, lt = U & ~(GT | EQ), le = U & ~GT,
// range [a, b] = ge(a) & le(b) (two state pairs in the same pass).  Every plane block is read exactly once per
// (column, value) work item; the values of one column are adjacent work items so the planes come from L2 after the
// first one (the same batching as aggregator::pipeline, agg_kernel.cuh).
//
// Plane blocks may be NULL / FULL / bit / GAP like any other block of the set; GAP planes are expanded through the
// 8 KB shared mask (block-wide run scatter), bit planes stream through 128-bit loads, four planes in flight.
// The result goes through the same epilogue as the aggregation kernel (popcount, digest, run count, kind, bit->GAP).
#pragma once
#include "agg_kernel.cuh"

namespace bmb200 {

struct ScanParams {
    AggParams out;             // set view + result buffers + work counter (group / goff unused; n_groups = n_values)
    uint32_t  plane0;          // plane j (bit j of the value) = set vector plane0 + j
    uint32_t  n_planes;        // <= 64
    uint32_t  universe;        // set vector holding the searchable index range (size mask / NOT-NULL plane); 0xffffffff = everything
    uint32_t  pred;            // BMB200_SCAN_*
    const uint64_t* values;    // device: n_values entries (RANGE: 2 * n_values, lo then hi of each range)
};

constexpr int kScanBatch = 4;  // planes whose loads are issued together

// selected (1-) runs of one GAP block straight from global memory into the zeroed mask, all 512 threads
__device__ __forceinline__ void gap_expand_block(uint32_t Ks, const uint16_t* __restrict__ g, int tid)
{
    const uint32_t hdr = g[0];
    const uint32_t len = hdr >> 3;
    const bool odd = (hdr & 1u) != 0u;                       // first run is a 1-run
    const uint32_t nsel = odd ? (len + 1u) >> 1 : len >> 1;
    const uint16_t* a0 = g + (odd ? 0 : 1);
    for (uint32_t j = tid; j < nsel; j += kAggThreads) {
        const uint32_t sv = a0[2u * j], ev = a0[2u * j + 1u];
        apply_run<true>(Ks, (odd && j == 0u) ? 0u : sv + 1u, ev);   // runs of one block are disjoint: XOR into zeros == OR
    }
}

__device__ __forceinline__ void scan_step(uint4& eq, uint4& gt, const uint4& P, bool vbit)
{
    if (vbit) { eq.x &= P.x; eq.y &= P.y; eq.z &= P.z; eq.w &= P.w; }
    else {
        gt.x |= eq.x & P.x; gt.y |= eq.y & P.y; gt.z |= eq.z & P.z; gt.w |= eq.w & P.w;
        eq.x &= ~P.x; eq.y &= ~P.y; eq.z &= ~P.z; eq.w &= ~P.w;
    }
}

__global__ void __launch_bounds__(kAggThreads, kCtasPerSm) scan_kernel(const ScanParams sp)
{
    __shared__ __align__(16) uint32_t K[kScanBatch][kBlockWords];   // expansion buffers: the GAP planes of one batch are expanded together
    __shared__ uint32_t s_desc[65];                          // descriptors of the planes (+ universe) of this column
    __shared__ uint32_t s_col;
    __shared__ uint32_t s_pc[kAggWarps], s_tr[kAggWarps], s_dg[kAggWarps];

    const AggParams& p = sp.out;
    const int tid = threadIdx.x;
    const uint32_t M = p.set.n_vec;
    uint4* K4 = reinterpret_cast<uint4*>(K[0]);
    const uint32_t Ks = smem_u32(K[0]);
    const bool is_range = (sp.pred == BMB200_SCAN_RANGE);

    uint32_t next_item = 0;
    if (tid == 0) next_item = atomicAdd(p.work_counter, 1u);
    for (;;) {
        __syncthreads();
        if (tid == 0) s_col = next_item;
        __syncthreads();
        const uint32_t item = s_col;
        if (item >= p.n_cols * p.n_groups) break;
        if (tid == 0) next_item = atomicAdd(p.work_counter, 1u);
        const uint32_t colx = item / p.n_groups, vi = item - colx * p.n_groups;
        const uint32_t col = vi * p.n_cols + colx;           // output slot (value-major)
        const uint32_t nb = p.nb_from + colx;
        const uint32_t* drow = p.set.desc + (size_t)nb * M;
        if (tid < (int)sp.n_planes) s_desc[tid] = drow[sp.plane0 + tid];
        if (tid == 64) s_desc[64] = (sp.universe == 0xffffffffu) ? BMB200_BLK_FULL : drow[sp.universe];
        const uint4* bseg = reinterpret_cast<const uint4*>(p.set.bit_pool) + p.set.bit_base[nb] * (size_t)(kBlockWords / 4) + tid;
        const uint16_t* gseg = p.set.gap_pool + p.set.gap_base[nb] * (size_t)kGapUnit;
        uint64_t va = sp.values[is_range ? 2u * vi : vi];
        uint64_t vb = is_range ? sp.values[2u * vi + 1u] : 0ull;
        if (is_range && vb < va) { const uint64_t t = va; va = vb; vb = t; }      // find_range swaps reversed bounds, :2871-2872
        __syncthreads();

        // one block of the column as this thread's 4 words: GAP blocks go through the shared mask (2 block barriers)
        auto load_gap = [&](uint32_t d) -> uint4 {
            K4[tid] = make_uint4(0u, 0u, 0u, 0u);
            __syncthreads();
            const uint32_t rel = d >> 2;
            gap_expand_block(Ks, gseg + (size_t)(rel & kRelMask) * kGapUnit + (rel >> 29), tid);
            __syncthreads();
            return K4[tid];
        };
        auto load_now = [&](uint32_t d) -> uint4 {           // NULL / FULL / bit (GAP handled by the caller)
            const uint32_t kind = d & 3u;
            if (kind == BMB200_BLK_BIT) return ld_stream_v4(bseg + (size_t)(d >> 2) * (kBlockWords / 4));
            return kind == BMB200_BLK_FULL ? make_uint4(~0u, ~0u, ~0u, ~0u) : make_uint4(0u, 0u, 0u, 0u);
        };

        const uint32_t du = s_desc[64];
        const uint4 U = ((du & 3u) == BMB200_BLK_GAP) ? load_gap(du) : load_now(du);
        uint4 eqA = U, gtA = make_uint4(0u, 0u, 0u, 0u), eqB = U, gtB = make_uint4(0u, 0u, 0u, 0u);
        // a value with bits above the top plane is greater than every element
        if (sp.n_planes < 64u && (va >> sp.n_planes)) { eqA = make_uint4(0u, 0u, 0u, 0u); va = 0ull; }
        if (sp.n_planes < 64u && (vb >> sp.n_planes)) { eqB = make_uint4(0u, 0u, 0u, 0u); vb = 0ull; }

        for (int jt = (int)sp.n_planes - 1; jt >= 0; jt -= kScanBatch) {
            uint4 P[kScanBatch];
            uint32_t d[kScanBatch];
            bool any_gap = false;
#pragma unroll
            for (int u = 0; u < kScanBatch; ++u) {           // issue the bit-plane loads of the batch together
                const int j = jt - u;
                d[u] = (j >= 0) ? s_desc[j] : BMB200_BLK_NULL;
                if ((d[u] & 3u) != BMB200_BLK_GAP) P[u] = load_now(d[u]); else any_gap = true;
            }
            if (any_gap) {                                   // uniform: the GAP planes of the batch share one pair of barriers
#pragma unroll
                for (int u = 0; u < kScanBatch; ++u)
                    if ((d[u] & 3u) == BMB200_BLK_GAP) reinterpret_cast<uint4*>(K[u])[tid] = make_uint4(0u, 0u, 0u, 0u);
                __syncthreads();
#pragma unroll
                for (int u = 0; u < kScanBatch; ++u)
                    if ((d[u] & 3u) == BMB200_BLK_GAP) {
                        const uint32_t rel = d[u] >> 2;
                        gap_expand_block(Ks + (uint32_t)u * kBlockWords * 4u, gseg + (size_t)(rel & kRelMask) * kGapUnit + (rel >> 29), tid);
                    }
                __syncthreads();
#pragma unroll
                for (int u = 0; u < kScanBatch; ++u)
                    if ((d[u] & 3u) == BMB200_BLK_GAP) P[u] = reinterpret_cast<const uint4*>(K[u])[tid];
                __syncthreads();                             // the buffers are rewritten by the next batch
            }
#pragma unroll
            for (int u = 0; u < kScanBatch; ++u) {
                const int j = jt - u;
                if (j < 0) break;
                scan_step(eqA, gtA, P[u], (va >> j) & 1ull);
                if (is_range) scan_step(eqB, gtB, P[u], (vb >> j) & 1ull);
            }
        }

        uint4 R;
        switch (sp.pred) {
        case BMB200_SCAN_EQ: R = eqA; break;
        case BMB200_SCAN_GT: R = gtA; break;
        case BMB200_SCAN_GE: R = make_uint4(gtA.x | eqA.x, gtA.y | eqA.y, gtA.z | eqA.z, gtA.w | eqA.w); break;
        case BMB200_SCAN_LT: R = make_uint4(U.x & ~(gtA.x | eqA.x), U.y & ~(gtA.y | eqA.y), U.z & ~(gtA.z | eqA.z), U.w & ~(gtA.w | eqA.w)); break;
        case BMB200_SCAN_LE: R = make_uint4(U.x & ~gtA.x, U.y & ~gtA.y, U.z & ~gtA.z, U.w & ~gtA.w); break;
        default:             // RANGE: ge(a) & le(b)
            R = make_uint4((gtA.x | eqA.x) & U.x & ~gtB.x, (gtA.y | eqA.y) & U.y & ~gtB.y,
                           (gtA.z | eqA.z) & U.z & ~gtB.z, (gtA.w | eqA.w) & U.w & ~gtB.w);
            break;
        }
        finish_block<true>(p, col, colx, vi, R, 2, K[0], s_pc, s_tr, s_dg);
    }
}

}  // namespace bmb200
