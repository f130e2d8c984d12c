// binop_kernel.cuh -- GAP x GAP block merge for the two-operand bvector ops (sm_100a).
//
// Replaces gap_buff_op / gap_operation_or / _and / _sub / _xor (src/bmfunc.h:3747,7423,7275,7469,7342) as they are used by
// bvector::combine_operation_block_* (src/bm.h:6945-7380): when both argument blocks are GAP blocks the reference merges the two
// run-end lists with a two-pointer loop and stores the result as a GAP block (clone_gap_block, src/bmblocks.h:866: all-zero -> no
// block, longer than the largest GAP capacity -> bit-block) -- no 8 KB block is ever materialised.
//
// The merge here has no serial dependency.  One warp per block column:
//   * both run-end lists are staged in shared memory (coalesced 16-bit loads);
//   * every run end e of A (and every run end of B that is not also an end of A) looks up, by binary search in the OTHER list,
//     the value of the other operand around e; the result value left and right of e follows from the parities of the two run
//     indexes, so "e is a run end of the result" is decided per boundary, independently;
//   * a warp scan over the keep-flags of both lists gives every kept boundary its slot: (kept ends of its own list before it) +
//     (kept ends of the other list below e), the second term read at the index the binary search already found;
//   * the output runs are then walked once more for the popcount and the 64-wave digest of the column's metadata.
// Work: (lenA + lenB) * log2(len) shared-memory probes per column instead of 2 x 8 KB of expansion + an 8 KB re-compression.
#pragma once
#include "agg_kernel.cuh"

namespace bmb200 {

constexpr int kMergeWarps = 4;
constexpr uint32_t kMergeCap = BMB200_GAP_MAX_WORDS + 8u;            // entries per list (index 0 unused, ends at 1..len, prefix arrays use len + 1)
constexpr size_t kMergeSmemPerWarp = 6u * kMergeCap * sizeof(uint16_t);
constexpr size_t kMergeSmem = kMergeWarps * kMergeSmemPerWarp;

struct MergeParams {
    SetView set;
    uint32_t va, vb;           // the two argument vectors
    uint32_t op;               // BINOP_*
    uint32_t nb_from, n_cols;
    uint8_t*  kind; uint32_t* popcnt; uint64_t* digest; uint32_t* nruns; uint16_t* gaps;
    unsigned long long* total;
};

__device__ __forceinline__ uint32_t binop_bit(uint32_t op, uint32_t a, uint32_t b)
{   // b arrives already inverted for SUB (gap_buff_op's vect2_mask = 1, src/bmfunc.h:7469)
    return op == BINOP_OR ? (a | b) : op == BINOP_XOR ? (a ^ b) : (a & b);
}
// first index j in [1, len] with ends[j] >= e (ends[len] = 65535 bounds the search)
__device__ __forceinline__ uint32_t merge_lower_bound(const uint16_t* ends, uint32_t len, uint32_t e)
{
    uint32_t lo = 1u, hi = len;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if ((uint32_t)ends[mid] < e) lo = mid + 1u; else hi = mid; }
    return lo;
}

__global__ void __launch_bounds__(kMergeWarps * 32) gap_merge_kernel(const MergeParams p)
{
    extern __shared__ __align__(16) uint8_t merge_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint16_t* A  = reinterpret_cast<uint16_t*>(merge_smem + (size_t)wib * kMergeSmemPerWarp);
    uint16_t* B  = A + kMergeCap;
    uint16_t* XA = B + kMergeCap;       // pass 1: index found in the other list | keep flag << 15 ; pass 2: exclusive prefix of the keep flags
    uint16_t* XB = XA + kMergeCap;
    uint16_t* PA = XB + kMergeCap;
    uint16_t* PB = PA + kMergeCap;
    const uint32_t warps_total = gridDim.x * kMergeWarps;
    for (uint32_t colx = blockIdx.x * kMergeWarps + wib; colx < p.n_cols; colx += warps_total) {
        const uint32_t nb = p.nb_from + colx;
        const uint32_t da = p.set.desc[(size_t)nb * p.set.n_vec + p.va], db = p.set.desc[(size_t)nb * p.set.n_vec + p.vb];
        if ((da & 3u) != BMB200_BLK_GAP || (db & 3u) != BMB200_BLK_GAP) continue;          // every other pairing goes through agg_kernel
        const uint16_t* ga = p.set.gap_pool + (p.set.gap_base[nb] + ((da >> 2) & BMB200_DESC_REL_MASK)) * (size_t)kGapUnit + (da >> 31);
        const uint16_t* gb = p.set.gap_pool + (p.set.gap_base[nb] + ((db >> 2) & BMB200_DESC_REL_MASK)) * (size_t)kGapUnit + (db >> 31);
        const uint32_t ha = ga[0], hb = gb[0], la = ha >> 3, lb = hb >> 3;
        const uint32_t fa = ha & 1u, fb = (hb & 1u) ^ (p.op == BINOP_SUB ? 1u : 0u);
        __syncwarp();
        for (uint32_t k = 1u + lane; k <= la; k += 32u) A[k] = ga[k];
        for (uint32_t k = 1u + lane; k <= lb; k += 32u) B[k] = gb[k];
        __syncwarp();
        // ---- pass 1: is boundary e a run end of the result?  (value of the result left of e) != (value right of e)
        for (uint32_t i = 1u + lane; i <= la; i += 32u) {
            const uint32_t e = A[i], j = merge_lower_bound(B, lb, e);
            const uint32_t a_cur = fa ^ ((i - 1u) & 1u), b_cur = fb ^ ((j - 1u) & 1u);
            const uint32_t b_nxt = ((uint32_t)B[j] == e) ? b_cur ^ 1u : b_cur;
            const uint32_t keep = (e != 65535u) && (binop_bit(p.op, a_cur, b_cur) != binop_bit(p.op, a_cur ^ 1u, b_nxt));
            XA[i] = (uint16_t)(j | (keep << 15));
        }
        for (uint32_t j = 1u + lane; j <= lb; j += 32u) {
            const uint32_t e = B[j], i = merge_lower_bound(A, la, e);
            uint32_t keep = 0u;
            if ((uint32_t)A[i] != e) {                                   // a shared end belongs to A's pass
                const uint32_t a_cur = fa ^ ((i - 1u) & 1u), b_cur = fb ^ ((j - 1u) & 1u);
                keep = binop_bit(p.op, a_cur, b_cur) != binop_bit(p.op, a_cur, b_cur ^ 1u);
            }
            XB[j] = (uint16_t)(i | (keep << 15));
        }
        __syncwarp();
        // ---- exclusive prefix sums of the keep flags (PA[i] = kept ends of A with index < i; one extra entry = the total)
        uint32_t carry = 0u;
        for (uint32_t base = 1u; base <= la + 1u; base += 32u) {
            const uint32_t i = base + lane;
            const uint32_t f = (i <= la) ? (uint32_t)(XA[i] >> 15) : 0u;
            uint32_t x = f;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
            if (i <= la + 1u) PA[i] = (uint16_t)(carry + x - f);
            carry += __shfl_sync(0xffffffffu, x, 31);
        }
        const uint32_t keptA = carry;
        carry = 0u;
        for (uint32_t base = 1u; base <= lb + 1u; base += 32u) {
            const uint32_t j = base + lane;
            const uint32_t f = (j <= lb) ? (uint32_t)(XB[j] >> 15) : 0u;
            uint32_t x = f;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
            if (j <= lb + 1u) PB[j] = (uint16_t)(carry + x - f);
            carry += __shfl_sync(0xffffffffu, x, 31);
        }
        const uint32_t runs = keptA + carry + 1u;                       // + the final run end 65535
        const uint32_t first = binop_bit(p.op, fa, fb);
        __syncwarp();
        uint8_t kd;
        if (runs > BMB200_GAP_THRESHOLD) kd = 0xffu;                    // clone_gap_block: gap_calc_level(len) < 0 -> bit-block; agg_kernel takes the column
        else if (runs == 1u && first == 0u) kd = BMB200_BLK_NULL;       // level 0 and gap_is_all_zero: nothing stored
        else kd = BMB200_BLK_GAP;
        uint32_t pc = 0u, dlo = 0u, dhi = 0u;
        if (kd == BMB200_BLK_GAP) {
            // ---- pass 2: every kept boundary writes itself to its slot (B ends below an A end e are exactly B[1 .. j-1])
            uint16_t* out = p.gaps + (size_t)colx * kGapMax;
            for (uint32_t i = 1u + lane; i <= la; i += 32u) {
                const uint32_t x = XA[i];
                if (x >> 15) out[1u + PA[i] + PB[x & 0x7fffu]] = A[i];
            }
            for (uint32_t j = 1u + lane; j <= lb; j += 32u) {
                const uint32_t x = XB[j];
                if (x >> 15) out[1u + PB[j] + PA[x & 0x7fffu]] = B[j];
            }
            if (lane == 0) {
                out[runs] = 65535u;
                const uint32_t lvl = runs <= 124u ? 0u : runs <= 252u ? 1u : runs <= 508u ? 2u : 3u;     // gap_calc_level src/bmfunc.h:5418
                out[0] = (uint16_t)(first | (lvl << 1) | (runs << 3));
            }
            __syncwarp();
            // ---- popcount + 64-wave digest of the merged block (its 1-runs)
            for (uint32_t r = 1u + lane; r <= runs; r += 32u) {
                if ((first ^ ((r - 1u) & 1u)) == 0u) continue;
                const uint32_t s = (r == 1u) ? 0u : (uint32_t)out[r - 1u] + 1u, e = out[r];
                pc += e - s + 1u;
                const uint32_t w0 = s >> 10, w1 = e >> 10;               // waves of 1024 bits
                const uint64_t m = ((w1 - w0 == 63u) ? ~0ull : ((1ull << (w1 - w0 + 1u)) - 1ull)) << w0;
                dlo |= (uint32_t)m; dhi |= (uint32_t)(m >> 32);
            }
            pc = warp_sum(pc);
            dlo = __reduce_or_sync(0xffffffffu, dlo); dhi = __reduce_or_sync(0xffffffffu, dhi);
        }
        if (lane == 0) {
            p.kind[colx] = kd;
            p.popcnt[colx] = pc;
            p.digest[colx] = ((uint64_t)dhi << 32) | dlo;
            p.nruns[colx] = runs;
            if (pc) atomicAdd(p.total, (unsigned long long)pc);
        }
    }
}

}  // namespace bmb200
