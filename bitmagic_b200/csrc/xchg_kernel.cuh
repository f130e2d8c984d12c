// xchg_kernel.cuh -- the exchange step of a sharded aggregation done by this library's own kernels over NVLink peer memory.
//
// Every rank owns one buffer `xbuf` = [2 slots][nranks][xwords] u32 + [2][nranks] sequence flags, exported through CUDA IPC and
// mapped by every other rank (one process per GPU).  After its aggregation kernel a rank PUSHES its (per-column popcounts |
// cardinality) row into slot s&1 of every peer with plain peer stores, fences, and publishes the exchange number s in that peer's
// flag word (st.release.sys).  Nobody runs a collective kernel: a rank's SMs are only involved in its own pushes, so the exchange
// costs one ~5 us launch on the aggregation stream instead of an all-gather kernel that cannot share the SMs with the persistent
// aggregation kernel (measured on 2 B200s: 45 us per step with ncclAllGather on a side stream).
// Reading side: xchg_wait_kernel spins (ld.acquire.sys) until every peer's flag of the slot shows the awaited number.
// Flow control with two slots: before pushing exchange s a rank waits until each peer has published s-1 -- that peer has then left
// exchange s-2 behind (bmb200_exchange_fetch only hands out the latest exchange), so its slot may be overwritten.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bmb200 {

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// spin until *flag >= want; false after ~timeout_ns (a peer that died must not hang this GPU)
__device__ __forceinline__ bool xchg_spin(const uint32_t* flag, uint32_t want, unsigned long long timeout_ns)
{
    if (ld_acquire_sys(flag) >= want) return true;
    const unsigned long long t0 = global_ns();
    for (;;) {
        if (ld_acquire_sys(flag) >= want) return true;
        __nanosleep(200);
        if (global_ns() - t0 > timeout_ns) return false;
    }
}

struct XchgParams {
    uint32_t* const* peers;        // [nranks] every rank's xbuf as mapped on this device (own entry = own xbuf)
    uint32_t nranks, rank;
    uint32_t xwords;               // words per rank and slot
    uint32_t slot, seq;            // this exchange: slot = seq & 1 (seq counts from 1)
    uint32_t n_cols, n;            // columns this rank produced, columns per rank in the exchange (n >= n_cols: zero padding)
    const uint32_t* popcnt;        // [n_cols]
    const unsigned long long* total;
    uint32_t* err;                 // set to 1 when a wait timed out; err + 2 .. : 4 x u64 globaltimer stamps of CTA 0 (start, after the
                                   // flow-control wait, after the stores + fence, end) for BMB200_TRACE
    unsigned long long timeout_ns;
};

__device__ __forceinline__ uint32_t* xchg_row(uint32_t* xbuf, uint32_t nranks, uint32_t xwords, uint32_t slot, uint32_t r)
{ return xbuf + ((size_t)slot * nranks + r) * xwords; }
__device__ __forceinline__ uint32_t* xchg_flag(uint32_t* xbuf, uint32_t nranks, uint32_t xwords, uint32_t slot, uint32_t r)
{ return xbuf + (size_t)2 * nranks * xwords + (size_t)slot * nranks + r; }

// one CTA per peer (including this rank itself: its own row goes through the same code)
__global__ void __launch_bounds__(256) xchg_push_kernel(XchgParams p)
{
    const uint32_t q = blockIdx.x;
    unsigned long long* stamp = reinterpret_cast<unsigned long long*>(p.err + 2);
    const bool probe = threadIdx.x == 0 && q == (p.rank ^ 1u) % p.nranks;     // the CTA that writes to a real peer
    if (probe) stamp[0] = global_ns();
    uint32_t* mine = p.peers[p.rank];
    uint32_t* dst_buf = p.peers[q];
    if (threadIdx.x == 0 && p.seq > 1u && q != p.rank) {
        // flow control: peer q has published exchange seq-1 here => it is done with exchange seq-2, whose slot this push overwrites
        if (!xchg_spin(xchg_flag(mine, p.nranks, p.xwords, p.slot ^ 1u, q), p.seq - 1u, p.timeout_ns)) atomicExch(p.err, 1u);
    }
    __syncthreads();
    if (probe) stamp[1] = global_ns();
    uint32_t* row = xchg_row(dst_buf, p.nranks, p.xwords, p.slot, p.rank);
    for (uint32_t i = threadIdx.x; i < p.n; i += blockDim.x) row[i] = i < p.n_cols ? p.popcnt[i] : 0u;
    if (threadIdx.x == 0) {
        const unsigned long long t = *p.total;
        row[p.n] = (uint32_t)t; row[p.n + 1u] = (uint32_t)(t >> 32);
    }
    __threadfence_system();
    __syncthreads();
    if (probe) stamp[2] = global_ns();
    if (threadIdx.x == 0) st_release_sys(xchg_flag(dst_buf, p.nranks, p.xwords, p.slot, p.rank), p.seq);
    if (probe) stamp[3] = global_ns();
}

// the launching stream continues once every rank's row of exchange `seq` has landed in this rank's buffer
__global__ void __launch_bounds__(64) xchg_wait_kernel(uint32_t* xbuf, uint32_t nranks, uint32_t xwords, uint32_t slot, uint32_t seq,
                                                       uint32_t* err, unsigned long long timeout_ns)
{
    for (uint32_t q = threadIdx.x; q < nranks; q += blockDim.x)
        if (!xchg_spin(xchg_flag(xbuf, nranks, xwords, slot, q), seq, timeout_ns)) atomicExch(err, 1u);
}

}  // namespace bmb200
