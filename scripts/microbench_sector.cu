// microbench_sector.cu -- random 32-byte sector reads from a 512 MiB buffer (the footprint of BASELINE config 4's bit-blocks):
// the sector-granular random-access bound SURVEY 8d asks rank / select to be quoted against (not the 6.6 TB/s streaming peak).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_bin/microbench_sector scripts/microbench_sector.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t x) { x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull; x ^= x >> 27; x *= 0x94d049bb133111ebull; x ^= x >> 31; return x; }
template <int DEP>   // DEP = 1: every load's address depends on the previous load (latency chain), 0: independent loads
__global__ void gather(const uint4* __restrict__ buf, uint64_t n_sectors, uint32_t per_thread, uint32_t* out)
{
    const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint32_t acc = 0; uint64_t h = mix(t + 1);
    for (uint32_t i = 0; i < per_thread; ++i) {
        const uint64_t s = h % n_sectors;
        const uint4 v = buf[s * 2];                       // first 16 bytes of the sector
        acc += v.x ^ v.w;
        h = mix(h + (DEP ? v.x : i) + 0x9e3779b97f4a7c15ull);
    }
    if (acc == 0x12345678u) out[0] = acc;
}
int main()
{
    const uint64_t bytes = 512ull << 20, n_sectors = bytes / 32;
    uint4* buf; uint32_t* out;
    cudaMalloc(&buf, bytes); cudaMalloc(&out, 4); cudaMemset(buf, 1, bytes);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int dep = 0; dep < 2; ++dep) {
        const uint32_t per = 64, threads = 256, blocks = 148 * 64;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(a);
            if (dep) gather<1><<<blocks, threads>>>(buf, n_sectors, per, out); else gather<0><<<blocks, threads>>>(buf, n_sectors, per, out);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            const double n = (double)per * threads * blocks;
            if (rep == 2) printf("{\"dependent_chain\": %d, \"sectors\": %.0f, \"ms\": %.3f, \"Gsectors_per_s\": %.2f, \"GBps_of_32B_sectors\": %.1f}\n", dep, n, ms, n / ms / 1e6, n * 32 / ms / 1e6);
        }
    }
    return 0;
}
