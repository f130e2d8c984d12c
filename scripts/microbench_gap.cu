// microbench_gap.cu -- ceiling of the "scatter GAP runs into an 8 KB shared-memory mask" step, isolated from
// HBM: the run-end stream sits in shared memory and is re-scattered many times.  Prints runs/cycle/SM and the
// equivalent GAP GB/s (4 bytes per 1-run) for several scatter flavours.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/microbench_gap scripts/microbench_gap.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

constexpr int kThreads = 512;
constexpr int kStreamWords = 4096;   // 16 KB of (start,end) u16 pairs staged in smem

__device__ __forceinline__ void red_or(uint32_t* p, uint32_t v)
{ asm volatile("red.shared.or.b32 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory"); }

template <int MODE>
__global__ void __launch_bounds__(kThreads, 2) scatter_kernel(const uint32_t* __restrict__ pairs, int n_pairs, int iters,
                                                              uint32_t* out, unsigned long long* cycles)
{
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t* K = smem;
    uint32_t* S = smem + 2048;
    uint8_t* B = reinterpret_cast<uint8_t*>(smem + 2048 + kStreamWords);
    const int tid = threadIdx.x;
    for (int i = tid; i < 2048; i += kThreads) K[i] = 0;
    for (int i = tid; i < n_pairs; i += kThreads) S[i] = pairs[i];
    if (MODE == 3) for (int i = tid; i < 65536 / 4; i += kThreads) reinterpret_cast<uint32_t*>(B)[i] = 0;
    __syncthreads();
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        for (int i = tid; i < n_pairs; i += kThreads) {
            const uint32_t w = S[i];
            const uint32_t s = w & 0xffffu, e = w >> 16;
            const uint32_t ws = s >> 5, we = e >> 5;
            if (MODE == 0) {                       // red.or, single-word fast path + general path
                if (ws == we) red_or(K + ws, (0xffffffffu << (s & 31)) & (0xffffffffu >> (31 - (e & 31))));
                else { red_or(K + ws, 0xffffffffu << (s & 31)); red_or(K + we, 0xffffffffu >> (31 - (e & 31)));
                       for (uint32_t x = ws + 1; x < we; ++x) red_or(K + x, 0xffffffffu); }
            } else if (MODE == 1) {                // test before set
                const uint32_t m = (0xffffffffu << (s & 31)) & (0xffffffffu >> (31 - (e & 31)));
                if (ws == we) { if ((K[ws] & m) != m) red_or(K + ws, m); }
                else { red_or(K + ws, 0xffffffffu << (s & 31)); red_or(K + we, 0xffffffffu >> (31 - (e & 31)));
                       for (uint32_t x = ws + 1; x < we; ++x) red_or(K + x, 0xffffffffu); }
            } else if (MODE == 2) {                // plain (racy) store: upper bound without atomics
                K[ws] = (0xffffffffu << (s & 31));
            } else if (MODE == 3) {                // byte map: one plain byte store per single-bit run
                for (uint32_t p = s; p <= e; ++p) B[p] = 1;
            } else if (MODE == 4) {                // atomicOr with return value (ATOMS)
                uint32_t old = atomicOr(K + ws, (0xffffffffu << (s & 31)) & (0xffffffffu >> (31 - (e & 31))));
                if (old == 0x12345u) K[0] = 1;
            }
        }
        __syncthreads();
    }
    const unsigned long long t1 = clock64();
    if (tid == 0) { cycles[blockIdx.x] = t1 - t0; }
    if (MODE == 3) { if (tid < 64) out[blockIdx.x * 64 + tid] = B[tid * 97]; }
    else if (tid < 64) out[blockIdx.x * 64 + tid] = K[tid * 31];
}

template <int MODE>
void run(const char* name, const std::vector<uint32_t>& pairs, int iters, int sm)
{
    uint32_t* d_pairs; uint32_t* d_out; unsigned long long* d_cyc;
    const int grid = sm * 2;
    cudaMalloc(&d_pairs, pairs.size() * 4); cudaMalloc(&d_out, grid * 64 * 4); cudaMalloc(&d_cyc, grid * 8);
    cudaMemcpy(d_pairs, pairs.data(), pairs.size() * 4, cudaMemcpyHostToDevice);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const size_t smem = (2048 + kStreamWords) * 4 + (MODE == 3 ? 65536 : 16);
    cudaFuncSetAttribute(scatter_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    scatter_kernel<MODE><<<grid, kThreads, smem>>>(d_pairs, (int)pairs.size(), 2, d_out, d_cyc);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    scatter_kernel<MODE><<<grid, kThreads, smem>>>(d_pairs, (int)pairs.size(), iters, d_out, d_cyc);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    std::vector<unsigned long long> cyc(grid);
    cudaMemcpy(cyc.data(), d_cyc, grid * 8, cudaMemcpyDeviceToHost);
    double mean = 0; for (auto c : cyc) mean += (double)c; mean /= grid;
    const double runs = (double)pairs.size() * iters;          // per CTA
    const double rpc_sm = runs * 2 / mean;                      // 2 CTAs per SM
    const double gbs = runs * grid * 4.0 / (ms * 1e-3) / 1e9;
    printf("%-28s %8.3f ms  %6.2f runs/cycle/SM  %8.1f GB/s GAP-equivalent (err=%s)\n", name, ms, rpc_sm, gbs,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(d_pairs); cudaFree(d_out); cudaFree(d_cyc);
}

int main()
{
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sm = prop.multiProcessorCount;
    printf("device %s, %d SMs\n", prop.name, sm);
    for (int density_runs : {64, 330, 1200}) {
        // stream of sorted single-bit runs: blocks of `density_runs` random positions in [0,65536), concatenated
        std::vector<uint32_t> pairs; srand(1);
        while ((int)pairs.size() + density_runs <= kStreamWords) {
            std::vector<uint32_t> pos;
            for (int i = 0; i < density_runs; ++i) pos.push_back(rand() & 0xffff);
            std::sort(pos.begin(), pos.end());
            for (auto p : pos) pairs.push_back(p | (p << 16));
        }
        printf("-- runs per block: %d (single-bit runs), %zu pairs staged\n", density_runs, pairs.size());
        run<0>("red.or", pairs, 400, sm);
        run<1>("test-then-red.or", pairs, 400, sm);
        run<2>("plain store (racy bound)", pairs, 400, sm);
        run<3>("byte map store", pairs, 400, sm);
        run<4>("atomicOr with return", pairs, 400, sm);
    }
    // longer runs: mean length 40 bits
    {
        std::vector<uint32_t> pairs; srand(2);
        while ((int)pairs.size() + 300 <= kStreamWords) {
            std::vector<uint32_t> pos;
            for (int i = 0; i < 300; ++i) pos.push_back(rand() & 0xffff);
            std::sort(pos.begin(), pos.end());
            for (auto p : pos) { uint32_t e = std::min<uint32_t>(65535u, p + (rand() % 80)); pairs.push_back(p | (e << 16)); }
        }
        printf("-- runs of mean length 40\n");
        run<0>("red.or", pairs, 400, sm);
        run<3>("byte map store", pairs, 100, sm);
    }
    return 0;
}
