// Experiment: does cudaLimitMaxL2FetchGranularity change random 32-byte gather throughput on B200?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}
template <int ILP, int BYTES>
__global__ void gather(const uint8_t* buf, uint64_t sectors, uint64_t loads, uint64_t seed, unsigned long long* sink) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (uint64_t)gridDim.x * blockDim.x;
    uint64_t acc = 0;
    for (uint64_t base = tid * ILP; base < loads; base += nth * ILP) {
        uint64_t v[ILP];
#pragma unroll
        for (int u = 0; u < ILP; u++) {
            const uint64_t s = __umul64hi(mix(seed + base + u), sectors);
            if (BYTES == 8) v[u] = __ldg((const unsigned long long*)(buf + s * 32));
            else { uint4 q = __ldg((const uint4*)(buf + s * 32)); v[u] = q.x ^ q.w; }
        }
#pragma unroll
        for (int u = 0; u < ILP; u++) acc ^= v[u];
    }
    if (acc == 0x1234567) atomicAdd(sink, 1ull);
}
template <int ILP>
__global__ void prefetch_only(const uint8_t* buf, uint64_t sectors, uint64_t loads, uint64_t seed) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t base = tid * ILP; base < loads; base += nth * ILP) {
#pragma unroll
        for (int u = 0; u < ILP; u++) {
            const uint64_t s = __umul64hi(mix(seed + base + u), sectors);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(buf + s * 32));
        }
    }
}
// demand loads of window i while prefetching window i+1 (same thread): does a prefetch one "iteration" ahead hide DRAM latency?
template <int ILP>
__global__ void gather_pf(const uint8_t* buf, uint64_t sectors, uint64_t loads, uint64_t seed, unsigned long long* sink) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (uint64_t)gridDim.x * blockDim.x;
    uint64_t acc = 0;
    for (uint64_t base = tid * ILP; base < loads; base += nth * ILP) {
        const uint64_t nb = base + nth * ILP;
#pragma unroll
        for (int u = 0; u < ILP; u++) {
            const uint64_t s = __umul64hi(mix(seed + nb + u), sectors);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(buf + s * 32));
        }
        uint64_t v[ILP];
#pragma unroll
        for (int u = 0; u < ILP; u++) {
            const uint64_t s = __umul64hi(mix(seed + base + u), sectors);
            v[u] = __ldg((const unsigned long long*)(buf + s * 32));
        }
#pragma unroll
        for (int u = 0; u < ILP; u++) acc ^= v[u];
    }
    if (acc == 0x1234567) atomicAdd(sink, 1ull);
}
int main() {
    const uint64_t loads = 1ull << 28;
    uint8_t* buf; unsigned long long* sink;
    const uint64_t maxb = 2688ull << 20;
    cudaMalloc(&buf, maxb); cudaMemset(buf, 1, maxb); cudaMalloc(&sink, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (uint64_t mb : {32ull, 128ull, 687ull, 2688ull}) {
        const uint64_t bytes = mb << 20;
        for (int blocks : {148 * 8, 148 * 16, 148 * 32}) {
            float best8 = 1e9;
            for (int it = 0; it < 4; it++) {
                float ms;
                cudaEventRecord(e0); gather<8, 8><<<blocks, 256>>>(buf, bytes / 32, loads, it, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms, e0, e1); if (it && ms < best8) best8 = ms;
            }
            float bestp = 1e9, bestg = 1e9;
            for (int it = 0; it < 3; it++) {
                float ms;
                cudaEventRecord(e0); prefetch_only<8><<<blocks, 256>>>(buf, bytes / 32, loads, it + 77); cudaEventRecord(e1); cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms, e0, e1); if (it && ms < bestp) bestp = ms;
                cudaEventRecord(e0); gather_pf<8><<<blocks, 256>>>(buf, bytes / 32, loads, it + 99, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms, e0, e1); if (it && ms < bestg) bestg = ms;
            }
            printf("working set %4llu MB blocks=%d: demand %.1f Gsect/s (%.2f ms) | prefetch-only %.1f G/s (%.2f ms) | demand+prefetch-ahead %.1f G/s (%.2f ms)\n",
                   (unsigned long long)mb, blocks, loads / best8 / 1e6, best8, loads / bestp / 1e6, bestp, loads / bestg / 1e6, bestg);
        }
    }
    return 0;
}
