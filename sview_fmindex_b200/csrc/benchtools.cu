// benchtools.cu -- measurement / self-check helpers (include/svfm_bench.h).  Not part of the product path: built into its
// own library, libsvfm_bench.so, which shares nothing with libsvfm.so but device pointers passed through the caller.
#include "common.cuh"
#include "../../include/svfm_bench.h"

namespace svfm {
thread_local std::string g_last_error;   // this library's own copies (common.cuh declares them)
std::atomic<uint64_t> g_launches{0};
namespace {

__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__host__ __device__ inline uint64_t hash_at(uint64_t seed, uint64_t i) {
    return splitmix64(seed + 0x9E3779B97F4A7C15ull * (i + 1));
}

__global__ void synth_text_kernel(uint8_t* text, uint64_t n, uint64_t seed, const uint8_t* alphabet, uint32_t alen,
                                  uint32_t rare, uint8_t rare_byte) {
    __shared__ uint8_t s_alpha[256];
    for (int i = threadIdx.x; i < (int)alen; i += blockDim.x) s_alpha[i] = alphabet[i];
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t h = hash_at(seed, i);
        uint8_t c = s_alpha[((h >> 32) * alen) >> 32];
        if (rare && hash_at(seed ^ 0xA5A5A5A5ull, i) % rare == 0) c = rare_byte;
        text[i] = c;
    }
}

__global__ void synth_patterns_kernel(const uint8_t* text, uint64_t n, uint8_t* pats, uint64_t* starts, uint64_t count,
                                      uint32_t len, uint64_t seed) {
    // one thread per pattern byte: coalesced stores
    const uint64_t total = count * len;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = t / len;
        const uint32_t j = (uint32_t)(t - i * len);
        const uint64_t s = hash_at(seed, i) % (n - len + 1);
        pats[t] = __ldg(text + s + j);
        if (starts && j == 0) starts[i] = s;
    }
}

constexpr int GATHER_ILP = 8;
__global__ void gather32_kernel(const uint8_t* __restrict__ buf, uint64_t sectors, uint64_t loads, uint64_t seed,
                                unsigned long long* sink) {
    // each thread keeps GATHER_ILP independent loads in flight, one 8-byte load per uniformly random 32-byte
    // sector (what one rank query reads of a checkpoint row); the sector is the unit of DRAM/L2 traffic
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;
    uint64_t acc = 0;
    for (uint64_t base = tid * GATHER_ILP; base < loads; base += nthreads * GATHER_ILP) {
        uint64_t v[GATHER_ILP];
#pragma unroll
        for (int u = 0; u < GATHER_ILP; u++) {
            // multiply-shift range reduction (a 64-bit modulo would make the kernel ALU-bound)
            const uint64_t s = __umul64hi(splitmix64(seed + base + u), sectors);
            v[u] = __ldg(reinterpret_cast<const unsigned long long*>(buf + s * 32));
        }
#pragma unroll
        for (int u = 0; u < GATHER_ILP; u++) acc ^= v[u];
    }
    if (acc == 0x12345679u) atomicAdd(sink, 1ull);
}

template <class P>
__global__ void verify_locate_kernel(const uint8_t* text, uint64_t n, const uint8_t* pats, uint32_t len, uint64_t n_pats,
                                     const uint64_t* starts, const uint64_t* out_offs, const P* positions,
                                     const uint8_t* table, unsigned long long* out /* [0..2] violations, [3] digest */) {
    __shared__ uint8_t s_table[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_table[i] = table ? table[i] : (uint8_t)i;
    __syncthreads();
    unsigned long long bad_pos = 0, missing = 0, empty = 0, digest = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pats; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t a = out_offs[i], b = out_offs[i + 1];
        if (a == b) empty++;
        bool found = starts == nullptr;
        const uint8_t* pat = pats + i * (uint64_t)len;
        for (uint64_t t = a; t < b; t++) {
            const uint64_t p = positions[t];
            digest += (p + 1) * (2 * i + 1);
            if (starts && p == starts[i]) found = true;
            bool ok = p + len <= n;
            for (uint32_t j = 0; ok && j < len; j++) ok = s_table[text[p + j]] == s_table[pat[j]];
            if (!ok) bad_pos++;
        }
        if (!found) missing++;
    }
    if (bad_pos) atomicAdd(&out[0], bad_pos);
    if (missing) atomicAdd(&out[1], missing);
    if (empty) atomicAdd(&out[2], empty);
    atomicAdd(&out[3], digest);
}

template <class P>
__global__ void count_digest_kernel(const P* counts, uint64_t n, unsigned long long* out) {
    unsigned long long sum = 0, digest = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long c = counts[i];
        sum += c;
        digest += c * (2 * i + 1);
    }
    atomicAdd(&out[0], sum);
    atomicAdd(&out[1], digest);
}

__global__ void flush_kernel(uint4* buf, uint64_t n16, uint32_t v) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x)
        buf[i] = make_uint4(v, v, v, v);
}

}  // namespace
}  // namespace svfm

using namespace svfm;

extern "C" {

int svfm_bench_synth_text(uint8_t* d_text, uint64_t n, uint64_t seed, const uint8_t* alphabet, uint32_t alphabet_len,
                          uint32_t rare, uint8_t rare_byte, void* stream) {
    if (!d_text || !alphabet || alphabet_len == 0 || alphabet_len > 256) return SVFM_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* d_alpha;
    SVFM_CUDA(cudaMalloc(&d_alpha, 256));
    SVFM_CUDA(cudaMemcpyAsync(d_alpha, alphabet, alphabet_len, cudaMemcpyHostToDevice, st));
    synth_text_kernel<<<148 * 16, 256, 0, st>>>(d_text, n, seed, d_alpha, alphabet_len, rare, rare_byte);
    g_launches++;
    SVFM_CUDA(cudaGetLastError());
    SVFM_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_alpha);
    return SVFM_OK;
}

int svfm_bench_synth_patterns(const uint8_t* d_text, uint64_t n, uint8_t* d_pats, uint64_t* d_starts, uint64_t count,
                              uint32_t len, uint64_t seed, void* stream) {
    if (!d_text || !d_pats || len == 0 || len > n) return SVFM_ERR_BAD_ARG;
    synth_patterns_kernel<<<148 * 16, 256, 0, (cudaStream_t)stream>>>(d_text, n, d_pats, d_starts, count, len, seed);
    g_launches++;
    SVFM_CUDA(cudaGetLastError());
    return SVFM_OK;
}

int svfm_bench_gather32(const uint8_t* d_buf, uint64_t bytes, uint64_t loads, uint32_t iters, uint64_t seed,
                        double* sectors_per_s, double* ms_best, void* stream) {
    if (!d_buf || bytes < 64 || !sectors_per_s || ((uintptr_t)d_buf % 32) != 0) return SVFM_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* d_sink;
    SVFM_CUDA(cudaMalloc(&d_sink, 8));
    SVFM_CUDA(cudaMemsetAsync(d_sink, 0, 8, st));
    cudaEvent_t e0, e1;
    SVFM_CUDA(cudaEventCreate(&e0));
    SVFM_CUDA(cudaEventCreate(&e1));
    const uint64_t sectors = bytes / 32;
    float best = 1e30f;
    for (uint32_t it = 0; it <= iters; it++) {
        SVFM_CUDA(cudaEventRecord(e0, st));
        gather32_kernel<<<148 * 16, 256, 0, st>>>(d_buf, sectors, loads, seed + it, d_sink);
        g_launches++;
        SVFM_CUDA(cudaEventRecord(e1, st));
        SVFM_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        SVFM_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (it > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_sink);
    *sectors_per_s = (double)loads / ((double)best * 1e-3);
    if (ms_best) *ms_best = best;
    return SVFM_OK;
}

int svfm_bench_verify_locate(const uint8_t* d_text, uint64_t n, const uint8_t* d_pats, uint32_t len, uint64_t n_pats,
                             const uint64_t* d_starts, const uint64_t* d_out_offs, const void* d_positions,
                             uint32_t pos_bits, const uint8_t* table256, uint64_t violations[3], uint64_t* digest,
                             void* stream) {
    if (!d_text || !d_pats || !d_out_offs || !violations) return SVFM_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* d_out;
    uint8_t* d_table = nullptr;
    SVFM_CUDA(cudaMalloc(&d_out, 4 * 8));
    SVFM_CUDA(cudaMemsetAsync(d_out, 0, 4 * 8, st));
    if (table256) {
        SVFM_CUDA(cudaMalloc(&d_table, 256));
        SVFM_CUDA(cudaMemcpyAsync(d_table, table256, 256, cudaMemcpyHostToDevice, st));
    }
    if (pos_bits == 32)
        verify_locate_kernel<uint32_t><<<148 * 16, 256, 0, st>>>(d_text, n, d_pats, len, n_pats, d_starts, d_out_offs,
                                                               (const uint32_t*)d_positions, d_table, d_out);
    else
        verify_locate_kernel<uint64_t><<<148 * 16, 256, 0, st>>>(d_text, n, d_pats, len, n_pats, d_starts, d_out_offs,
                                                               (const uint64_t*)d_positions, d_table, d_out);
    g_launches++;
    unsigned long long h[4];
    SVFM_CUDA(cudaMemcpyAsync(h, d_out, sizeof(h), cudaMemcpyDeviceToHost, st));
    SVFM_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_out);
    if (d_table) cudaFree(d_table);
    violations[0] = h[0]; violations[1] = h[1]; violations[2] = h[2];
    if (digest) *digest = h[3];
    return SVFM_OK;
}

int svfm_bench_count_digest(const void* d_counts, uint32_t pos_bits, uint64_t n, uint64_t* sum, uint64_t* digest,
                            void* stream) {
    if (!d_counts) return SVFM_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* d_out;
    SVFM_CUDA(cudaMalloc(&d_out, 2 * 8));
    SVFM_CUDA(cudaMemsetAsync(d_out, 0, 2 * 8, st));
    if (pos_bits == 32) count_digest_kernel<uint32_t><<<148 * 8, 256, 0, st>>>((const uint32_t*)d_counts, n, d_out);
    else count_digest_kernel<uint64_t><<<148 * 8, 256, 0, st>>>((const uint64_t*)d_counts, n, d_out);
    g_launches++;
    unsigned long long h[2];
    SVFM_CUDA(cudaMemcpyAsync(h, d_out, sizeof(h), cudaMemcpyDeviceToHost, st));
    SVFM_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_out);
    if (sum) *sum = h[0];
    if (digest) *digest = h[1];
    return SVFM_OK;
}

int svfm_bench_flush_l2(uint8_t* d_buf, uint64_t bytes, void* stream) {
    if (!d_buf) return SVFM_ERR_BAD_ARG;
    static uint32_t v = 1;
    flush_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<uint4*>(d_buf), bytes / 16, v++);
    g_launches++;
    SVFM_CUDA(cudaGetLastError());
    return SVFM_OK;
}

}  // extern "C"

extern "C" const char* svfm_bench_last_error(void) { return svfm::g_last_error.c_str(); }
